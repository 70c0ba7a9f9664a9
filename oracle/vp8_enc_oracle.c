/*
 * vp8_enc_oracle.c - CPU restatement of the reference ENCODER's in-loop reconstruction for the two whole-macroblock
 * prediction modes of its front end:
 *
 *   enc_vp8_encode_dc_pred_inloop        /root/reference/src/enc-m08_recon/enc_recon.c:855-1085
 *   enc_vp8_encode_i16x16_uv_sad_inloop  /root/reference/src/enc-m08_recon/enc_recon.c:1189-1483
 *
 * i.e. per macroblock, in raster order: predict from the RECONSTRUCTED neighbours (DC only, or the best of DC/V/H/TM by
 * squared error against the source), forward DCT of source minus prediction (enc-m05_intra/enc_transform.c:7-44), forward
 * WHT of the sixteen luma DCs (:46-72), quantisation (enc-m06_quant/enc_quant.c:62-83), then the decoder's own
 * dequantisation / inverse WHT / inverse DCT / add (enc_recon.c:723-788, 848-853) so that the next macroblock predicts
 * from what a decoder will see.
 *
 * TEST INFRASTRUCTURE ONLY (see vp8_oracle.h). Written as one routine over a table of three planes rather than the
 * reference's per-plane code. Pinned against the reference library (oracle/_ref/libref_enc.so) by tests/test_enc_oracle.py
 * and, where /root/reference is absent, against the digests in tests/golden/enc.json.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vp8_oracle.h"

/* RFC 6386 14.1 tables (the same data as enc_quant.c:15-36) and the quality -> qindex map (enc_quality_table.c:5-13). */
static const uint16_t kDc[128] = {
    4,   5,   6,   7,   8,   9,   10,  10,  11,  12,  13,  14,  15,  16,  17,  17,  18,  19,  20,  20,  21,  21,  22,  22,  23,  23,
    24,  25,  25,  26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  46,
    47,  48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  59,  60,  61,  62,  63,  64,  65,  66,  67,  68,  69,  70,  71,  72,
    73,  74,  75,  76,  76,  77,  78,  79,  80,  81,  82,  83,  84,  85,  86,  87,  88,  89,  91,  93,  95,  96,  98,  100, 101, 102,
    104, 106, 108, 110, 112, 114, 116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151, 154, 157};
static const uint16_t kAc[128] = {
    4,   5,   6,   7,   8,   9,   10,  11,  12,  13,  14,  15,  16,  17,  18,  19,  20,  21,  22,  23,  24,  25,  26,  27,  28,  29,
    30,  31,  32,  33,  34,  35,  36,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  47,  48,  49,  50,  51,  52,  53,  54,  55,
    56,  57,  58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,  82,  84,  86,  88,  90,  92,  94,  96,  98,  100, 102, 104,
    106, 108, 110, 112, 114, 116, 119, 122, 125, 128, 131, 134, 137, 140, 143, 146, 149, 152, 155, 158, 161, 164, 167, 170, 173, 177,
    181, 185, 189, 193, 197, 201, 205, 209, 213, 217, 221, 225, 229, 234, 239, 245, 249, 254, 259, 264, 269, 274, 279, 284};
static const uint8_t kQIndexOfQuality[101] = {
    127, 103, 96, 92, 89, 86, 83, 81, 79, 77, 75, 73, 72, 70, 69, 68, 66, 65, 64, 63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 51, 50,
    49,  48,  48, 47, 46, 45, 45, 44, 43, 43, 42, 41, 41, 40, 40, 39, 38, 38, 37, 37, 36, 36, 35, 35, 34, 33, 33, 32, 32, 31, 31, 30, 30, 29,
    29,  28,  28, 28, 27, 27, 26, 26, 24, 23, 22, 21, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9,  8,  7,  6,  5,  4,  3,  2,  1,  0,  0};

/* {y1dc, y1ac, uvdc, uvac, y2dc, y2ac} for a quality, all deltas zero (enc_quant.c:45-60). */
int orc_enc_quant(int quality, int q[6]) {
	if (quality < 0) quality = 0;
	if (quality > 100) quality = 100;
	const int qi = kQIndexOfQuality[quality];
	q[0] = kDc[qi];
	q[1] = kAc[qi];
	q[2] = kDc[qi] > 132 ? 132 : kDc[qi];
	q[3] = kAc[qi];
	q[4] = kDc[qi] * 2;
	q[5] = kAc[qi] * 155 / 100;
	if (q[5] < 8) q[5] = 8;
	return qi;
}

static int clip255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* round-half-away division, saturated to int16 (enc_quant.c:62-76) */
static int16_t quantise(int c, int step) {
	if (step <= 0) return 0;
	const int mag = (abs(c) + (step >> 1)) / step;
	const int r = c < 0 ? -mag : mag;
	return (int16_t)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
}

/* d[16] = source - prediction, raster; libwebp-style forward DCT (enc_transform.c:7-44) */
static void fdct(const int d[16], int16_t out[16]) {
	int t[16];
	for (int r = 0; r < 4; r++) {
		const int s03 = d[4 * r] + d[4 * r + 3], s12 = d[4 * r + 1] + d[4 * r + 2];
		const int d12 = d[4 * r + 1] - d[4 * r + 2], d03 = d[4 * r] - d[4 * r + 3];
		t[4 * r] = (s03 + s12) * 8;
		t[4 * r + 1] = (d12 * 2217 + d03 * 5352 + 1812) >> 9;
		t[4 * r + 2] = (s03 - s12) * 8;
		t[4 * r + 3] = (d03 * 2217 - d12 * 5352 + 937) >> 9;
	}
	for (int c = 0; c < 4; c++) {
		const int s03 = t[c] + t[12 + c], s12 = t[4 + c] + t[8 + c];
		const int d12 = t[4 + c] - t[8 + c], d03 = t[c] - t[12 + c];
		out[c] = (int16_t)((s03 + s12 + 7) >> 4);
		out[4 + c] = (int16_t)(((d12 * 2217 + d03 * 5352 + 12000) >> 16) + (d03 != 0));
		out[8 + c] = (int16_t)((s03 - s12 + 7) >> 4);
		out[12 + c] = (int16_t)((d03 * 2217 - d12 * 5352 + 51000) >> 16);
	}
}

/* forward WHT over the DCs of the sixteen luma blocks, dc[n] = block n in raster order (enc_transform.c:46-72) */
static void fwht(const int16_t dc[16], int16_t out[16]) {
	int t[16];
	for (int r = 0; r < 4; r++) {
		const int a0 = dc[4 * r] + dc[4 * r + 2], a1 = dc[4 * r + 1] + dc[4 * r + 3];
		const int a2 = dc[4 * r + 1] - dc[4 * r + 3], a3 = dc[4 * r] - dc[4 * r + 2];
		t[4 * r] = a0 + a1;
		t[4 * r + 1] = a3 + a2;
		t[4 * r + 2] = a3 - a2;
		t[4 * r + 3] = a0 - a1;
	}
	for (int c = 0; c < 4; c++) {
		const int a0 = t[c] + t[8 + c], a1 = t[4 + c] + t[12 + c];
		const int a2 = t[4 + c] - t[12 + c], a3 = t[c] - t[8 + c];
		out[c] = (int16_t)((a0 + a1) >> 1);
		out[4 + c] = (int16_t)((a3 + a2) >> 1);
		out[8 + c] = (int16_t)((a3 - a2) >> 1);
		out[12 + c] = (int16_t)((a0 - a1) >> 1);
	}
}

/* the decoder's inverse transforms with their int16 stores (enc_recon.c:723-788) */
static void iwht(const int16_t in[16], int16_t out[16]) {
	int16_t t[16];
	for (int c = 0; c < 4; c++) {
		const int a = in[c] + in[12 + c], b = in[4 + c] + in[8 + c], cc = in[4 + c] - in[8 + c], d = in[c] - in[12 + c];
		t[c] = (int16_t)(a + b);
		t[4 + c] = (int16_t)(cc + d);
		t[8 + c] = (int16_t)(a - b);
		t[12 + c] = (int16_t)(d - cc);
	}
	for (int r = 0; r < 4; r++) {
		const int a = t[4 * r] + t[4 * r + 3], b = t[4 * r + 1] + t[4 * r + 2], cc = t[4 * r + 1] - t[4 * r + 2], d = t[4 * r] - t[4 * r + 3];
		out[4 * r] = (int16_t)((a + b + 3) >> 3);
		out[4 * r + 1] = (int16_t)((cc + d + 3) >> 3);
		out[4 * r + 2] = (int16_t)((a - b + 3) >> 3);
		out[4 * r + 3] = (int16_t)((d - cc + 3) >> 3);
	}
}

static void butterfly(int x0, int x1, int x2, int x3, int o[4]) {
	const int e = x0 + x2, g = x0 - x2;
	const int lo = ((x1 * 35468) >> 16) - (x3 + ((x3 * 20091) >> 16));
	const int hi = (x1 + ((x1 * 20091) >> 16)) + ((x3 * 35468) >> 16);
	o[0] = e + hi;
	o[1] = g + lo;
	o[2] = g - lo;
	o[3] = e - hi;
}

static void idct(const int16_t in[16], int16_t out[16]) {
	int16_t t[16];
	int o[4];
	for (int c = 0; c < 4; c++) {
		butterfly(in[c], in[4 + c], in[8 + c], in[12 + c], o);
		for (int k = 0; k < 4; k++) t[4 * k + c] = (int16_t)o[k];
	}
	for (int r = 0; r < 4; r++) {
		butterfly(t[4 * r], t[4 * r + 1], t[4 * r + 2], t[4 * r + 3], o);
		for (int k = 0; k < 4; k++) out[4 * r + k] = (int16_t)((o[k] + 4) >> 3);
	}
}

typedef struct {
	const uint8_t* src; /* source plane, cropped size; reads beyond the frame repeat the last row / column (enc_recon.c:563-582) */
	uint32_t src_stride, src_w, src_h;
	uint8_t* rec; /* reconstruction, macroblock-aligned (enc_recon.c:790-838) */
	uint32_t rec_stride;
	int n; /* 16 luma, 8 chroma */
} EncPlane;

static int src_px(const EncPlane* p, uint32_t x, uint32_t y) {
	if (x >= p->src_w) x = p->src_w - 1;
	if (y >= p->src_h) y = p->src_h - 1;
	return p->src[(size_t)y * p->src_stride + x];
}

/* n x n prediction of the block at (x0, y0) from the reconstructed row above / column left, whole-macroblock modes
 * 0 DC, 1 V, 2 H, 3 TM with the frame-border rules of enc_recon.c:338-409 (127 above, 129 left). */
static void predict(const EncPlane* p, uint32_t x0, uint32_t y0, int mode, uint8_t* pred) {
	const int n = p->n, have_a = y0 > 0, have_l = x0 > 0;
	int A[16], L[16];
	for (int i = 0; i < n; i++) {
		A[i] = have_a ? p->rec[(size_t)(y0 - 1) * p->rec_stride + x0 + i] : 127;
		L[i] = have_l ? p->rec[(size_t)(y0 + i) * p->rec_stride + x0 - 1] : 129;
	}
	const int P = (have_a && have_l) ? p->rec[(size_t)(y0 - 1) * p->rec_stride + x0 - 1] : (have_a ? 129 : 127);
	int dc = 128;
	if (have_a || have_l) {
		int s = 0;
		for (int i = 0; i < n; i++) s += (have_a ? A[i] : 0) + (have_l ? L[i] : 0);
		if (!(have_a && have_l)) s += s;
		dc = (s + n) >> (n == 16 ? 5 : 4);
	}
	for (int r = 0; r < n; r++)
		for (int c = 0; c < n; c++) {
			int v;
			switch (mode) {
				case 1: v = A[c]; break;
				case 2: v = L[r]; break;
				case 3: v = clip255(L[r] + A[c] - P); break;
				default: v = dc; break;
			}
			pred[r * n + c] = (uint8_t)v;
		}
}

static uint32_t sq_error(const EncPlane* p, uint32_t x0, uint32_t y0, const uint8_t* pred) {
	uint32_t e = 0;
	for (int r = 0; r < p->n; r++)
		for (int c = 0; c < p->n; c++) {
			const int d = src_px(p, x0 + c, y0 + r) - pred[r * p->n + c];
			e += (uint32_t)(d * d);
		}
	return e;
}

/* Encodes one picture. search = 0: DC prediction everywhere (dc_pred_inloop); 1: best of the four modes by squared
 * error, luma and chroma (U+V together) separately, first minimum wins (i16x16_uv_sad_inloop).
 * coeffs: mb_total * 400 int16, per macroblock Y2[16] Y[16][16] U[4][16] V[4][16]; y_modes / uv_modes: mb_total bytes (may be
 * NULL when search = 0); rec_y/u/v: macroblock-aligned planes (stride 16*mb_cols / 8*mb_cols), may be NULL.
 * Returns the qindex, or -1. */
int orc_enc_i16_inloop(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint32_t width, uint32_t height, uint32_t y_stride,
                       uint32_t uv_stride, int quality, int search, uint8_t* y_modes, uint8_t* uv_modes, int16_t* coeffs, uint8_t* rec_y,
                       uint8_t* rec_u, uint8_t* rec_v) {
	if (!y || !u || !v || !coeffs || width == 0 || height == 0) return -1;
	const uint32_t cols = (width + 15) >> 4, rows = (height + 15) >> 4;
	int q[6];
	const int qindex = orc_enc_quant(quality, q);
	uint8_t* own[3] = {NULL, NULL, NULL};
	if (!rec_y) rec_y = own[0] = (uint8_t*)malloc((size_t)cols * 16 * rows * 16);
	if (!rec_u) rec_u = own[1] = (uint8_t*)malloc((size_t)cols * 8 * rows * 8);
	if (!rec_v) rec_v = own[2] = (uint8_t*)malloc((size_t)cols * 8 * rows * 8);
	if (!rec_y || !rec_u || !rec_v) {
		for (int i = 0; i < 3; i++) free(own[i]);
		return -1;
	}
	EncPlane pl[3] = {{y, y_stride, width, height, rec_y, cols * 16, 16},
	                  {u, uv_stride, (width + 1) >> 1, (height + 1) >> 1, rec_u, cols * 8, 8},
	                  {v, uv_stride, (width + 1) >> 1, (height + 1) >> 1, rec_v, cols * 8, 8}};
	uint8_t pred[3][256], trial[2][256];
	for (uint32_t my = 0; my < rows; my++)
		for (uint32_t mx = 0; mx < cols; mx++) {
			const size_t mb = (size_t)my * cols + mx;
			int16_t* out = coeffs + mb * 400;
			int ym = 0, cm = 0;
			if (search) {
				uint32_t best = 0xffffffffu;
				for (int m = 0; m < 4; m++) {
					predict(&pl[0], mx * 16, my * 16, m, trial[0]);
					const uint32_t e = sq_error(&pl[0], mx * 16, my * 16, trial[0]);
					if (e < best) best = e, ym = m;
				}
				best = 0xffffffffu;
				for (int m = 0; m < 4; m++) {
					predict(&pl[1], mx * 8, my * 8, m, trial[0]);
					predict(&pl[2], mx * 8, my * 8, m, trial[1]);
					const uint32_t e = sq_error(&pl[1], mx * 8, my * 8, trial[0]) + sq_error(&pl[2], mx * 8, my * 8, trial[1]);
					if (e < best) best = e, cm = m;
				}
			}
			if (y_modes) y_modes[mb] = (uint8_t)ym;
			if (uv_modes) uv_modes[mb] = (uint8_t)cm;
			for (int k = 0; k < 3; k++) predict(&pl[k], mx * pl[k].n, my * pl[k].n, k ? cm : ym, pred[k]);

			/* forward: 16 + 4 + 4 blocks, luma DCs through the WHT */
			int16_t* blk[3] = {out + 16, out + 16 + 256, out + 16 + 256 + 64};
			int16_t dcs[16];
			for (int k = 0; k < 3; k++) {
				const int n = pl[k].n, per_row = n / 4;
				for (int b = 0; b < per_row * per_row; b++) {
					const int bx = (b % per_row) * 4, by = (b / per_row) * 4;
					int d[16];
					for (int r = 0; r < 4; r++)
						for (int c = 0; c < 4; c++)
							d[4 * r + c] = src_px(&pl[k], mx * n + bx + c, my * n + by + r) - pred[k][(by + r) * n + bx + c];
					fdct(d, blk[k] + 16 * b);
					if (k == 0) {
						dcs[b] = blk[0][16 * b];
						blk[0][16 * b] = 0;
					}
					for (int i = 0; i < 16; i++) blk[k][16 * b + i] = quantise(blk[k][16 * b + i], k ? (i ? q[3] : q[2]) : (i ? q[1] : q[0]));
				}
			}
			fwht(dcs, out);
			for (int i = 0; i < 16; i++) out[i] = quantise(out[i], i ? q[5] : q[4]);

			/* backward, as a decoder will: Y2 -> luma DCs, then every block */
			int16_t y2d[16], ydc[16];
			for (int i = 0; i < 16; i++) y2d[i] = (int16_t)(out[i] * (i ? q[5] : q[4]));
			iwht(y2d, ydc);
			for (int k = 0; k < 3; k++) {
				const int n = pl[k].n, per_row = n / 4;
				for (int b = 0; b < per_row * per_row; b++) {
					const int bx = (b % per_row) * 4, by = (b / per_row) * 4;
					int16_t c16[16], res[16];
					for (int i = 0; i < 16; i++) {
						const int raw = (k == 0 && i == 0) ? ydc[b] : blk[k][16 * b + i];
						c16[i] = (int16_t)(raw * (k ? (i ? q[3] : q[2]) : (i ? q[1] : q[0])));
					}
					idct(c16, res);
					for (int r = 0; r < 4; r++)
						for (int c = 0; c < 4; c++)
							pl[k].rec[(size_t)(my * n + by + r) * pl[k].rec_stride + mx * n + bx + c] =
							    (uint8_t)clip255(pred[k][(by + r) * n + bx + c] + res[4 * r + c]);
				}
			}
		}
	for (int i = 0; i < 3; i++) free(own[i]);
	return qindex;
}

/* ------------------------------------------------------------------------------------------------------------------
 * enc_vp8_encode_bpred_uv_sad_inloop (/root/reference/src/enc-m08_recon/enc_recon.c:1507-1831): every macroblock B_PRED.
 * Chroma: best whole-block mode by SAD (U + V). Luma: the sixteen sub-blocks in raster order, each the best of the ten
 * 4x4 modes by SAD against predictors built from the reconstruction so far (:197-336), then forward DCT, quantisation with
 * the y1 steps (no Y2 block: its sixteen coefficients stay zero), dequantisation, inverse DCT, add. */
static const uint8_t kTap[8][16] = { /* modes 2..9: first tap over E[16], | 0x10 for the two-tap average */
    {6, 7, 8, 9, 6, 7, 8, 9, 6, 7, 8, 9, 6, 7, 8, 9},
    {4, 4, 4, 4, 3, 3, 3, 3, 2, 2, 2, 2, 1, 1, 1, 1},
    {7, 8, 9, 10, 8, 9, 10, 11, 9, 10, 11, 12, 10, 11, 12, 13},
    {5, 6, 7, 8, 4, 5, 6, 7, 3, 4, 5, 6, 2, 3, 4, 5},
    {0x16, 0x17, 0x18, 0x19, 5, 6, 7, 8, 4, 0x16, 0x17, 0x18, 3, 5, 6, 7},
    {0x17, 0x18, 0x19, 0x1a, 7, 8, 9, 10, 0x18, 0x19, 0x1a, 11, 8, 9, 10, 12},
    {0x15, 5, 6, 7, 0x14, 4, 0x15, 5, 0x13, 3, 0x14, 4, 0x12, 2, 0x13, 3},
    {0x14, 3, 0x13, 2, 0x13, 2, 0x12, 1, 0x12, 1, 0, 0, 0, 0, 0, 0},
};

/* E[0..2] = L3, E[3] = L2, E[4] = L1, E[5] = L0, E[6] = P, E[7..14] = A0..A7, E[15] = A7 */
static void sub_predict(const uint8_t E[16], int mode, uint8_t pred[16]) {
	for (int p = 0; p < 16; p++) {
		int v;
		if (mode == 0) v = (E[7] + E[8] + E[9] + E[10] + E[2] + E[3] + E[4] + E[5] + 4) >> 3;
		else if (mode == 1) v = clip255(E[5 - (p >> 2)] + E[7 + (p & 3)] - E[6]);
		else {
			const int t = kTap[mode - 2][p], i = t & 15;
			v = (t & 16) ? (E[i] + E[i + 1] + 1) >> 1 : (E[i] + 2 * E[i + 1] + E[i + 2] + 2) >> 2;
		}
		pred[p] = (uint8_t)v;
	}
}

static uint32_t abs_error(const EncPlane* p, uint32_t x0, uint32_t y0, int n, const uint8_t* pred) {
	uint32_t e = 0;
	for (int r = 0; r < n; r++)
		for (int c = 0; c < n; c++) e += (uint32_t)abs(src_px(p, x0 + c, y0 + r) - pred[r * n + c]);
	return e;
}

/* b_modes: mb_total * 16 bytes; y_modes (always 4) and uv_modes: mb_total bytes. Returns the qindex, or -1. */
int orc_enc_bpred_inloop(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint32_t width, uint32_t height, uint32_t y_stride,
                         uint32_t uv_stride, int quality, uint8_t* y_modes, uint8_t* b_modes, uint8_t* uv_modes, int16_t* coeffs,
                         uint8_t* rec_y, uint8_t* rec_u, uint8_t* rec_v) {
	if (!y || !u || !v || !coeffs || !b_modes || width == 0 || height == 0) return -1;
	const uint32_t cols = (width + 15) >> 4, rows = (height + 15) >> 4;
	int q[6];
	const int qindex = orc_enc_quant(quality, q);
	uint8_t* own[3] = {NULL, NULL, NULL};
	if (!rec_y) rec_y = own[0] = (uint8_t*)malloc((size_t)cols * 16 * rows * 16);
	if (!rec_u) rec_u = own[1] = (uint8_t*)malloc((size_t)cols * 8 * rows * 8);
	if (!rec_v) rec_v = own[2] = (uint8_t*)malloc((size_t)cols * 8 * rows * 8);
	if (!rec_y || !rec_u || !rec_v) {
		for (int i = 0; i < 3; i++) free(own[i]);
		return -1;
	}
	EncPlane pl[3] = {{y, y_stride, width, height, rec_y, cols * 16, 16},
	                  {u, uv_stride, (width + 1) >> 1, (height + 1) >> 1, rec_u, cols * 8, 8},
	                  {v, uv_stride, (width + 1) >> 1, (height + 1) >> 1, rec_v, cols * 8, 8}};
	const uint32_t ys = cols * 16;
	for (uint32_t my = 0; my < rows; my++)
		for (uint32_t mx = 0; mx < cols; mx++) {
			const size_t mb = (size_t)my * cols + mx;
			int16_t* out = coeffs + mb * 400;
			memset(out, 0, 400 * sizeof(int16_t));
			if (y_modes) y_modes[mb] = 4;
			/* chroma mode: sum of absolute differences over U and V, first minimum */
			uint8_t pc[2][64], trial[2][64];
			int cm = 0;
			uint32_t best = 0xffffffffu;
			for (int m = 0; m < 4; m++) {
				predict(&pl[1], mx * 8, my * 8, m, trial[0]);
				predict(&pl[2], mx * 8, my * 8, m, trial[1]);
				const uint32_t e = abs_error(&pl[1], mx * 8, my * 8, 8, trial[0]) + abs_error(&pl[2], mx * 8, my * 8, 8, trial[1]);
				if (e < best) best = e, cm = m;
			}
			if (uv_modes) uv_modes[mb] = (uint8_t)cm;
			predict(&pl[1], mx * 8, my * 8, cm, pc[0]);
			predict(&pl[2], mx * 8, my * 8, cm, pc[1]);
			/* luma sub-blocks */
			for (int sb = 0; sb < 16; sb++) {
				const uint32_t sx = mx * 16 + (sb & 3) * 4, sy = my * 16 + (sb >> 2) * 4;
				uint8_t E[16];
				for (int i = 0; i < 4; i++) E[5 - i] = sx ? rec_y[(size_t)(sy + i) * ys + sx - 1] : 129;
				E[0] = E[1] = E[2];
				E[6] = sy == 0 ? 127 : (sx == 0 ? 129 : rec_y[(size_t)(sy - 1) * ys + sx - 1]);
				for (uint32_t i = 0; i < 8; i++) {
					int a = 127;
					if (sy > 0) {
						uint32_t row = sy - 1, col = sx + i;
						if ((sb & 3) == 3 && i >= 4) { /* above-right of the last column: always from the macroblock row above */
							row = my * 16 - 1;
							col = mx * 16 + 16 + (i - 4);
						}
						if (col >= ys) col = ys - 1;
						a = ((sb & 3) == 3 && i >= 4 && my == 0) ? 127 : rec_y[(size_t)row * ys + col];
					}
					E[7 + i] = (uint8_t)a;
				}
				E[15] = E[14];
				uint8_t src[16], pred[16], cand[16];
				for (int p = 0; p < 16; p++) src[p] = (uint8_t)src_px(&pl[0], sx + (p & 3), sy + (p >> 2));
				int bm = 0;
				best = 0xffffffffu;
				for (int m = 0; m < 10; m++) {
					sub_predict(E, m, cand);
					uint32_t e = 0;
					for (int p = 0; p < 16; p++) e += (uint32_t)abs(src[p] - cand[p]);
					if (e < best) best = e, bm = m;
				}
				b_modes[mb * 16 + sb] = (uint8_t)bm;
				sub_predict(E, bm, pred);
				int d[16];
				int16_t* c16 = out + 16 + 16 * sb, deq[16], res[16];
				for (int p = 0; p < 16; p++) d[p] = src[p] - pred[p];
				fdct(d, c16);
				for (int i = 0; i < 16; i++) {
					c16[i] = quantise(c16[i], i ? q[1] : q[0]);
					deq[i] = (int16_t)(c16[i] * (i ? q[1] : q[0]));
				}
				idct(deq, res);
				for (int p = 0; p < 16; p++) rec_y[(size_t)(sy + (p >> 2)) * ys + sx + (p & 3)] = (uint8_t)clip255(pred[p] + res[p]);
			}
			/* chroma blocks */
			for (int k = 1; k < 3; k++)
				for (int b = 0; b < 4; b++) {
					const int bx = (b & 1) * 4, by = (b >> 1) * 4;
					int d[16];
					int16_t* c16 = out + 16 + 256 + 64 * (k - 1) + 16 * b, deq[16], res[16];
					for (int r = 0; r < 4; r++)
						for (int c = 0; c < 4; c++) d[4 * r + c] = src_px(&pl[k], mx * 8 + bx + c, my * 8 + by + r) - pc[k - 1][(by + r) * 8 + bx + c];
					fdct(d, c16);
					for (int i = 0; i < 16; i++) {
						c16[i] = quantise(c16[i], i ? q[3] : q[2]);
						deq[i] = (int16_t)(c16[i] * (i ? q[3] : q[2]));
					}
					idct(deq, res);
					for (int r = 0; r < 4; r++)
						for (int c = 0; c < 4; c++)
							pl[k].rec[(size_t)(my * 8 + by + r) * pl[k].rec_stride + mx * 8 + bx + c] =
							    (uint8_t)clip255(pc[k - 1][(by + r) * 8 + bx + c] + res[4 * r + c]);
				}
		}
	for (int i = 0; i < 3; i++) free(own[i]);
	return qindex;
}
