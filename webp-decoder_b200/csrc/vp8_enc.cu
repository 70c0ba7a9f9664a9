// vp8_enc.cu - the reference ENCODER's in-loop reconstruction for whole-macroblock prediction on the GPU (include/vp8_enc.h;
// reference src/enc-m08_recon/enc_recon.c:855-1085 dc_pred_inloop, :1189-1483 i16x16_uv_sad_inloop): the encoder-side
// neighbour of the decoder's pixel path, built from the same device transforms (vp8_common.cuh idct4x4 / iwht4x4).
//
// One CTA per picture, one warp per macroblock ROW (rows dealt round-robin), all 32 lanes on one macroblock:
//   lanes 0..15 the luma 4x4 blocks in raster order, 16..19 U, 20..23 V, lane 24 the Y2 block.
// MB(x, y) predicts from the reconstructed row above and column left (no above-right in the whole-macroblock modes), so the
// rows of a picture form a wavefront one macroblock apart; the only synchronisation is one progress stamp per row in shared
// memory. The reconstruction lives in macroblock-aligned planes in global memory exactly like the reference's
// EncVp8ReconPlanes (enc_recon.c:790-838): the row above is read from there through L2 (ld.global.cg - another warp wrote
// it), the column left comes from the warp's own shared memory.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/vp8_enc.h"
#include "vp8_common.cuh"

extern "C" int vp8_set_error(int err, const char* what, int cuda_error); // vp8_gpu.cu: errno + the text vp8_gpu_last_error() returns

namespace {

constexpr int kEncWarps = 8;
constexpr int kEncMaxRows = 1024; // 16383 / 16

struct EncImgDesc {
	const uint8_t *src_y, *src_u, *src_v; // source planes on the device, tight (stride = plane width)
	uint8_t *rec_y, *rec_u, *rec_v;       // reconstruction, macroblock-aligned
	int16_t* coeffs;                      // mb_total * 400
	uint8_t *y_modes, *uv_modes;          // mb_total each
	uint8_t* b_modes;                     // mb_total * 16 (sub-block front end only)
	uint32_t width, height, mb_cols, mb_rows;
	int32_t q[6];  // y1dc y1ac uvdc uvac y2dc y2ac
	uint32_t qm[6]; // 2^32 / q + 1: n / q == umulhi(n, qm) for every n this path can meet (checked exhaustively up to 70000 for q < 1000)
};

struct EncWarpWs {
	uint8_t lcol[32];  // right column of the previous macroblock: Y 0..15, U 16..23, V 24..31
	int16_t dcs[16];   // luma DCs on their way to the Y2 lane
	int16_t ydc[16];   // and back, after quantise / dequantise / inverse WHT
};

// round-half-away division saturated to int16 (enc_quant.c:62-76); the division is a multiplication by the step's reciprocal
__device__ __forceinline__ int quantise(int c, int step, uint32_t magic) {
	const int mag = (int)__umulhi((uint32_t)(abs(c) + (step >> 1)), magic);
	return max(min(c < 0 ? -mag : mag, 32767), -32768);
}

// libwebp-style forward DCT of source minus prediction (enc_transform.c:7-44)
__device__ __forceinline__ void fdct4x4(const int (&d)[16], int (&out)[16]) {
	int t[16];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const int s03 = d[4 * r] + d[4 * r + 3], s12 = d[4 * r + 1] + d[4 * r + 2];
		const int d12 = d[4 * r + 1] - d[4 * r + 2], d03 = d[4 * r] - d[4 * r + 3];
		t[4 * r] = (s03 + s12) * 8;
		t[4 * r + 1] = (d12 * 2217 + d03 * 5352 + 1812) >> 9;
		t[4 * r + 2] = (s03 - s12) * 8;
		t[4 * r + 3] = (d03 * 2217 - d12 * 5352 + 937) >> 9;
	}
#pragma unroll
	for (int c = 0; c < 4; c++) {
		const int s03 = t[c] + t[12 + c], s12 = t[4 + c] + t[8 + c];
		const int d12 = t[4 + c] - t[8 + c], d03 = t[c] - t[12 + c];
		out[c] = s16((s03 + s12 + 7) >> 4);
		out[4 + c] = s16(((d12 * 2217 + d03 * 5352 + 12000) >> 16) + (d03 != 0));
		out[8 + c] = s16((s03 - s12 + 7) >> 4);
		out[12 + c] = s16((d03 * 2217 - d12 * 5352 + 51000) >> 16);
	}
}

// forward WHT of the sixteen luma DCs (enc_transform.c:46-72)
__device__ __forceinline__ void fwht4x4(const int (&dc)[16], int (&out)[16]) {
	int t[16];
#pragma unroll
	for (int r = 0; r < 4; r++) {
		const int a0 = dc[4 * r] + dc[4 * r + 2], a1 = dc[4 * r + 1] + dc[4 * r + 3];
		const int a2 = dc[4 * r + 1] - dc[4 * r + 3], a3 = dc[4 * r] - dc[4 * r + 2];
		t[4 * r] = a0 + a1;
		t[4 * r + 1] = a3 + a2;
		t[4 * r + 2] = a3 - a2;
		t[4 * r + 3] = a0 - a1;
	}
#pragma unroll
	for (int c = 0; c < 4; c++) {
		const int a0 = t[c] + t[8 + c], a1 = t[4 + c] + t[12 + c];
		const int a2 = t[4 + c] - t[12 + c], a3 = t[c] - t[8 + c];
		out[c] = s16((a0 + a1) >> 1);
		out[4 + c] = s16((a3 + a2) >> 1);
		out[8 + c] = s16((a3 - a2) >> 1);
		out[12 + c] = s16((a0 - a1) >> 1);
	}
}

// prediction of this lane's 4x4 block, whole-macroblock mode 0 DC, 1 V, 2 H, 3 TM (enc_recon.c:338-409, 447-520)
__device__ __forceinline__ void predict4x4(int mode, int dc, uint32_t aw, uint32_t lw, int p, int (&pred)[16]) {
#pragma unroll
	for (int r = 0; r < 4; r++)
#pragma unroll
		for (int c = 0; c < 4; c++) {
			const int a = (aw >> (8 * c)) & 255, l = (lw >> (8 * r)) & 255;
			pred[4 * r + c] = mode == 1 ? a : mode == 2 ? l : mode == 3 ? clip255(l + a - p) : dc;
		}
}

template <bool SEARCH>
__global__ void __launch_bounds__(kEncWarps * 32) vp8_enc_i16(const EncImgDesc* __restrict__ descs, int n_images) {
	__shared__ int prog_s[kEncMaxRows];
	__shared__ EncWarpWs ws_s[kEncWarps];
	volatile int* prog = prog_s;
	constexpr uint32_t FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	EncWarpWs& ws = ws_s[warp];

	// lane roles
	const int plane = lane < 16 ? 0 : (lane < 20 ? 1 : (lane < 24 ? 2 : 3)); // 3: no block
	const int b = plane == 0 ? lane : (lane & 3);
	const int bx = plane == 0 ? (b & 3) * 4 : (b & 1) * 4, by = plane == 0 ? (b >> 2) * 4 : (b >> 1) * 4;
	const int n = plane == 0 ? 16 : 8;

	for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
		__syncthreads(); // previous picture retired before its stamps are reused
		const EncImgDesc& d = descs[img];
		const int cols = d.mb_cols, rows = d.mb_rows;
		for (int i = threadIdx.x; i < rows; i += blockDim.x) prog_s[i] = 0;
		__syncthreads();

		const uint8_t* src = plane == 0 ? d.src_y : (plane == 1 ? d.src_u : d.src_v);
		uint8_t* rec = plane == 0 ? d.rec_y : (plane == 1 ? d.rec_u : d.rec_v);
		const int pw = plane == 0 ? (int)d.width : (int)((d.width + 1) >> 1), ph = plane == 0 ? (int)d.height : (int)((d.height + 1) >> 1);
		const int rstride = cols * n;
		const int q_dc = plane == 0 ? d.q[0] : d.q[2], q_ac = plane == 0 ? d.q[1] : d.q[3];
		const uint32_t m_dc = plane == 0 ? d.qm[0] : d.qm[2], m_ac = plane == 0 ? d.qm[1] : d.qm[3];
		uint8_t* const lcol = ws.lcol + (plane == 0 ? 0 : (plane == 1 ? 16 : 24));

		for (int row = warp; row < rows; row += kEncWarps) {
			for (int x = 0; x < cols; x++) {
				// ---- the row above has to be one macroblock ahead
				if (row > 0) {
					if (lane == 0)
						while (prog[row - 1] < x + 1) __nanosleep(32);
					__syncwarp();
					__threadfence();
				}
				const bool have_a = row > 0, have_l = x > 0;
				const int x0 = x * n, y0 = row * n;

				// ---- neighbours of this lane's block and the plane's DC predictor (enc_recon.c:541-561)
				uint32_t aw = 0x7f7f7f7fu, lw = 0x81818181u;
				int p = have_a ? 129 : 127;
				int part = 0;
				if (plane < 3) {
					if (have_a) aw = __ldcg(reinterpret_cast<const uint32_t*>(rec + (size_t)(y0 - 1) * rstride + x0 + bx));
					if (have_l) lw = *reinterpret_cast<const uint32_t*>(lcol + by);
					if (have_a && have_l) p = __ldcg(rec + (size_t)(y0 - 1) * rstride + x0 - 1);
					if (have_a && by == 0) part += (int)sum4(aw);
					if (have_l && bx == 0) part += (int)sum4(lw);
				}
				const uint32_t s_yu = __reduce_add_sync(FULL, plane == 0 ? (uint32_t)part : (plane == 1 ? (uint32_t)part << 16 : 0u));
				const uint32_t s_v = __reduce_add_sync(FULL, plane == 2 ? (uint32_t)part : 0u);
				int dc = 128;
				{
					int s = plane == 0 ? (int)(s_yu & 0xffffu) : (plane == 1 ? (int)(s_yu >> 16) : (int)s_v);
					if (have_a != have_l) s += s;
					if (have_a || have_l) dc = (s + n) >> (plane == 0 ? 5 : 4);
				}

				// ---- source block, rows / columns beyond the picture repeat the last one (enc_recon.c:563-582)
				int sp[16];
				if (plane < 3) {
#pragma unroll
					for (int r = 0; r < 4; r++) {
						const uint8_t* srow = src + (size_t)min(y0 + by + r, ph - 1) * pw;
#pragma unroll
						for (int c = 0; c < 4; c++) sp[4 * r + c] = __ldg(srow + min(x0 + bx + c, pw - 1));
					}
				} else {
#pragma unroll
					for (int i = 0; i < 16; i++) sp[i] = 0;
				}

				// ---- mode decision: squared error of each predictor over the macroblock, first minimum wins
				int ymode = 0, cmode = 0;
				if (SEARCH) {
					uint32_t best_y = 0xffffffffu, best_c = 0xffffffffu;
#pragma unroll 1
					for (int m = 0; m < 4; m++) {
						int pr[16];
						predict4x4(m, dc, aw, lw, p, pr);
						uint32_t e = 0;
#pragma unroll
						for (int i = 0; i < 16; i++) e += (uint32_t)((sp[i] - pr[i]) * (sp[i] - pr[i]));
						const uint32_t ey = __reduce_add_sync(FULL, plane == 0 ? e : 0u);
						const uint32_t ec = __reduce_add_sync(FULL, (plane == 1 || plane == 2) ? e : 0u);
						if (ey < best_y) best_y = ey, ymode = m;
						if (ec < best_c) best_c = ec, cmode = m;
					}
				}
				int pr[16];
				predict4x4(plane == 0 ? ymode : cmode, dc, aw, lw, p, pr);

				// ---- forward: DCT, luma DCs to the Y2 lane, quantise
				int cf[16];
				{
					int dd[16];
#pragma unroll
					for (int i = 0; i < 16; i++) dd[i] = sp[i] - pr[i];
					fdct4x4(dd, cf);
				}
				if (plane == 0) {
					ws.dcs[b] = (int16_t)cf[0];
					cf[0] = 0;
				}
#pragma unroll
				for (int i = 0; i < 16; i++) cf[i] = quantise(cf[i], i ? q_ac : q_dc, i ? m_ac : m_dc);
				__syncwarp();
				const size_t mb = (size_t)row * cols + x;
				int16_t* const out = d.coeffs + mb * 400;
				if (lane == 24) {
					int dcv[16], y2[16];
#pragma unroll
					for (int i = 0; i < 16; i++) dcv[i] = ws.dcs[i];
					fwht4x4(dcv, y2);
					uint32_t pk[8];
#pragma unroll
					for (int i = 0; i < 16; i++) y2[i] = quantise(y2[i], i ? d.q[5] : d.q[4], i ? d.qm[5] : d.qm[4]);
#pragma unroll
					for (int i = 0; i < 8; i++) pk[i] = (uint32_t)(y2[2 * i] & 0xffff) | ((uint32_t)y2[2 * i + 1] << 16);
					reinterpret_cast<uint4*>(out)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
					reinterpret_cast<uint4*>(out)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
					// backward, as a decoder will (enc_recon.c:1037-1040)
					int deq[16], back[16];
#pragma unroll
					for (int i = 0; i < 16; i++) deq[i] = s16(y2[i] * (i ? d.q[5] : d.q[4]));
					iwht4x4(deq, back);
#pragma unroll
					for (int i = 0; i < 16; i++) ws.ydc[i] = (int16_t)back[i];
					if (d.y_modes) d.y_modes[mb] = (uint8_t)ymode;
					if (d.uv_modes) d.uv_modes[mb] = (uint8_t)cmode;
				}
				if (plane < 3) {
					uint32_t pk[8];
#pragma unroll
					for (int i = 0; i < 8; i++) pk[i] = (uint32_t)(cf[2 * i] & 0xffff) | ((uint32_t)cf[2 * i + 1] << 16);
					uint4* o = reinterpret_cast<uint4*>(out + 16 + (plane == 0 ? 0 : (plane == 1 ? 256 : 320)) + 16 * b);
					o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
					o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
				}
				__syncwarp();

				// ---- backward: dequantise (int16 wrap), inverse DCT, add. The reference multiplies the luma DC that comes
				//      back from the inverse WHT by the y1 DC step once more (enc_recon.c:1044-1046); so do we.
				if (plane < 3) {
					int deq[16], res[16];
#pragma unroll
					for (int i = 0; i < 16; i++) deq[i] = s16(cf[i] * (i ? q_ac : q_dc));
					if (plane == 0) deq[0] = s16(ws.ydc[b] * q_dc);
					idct4x4(deq, res);
					uint32_t right = 0;
#pragma unroll
					for (int r = 0; r < 4; r++) {
						uint32_t w = 0;
#pragma unroll
						for (int c = 0; c < 4; c++) w |= (uint32_t)add_clip255(pr[4 * r + c], res[4 * r + c]) << (8 * c);
						__stcg(reinterpret_cast<uint32_t*>(rec + (size_t)(y0 + by + r) * rstride + x0 + bx), w);
						right |= (w >> 24) << (8 * r);
					}
					if (bx + 4 == n) *reinterpret_cast<uint32_t*>(lcol + by) = right;
				}
				__syncwarp();
				if (lane == 0) {
					__threadfence();
					prog[row] = x + 1;
				}
			}
		}
	}
}


// ------------------------------------------------------------------------------------------------ 4x4 sub-block front end
// enc_vp8_encode_bpred_uv_sad_inloop (enc_recon.c:1507-1831): every macroblock B_PRED. Same wavefront, but the luma of a
// macroblock is sixteen sub-blocks in raster order, each predicted from the reconstruction so far - including the row above
// up to four pixels to the right of the macroblock (RFC 6386 11.4), so a row trails the row above by TWO macroblocks, as in
// the decoder. Per sub-block the warp scores all ten modes at once: lanes 0..15 = the pixels under modes 0..4, lanes 16..31
// under modes 5..9, five REDUX sums carry two modes each.
struct EncBpWs {
	uint8_t tile[17 * 24]; // luma of the macroblock being built: row r, column c at (r + 1) * 24 + 4 + c, r = -1..15, c = -1..19
	uint8_t lcol[32];      // right column of the previous macroblock: Y 0..15, U 16..23, V 24..31
	uint8_t edge[16];      // the sub-block's edge vector E (vp8_common.cuh: L3 L3 L3 L2 L1 L0 P A0..A7 A7)
	int16_t buf[16];       // residual -> coefficients -> dequantised -> inverse, between the pixel lanes and lane 0
	uint8_t bmode[16];
};

__device__ __forceinline__ int predict_sub_px(int mode, int p, const uint8_t* E) {
	if (mode == 0) return (E[7] + E[8] + E[9] + E[10] + E[2] + E[3] + E[4] + E[5] + 4) >> 3;
	if (mode == 1) return clip255(E[5 - (p >> 2)] + E[7 + (p & 3)] - E[6]);
	const int t = c_bpred_taps[mode * 16 + p], i = t & 15;
	return (t & 16) ? (E[i] + E[i + 1] + 1) >> 1 : (E[i] + 2 * E[i + 1] + E[i + 2] + 2) >> 2;
}

__global__ void __launch_bounds__(kEncWarps * 32) vp8_enc_bpred(const EncImgDesc* __restrict__ descs, int n_images) {
	__shared__ int prog_s[kEncMaxRows];
	__shared__ EncBpWs ws_s[kEncWarps];
	volatile int* prog = prog_s;
	constexpr uint32_t FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	EncBpWs& ws = ws_s[warp];

	// chroma lanes as in vp8_enc_i16: 16..19 U, 20..23 V
	const int plane = lane < 16 ? 0 : (lane < 20 ? 1 : (lane < 24 ? 2 : 3));
	const int cb = lane & 3, cbx = (cb & 1) * 4, cby = (cb >> 1) * 4;
	// luma pixel lanes
	const int p = lane & 15, pr = p >> 2, pc = p & 3, grp = lane >> 4;
	const int e_dy = p <= 2 ? 3 : (p <= 5 ? 5 - p : -1), e_dx = p <= 6 ? -1 : (p == 15 ? 7 : p - 7);

	for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
		__syncthreads();
		const EncImgDesc& d = descs[img];
		const int cols = d.mb_cols, rows = d.mb_rows;
		for (int i = threadIdx.x; i < rows; i += blockDim.x) prog_s[i] = 0;
		__syncthreads();
		const int w = (int)d.width, h = (int)d.height, cw = (w + 1) >> 1, ch = (h + 1) >> 1;
		const int ys = cols * 16, cs = cols * 8;
		const uint8_t* csrc = plane == 1 ? d.src_u : d.src_v;
		uint8_t* crec = plane == 1 ? d.rec_u : d.rec_v;
		uint8_t* const clcol = ws.lcol + (plane == 1 ? 16 : 24);

		for (int row = warp; row < rows; row += kEncWarps) {
			for (int x = 0; x < cols; x++) {
				if (row > 0) {
					if (lane == 0)
						while (prog[row - 1] < min(x + 2, cols)) __nanosleep(32);
					__syncwarp();
					__threadfence();
				}
				const bool have_a = row > 0, have_l = x > 0;
				const int x0 = 16 * x, y0 = 16 * row;
				const size_t mb = (size_t)row * cols + x;
				int16_t* const out = d.coeffs + mb * 400;

				// ---- borders of the luma tile: row -1 (columns -1..19) and column -1
				if (lane < 21) {
					const int c = lane - 1;
					int v = 127;
					if (have_a) v = (c < 0 && !have_l) ? 129 : __ldcg(d.rec_y + (size_t)(y0 - 1) * ys + min(x0 + c, ys - 1));
					ws.tile[4 + c] = (uint8_t)v;
					if (c >= 16) ws.tile[4 * 24 + 4 + c] = ws.tile[8 * 24 + 4 + c] = ws.tile[12 * 24 + 4 + c] = (uint8_t)v; // above-right of rows 4, 8, 12
				}
				if (lane < 16) ws.tile[(lane + 1) * 24 + 3] = have_l ? ws.lcol[lane] : 129;

				// ---- chroma: mode by sum of absolute differences over U and V (first minimum), then as in vp8_enc_i16
				uint32_t aw = 0x7f7f7f7fu, lw = 0x81818181u;
				int cp = have_a ? 129 : 127, part = 0;
				const bool chroma = plane == 1 || plane == 2;
				if (chroma) {
					if (have_a) aw = __ldcg(reinterpret_cast<const uint32_t*>(crec + (size_t)(8 * row - 1) * cs + 8 * x + cbx));
					if (have_l) lw = *reinterpret_cast<const uint32_t*>(clcol + cby);
					if (have_a && have_l) cp = __ldcg(crec + (size_t)(8 * row - 1) * cs + 8 * x - 1);
					if (have_a && cby == 0) part += (int)sum4(aw);
					if (have_l && cbx == 0) part += (int)sum4(lw);
				}
				const uint32_t s_uv = __reduce_add_sync(FULL, plane == 1 ? (uint32_t)part : (plane == 2 ? (uint32_t)part << 16 : 0u));
				int dc = 128;
				{
					int s = plane == 1 ? (int)(s_uv & 0xffffu) : (int)(s_uv >> 16);
					if (have_a != have_l) s += s;
					if (have_a || have_l) dc = (s + 8) >> 4;
				}
				int sp[16];
#pragma unroll
				for (int i = 0; i < 16; i++) sp[i] = 0;
				if (chroma) {
#pragma unroll
					for (int r = 0; r < 4; r++) {
						const uint8_t* srow = csrc + (size_t)min(8 * row + cby + r, ch - 1) * cw;
#pragma unroll
						for (int c = 0; c < 4; c++) sp[4 * r + c] = __ldg(srow + min(8 * x + cbx + c, cw - 1));
					}
				}
				int cmode = 0;
				{
					uint32_t best = 0xffffffffu;
#pragma unroll 1
					for (int m = 0; m < 4; m++) {
						int pr4[16];
						predict4x4(m, dc, aw, lw, cp, pr4);
						uint32_t e = 0;
#pragma unroll
						for (int i = 0; i < 16; i++) e += (uint32_t)abs(sp[i] - pr4[i]);
						const uint32_t ec = __reduce_add_sync(FULL, chroma ? e : 0u);
						if (ec < best) best = ec, cmode = m;
					}
				}
				if (chroma) {
					int pr4[16], cf[16], dd[16];
					predict4x4(cmode, dc, aw, lw, cp, pr4);
#pragma unroll
					for (int i = 0; i < 16; i++) dd[i] = sp[i] - pr4[i];
					fdct4x4(dd, cf);
					uint32_t pk[8];
#pragma unroll
					for (int i = 0; i < 16; i++) cf[i] = quantise(cf[i], i ? d.q[3] : d.q[2], i ? d.qm[3] : d.qm[2]);
#pragma unroll
					for (int i = 0; i < 8; i++) pk[i] = (uint32_t)(cf[2 * i] & 0xffff) | ((uint32_t)cf[2 * i + 1] << 16);
					uint4* o = reinterpret_cast<uint4*>(out + 16 + (plane == 1 ? 256 : 320) + 16 * cb);
					o[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
					o[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
					int deq[16], res[16];
#pragma unroll
					for (int i = 0; i < 16; i++) deq[i] = s16(cf[i] * (i ? d.q[3] : d.q[2]));
					idct4x4(deq, res);
					uint32_t right = 0;
#pragma unroll
					for (int r = 0; r < 4; r++) {
						uint32_t wv = 0;
#pragma unroll
						for (int c = 0; c < 4; c++) wv |= (uint32_t)add_clip255(pr4[4 * r + c], res[4 * r + c]) << (8 * c);
						__stcg(reinterpret_cast<uint32_t*>(crec + (size_t)(8 * row + cby + r) * cs + 8 * x + cbx), wv);
						right |= (wv >> 24) << (8 * r);
					}
					if (cbx == 4) *reinterpret_cast<uint32_t*>(clcol + cby) = right;
				}
				if (lane == 24) { // no Y2 block in a B_PRED macroblock: its sixteen coefficients stay zero
					reinterpret_cast<uint4*>(out)[0] = make_uint4(0, 0, 0, 0);
					reinterpret_cast<uint4*>(out)[1] = make_uint4(0, 0, 0, 0);
					if (d.y_modes) d.y_modes[mb] = 4;
					if (d.uv_modes) d.uv_modes[mb] = (uint8_t)cmode;
				}
				__syncwarp();

				// ---- luma: sixteen sub-blocks in raster order
#pragma unroll 1
				for (int sb = 0; sb < 16; sb++) {
					const int r4 = (sb >> 2) * 4, c4 = (sb & 3) * 4;
					if (lane < 16) ws.edge[p] = ws.tile[(r4 + e_dy + 1) * 24 + 4 + c4 + e_dx];
					const int s = __ldg(d.src_y + (size_t)min(y0 + r4 + pr, h - 1) * w + min(x0 + c4 + pc, w - 1));
					__syncwarp();
					uint32_t tot[5];
#pragma unroll
					for (int k = 0; k < 5; k++) {
						const uint32_t ad = (uint32_t)abs(s - predict_sub_px(5 * grp + k, p, ws.edge));
						tot[k] = __reduce_add_sync(FULL, ad << (16 * grp));
					}
					int bm = 0;
					uint32_t best = 0xffffffffu;
#pragma unroll
					for (int m = 0; m < 10; m++) {
						const uint32_t e = m < 5 ? (tot[m] & 0xffffu) : (tot[m - 5] >> 16);
						if (e < best) best = e, bm = m;
					}
					const int pred = predict_sub_px(bm, p, ws.edge);
					if (lane < 16) ws.buf[p] = (int16_t)(s - pred);
					__syncwarp();
					if (lane == 0) {
						int dd[16], cf[16];
#pragma unroll
						for (int i = 0; i < 16; i++) dd[i] = ws.buf[i];
						fdct4x4(dd, cf);
#pragma unroll
						for (int i = 0; i < 16; i++) ws.buf[i] = (int16_t)cf[i];
						ws.bmode[sb] = (uint8_t)bm;
					}
					__syncwarp();
					if (lane < 16) { // one coefficient per lane: quantise, store, dequantise
						const int step = lane ? d.q[1] : d.q[0];
						const int qc = quantise(ws.buf[lane], step, lane ? d.qm[1] : d.qm[0]);
						out[16 + 16 * sb + lane] = (int16_t)qc;
						ws.buf[lane] = (int16_t)(qc * step);
					}
					__syncwarp();
					if (lane == 0) {
						int deq[16], res[16];
#pragma unroll
						for (int i = 0; i < 16; i++) deq[i] = ws.buf[i];
						idct4x4(deq, res);
#pragma unroll
						for (int i = 0; i < 16; i++) ws.buf[i] = (int16_t)res[i];
					}
					__syncwarp();
					if (lane < 16) ws.tile[(r4 + pr + 1) * 24 + 4 + c4 + pc] = (uint8_t)add_clip255(pred, ws.buf[p]);
					__syncwarp();
				}

				// ---- the macroblock's luma leaves for the reconstruction plane; right column and modes are kept / stored
#pragma unroll
				for (int j = 0; j < 2; j++) {
					const int i = lane + 32 * j, r = i >> 2, wc = i & 3;
					__stcg(reinterpret_cast<uint32_t*>(d.rec_y + (size_t)(y0 + r) * ys + x0 + 4 * wc),
					       *reinterpret_cast<const uint32_t*>(ws.tile + (r + 1) * 24 + 4 + 4 * wc));
				}
				if (lane < 16) {
					ws.lcol[lane] = ws.tile[(lane + 1) * 24 + 4 + 15];
					d.b_modes[mb * 16 + lane] = ws.bmode[lane];
				}
				__syncwarp();
				if (lane == 0) {
					__threadfence();
					prog[row] = x + 1;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ host side
// RFC 6386 14.1 (the data of enc_quant.c:15-36) and libwebp's quality -> qindex map (enc_quality_table.c:5-13)
const uint16_t kDc[128] = {
    4,   5,   6,   7,   8,   9,   10,  10,  11,  12,  13,  14,  15,  16,  17,  17,  18,  19,  20,  20,  21,  21,  22,  22,  23,  23,
    24,  25,  25,  26,  27,  28,  29,  30,  31,  32,  33,  34,  35,  36,  37,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  46,
    47,  48,  49,  50,  51,  52,  53,  54,  55,  56,  57,  58,  59,  60,  61,  62,  63,  64,  65,  66,  67,  68,  69,  70,  71,  72,
    73,  74,  75,  76,  76,  77,  78,  79,  80,  81,  82,  83,  84,  85,  86,  87,  88,  89,  91,  93,  95,  96,  98,  100, 101, 102,
    104, 106, 108, 110, 112, 114, 116, 118, 122, 124, 126, 128, 130, 132, 134, 136, 138, 140, 143, 145, 148, 151, 154, 157};
const uint16_t kAc[128] = {
    4,   5,   6,   7,   8,   9,   10,  11,  12,  13,  14,  15,  16,  17,  18,  19,  20,  21,  22,  23,  24,  25,  26,  27,  28,  29,
    30,  31,  32,  33,  34,  35,  36,  37,  38,  39,  40,  41,  42,  43,  44,  45,  46,  47,  48,  49,  50,  51,  52,  53,  54,  55,
    56,  57,  58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,  82,  84,  86,  88,  90,  92,  94,  96,  98,  100, 102, 104,
    106, 108, 110, 112, 114, 116, 119, 122, 125, 128, 131, 134, 137, 140, 143, 146, 149, 152, 155, 158, 161, 164, 167, 170, 173, 177,
    181, 185, 189, 193, 197, 201, 205, 209, 213, 217, 221, 225, 229, 234, 239, 245, 249, 254, 259, 264, 269, 274, 279, 284};
const uint8_t kQIndexOfQuality[101] = {
    127, 103, 96, 92, 89, 86, 83, 81, 79, 77, 75, 73, 72, 70, 69, 68, 66, 65, 64, 63, 62, 61, 60, 59, 58, 57, 56, 55, 54, 53, 52, 51, 51, 50,
    49,  48,  48, 47, 46, 45, 45, 44, 43, 43, 42, 41, 41, 40, 40, 39, 38, 38, 37, 37, 36, 36, 35, 35, 34, 33, 33, 32, 32, 31, 31, 30, 30, 29,
    29,  28,  28, 28, 27, 27, 26, 26, 24, 23, 22, 21, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9,  8,  7,  6,  5,  4,  3,  2,  1,  0,  0};

int enc_quant(int quality, int32_t q[6]) { // enc_quant.c:38-60, all deltas zero
	const int qi = kQIndexOfQuality[std::min(std::max(quality, 0), 100)];
	q[0] = kDc[qi];
	q[1] = kAc[qi];
	q[2] = std::min<int>(kDc[qi], 132);
	q[3] = kAc[qi];
	q[4] = kDc[qi] * 2;
	q[5] = std::max(kAc[qi] * 155 / 100, 8);
	return qi;
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

// grow-only device workspace and stream of the calling thread
struct EncState {
	int device = -1;
	cudaStream_t stream = nullptr;
	uint8_t* dev = nullptr;
	size_t dev_bytes = 0;
	EncImgDesc* pin_desc = nullptr;
	size_t pin_desc_n = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	double last_ms = 0;
	int sm_count = 0;
};
thread_local EncState g_enc;

#define ECU(call)                                                          \
	do {                                                                   \
		cudaError_t e__ = (call);                                          \
		if (e__ != cudaSuccess) return vp8_set_error(EIO, #call, (int)e__); \
	} while (0)

int enc_prepare(int device) {
	if (device < 0) {
		device = 0;
		if (const char* e = getenv("VP8_GPU_DEVICE")) device = atoi(e);
	}
	EncState& s = g_enc;
	ECU(cudaSetDevice(device));
	if (s.device != device) {
		if (s.dev) cudaFree(s.dev);
		s.dev = nullptr;
		s.dev_bytes = 0;
		if (s.stream) cudaStreamDestroy(s.stream);
		s.stream = nullptr;
		if (s.ev0) cudaEventDestroy(s.ev0);
		if (s.ev1) cudaEventDestroy(s.ev1);
		s.ev0 = s.ev1 = nullptr;
		s.device = device;
	}
	if (!s.stream) ECU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
	if (!s.ev0) ECU(cudaEventCreate(&s.ev0));
	if (!s.ev1) ECU(cudaEventCreate(&s.ev1));
	if (!s.sm_count) ECU(cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, device));
	return 0;
}

} // namespace

size_t vp8_gpu_enc_mb_total(uint32_t width, uint32_t height) { return (size_t)((width + 15) >> 4) * ((height + 15) >> 4); }
double vp8_gpu_enc_last_kernel_ms(void) { return g_enc.last_ms; }

// kind: 0 DC prediction, 1 whole-macroblock mode search, 2 sub-block (B_PRED) mode search
static int enc_run(int device, const EncYuv420Image* const* yuv, int n, int quality, int kind, int16_t* const* coeffs, uint8_t* const* y_modes,
                   uint8_t* const* b_modes, uint8_t* const* uv_modes, uint8_t* const* rec_y, uint8_t* const* rec_u, uint8_t* const* rec_v,
                   uint8_t* qindex_out) {
	if (!yuv || !coeffs || !qindex_out || n <= 0 || (kind == 2 && !b_modes)) return vp8_set_error(EINVAL, "bad arguments", 0);
	for (int i = 0; i < n; i++) {
		const EncYuv420Image* im = yuv[i];
		if (!im || !im->y || !im->u || !im->v || im->width == 0 || im->height == 0 || !coeffs[i] || (kind == 2 && !b_modes[i]))
			return vp8_set_error(EINVAL, "bad picture", 0);
		if (im->width > 16383 || im->height > 16383) return vp8_set_error(EINVAL, "picture too large", 0);
		if (im->y_stride < im->width || im->uv_stride < (im->width + 1) / 2) return vp8_set_error(EINVAL, "bad stride", 0);
	}
	if (enc_prepare(device)) return -1;
	EncState& s = g_enc;
	std::vector<EncImgDesc> h(n);
	int32_t q[6];
	*qindex_out = (uint8_t)enc_quant(quality, q);

	// ---- device layout: per picture [src y|u|v][rec y|u|v][coeffs][y_modes][uv_modes], then the descriptors
	struct Off {
		size_t src[3], rec[3], coeffs, ym, cm, bm;
	};
	std::vector<Off> off(n);
	size_t total = 0;
	for (int i = 0; i < n; i++) {
		const uint32_t w = yuv[i]->width, hgt = yuv[i]->height, cw = (w + 1) / 2, ch = (hgt + 1) / 2;
		const size_t cols = (w + 15) >> 4, rows = (hgt + 15) >> 4, mb = cols * rows;
		Off& o = off[i];
		o.src[0] = total, total += up256((size_t)w * hgt);
		o.src[1] = total, total += up256((size_t)cw * ch);
		o.src[2] = total, total += up256((size_t)cw * ch);
		o.rec[0] = total, total += up256(mb * 256);
		o.rec[1] = total, total += up256(mb * 64);
		o.rec[2] = total, total += up256(mb * 64);
		o.coeffs = total, total += up256(mb * 800);
		o.ym = total, total += up256(mb);
		o.cm = total, total += up256(mb);
		o.bm = total, total += kind == 2 ? up256(mb * 16) : 0;
	}
	const size_t desc_off = total;
	total += up256(sizeof(EncImgDesc) * (size_t)n);
	if (s.dev_bytes < total) {
		if (s.dev) cudaFree(s.dev);
		s.dev = nullptr;
		s.dev_bytes = 0;
		cudaError_t e = cudaMalloc((void**)&s.dev, total);
		if (e != cudaSuccess) return vp8_set_error(ENOMEM, "device workspace of the encoder", (int)e);
		s.dev_bytes = total;
	}
	if (s.pin_desc_n < (size_t)n) {
		if (s.pin_desc) cudaFreeHost(s.pin_desc);
		s.pin_desc = nullptr;
		s.pin_desc_n = 0;
		ECU(cudaHostAlloc((void**)&s.pin_desc, sizeof(EncImgDesc) * (size_t)n, cudaHostAllocDefault));
		s.pin_desc_n = n;
	}

	// ---- uploads (planes with a stride go row by row through cudaMemcpy2D)
	for (int i = 0; i < n; i++) {
		const EncYuv420Image* im = yuv[i];
		const uint32_t w = im->width, hgt = im->height, cw = (w + 1) / 2, ch = (hgt + 1) / 2;
		const uint8_t* sp[3] = {im->y, im->u, im->v};
		const size_t pw[3] = {w, cw, cw}, ph[3] = {hgt, ch, ch}, st[3] = {im->y_stride, im->uv_stride, im->uv_stride};
		for (int k = 0; k < 3; k++)
			ECU(cudaMemcpy2DAsync(s.dev + off[i].src[k], pw[k], sp[k], st[k], pw[k], ph[k], cudaMemcpyHostToDevice, s.stream));
		EncImgDesc& d = s.pin_desc[i];
		d.src_y = s.dev + off[i].src[0];
		d.src_u = s.dev + off[i].src[1];
		d.src_v = s.dev + off[i].src[2];
		d.rec_y = s.dev + off[i].rec[0];
		d.rec_u = s.dev + off[i].rec[1];
		d.rec_v = s.dev + off[i].rec[2];
		d.coeffs = reinterpret_cast<int16_t*>(s.dev + off[i].coeffs);
		d.y_modes = s.dev + off[i].ym;
		d.uv_modes = s.dev + off[i].cm;
		d.b_modes = s.dev + off[i].bm;
		d.width = w;
		d.height = hgt;
		d.mb_cols = (w + 15) >> 4;
		d.mb_rows = (hgt + 15) >> 4;
		memcpy(d.q, q, sizeof(q));
		for (int k = 0; k < 6; k++) d.qm[k] = (uint32_t)((1ull << 32) / (uint32_t)q[k] + 1);
	}
	ECU(cudaMemcpyAsync(s.dev + desc_off, s.pin_desc, sizeof(EncImgDesc) * (size_t)n, cudaMemcpyHostToDevice, s.stream));

	// ---- one launch: a CTA per picture, as many resident CTAs as the device holds
	const int grid = std::min(n, s.sm_count * 4);
	ECU(cudaEventRecord(s.ev0, s.stream));
	const EncImgDesc* dd = reinterpret_cast<const EncImgDesc*>(s.dev + desc_off);
	if (kind == 2) vp8_enc_bpred<<<grid, kEncWarps * 32, 0, s.stream>>>(dd, n);
	else if (kind == 1) vp8_enc_i16<true><<<grid, kEncWarps * 32, 0, s.stream>>>(dd, n);
	else vp8_enc_i16<false><<<grid, kEncWarps * 32, 0, s.stream>>>(dd, n);
	ECU(cudaGetLastError());
	ECU(cudaEventRecord(s.ev1, s.stream));

	// ---- results
	for (int i = 0; i < n; i++) {
		const size_t mb = vp8_gpu_enc_mb_total(yuv[i]->width, yuv[i]->height);
		ECU(cudaMemcpyAsync(coeffs[i], s.dev + off[i].coeffs, mb * 800, cudaMemcpyDeviceToHost, s.stream));
		if (y_modes && y_modes[i]) ECU(cudaMemcpyAsync(y_modes[i], s.dev + off[i].ym, mb, cudaMemcpyDeviceToHost, s.stream));
		if (uv_modes && uv_modes[i]) ECU(cudaMemcpyAsync(uv_modes[i], s.dev + off[i].cm, mb, cudaMemcpyDeviceToHost, s.stream));
		if (kind == 2) ECU(cudaMemcpyAsync(b_modes[i], s.dev + off[i].bm, mb * 16, cudaMemcpyDeviceToHost, s.stream));
		uint8_t* const* rp[3] = {rec_y, rec_u, rec_v};
		for (int k = 0; k < 3; k++)
			if (rp[k] && rp[k][i]) ECU(cudaMemcpyAsync(rp[k][i], s.dev + off[i].rec[k], mb * (k ? 64 : 256), cudaMemcpyDeviceToHost, s.stream));
	}
	ECU(cudaStreamSynchronize(s.stream));
	float ms = 0;
	if (cudaEventElapsedTime(&ms, s.ev0, s.ev1) == cudaSuccess) s.last_ms = ms;
	return 0;
}

int vp8_gpu_enc_i16_inloop(int device, const EncYuv420Image* const* yuv, int n, int quality, int search, int16_t* const* coeffs,
                           uint8_t* const* y_modes, uint8_t* const* uv_modes, uint8_t* const* rec_y, uint8_t* const* rec_u,
                           uint8_t* const* rec_v, uint8_t* qindex_out) {
	return enc_run(device, yuv, n, quality, search ? 1 : 0, coeffs, y_modes, nullptr, uv_modes, rec_y, rec_u, rec_v, qindex_out);
}

int vp8_gpu_enc_bpred_inloop(int device, const EncYuv420Image* const* yuv, int n, int quality, int16_t* const* coeffs, uint8_t* const* y_modes,
                             uint8_t* const* b_modes, uint8_t* const* uv_modes, uint8_t* const* rec_y, uint8_t* const* rec_u,
                             uint8_t* const* rec_v, uint8_t* qindex_out) {
	return enc_run(device, yuv, n, quality, 2, coeffs, y_modes, b_modes, uv_modes, rec_y, rec_u, rec_v, qindex_out);
}

// ------------------------------------------------------------------------------------------------ reference entry points
static int enc_one(const EncYuv420Image* yuv, int quality, int kind, uint8_t** y_modes_out, size_t* y_modes_count_out, uint8_t** b_modes_out,
                   size_t* b_modes_count_out, uint8_t** uv_modes_out, size_t* uv_modes_count_out, int16_t** coeffs_out, size_t* coeffs_count_out,
                   uint8_t* qindex_out) {
	*coeffs_out = nullptr;
	*coeffs_count_out = 0;
	*qindex_out = 0;
	if (y_modes_out) *y_modes_out = nullptr, *y_modes_count_out = 0;
	if (b_modes_out) *b_modes_out = nullptr, *b_modes_count_out = 0;
	if (uv_modes_out) *uv_modes_out = nullptr, *uv_modes_count_out = 0;
	if (!yuv || !yuv->y || !yuv->u || !yuv->v || yuv->width == 0 || yuv->height == 0) return vp8_set_error(EINVAL, "bad picture", 0);
	const size_t mb = vp8_gpu_enc_mb_total(yuv->width, yuv->height);
	int16_t* co = (int16_t*)malloc(mb * 400 * sizeof(int16_t));
	uint8_t* ym = y_modes_out ? (uint8_t*)malloc(mb) : nullptr;
	uint8_t* bm = b_modes_out ? (uint8_t*)malloc(mb * 16) : nullptr;
	uint8_t* cm = uv_modes_out ? (uint8_t*)malloc(mb) : nullptr;
	if (!co || (y_modes_out && !ym) || (b_modes_out && !bm) || (uv_modes_out && !cm)) {
		free(co), free(ym), free(bm), free(cm);
		return vp8_set_error(ENOMEM, "encoder output arrays", 0);
	}
	if (enc_run(-1, &yuv, 1, quality, kind, &co, ym ? &ym : nullptr, bm ? &bm : nullptr, cm ? &cm : nullptr, nullptr, nullptr, nullptr, qindex_out)) {
		const int saved = errno;
		free(co), free(ym), free(bm), free(cm);
		errno = saved;
		return -1;
	}
	*coeffs_out = co;
	*coeffs_count_out = mb * 400;
	if (y_modes_out) *y_modes_out = ym, *y_modes_count_out = mb;
	if (b_modes_out) *b_modes_out = bm, *b_modes_count_out = mb * 16;
	if (uv_modes_out) *uv_modes_out = cm, *uv_modes_count_out = mb;
	return 0;
}

int enc_vp8_encode_bpred_uv_sad_inloop(const EncYuv420Image* yuv, int quality, uint8_t** y_modes_out, size_t* y_modes_count_out,
                                       uint8_t** b_modes_out, size_t* b_modes_count_out, uint8_t** uv_modes_out, size_t* uv_modes_count_out,
                                       int16_t** coeffs_out, size_t* coeffs_count_out, uint8_t* qindex_out) {
	if (!y_modes_out || !y_modes_count_out || !b_modes_out || !b_modes_count_out || !uv_modes_out || !uv_modes_count_out || !coeffs_out ||
	    !coeffs_count_out || !qindex_out)
		return vp8_set_error(EINVAL, "null output", 0);
	return enc_one(yuv, quality, 2, y_modes_out, y_modes_count_out, b_modes_out, b_modes_count_out, uv_modes_out, uv_modes_count_out, coeffs_out,
	               coeffs_count_out, qindex_out);
}

int enc_vp8_encode_dc_pred_inloop(const EncYuv420Image* yuv, int quality, int16_t** coeffs_out, size_t* coeffs_count_out, uint8_t* qindex_out) {
	if (!coeffs_out || !coeffs_count_out || !qindex_out) return vp8_set_error(EINVAL, "null output", 0);
	return enc_one(yuv, quality, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, coeffs_out, coeffs_count_out, qindex_out);
}

int enc_vp8_encode_i16x16_uv_sad_inloop(const EncYuv420Image* yuv, int quality, uint8_t** y_modes_out, size_t* y_modes_count_out,
                                        uint8_t** uv_modes_out, size_t* uv_modes_count_out, int16_t** coeffs_out,
                                        size_t* coeffs_count_out, uint8_t* qindex_out) {
	if (!y_modes_out || !y_modes_count_out || !uv_modes_out || !uv_modes_count_out || !coeffs_out || !coeffs_count_out || !qindex_out)
		return vp8_set_error(EINVAL, "null output", 0);
	return enc_one(yuv, quality, 1, y_modes_out, y_modes_count_out, nullptr, nullptr, uv_modes_out, uv_modes_count_out, coeffs_out, coeffs_count_out,
	               qindex_out);
}

int enc_vp8_encode_i16x16_sad_inloop(const EncYuv420Image* yuv, int quality, uint8_t** y_modes_out, size_t* y_modes_count_out,
                                     int16_t** coeffs_out, size_t* coeffs_count_out, uint8_t* qindex_out) {
	if (!y_modes_out || !y_modes_count_out || !coeffs_out || !coeffs_count_out || !qindex_out) return vp8_set_error(EINVAL, "null output", 0);
	return enc_one(yuv, quality, 1, y_modes_out, y_modes_count_out, nullptr, nullptr, nullptr, nullptr, coeffs_out, coeffs_count_out, qindex_out);
}
