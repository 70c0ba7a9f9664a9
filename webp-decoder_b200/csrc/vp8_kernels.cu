// vp8_kernels.cu - sm_100a kernels for the VP8 key-frame pixel path.
//
//   vp8_mb_wavefront<NW, RECON, FILTER>
//       dequantisation + inverse WHT/DCT + intra prediction + residual add   (reference m06, vp8_recon.c:423-684)
//       fused with the in-loop deblocking filter                             (reference m07, vp8_loopfilter.c:201-283)
//       and the crop to the visible frame                                    (vp8_recon.c:693-707).
//   vp8_i420_to_rgb
//       libwebp-exact fixed-point YUV->RGB with fancy upsampling              (reference m08, yuv2rgb_ppm.c:30-121,164-201).
//
// Execution model of the wavefront kernel
// ---------------------------------------
// One CTA owns one image at a time; many images are in flight per launch (grid-stride over the batch).
// Inside the CTA, warp w owns macroblock rows w, w+NW, ... and walks each row left to right, one macroblock
// per iteration, all 32 lanes cooperating on that macroblock. MB(x,y) needs MB(x-1,y) (same warp, previous
// iteration) and MB(x,y-1), MB(x+1,y-1) (the warp one row up), so the only synchronisation is a per-row progress
// stamp in shared memory that the row below spins on: the rows of an image form a diagonal wavefront without any
// CTA-wide barrier and without any inter-CTA dependency (nothing to deadlock on, nothing to co-schedule).
//
// Nothing but coefficients/modes is read from HBM and nothing but final pixels is written:
//   * the unfiltered bottom pixel row of every macroblock (what the row below predicts from) lives in a
//     shared-memory line buffer (tu_*), the unfiltered right column stays in the warp's own tile;
//   * the loop filter runs on a 20x20 (+2x 12x12) tile whose 4-pixel top/left aprons are the already filtered
//     neighbours: the left apron is the warp's previous tile, the top apron comes from a second shared-memory
//     line buffer (tf_*) holding the last four filtered rows of the macroblock row above;
//   * a macroblock's pixels are stored to HBM only once they can no longer change: the block of rows -4..11 /
//     columns -4..11 relative to the macroblock (the right 4 columns and bottom 4 rows wait for the neighbours'
//     edge filters). The 4-pixel shift keeps every store 32-bit aligned.
#include <cuda_runtime.h>
#include <stdint.h>

#include "vp8_dev.h"

#include "vp8_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ the kernel
#ifndef VP8_MIN_CTAS
#define VP8_MIN_CTAS(NW) ((NW) == 4 ? 7 : (NW) == 8 ? 3 : 1)
#endif

template <int NW, bool RECON, bool FILTER>
__global__ void __launch_bounds__(NW * 32, VP8_MIN_CTAS(NW)) vp8_mb_wavefront(const Vp8ImgDesc* __restrict__ descs, int n_images, int line_px) {
	extern __shared__ __align__(16) uint8_t smem[];
	volatile int* prog = reinterpret_cast<volatile int*>(smem);
	Vp8ImgDesc* sd = reinterpret_cast<Vp8ImgDesc*>(smem + 256);
	uint32_t* btab = reinterpret_cast<uint32_t*>(smem + 512);
	// line buffers: unfiltered bottom rows (tu), last four filtered rows (tf)
	uint8_t* tu_y = smem + kSmemFixed;
	uint8_t* tu_u = tu_y + line_px;
	uint8_t* tu_v = tu_u + line_px / 2;
	uint8_t* tf_y = tu_v + line_px / 2;
	uint8_t* tf_u = tf_y + 4 * line_px;
	uint8_t* tf_v = tf_u + 2 * line_px;
	const int line_c = line_px / 2;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	WarpWs& ws = reinterpret_cast<WarpWs*>(smem + kSmemFixed + 10 * line_px)[warp];

	static_assert(sizeof(Vp8ImgDesc) <= 256, "descriptor must fit its shared-memory slot");
	static_assert(2 * NW <= kProgRing, "progress ring too small");

	// B_PRED lane table: btab[half][mode 0..10][pixel] = lane(a) | lane(b) << 8 | lane(c) << 16 | kind << 24, where the
	// prediction is (a + 2b + c + 2) >> 2 over edge-vector entries held by lanes half*16 + index (a 2-tap average is
	// the same formula with c = a), kind 1 = B_TM (clip(a + b - c) with a = L[r], b = A[c], c = P), kind 2 = B_DC,
	// kind 3 = out-of-range mode (constant 128, reference vp8_recon.c:352-356).
	if (RECON) {
		for (int i = tid; i < kBtabWords; i += NW * 32) {
			const int h = i / 176, m = (i % 176) / 16, p = i % 16, base = h * 16;
			uint32_t a = 0, b = 0, c = 0, kind = 0;
			if (m == 0) kind = 2;
			else if (m == 1) { a = 5 - (p >> 2); b = 7 + (p & 3); c = 6; kind = 1; }
			else if (m == 10) kind = 3;
			else {
				const int tap = c_bpred_taps[m * 16 + p], t0 = tap & 15;
				a = t0; b = t0 + 1; c = (tap & 16) ? t0 : t0 + 2;
			}
			btab[i] = (base + a) | ((base + b) << 8) | ((base + c) << 16) | (kind << 24);
		}
	}

	// lane roles that never change
	//   transform / block-per-lane prediction: lanes 0..15 luma block, 16..19 U block, 20..23 V block, 24 Y2
	const bool is_luma_lane = lane < 16;
	const int cb = lane & 3; // chroma block index within its plane
	const int blk_bx = is_luma_lane ? (lane & 3) * 4 : (cb & 1) * 4;
	const int blk_by = is_luma_lane ? (lane >> 2) * 4 : (cb >> 1) * 4;
	//   B_PRED pixel-per-lane: two sub-blocks in flight, 16 lanes each. At step s half h works on sub-block
	//   (row (s>>1)-h, column (s&1)+2h): every address is a per-lane base plus a per-step compile-time constant.
	const int half = lane >> 4, px_i = lane & 15, px_r = px_i >> 2, px_c = px_i & 3;
	const int e_dy = px_i <= 2 ? 3 : (px_i <= 5 ? 5 - px_i : -1);
	const int e_dx = px_i <= 6 ? -1 : (px_i == 15 ? 7 : px_i - 7);
	uint8_t* const bp_edge = ws.rt_y + half * (8 - 96) + e_dy * 24 + e_dx;      // + step constant -> this lane's edge byte
	uint8_t* const bp_out = ws.rt_y + half * (8 - 96) + px_r * 24 + px_c;        // + step constant -> this lane's pixel
	const int16_t* const bp_res = &ws.res[0][0] + half * (-2 * 16) + px_i;       // + step constant -> this lane's residual
	const uint32_t* const bp_tab = btab + half * 176 + px_i;                     // + mode*16 -> this lane's table word
	const int bp_blk = -2 * half;                                                // + step constant -> sub-block index
	const bool dc_tap = (px_i >= 2 && px_i <= 5) || (px_i >= 7 && px_i <= 10);
	//   coefficient stream of this lane (lanes 0..24), in int16 units per macroblock
	const int cstep = is_luma_lane ? 256 : (lane < 24 ? 64 : 16);

	for (int img = blockIdx.x; img < n_images; img += gridDim.x) {
		__syncthreads(); // previous image fully retired before its line buffers and descriptor are reused
		{
			const uint32_t* src = reinterpret_cast<const uint32_t*>(descs + img);
			uint32_t* dst = reinterpret_cast<uint32_t*>(sd);
			for (int i = tid; i < (int)(sizeof(Vp8ImgDesc) / 4); i += NW * 32) dst[i] = src[i];
			if (tid < kProgRing) prog[tid] = 0;
		}
		__syncthreads();

		const int cols = sd->mb_cols, rows = sd->mb_rows;
		OutPlane oy{sd->out_y, sd->out_stride_y, sd->out_w, sd->out_h,
		            ((reinterpret_cast<uintptr_t>(sd->out_y) | sd->out_stride_y) & 3) == 0};
		const uint32_t ocw = (sd->out_w + 1) >> 1, och = (sd->out_h + 1) >> 1;
		OutPlane ou{sd->out_u, sd->out_stride_uv, ocw, och,
		            ((reinterpret_cast<uintptr_t>(sd->out_u) | sd->out_stride_uv) & 3) == 0};
		OutPlane ov{sd->out_v, sd->out_stride_uv, ocw, och,
		            ((reinterpret_cast<uintptr_t>(sd->out_v) | sd->out_stride_uv) & 3) == 0};
		const bool words_ok = oy.word_ok && ou.word_ok && ov.word_ok;
		const bool lf_simple = sd->lf_simple != 0;
		const uint8_t* const g_ymode = sd->ymode;
		const uint8_t* const g_seg = sd->segment_id;

		for (int y = warp; y < rows; y += NW) {
			const bool last_row = (y == rows - 1);
			const size_t mb_row0 = (size_t)y * cols;

			// ---- row start: out-of-frame left neighbours (129) and corner (127 on the top row, else 129)
			const int16_t* cptr = nullptr;
			if (RECON) {
				ws.lcol[lane] = 129;
				if (lane < 16) ws.rt_y[(lane + 1) * 24 + 3] = 129;
				else if (lane < 24) ws.rt_u[(lane - 16 + 1) * 12 + 3] = 129;
				else ws.rt_v[(lane - 24 + 1) * 12 + 3] = 129;
				const uint8_t corner = (y == 0) ? 127 : 129;
				if (lane == 0) ws.rt_y[3] = corner;
				if (lane == 1) ws.rt_u[3] = corner;
				if (lane == 2) ws.rt_v[3] = corner;
				// coefficient stream of this lane, and the first macroblock's blocks on their way to shared memory
				if (is_luma_lane) cptr = sd->coeff_y + (mb_row0 * 16 + lane) * 16;
				else if (lane < 20) cptr = sd->coeff_u + (mb_row0 * 4 + cb) * 16;
				else if (lane < 24) cptr = sd->coeff_v + (mb_row0 * 4 + cb) * 16;
				else cptr = sd->coeff_y2 + mb_row0 * 16; // lane 24; lanes 25..31 carry it along unused
				if (lane < 25) {
					cp_async16(&ws.coef[lane], cptr);
					cp_async16(&ws.coef[25 + lane], cptr + 8);
				}
				cp_async_commit();
			}
			int ymode_next = g_ymode[mb_row0];
			int seg_next = g_seg ? g_seg[mb_row0] : 0;

			for (int x = 0; x < cols; x++) {
				const size_t mb = mb_row0 + x;
				const bool last_col = (x == cols - 1);

				// ---- per-macroblock syntax (this macroblock's was fetched one iteration ago)
				const int ymode = ymode_next;
				const bool bpred = (ymode == 4);
				const int seg = seg_next & 3;
				if (!last_col) {
					ymode_next = g_ymode[mb + 1];
					if (g_seg) seg_next = g_seg[mb + 1];
				}
				int uvmode = 0, bmode = 0;
				if (RECON) {
					uvmode = sd->uv_mode[mb];
					if (bpred && lane < 16) bmode = min((int)sd->bmode[mb * 16 + lane], 10);
				}

				// ---- wait for MB(x+1, y-1) (or the end of the row above)
				if (y > 0) {
					if (lane == 0) {
						const int target = y * kStampRow + min(x + 2, cols);
						while (prog[(y - 1) & (kProgRing - 1)] < target) __nanosleep(64);
						__threadfence_block();
					}
					__syncwarp();
				}

				if (RECON) {
					// ================================================================== m06: reconstruction
					// ---- top border (row -1) from the unfiltered line buffer, 127 above the frame
					if (lane < 9) {
						uint32_t w = 0x7f7f7f7fu;
						if (y > 0) {
							if (lane < 5) {
								if (lane == 4 && last_col) w = 0x01010101u * tu_y[16 * x + 15];
								else w = ld32(tu_y + 16 * x + 4 * lane);
							} else if (lane < 7) {
								w = ld32(tu_u + 8 * x + 4 * (lane - 5));
							} else {
								w = ld32(tu_v + 8 * x + 4 * (lane - 7));
							}
						}
						if (lane < 5) {
							st32(ws.rt_y + 4 + 4 * lane, w);
							if (lane == 4) { // above-right of sub-block column 3 always comes from the MB row above
								st32(ws.rt_y + 4 * 24 + 20, w);
								st32(ws.rt_y + 8 * 24 + 20, w);
								st32(ws.rt_y + 12 * 24 + 20, w);
							}
						} else if (lane < 7) {
							st32(ws.rt_u + 4 + 4 * (lane - 5), w);
						} else {
							st32(ws.rt_v + 4 + 4 * (lane - 7), w);
						}
					}

					// ---- coefficients: landed by cp.async, pick up this lane's block, then refill for MB x+1
					cp_async_wait_all();
					__syncwarp();
					uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
					if (lane < 25) {
						c0 = ws.coef[lane];
						c1 = ws.coef[25 + lane];
					}
					__syncwarp();
					if (!last_col) {
						cptr += cstep; // loop-carried on purpose: keeps the stream pointer in registers
						if (lane < 25) {
							cp_async16(&ws.coef[lane], cptr);
							cp_async16(&ws.coef[25 + lane], cptr + 8);
						}
						cp_async_commit();
					}

					// ---- dequantise + inverse transforms, one 4x4 block per lane (lanes 0..24)
					int r[16];
					bool any = false;
					{
						int v[16];
						const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
						const int16_t* dq = sd->dq[seg];
						const int dcq = is_luma_lane ? dq[0] : (lane < 24 ? dq[2] : dq[4]);
						const int acq = is_luma_lane ? dq[1] : (lane < 24 ? dq[3] : dq[5]);
#pragma unroll
						for (int i = 0; i < 8; i++) {
							v[2 * i] = s16(s16(cw[i]) * (i == 0 ? dcq : acq));
							v[2 * i + 1] = s16(((int)cw[i] >> 16) * acq);
						}
						if (!bpred) {
							// Y2: lane 24 runs the WHT, luma lanes take their DC from it (vp8_recon.c:563-586)
							if (lane == 24) {
								if ((c0.x | c0.y | c0.z | c0.w | c1.x | c1.y | c1.z | c1.w) == 0) {
									uint4* z = reinterpret_cast<uint4*>(ws.res[0]);
									z[0] = make_uint4(0, 0, 0, 0);
									z[1] = make_uint4(0, 0, 0, 0);
								} else {
									int d[16];
									iwht4x4(v, d);
#pragma unroll
									for (int i = 0; i < 16; i++) ws.res[0][i] = (int16_t)d[i];
								}
							}
							__syncwarp();
							if (is_luma_lane) v[0] = ws.res[0][lane];
						}
						uint32_t ac_or = 0;
#pragma unroll
						for (int i = 1; i < 16; i++) ac_or |= (uint32_t)v[i];
						any = (lane < 24) && ((ac_or | (uint32_t)v[0]) != 0);
						if (any) {
							if (ac_or == 0) {
								const int dc = s16((v[0] + 4) >> 3); // DC-only block: flat residual
#pragma unroll
								for (int i = 0; i < 16; i++) r[i] = dc;
							} else {
								idct4x4(v, r);
							}
						} else {
#pragma unroll
							for (int i = 0; i < 16; i++) r[i] = 0;
						}
					}
					__syncwarp(); // borders visible; res[0] consumed

					// ---- block-per-lane prediction: i16 luma (lanes 0..15) and chroma (lanes 16..23)
					if (bpred && is_luma_lane) {
						// park the residuals for the pixel-per-lane pass
						uint32_t pk[8];
#pragma unroll
						for (int i = 0; i < 8; i++) pk[i] = (uint32_t)(r[2 * i] & 0xffff) | ((uint32_t)r[2 * i + 1] << 16);
						uint4* dst = reinterpret_cast<uint4*>(ws.res[lane]);
						dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
						dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
					} else if (lane < 24) {
						uint8_t* tile;
						const uint8_t* lc;
						int stride, nwords, mode;
						if (is_luma_lane) { tile = ws.rt_y; lc = ws.lcol; stride = 24; nwords = 4; mode = ymode; }
						else if (lane < 20) { tile = ws.rt_u; lc = ws.lcol + 16; stride = 12; nwords = 2; mode = uvmode; }
						else { tile = ws.rt_v; lc = ws.lcol + 24; stride = 12; nwords = 2; mode = uvmode; }
						uint32_t pw[4];
						if (mode == 1) { // V: above row
							const uint32_t a = ld32(tile + 4 + blk_bx);
							pw[0] = pw[1] = pw[2] = pw[3] = a;
						} else if (mode == 2) { // H: left column
							const uint32_t l = ld32(lc + blk_by);
#pragma unroll
							for (int k = 0; k < 4; k++) pw[k] = ((l >> (8 * k)) & 255u) * 0x01010101u;
						} else if (mode == 3) { // TM: clip(L + A - P)
							const uint32_t a = ld32(tile + 4 + blk_bx), l = ld32(lc + blk_by);
							const int p = tile[3];
#pragma unroll
							for (int k = 0; k < 4; k++) {
								const int base = (int)((l >> (8 * k)) & 255u) - p;
								uint32_t w = 0;
#pragma unroll
								for (int j = 0; j < 4; j++) w |= (uint32_t)add_clip255(base, (int)((a >> (8 * j)) & 255u)) << (8 * j);
								pw[k] = w;
							}
						} else { // DC (also any out-of-range mode): vp8_recon.c:152-176
							uint32_t sum = 0;
							const bool have_a = y > 0, have_l = x > 0;
							for (int k = 0; k < nwords; k++) {
								if (have_a) sum += sum4(ld32(tile + 4 + 4 * k));
								if (have_l) sum += sum4(ld32(lc + 4 * k));
							}
							const int lg = (nwords == 4) ? 4 : 3; // log2(n)
							uint32_t dcv;
							if (have_a && have_l) dcv = (sum + (1u << lg)) >> (lg + 1);
							else if (have_a || have_l) dcv = (sum + (1u << (lg - 1))) >> lg;
							else dcv = 128;
							pw[0] = pw[1] = pw[2] = pw[3] = dcv * 0x01010101u;
						}
						uint8_t* dst = tile + (blk_by + 1) * stride + 4 + blk_bx;
#pragma unroll
						for (int k = 0; k < 4; k++) {
							uint32_t w = pw[k];
							if (any) {
								uint32_t o = 0;
#pragma unroll
								for (int j = 0; j < 4; j++) o |= (uint32_t)add_clip255((int)((w >> (8 * j)) & 255u), r[4 * k + j]) << (8 * j);
								w = o;
							}
							st32(dst + k * stride, w);
						}
					}
					__syncwarp();

					// ---- B_PRED luma: sub-block wavefront, step s handles sub-blocks with col + 2*row == s
					if (bpred) {
						// which sub-blocks need the two non-tap predictors (bit = sub-block index)
						const uint32_t special = __ballot_sync(0xffffffffu, lane < 16 && (bmode <= 1 || bmode == 10));
#if VP8_BPRED_LOOP
#pragma unroll 1
						for (int s2 = 0; s2 < 5; s2++) {
							// steps 2*s2 and 2*s2+1: half 0 works on sub-block row s2 (columns 0,1), half 1 on row s2-1 (columns 2,3)
							const bool active = half == 0 ? s2 < 4 : s2 > 0;
							// bit k set: step k of this pair has a sub-block that needs B_DC / B_TM / out-of-range handling
							const uint32_t sp2 = (s2 < 4 ? (special >> (4 * s2)) & 3u : 0u) | (s2 > 0 ? (special >> (4 * s2 - 2)) & 3u : 0u);
#pragma unroll
							for (int k = 0; k < 2; k++) {
								const int blk0 = 4 * s2 + k;                                  // half 0's sub-block; half 1 has blk0 - 2
								const int tile_c = (4 * s2 + 1) * 24 + 4 + 4 * k;            // pixel (0,0) of half 0's sub-block
								const int mode = __shfl_sync(0xffffffffu, bmode, blk0 + bp_blk);
								uint32_t t = 0;
								int e = 0, rs = 0;
								if (active) {
									t = bp_tab[mode * 16];
									e = bp_edge[tile_c];
									rs = bp_res[blk0 * 16];
								}
								const int a = __shfl_sync(0xffffffffu, e, t);
								const int b = __shfl_sync(0xffffffffu, e, t >> 8);
								const int c = __shfl_sync(0xffffffffu, e, t >> 16);
								int v = (a + c + 2 + 2 * b) >> 2;
								if (sp2 & (1u << k)) {
									const uint32_t kind = t >> 24;
									if (kind == 1) v = clip255(a + b - c);
									// B_DC: (sum of A0..A3 and L0..L3 + 4) >> 3, one warp-wide reduction per half
									const int mine = dc_tap ? e : 0;
									const int s_lo = __reduce_add_sync(0xffffffffu, half == 0 ? mine : 0);
									const int s_hi = __reduce_add_sync(0xffffffffu, half == 1 ? mine : 0);
									if (kind == 2) v = ((half ? s_hi : s_lo) + 4) >> 3;
									if (kind == 3) v = 128;
								}
								if (active) bp_out[tile_c] = (uint8_t)add_clip255(v, rs);
								__syncwarp();
							}
						}
					}
#else
#pragma unroll
						for (int s = 0; s < 10; s++) {
							const int blk0 = (s >> 1) * 4 + (s & 1);                   // half 0's sub-block; half 1 has blk0 - 2
							const int tile_c = ((s >> 1) * 4 + 1) * 24 + 4 + (s & 1) * 4; // pixel (0,0) of half 0's sub-block
							const bool active = (s < 2) ? (half == 0) : (s >= 8 ? (half == 1) : true);
							const int mode = __shfl_sync(0xffffffffu, bmode, blk0 + bp_blk);
							uint32_t t = 0;
							int e = 0, rs = 0;
							if (active) {
								t = bp_tab[mode * 16];
								e = bp_edge[tile_c];
								rs = bp_res[blk0 * 16];
							}
							const int a = __shfl_sync(0xffffffffu, e, t);
							const int b = __shfl_sync(0xffffffffu, e, t >> 8);
							const int c = __shfl_sync(0xffffffffu, e, t >> 16);
							int v = (a + c + 2 + 2 * b) >> 2;
							const uint32_t here = (s < 2 ? 0u : (1u << (blk0 - 2))) | (s >= 8 ? 0u : (1u << blk0));
							if (special & here) {
								const uint32_t kind = t >> 24;
								if (kind == 1) v = clip255(a + b - c);
								// B_DC: (sum of A0..A3 and L0..L3 + 4) >> 3, one warp-wide reduction per half
								const int mine = dc_tap ? e : 0;
								const int s_lo = __reduce_add_sync(0xffffffffu, half == 0 ? mine : 0);
								const int s_hi = __reduce_add_sync(0xffffffffu, half == 1 ? mine : 0);
								if (kind == 2) v = ((half ? s_hi : s_lo) + 4) >> 3;
								if (kind == 3) v = 128;
							}
							if (active) bp_out[tile_c] = (uint8_t)add_clip255(v, rs);
							__syncwarp();
						}
					}
#endif
				} else {
					// ================================================================== stand-alone m07: load the MB
					const uint8_t* sy = sd->src_y + (size_t)(16 * y) * sd->src_stride_y + 16 * x;
#pragma unroll
					for (int k = 0; k < 2; k++) {
						const int i = lane + 32 * k, row = i >> 2, wc = i & 3;
						st32(ws.rt_y + (row + 1) * 24 + 4 + 4 * wc, ld32(sy + (size_t)row * sd->src_stride_y + 4 * wc));
					}
					{
						const int j = lane & 15, row = j >> 1, wc = j & 1;
						const uint8_t* sp = (lane < 16 ? sd->src_u : sd->src_v) + (size_t)(8 * y + row) * sd->src_stride_uv + 8 * x + 4 * wc;
						st32((lane < 16 ? ws.rt_u : ws.rt_v) + (row + 1) * 12 + 4 + 4 * wc, ld32(sp));
					}
					__syncwarp();
				}

				// ---- snapshot of the unfiltered tile that the neighbours will need (read phase)
				uint32_t edge_b = 0, bot_w = 0, corner_b = 0;
				if (RECON) {
					if (lane < 16) edge_b = ws.rt_y[(lane + 1) * 24 + 4 + 15];
					else if (lane < 24) edge_b = ws.rt_u[(lane - 16 + 1) * 12 + 4 + 7];
					else edge_b = ws.rt_v[(lane - 24 + 1) * 12 + 4 + 7];
					if (lane < 4) bot_w = ld32(ws.rt_y + 16 * 24 + 4 + 4 * lane);
					else if (lane < 6) bot_w = ld32(ws.rt_u + 8 * 12 + 4 + 4 * (lane - 4));
					else if (lane < 8) bot_w = ld32(ws.rt_v + 8 * 12 + 4 + 4 * (lane - 6));
					if (lane == 8) corner_b = ws.rt_y[4 + 15];
					if (lane == 9) corner_b = ws.rt_u[4 + 7];
					if (lane == 10) corner_b = ws.rt_v[4 + 7];
				}

				if (!FILTER) {
					// ================================================================== unfiltered output (-yuv)
					const int row0 = lane >> 2, wc = lane & 3;
					const int j = lane & 15, crow = j >> 1, cwc = j & 1;
					const uint32_t w0 = ld32(ws.rt_y + (row0 + 1) * 24 + 4 + 4 * wc), w1 = ld32(ws.rt_y + (row0 + 9) * 24 + 4 + 4 * wc);
					const uint32_t wcx = ld32((lane < 16 ? ws.rt_u : ws.rt_v) + (crow + 1) * 12 + 4 + 4 * cwc);
					if (words_ok && !last_col && !last_row) {
						uint8_t* d = oy.p + (size_t)(16 * y + row0) * oy.stride + 16 * x + 4 * wc;
						st32(d, w0);
						st32(d + (size_t)8 * oy.stride, w1);
						st32((lane < 16 ? ou.p : ov.p) + (size_t)(8 * y + crow) * ou.stride + 8 * x + 4 * cwc, wcx);
					} else {
						put_word(oy, 16 * x + 4 * wc, 16 * y + row0, w0);
						put_word(oy, 16 * x + 4 * wc, 16 * y + row0 + 8, w1);
						put_word(lane < 16 ? ou : ov, 8 * x + 4 * cwc, 8 * y + crow, wcx);
					}
				} else {
					// ================================================================== m07: loop filter
					// ---- assemble the filter tile: left apron = previous tile's right 4 columns (all 20 / 12 rows, so the
					//      corner above-left travels along), top apron = tf lines, interior = the fresh reconstruction
					uint32_t la, la2 = 0, ta = 0, in0, in1, inc;
					if (lane < 16) la = ld32(ws.ft_y + (lane + 4) * 20 + 16);
					else if (lane < 24) la = ld32(ws.ft_u + (lane - 16 + 4) * 12 + 8);
					else la = ld32(ws.ft_v + (lane - 24 + 4) * 12 + 8);
					if (lane < 4) la2 = ld32(ws.ft_y + lane * 20 + 16);
					else if (lane < 8) la2 = ld32(ws.ft_u + (lane - 4) * 12 + 8);
					else if (lane < 12) la2 = ld32(ws.ft_v + (lane - 8) * 12 + 8);
					if (y > 0) {
						if (lane < 16) ta = ld32(tf_y + (lane >> 2) * line_px + 16 * x + 4 * (lane & 3));
						else {
							const int j = lane & 7; // 4 rows x 2 words
							ta = ld32((lane < 24 ? tf_u : tf_v) + (j >> 1) * line_c + 8 * x + 4 * (j & 1));
						}
					}
					in0 = ld32(ws.rt_y + ((lane >> 2) + 1) * 24 + 4 + 4 * (lane & 3));
					in1 = ld32(ws.rt_y + ((lane >> 2) + 9) * 24 + 4 + 4 * (lane & 3));
					{
						const int j = lane & 15;
						inc = ld32((lane < 16 ? ws.rt_u : ws.rt_v) + ((j >> 1) + 1) * 12 + 4 + 4 * (j & 1));
					}
					__syncwarp();
					if (lane < 16) st32(ws.ft_y + (lane + 4) * 20, la);
					else if (lane < 24) st32(ws.ft_u + (lane - 16 + 4) * 12, la);
					else st32(ws.ft_v + (lane - 24 + 4) * 12, la);
					if (lane < 4) st32(ws.ft_y + lane * 20, la2);
					else if (lane < 8) st32(ws.ft_u + (lane - 4) * 12, la2);
					else if (lane < 12) st32(ws.ft_v + (lane - 8) * 12, la2);
					if (lane < 16) st32(ws.ft_y + (lane >> 2) * 20 + 4 + 4 * (lane & 3), ta);
					else {
						const int j = lane & 7;
						st32((lane < 24 ? ws.ft_u : ws.ft_v) + (j >> 1) * 12 + 4 + 4 * (j & 1), ta);
					}
					st32(ws.ft_y + ((lane >> 2) + 4) * 20 + 4 + 4 * (lane & 3), in0);
					st32(ws.ft_y + ((lane >> 2) + 12) * 20 + 4 + 4 * (lane & 3), in1);
					{
						const int j = lane & 15;
						st32((lane < 16 ? ws.ft_u : ws.ft_v) + ((j >> 1) + 4) * 12 + 4 + 4 * (j & 1), inc);
					}
					__syncwarp();

					// ---- edge filters in the reference's order (vp8_loopfilter.c:226-277)
					const uint8_t* lfp = sd->lf[seg][bpred ? 1 : 0];
					const int level = lfp[0], interior = lfp[1], hev_thr = lfp[2];
					if (level > 0) {
						const bool inner = bpred || (sd->has_coeff && sd->has_coeff[mb]);
						const int lim_mb = 2 * (level + 2) + interior, lim_in = 2 * level + interior;
						// this lane's line (row for column edges, column for row edges) in luma or chroma
						uint8_t* t;
						int stride, n;
						if (lane < 16) { t = ws.ft_y + 4 * 20 + 4; stride = 20; n = lane; }
						else if (lane < 24) { t = ws.ft_u + 4 * 12 + 4; stride = 12; n = lane - 16; }
						else { t = ws.ft_v + 4 * 12 + 4; stride = 12; n = lane - 24; }
#if VP8_LF_GENERIC
						// pass p: p < 4 edges between columns at 4*(p&3), p >= 4 edges between rows. Edge 0 is the macroblock edge
						// (needs the neighbour), edges 4/8/12 are inner edges (luma has three, chroma only the one at 4; the
						// simple filter touches luma only).
#pragma unroll 1
						for (int p = 0; p < 8; p++) {
							const int e = (p & 3) * 4;
							const bool rows_dir = p >= 4;
							if (e == 0 ? (rows_dir ? y == 0 : x == 0) : !inner) continue;
							if (lane < 16 || (e <= 4 && !lf_simple)) {
								uint8_t* q = rows_dir ? t + e * stride + n : t + n * stride + e;
								lf_line(q, rows_dir ? stride : 1, lf_simple ? EDGE_SIMPLE : (e == 0 ? EDGE_MB : EDGE_INNER),
								        e == 0 ? lim_mb : lim_in, interior, hev_thr);
							}
							__syncwarp();
						}
					}

#else
						if (!lf_simple) {
							if (x > 0) {
								lf_across_columns<EDGE_MB>(t + n * stride, lim_mb, interior, hev_thr);
								__syncwarp();
							}
							if (inner) {
								lf_across_columns<EDGE_INNER>(t + n * stride + 4, lim_in, interior, hev_thr);
								__syncwarp();
								if (lane < 16) lf_across_columns<EDGE_INNER>(t + n * stride + 8, lim_in, interior, hev_thr);
								__syncwarp();
								if (lane < 16) lf_across_columns<EDGE_INNER>(t + n * stride + 12, lim_in, interior, hev_thr);
								__syncwarp();
							}
							if (y > 0) {
								lf_across_rows<EDGE_MB>(t + n, stride, lim_mb, interior, hev_thr);
								__syncwarp();
							}
							if (inner) {
								lf_across_rows<EDGE_INNER>(t + 4 * stride + n, stride, lim_in, interior, hev_thr);
								__syncwarp();
								if (lane < 16) lf_across_rows<EDGE_INNER>(t + 8 * stride + n, stride, lim_in, interior, hev_thr);
								__syncwarp();
								if (lane < 16) lf_across_rows<EDGE_INNER>(t + 12 * stride + n, stride, lim_in, interior, hev_thr);
								__syncwarp();
							}
						} else if (lane < 16) {
							// simple filter: luma only (vp8_loopfilter.c:228-244)
							if (x > 0) lf_across_columns<EDGE_SIMPLE>(t + n * stride, lim_mb, 0, 0);
							__syncwarp(0xffffu);
							if (inner) {
								for (int e = 4; e < 16; e += 4) {
									lf_across_columns<EDGE_SIMPLE>(t + n * stride + e, lim_in, 0, 0);
									__syncwarp(0xffffu);
								}
							}
							if (y > 0) lf_across_rows<EDGE_SIMPLE>(t + n, stride, lim_mb, 0, 0);
							__syncwarp(0xffffu);
							if (inner) {
								for (int e = 4; e < 16; e += 4) {
									lf_across_rows<EDGE_SIMPLE>(t + e * stride + n, stride, lim_in, 0, 0);
									__syncwarp(0xffffu);
								}
							}
						}
						__syncwarp();
					}

#endif
					// ---- store what can no longer change: the 16x16 (8x8) block whose origin is 4 pixels up and left of the
					//      macroblock; the last column / row of macroblocks also flush the strips nobody else will
					{
						const int row0 = lane >> 2, wc = lane & 3;                      // luma: rows row0, row0+8; word wc
						const int j = lane & 15, crow = j >> 1, cwc = j & 1;            // chroma: plane lane>>4
						const uint32_t w0 = ld32(ws.ft_y + row0 * 20 + 4 * wc), w1 = ld32(ws.ft_y + (row0 + 8) * 20 + 4 * wc);
						const uint32_t wcx = ld32((lane < 16 ? ws.ft_u : ws.ft_v) + crow * 12 + 4 * cwc);
						if (words_ok && x > 0 && y > 0 && !last_col && !last_row) {
							// whole block inside the frame: only the last column / row of macroblocks can be cropped
							uint8_t* d = oy.p + (size_t)(16 * y - 4 + row0) * oy.stride + (16 * x - 4 + 4 * wc);
							st32(d, w0);
							st32(d + (size_t)8 * oy.stride, w1);
							st32((lane < 16 ? ou.p : ov.p) + (size_t)(8 * y - 4 + crow) * ou.stride + (8 * x - 4 + 4 * cwc), wcx);
						} else {
							put_word(oy, 16 * x - 4 + 4 * wc, 16 * y - 4 + row0, w0);
							put_word(oy, 16 * x - 4 + 4 * wc, 16 * y + 4 + row0, w1);
							put_word(lane < 16 ? ou : ov, 8 * x - 4 + 4 * cwc, 8 * y - 4 + crow, wcx);
						}
						if (last_col) { // right strip: columns 12..15 (4..7), rows -4..11 (-4..3)
							if (lane < 16) put_word(oy, 16 * x + 12, 16 * y - 4 + lane, ld32(ws.ft_y + lane * 20 + 16));
							else {
								const int k = lane & 7;
								put_word(lane < 24 ? ou : ov, 8 * x + 4, 8 * y - 4 + k, ld32((lane < 24 ? ws.ft_u : ws.ft_v) + k * 12 + 8));
							}
						}
						if (last_row) { // bottom strip: rows 12..15 (4..7), columns -4..15 (-4..7)
							if (lane < 20) {
								const int rr = lane / 5, ww = lane % 5;
								if (ww < 4 || last_col) put_word(oy, 16 * x - 4 + 4 * ww, 16 * y + 12 + rr, ld32(ws.ft_y + (16 + rr) * 20 + 4 * ww));
							}
							if (lane < 24) {
								const int pl = lane / 12, k = lane % 12, rr = k / 3, ww = k % 3;
								if (ww < 2 || last_col)
									put_word(pl ? ov : ou, 8 * x - 4 + 4 * ww, 8 * y + 4 + rr, ld32((pl ? ws.ft_v : ws.ft_u) + (8 + rr) * 12 + 4 * ww));
							}
						}
					}
					if (!last_row) { // hand the bottom 4 filtered rows to the row below: columns -4..11 (+12..15 on the last column)
						if (lane < 16) {
							if (x > 0 || (lane & 3)) st32(tf_y + (lane >> 2) * line_px + 16 * x - 4 + 4 * (lane & 3), ld32(ws.ft_y + ((lane >> 2) + 16) * 20 + 4 * (lane & 3)));
						} else {
							const int k = lane & 7; // 4 rows x 2 words per plane
							if (x > 0 || (k & 1))
								st32((lane < 24 ? tf_u : tf_v) + (k >> 1) * line_c + 8 * x - 4 + 4 * (k & 1), ld32((lane < 24 ? ws.ft_u : ws.ft_v) + ((k >> 1) + 8) * 12 + 4 * (k & 1)));
						}
						if (last_col) {
							if (lane < 4) st32(tf_y + lane * line_px + 16 * x + 12, ld32(ws.ft_y + (lane + 16) * 20 + 16));
							else if (lane < 12) {
								const int k = lane - 4;
								st32(((k >> 2) ? tf_v : tf_u) + (k & 3) * line_c + 8 * x + 4, ld32(((k >> 2) ? ws.ft_v : ws.ft_u) + ((k & 3) + 8) * 12 + 8));
							}
						}
					}
				}

				// ---- hand the unfiltered borders on (write phase) and publish progress
				if (RECON) {
					__syncwarp();
					ws.lcol[lane] = (uint8_t)edge_b;
					if (lane < 16) ws.rt_y[(lane + 1) * 24 + 3] = (uint8_t)edge_b;
					else if (lane < 24) ws.rt_u[(lane - 16 + 1) * 12 + 3] = (uint8_t)edge_b;
					else ws.rt_v[(lane - 24 + 1) * 12 + 3] = (uint8_t)edge_b;
					if (lane == 8) ws.rt_y[3] = (uint8_t)corner_b;
					if (lane == 9) ws.rt_u[3] = (uint8_t)corner_b;
					if (lane == 10) ws.rt_v[3] = (uint8_t)corner_b;
					if (!last_row) {
						if (lane < 4) st32(tu_y + 16 * x + 4 * lane, bot_w);
						else if (lane < 6) st32(tu_u + 8 * x + 4 * (lane - 4), bot_w);
						else if (lane < 8) st32(tu_v + 8 * x + 4 * (lane - 6), bot_w);
					}
				}
				__syncwarp();
				if (lane == 0) {
					__threadfence_block();
					prog[y & (kProgRing - 1)] = (y + 1) * kStampRow + x + 1;
				}
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ m08: YUV -> RGB
// Reference mult_hi / vp8_clip8 / vp8_yuv_to_rgb (yuv2rgb_ppm.c:19-41).
// (v & ~16383) == 0 ? v >> 6 : (v < 0 ? 0 : 255)  ==  clamp(v, 0, 16383) >> 6   (16383 >> 6 == 255)
__device__ __forceinline__ uint32_t fix_clip(int v) { return (uint32_t)__vimin_s32_relu(v, 16383) >> 6; }

__device__ __forceinline__ void yuv_to_rgb(int Y, int U, int V, uint32_t& R, uint32_t& G, uint32_t& B) {
	const int yy = (Y * 19077) >> 8;
	R = fix_clip(yy + ((V * 26149) >> 8) - 14234);
	G = fix_clip(yy - ((U * 6419) >> 8) - ((V * 13320) >> 8) + 8708);
	B = fix_clip(yy + ((U * 33050) >> 8) - 17685);
}

// Interior sample of the fancy upsampler: near/far rows N,F; near/far columns nc,fc.
__device__ __forceinline__ int fancy(int Nn, int Nf, int Fn, int Ff) {
	const int sum = Nn + Nf + Fn + Ff + 8;
	const int diag = (sum + 2 * (Nf + Fn)) >> 3;
	return (diag + Nn) >> 1;
}

constexpr int kRgbThreads = 256;

// 4 pixels px0..px0+3 of row py, any geometry (row ends, odd widths, unaligned planes). px0 is a multiple of 4.
__device__ __noinline__ void rgb_group4(const Vp8RgbDesc& d, uint32_t py, uint32_t px0, const uint8_t* un, const uint8_t* uf,
                                           const uint8_t* vn, const uint8_t* vf) {
	const uint32_t w = d.width, cw = (w + 1) >> 1;
	const uint8_t* yrow = d.y + (size_t)py * d.stride_y;
	// chroma columns j-1 .. j+2 with j = px0/2 cover all four pixels; clamp indices (clamped taps are unused)
	const int j = (int)(px0 >> 1);
	int un4[4], uf4[4], vn4[4], vf4[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const int c = min(max(j - 1 + k, 0), (int)cw - 1);
		un4[k] = un[c];
		uf4[k] = uf[c];
		vn4[k] = vn[c];
		vf4[k] = vf[c];
	}
	uint8_t out[12];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t px = px0 + k;
		if (px >= w) break;
		int U, V;
		if (px == 0) {
			U = (3 * un[0] + uf[0] + 2) >> 2;
			V = (3 * vn[0] + vf[0] + 2) >> 2;
		} else if (px == w - 1 && (w & 1) == 0) {
			U = (3 * un[cw - 1] + uf[cw - 1] + 2) >> 2;
			V = (3 * vn[cw - 1] + vf[cw - 1] + 2) >> 2;
		} else {
			// k=0: near j, far j-1; k=1: near j, far j+1; k=2: near j+1, far j; k=3: near j+1, far j+2  (array index = col-j+1)
			const int ni = (k < 2) ? 1 : 2, fi = (k == 0) ? 0 : (k == 1 ? 2 : (k == 2 ? 1 : 3));
			U = fancy(un4[ni], un4[fi], uf4[ni], uf4[fi]);
			V = fancy(vn4[ni], vn4[fi], vf4[ni], vf4[fi]);
		}
		uint32_t R, G, B;
		yuv_to_rgb(yrow[px], U, V, R, G, B);
		out[3 * k] = (uint8_t)R;
		out[3 * k + 1] = (uint8_t)G;
		out[3 * k + 2] = (uint8_t)B;
	}
	uint8_t* dst = d.rgb + ((size_t)py * w + px0) * 3;
	if (px0 + 4 <= w && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
		uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
		d32[0] = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
		d32[1] = out[4] | (out[5] << 8) | (out[6] << 16) | ((uint32_t)out[7] << 24);
		d32[2] = out[8] | (out[9] << 8) | (out[10] << 16) | ((uint32_t)out[11] << 24);
	} else {
		for (uint32_t k = 0; k < 12 && px0 * 3 + k < w * 3; k++) dst[k] = out[k];
	}
}

// One thread = 8 horizontally adjacent pixels of one image. grid = (blocks per image, images).
// Interior groups of aligned images take the vector path: one 8-byte luma load, one 4-byte load per chroma row and plane
// (+ the two neighbour bytes), 24 output bytes as three 8-byte stores. Row ends and unaligned images fall back to
// rgb_group4, which handles every geometry.
__global__ void __launch_bounds__(kRgbThreads, 4) vp8_i420_to_rgb(const Vp8RgbDesc* __restrict__ descs) {
	const Vp8RgbDesc d = descs[blockIdx.y];
	const uint32_t w = d.width, h = d.height, ch = (h + 1) >> 1;
	const uint32_t groups_per_row = (w + 7) / 8;
	const uint32_t g = blockIdx.x * kRgbThreads + threadIdx.x;
	if (g >= groups_per_row * h) return;
	const uint32_t py = g / groups_per_row, gx = g % groups_per_row, px0 = gx * 8;

	// near / far chroma rows (reference yuv420_write_ppm_fd row pairing, yuv2rgb_ppm.c:164-201)
	const uint32_t nrow = py >> 1;
	uint32_t frow;
	if (py == 0) frow = 0;
	else if (py & 1) frow = min(nrow + 1, ch - 1);
	else frow = nrow - 1;
	const uint8_t* un = d.u + (size_t)nrow * d.stride_uv;
	const uint8_t* uf = d.u + (size_t)frow * d.stride_uv;
	const uint8_t* vn = d.v + (size_t)nrow * d.stride_uv;
	const uint8_t* vf = d.v + (size_t)frow * d.stride_uv;

	const bool aligned = ((reinterpret_cast<uintptr_t>(d.y) | d.stride_y) & 7) == 0 &&
	                     ((reinterpret_cast<uintptr_t>(d.u) | reinterpret_cast<uintptr_t>(d.v) | d.stride_uv) & 3) == 0 &&
	                     (reinterpret_cast<uintptr_t>(d.rgb) & 7) == 0 && (w & 7) == 0;
	if (!aligned) {
		rgb_group4(d, py, px0, un, uf, vn, vf);
		if (px0 + 4 < w) rgb_group4(d, py, px0 + 4, un, uf, vn, vf);
		return;
	}

	const uint32_t j = px0 >> 1; // first chroma column of the group, a multiple of 4; columns j-1 .. j+4 are needed
	const bool first = gx == 0, last = gx == groups_per_row - 1;
	const uint2 yw = *reinterpret_cast<const uint2*>(d.y + (size_t)py * d.stride_y + px0);
	int n[2][6], f[2][6]; // [plane][column j-1+i]; the out-of-row neighbours of the first / last group are never used
	{
		const uint8_t* N[2] = {un, vn};
		const uint8_t* F[2] = {uf, vf};
#pragma unroll
		for (int p = 0; p < 2; p++) {
			const uint32_t a = *reinterpret_cast<const uint32_t*>(N[p] + j), b = *reinterpret_cast<const uint32_t*>(F[p] + j);
			n[p][0] = first ? 0 : N[p][j - 1];
			f[p][0] = first ? 0 : F[p][j - 1];
			n[p][5] = last ? 0 : N[p][j + 4];
			f[p][5] = last ? 0 : F[p][j + 4];
#pragma unroll
			for (int i = 0; i < 4; i++) {
				n[p][1 + i] = (a >> (8 * i)) & 255;
				f[p][1 + i] = (b >> (8 * i)) & 255;
			}
		}
	}
	uint32_t o[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
	for (int k = 0; k < 8; k++) {
		// even pixel: near column j+k/2, far one to the left; odd pixel: near column j+(k-1)/2, far one to the right
		const int ni = (k >> 1) + 1, fi = (k & 1) ? ni + 1 : ni - 1;
		int U = fancy(n[0][ni], n[0][fi], f[0][ni], f[0][fi]);
		int V = fancy(n[1][ni], n[1][fi], f[1][ni], f[1][fi]);
		// the first pixel of a row and (even widths) the last one have no far column: (3 near + far-row + 2) >> 2
		if ((k == 0 && first) || (k == 7 && last)) {
			U = (3 * n[0][ni] + f[0][ni] + 2) >> 2;
			V = (3 * n[1][ni] + f[1][ni] + 2) >> 2;
		}
		const int Y = (k < 4 ? yw.x >> (8 * k) : yw.y >> (8 * (k - 4))) & 255;
		uint32_t R, G, B;
		yuv_to_rgb(Y, U, V, R, G, B);
		o[(3 * k) >> 2] |= R << (8 * ((3 * k) & 3));
		o[(3 * k + 1) >> 2] |= G << (8 * ((3 * k + 1) & 3));
		o[(3 * k + 2) >> 2] |= B << (8 * ((3 * k + 2) & 3));
	}
	uint2* dst = reinterpret_cast<uint2*>(d.rgb + ((size_t)py * w + px0) * 3);
	dst[0] = make_uint2(o[0], o[1]);
	dst[1] = make_uint2(o[2], o[3]);
	dst[2] = make_uint2(o[4], o[5]);
}

template <int NW, bool RECON, bool FILTER>
int launch_t(const Vp8ImgDesc* descs, int n, int line_px, int grid, size_t smem, cudaStream_t st) {
	auto k = vp8_mb_wavefront<NW, RECON, FILTER>;
	cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return (int)e;
	k<<<grid, NW * 32, smem, st>>>(descs, n, line_px);
	return (int)cudaGetLastError();
}

template <int NW, bool RECON, bool FILTER>
int occupancy_t(size_t smem) {
	auto k = vp8_mb_wavefront<NW, RECON, FILTER>;
	if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
	int nb = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, NW * 32, smem) != cudaSuccess) return 0;
	return nb;
}

} // namespace

int vp8_wavefront_smem_bytes(int mode, int warps_per_image, int max_mb_cols) {
	(void)mode;
	return kSmemFixed + 10 * 16 * max_mb_cols + warps_per_image * (int)sizeof(WarpWs);
}

#define VP8_DISPATCH(FN, ...)                                                            \
	switch (mode * 100 + warps_per_image) {                                              \
		case 4: return FN<4, true, false>(__VA_ARGS__);                                   \
		case 8: return FN<8, true, false>(__VA_ARGS__);                                   \
		case 16: return FN<16, true, false>(__VA_ARGS__);                                 \
		case 32: return FN<32, true, false>(__VA_ARGS__);                                 \
		case 104: return FN<4, true, true>(__VA_ARGS__);                                  \
		case 108: return FN<8, true, true>(__VA_ARGS__);                                  \
		case 116: return FN<16, true, true>(__VA_ARGS__);                                 \
		case 132: return FN<32, true, true>(__VA_ARGS__);                                 \
		case 204: return FN<4, false, true>(__VA_ARGS__);                                 \
		case 208: return FN<8, false, true>(__VA_ARGS__);                                 \
		case 216: return FN<16, false, true>(__VA_ARGS__);                                \
		case 232: return FN<32, false, true>(__VA_ARGS__);                                \
		default: return -1;                                                               \
	}

int vp8_launch_wavefront(int mode, int warps_per_image, const Vp8ImgDesc* descs_dev, int n_images, int max_mb_cols,
                         int grid_ctas, void* stream) {
	const size_t smem = (size_t)vp8_wavefront_smem_bytes(mode, warps_per_image, max_mb_cols);
	const int line_px = 16 * max_mb_cols;
	cudaStream_t st = (cudaStream_t)stream;
	VP8_DISPATCH(launch_t, descs_dev, n_images, line_px, grid_ctas, smem, st)
}

int vp8_wavefront_max_ctas_per_sm(int mode, int warps_per_image, int max_mb_cols) {
	const size_t smem = (size_t)vp8_wavefront_smem_bytes(mode, warps_per_image, max_mb_cols);
	VP8_DISPATCH(occupancy_t, smem)
}

int vp8_launch_rgb(const Vp8RgbDesc* descs_dev, int n_images, uint32_t max_blocks_per_image, void* stream) {
	if (n_images <= 0) return 0;
	dim3 grid(max_blocks_per_image, (unsigned)n_images);
	vp8_i420_to_rgb<<<grid, kRgbThreads, 0, (cudaStream_t)stream>>>(descs_dev);
	return (int)cudaGetLastError();
}
