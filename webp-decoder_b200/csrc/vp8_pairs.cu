// vp8_pairs.cu - second-generation wavefront kernel: every HALF-warp is a macroblock engine.
//
// Same arithmetic, same single pass over HBM and the same shared-memory line-buffer scheme as vp8_mb_wavefront
// (vp8_kernels.cu, read that header first); what changes is the lane mapping, chosen from the ncu profile of that
// kernel (profiles/README.md: instruction-issue bound, ~70 % of lanes active, every per-macroblock overhead paid once
// per warp):
//
//   * a warp walks TWO macroblock rows in staggered lockstep: lanes 0..15 do MB(t, y), lanes 16..31 do MB(t-2, y+1).
//     Row y+1 trails row y by two macroblocks, which is exactly its dependency (left, top, top-right), so the pair
//     needs no flag between its rows; only every second row boundary spins on a progress stamp.
//   * 16 lanes per macroblock: transforms two 4x4 blocks per lane, B_PRED one sub-block per step with one lane per
//     pixel (16 raster steps), loop filter one lane per pixel row / column (luma) or per U/V row / column.
//     Every phase now runs with (close to) all 32 lanes busy, and border / progress / store bookkeeping is paid once
//     per two macroblocks.
//   * the four filtered rows handed to the row below (tf) live in an L2-resident global scratch line per CTA instead
//     of shared memory: shared memory per image drops from 10 to 2 bytes per pixel column, which is what lets 7-8
//     images stay resident per SM although every warp now carries two macroblock workspaces.
#include "vp8_common.cuh"

namespace {

// Per-half-warp workspace (one macroblock). Tile layouts as in WarpWs (vp8_common.cuh).
//   coef: this macroblock's coefficient blocks as landed by cp.async. Lane l < 12 owns blocks 2l, 2l+1 (lanes 0..7 luma,
//         8..9 U, 10..11 V), lane 12 owns Y2; 16-byte chunk q of lane l sits at [q*13 + l] (q = 2*block_in_lane + half
//         of the block), so every read is conflict-free.
struct __align__(16) HalfWs {
	uint8_t rt_y[17 * 24];
	uint8_t rt_u[9 * 12];
	uint8_t rt_v[9 * 12];
	uint8_t lcol[32];
	int16_t res[16][16];
	uint8_t ft_y[20 * 20];
	uint8_t ft_u[12 * 12];
	uint8_t ft_v[12 * 12];
	uint4 coef[52];
};
static_assert(sizeof(HalfWs) % 16 == 0 && offsetof(HalfWs, res) % 16 == 0 && offsetof(HalfWs, coef) % 16 == 0, "HalfWs alignment");

__device__ __forceinline__ uint32_t cluster_ctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ldcg32(const uint8_t* p) { return __ldcg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ void stcg32(uint8_t* p, uint32_t v) { __stcg(reinterpret_cast<uint32_t*>(p), v); }
// unfiltered line buffer: shared memory normally, L2 (never L1) when the image is spread over a cluster
template <bool CL>
__device__ __forceinline__ uint32_t ld_line(const uint8_t* p) { return CL ? ldcg32(p) : ld32(p); }
template <bool CL>
__device__ __forceinline__ void st_line(uint8_t* p, uint32_t v) {
	if (CL) stcg32(p, v);
	else st32(p, v);
}

// Compact layout: request this lane's (up to two) present blocks of one macroblock; returns their presence bits.
__device__ __forceinline__ uint32_t fetch_compact(HalfWs& ws, const uint8_t* packed, uint32_t mask, uint32_t first, int bit0, int hl) {
	if (hl >= 13) return 0;
	const uint32_t nz = (mask >> bit0) & (hl < 12 ? 3u : 1u);
	uint32_t idx = first + __popc(mask & ((1u << bit0) - 1u));
	if (nz & 1) {
		const uint8_t* src = packed + (size_t)idx * 32;
		cp_async16(&ws.coef[0 * 13 + hl], src);
		cp_async16(&ws.coef[1 * 13 + hl], src + 16);
		idx++;
	}
	if (nz & 2) {
		const uint8_t* src = packed + (size_t)idx * 32;
		cp_async16(&ws.coef[2 * 13 + hl], src);
		cp_async16(&ws.coef[3 * 13 + hl], src + 16);
	}
	return nz;
}

constexpr int kClusterProg = 1024; // progress stamps per image in cluster mode (VP8 frames have at most 1024 macroblock rows)

#ifndef VP8P_BPRED_UNROLL
#define VP8P_BPRED_UNROLL 16 // sub-block steps unrolled per loop iteration (16 = fully unrolled)
#endif
#ifndef VP8P_BLOCK_UNROLL
#define VP8P_BLOCK_UNROLL 2 // 2: both 4x4 blocks of a lane unrolled, 1: looped
#endif
#ifndef VP8P_LF_COMPACT
#define VP8P_LF_COMPACT 0 // 1: ONE byte-addressed filter body looped over the 12 edge passes (smallest code)
#endif
#ifndef VP8P_LF_LOOP
#define VP8P_LF_LOOP 0 // 1: inner edges 4/8/12 as a loop, 0: unrolled
#endif
#define VP8P_STR2(x) #x
#define VP8P_STR(x) VP8P_STR2(x)
#define VP8P_UNROLL(n) _Pragma(VP8P_STR(unroll n))
#ifndef VP8_PAIR_MIN_CTAS
#define VP8_PAIR_MIN_CTAS(NW) ((NW) == 4 ? 7 : (NW) == 8 ? 3 : 1)
#endif

// CL = false: one CTA per image, unfiltered line buffer and progress stamps in shared memory.
// CL = true : one thread-block CLUSTER per image (launched with a cluster dimension, so its CTAs are co-scheduled and may
//             spin on each other): row pairs are dealt round-robin to the warps of all CTAs of the cluster; the unfiltered
//             line buffer and the progress stamps live next to the filtered line in the L2-resident scratch and are
//             published / acquired with gpu-scope fences. This is what lets ONE big frame use several SMs.
template <int NW, bool RECON, bool FILTER, bool CL>
__global__ void __launch_bounds__(NW * 32, CL ? 1 : VP8_PAIR_MIN_CTAS(NW))
vp8_mb_pairs(const Vp8ImgDesc* __restrict__ descs, int n_images, int line_px, uint8_t* __restrict__ tf_scratch) {
	extern __shared__ __align__(16) uint8_t smem[];
	volatile int* prog = reinterpret_cast<volatile int*>(smem);
	Vp8ImgDesc* sd = reinterpret_cast<Vp8ImgDesc*>(smem + 256);
	uint32_t* btab = reinterpret_cast<uint32_t*>(smem + 512);
	const int line_c = line_px / 2;
	const int c_rank = CL ? (int)cluster_ctarank() : 0, c_size = CL ? (int)cluster_nctarank() : 1;
	const int slot = CL ? blockIdx.x / c_size : blockIdx.x, n_slots = CL ? gridDim.x / c_size : gridDim.x;
	// per-slot scratch in global memory (L2): [tf: 8 x line][tu: 2 x line][progress stamps: one per macroblock row]
	uint8_t* const sc = tf_scratch + (size_t)slot * ((size_t)10 * line_px + kClusterProg * 4);
	uint8_t* tf_y = sc; // last four filtered rows of the row above
	uint8_t* tf_u = tf_y + 4 * line_px;
	uint8_t* tf_v = tf_u + 2 * line_px;
	// unfiltered bottom rows of the row above: shared memory, or the scratch when the image is spread over a cluster
	uint8_t* tu_y = CL ? sc + 8 * line_px : smem + kSmemFixed;
	uint8_t* tu_u = tu_y + line_px;
	uint8_t* tu_v = tu_u + line_px / 2;
	if (CL) prog = reinterpret_cast<volatile int*>(sc + (size_t)10 * line_px);
	// a cluster can have every row of the image in flight: one stamp per row, no ring
	constexpr int prog_mask = CL ? kClusterProg - 1 : kProgRing - 1;

	constexpr uint32_t FULL = 0xffffffffu;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int hl = lane & 15, half = lane >> 4, hbit = lane & 16;
	HalfWs& ws = reinterpret_cast<HalfWs*>(smem + kSmemFixed + (CL ? 0 : 2 * line_px))[warp * 2 + half];

	static_assert(sizeof(Vp8ImgDesc) <= 256, "descriptor must fit its shared-memory slot");
	static_assert(4 * NW <= kProgRing, "progress ring too small");

	// B_PRED lane table, see vp8_kernels.cu: btab[half][mode 0..10][pixel] = lane(a) | lane(b)<<8 | lane(c)<<16 | kind<<24
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

	// ---- lane roles inside a half (hl = 0..15)
	//   transforms / block prediction: hl < 8 luma blocks 2hl, 2hl+1 (same block row); 8,9: U; 10,11: V; 12: Y2
	const bool tl_luma = hl < 8, tl_chroma = hl >= 8 && hl < 12;
	const int tl_by = tl_luma ? (hl >> 1) * 4 : (hl & 1) * 4;   // block row offset inside the plane tile
	const int tl_bx0 = tl_luma ? (hl & 1) * 8 : 0;              // column offset of the lane's first block (second: +4)
	const int cstep = tl_luma ? 256 : (tl_chroma ? 64 : 16);    // int16 per macroblock in this lane's coefficient stream
	const int bit0 = tl_luma ? 2 * hl : (hl < 10 ? 16 + 2 * (hl - 8) : (hl < 12 ? 20 + 2 * (hl - 10) : 24)); // compact layout: mask bit of the lane's first block
	//   B_PRED: lane = pixel of the current sub-block
	const int px_r = hl >> 2, px_c = hl & 3;
	const int e_dy = hl <= 2 ? 3 : (hl <= 5 ? 5 - hl : -1);
	const int e_dx = hl <= 6 ? -1 : (hl == 15 ? 7 : hl - 7);
	uint8_t* const bp_edge = ws.rt_y + e_dy * 24 + e_dx;
	uint8_t* const bp_out = ws.rt_y + px_r * 24 + px_c;
	const int16_t* const bp_res = &ws.res[0][0] + hl;
	const uint8_t* const bp_tab = reinterpret_cast<const uint8_t*>(btab + half * 176 + hl);
	const bool dc_tap = (hl >= 2 && hl <= 5) || (hl >= 7 && hl <= 10);
	//   loop filter: luma line hl; chroma plane hl>>3, line hl&7

	for (int img = slot; img < n_images; img += n_slots) {
		// previous image fully retired before its line buffers, stamps and descriptor are reused
		if (CL) cluster_sync_all();
		else __syncthreads();
		{
			const uint32_t* src = reinterpret_cast<const uint32_t*>(descs + img);
			uint32_t* dst = reinterpret_cast<uint32_t*>(sd);
			for (int i = tid; i < (int)(sizeof(Vp8ImgDesc) / 4); i += NW * 32) dst[i] = src[i];
			if (CL) {
				if (c_rank == 0)
					for (int i = tid; i < kClusterProg; i += NW * 32) prog[i] = 0;
			} else if (tid < kProgRing) {
				prog[tid] = 0;
			}
		}
		if (CL) {
			__threadfence();
			cluster_sync_all();
		} else {
			__syncthreads();
		}

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
		const uint8_t* const g_hc = sd->has_coeff;
		// compact layout: packed blocks / per-macroblock mask / first-block index live behind the re-purposed coeff pointers;
		// they are re-read from the shared-memory descriptor at each use to keep registers free
#define G_COMPACT (sd->compact != 0)
#define G_PACKED reinterpret_cast<const uint8_t*>(sd->coeff_y)
#define G_MASK reinterpret_cast<const uint32_t*>(sd->coeff_u)
#define G_FIRST reinterpret_cast<const uint32_t*>(sd->coeff_v)

		for (int p = c_rank * NW + warp; 2 * p < rows; p += NW * c_size) {
			const int y = 2 * p + half;
			const bool row_ok = y < rows;
			const bool last_row = (y == rows - 1);
			const size_t mb_row0 = (size_t)y * cols;

			// ---- row start: out-of-frame left neighbours (129) and corner (127 on the top row, else 129)
			const int16_t* cptr = sd->coeff_y2;
			int x_pref = 0; // next macroblock column whose coefficients have not been requested yet
			uint32_t staged_nz = 3; // compact layout: which of the lane's two staged blocks are present (dense layout: both)
			if (RECON) {
				ws.lcol[hl] = 129;
				ws.lcol[16 + hl] = 129;
				ws.rt_y[(hl + 1) * 24 + 3] = 129;
				if (hl < 8) ws.rt_u[(hl + 1) * 12 + 3] = 129;
				else ws.rt_v[(hl - 8 + 1) * 12 + 3] = 129;
				const uint8_t corner = (y == 0) ? 127 : 129;
				if (hl == 0) ws.rt_y[3] = corner;
				if (hl == 1) ws.rt_u[3] = corner;
				if (hl == 2) ws.rt_v[3] = corner;
				if (row_ok && !G_COMPACT) {
					if (tl_luma) cptr = sd->coeff_y + (mb_row0 * 16 + 2 * hl) * 16;
					else if (hl < 10) cptr = sd->coeff_u + (mb_row0 * 4 + 2 * (hl - 8)) * 16;
					else if (hl < 12) cptr = sd->coeff_v + (mb_row0 * 4 + 2 * (hl - 10)) * 16;
					else cptr = sd->coeff_y2 + mb_row0 * 16;
				}
				if (half == 0 && row_ok) { // row y starts at step 0: its first macroblock's blocks go on their way now
					if (G_COMPACT) {
						staged_nz = fetch_compact(ws, G_PACKED, G_MASK[mb_row0], G_FIRST[mb_row0], bit0, hl);
					} else if (hl < 13) {
						cp_async16(&ws.coef[0 * 13 + hl], cptr);
						cp_async16(&ws.coef[1 * 13 + hl], cptr + 8);
						if (hl < 12) {
							cp_async16(&ws.coef[2 * 13 + hl], cptr + 16);
							cp_async16(&ws.coef[3 * 13 + hl], cptr + 24);
						}
					}
					cptr += cstep;
					x_pref = 1;
				}
				cp_async_commit();
			}

			for (int t = 0; t < cols + 2; t++) {
				const int x = t - 2 * half;
				const bool v = row_ok && x >= 0 && x < cols; // this half has a macroblock in this step
				const bool last_col = (x == cols - 1);
				const size_t mb = mb_row0 + (size_t)(v ? x : 0);

				// ---- per-macroblock syntax
				int ymode = 0, seg = 0, uvmode = 0, bmode = 0;
				bool bpred = false, inner = false;
				if (v) {
					ymode = g_ymode[mb];
					bpred = (ymode == 4);
					if (g_seg) seg = g_seg[mb] & 3;
					inner = bpred || (g_hc && g_hc[mb]);
					if (RECON) {
						uvmode = sd->uv_mode[mb];
						if (bpred) bmode = min((int)sd->bmode[mb * 16 + hl], 10);
					}
				}

				// ---- row y (half 0) waits for MB(x+1, y-1), which another warp produces; row y+1 trails row y by design
				if (p > 0 && t < cols) {
					if (lane == 0) {
						const int target = 2 * p * kStampRow + min(t + 2, cols);
						while (prog[(2 * p - 1) & prog_mask] < target) __nanosleep(64);
						if (CL) __threadfence();
						else __threadfence_block();
					}
				}
				__syncwarp();

				// ---- filtered rows of the row above: requested now, needed only when the filter tile is assembled
				uint32_t ta_y = 0, ta_c = 0;
				if (FILTER && v && y > 0) {
					ta_y = ldcg32(tf_y + (hl >> 2) * line_px + 16 * x + 4 * (hl & 3));
					ta_c = ldcg32((hl < 8 ? tf_u : tf_v) + ((hl & 7) >> 1) * line_c + 8 * x + 4 * (hl & 1));
				}

				if (RECON) {
					// ================================================================== m06: reconstruction
					// ---- top border (row -1) from the unfiltered line buffer, 127 above the frame
					if (v && hl < 9) {
						uint32_t w = 0x7f7f7f7fu;
						if (y > 0) {
							if (hl < 5) {
								if (hl == 4 && last_col) w = 0x01010101u * (ld_line<CL>(tu_y + 16 * x + 12) >> 24);
								else w = ld_line<CL>(tu_y + 16 * x + 4 * hl);
							} else if (hl < 7) {
								w = ld_line<CL>(tu_u + 8 * x + 4 * (hl - 5));
							} else {
								w = ld_line<CL>(tu_v + 8 * x + 4 * (hl - 7));
							}
						}
						if (hl < 5) {
							st32(ws.rt_y + 4 + 4 * hl, w);
							if (hl == 4) { // above-right of sub-block column 3 always comes from the MB row above
								st32(ws.rt_y + 4 * 24 + 20, w);
								st32(ws.rt_y + 8 * 24 + 20, w);
								st32(ws.rt_y + 12 * 24 + 20, w);
							}
						} else if (hl < 7) {
							st32(ws.rt_u + 4 + 4 * (hl - 5), w);
						} else {
							st32(ws.rt_v + 4 + 4 * (hl - 7), w);
						}
					}

					// ---- coefficients have landed
					cp_async_wait_all();
					__syncwarp();

					const int16_t* dq = sd->dq[seg];
					// ---- Y2: lane 12 runs the WHT, luma lanes take their DC from it (vp8_recon.c:563-586)
					if (v && !bpred && hl == 12) {
						const uint4 c0 = ws.coef[0 * 13 + 12], c1 = ws.coef[1 * 13 + 12];
						uint4* z = reinterpret_cast<uint4*>(ws.res[0]);
						if (!(staged_nz & 1) || (c0.x | c0.y | c0.z | c0.w | c1.x | c1.y | c1.z | c1.w) == 0) {
							z[0] = make_uint4(0, 0, 0, 0);
							z[1] = make_uint4(0, 0, 0, 0);
						} else {
							const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
							const int dcq = dq[4], acq = dq[5];
							int vv[16], d[16];
#pragma unroll
							for (int i = 0; i < 8; i++) {
								vv[2 * i] = s16(s16(cw[i]) * (i == 0 ? dcq : acq));
								vv[2 * i + 1] = s16(((int)cw[i] >> 16) * acq);
							}
							iwht4x4(vv, d);
							uint32_t pk[8];
#pragma unroll
							for (int i = 0; i < 8; i++) pk[i] = (uint32_t)(d[2 * i] & 0xffff) | ((uint32_t)d[2 * i + 1] << 16);
							z[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
							z[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
						}
					}
					__syncwarp();

					// ---- two 4x4 blocks per lane: dequantise, inverse DCT, then predict + add (i16 luma, chroma) or park the
					//      residual for the sub-block pass (B_PRED luma)
					if (v && hl < 12) {
						const int dcq = tl_luma ? dq[0] : dq[2], acq = tl_luma ? dq[1] : dq[3];
						uint8_t* tile;
						const uint8_t* lc;
						int stride, nwords, mode;
						if (tl_luma) { tile = ws.rt_y; lc = ws.lcol; stride = 24; nwords = 4; mode = ymode; }
						else if (hl < 10) { tile = ws.rt_u; lc = ws.lcol + 16; stride = 12; nwords = 2; mode = uvmode; }
						else { tile = ws.rt_v; lc = ws.lcol + 24; stride = 12; nwords = 2; mode = uvmode; }
						const bool park = tl_luma && bpred;

						// prediction that does not depend on the block column: left word, DC value
						uint32_t lw = 0, dcw = 0;
						int tm_p = 0;
						if (!park) {
							lw = ld32(lc + tl_by);
							tm_p = tile[3];
							if (mode != 1 && mode != 2 && mode != 3) { // DC (also any out-of-range mode): vp8_recon.c:152-176
								uint32_t sum = 0;
								const bool have_a = y > 0, have_l = x > 0;
								for (int k = 0; k < nwords; k++) {
									if (have_a) sum += sum4(ld32(tile + 4 + 4 * k));
									if (have_l) sum += sum4(ld32(lc + 4 * k));
								}
								const int lg = (nwords == 4) ? 4 : 3;
								uint32_t dcv;
								if (have_a && have_l) dcv = (sum + (1u << lg)) >> (lg + 1);
								else if (have_a || have_l) dcv = (sum + (1u << (lg - 1))) >> lg;
								else dcv = 128;
								dcw = dcv * 0x01010101u;
							}
						}

						VP8P_UNROLL(VP8P_BLOCK_UNROLL)
						for (int k = 0; k < 2; k++) {
							uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
							if (staged_nz & (1u << k)) {
								c0 = ws.coef[(2 * k) * 13 + hl];
								c1 = ws.coef[(2 * k + 1) * 13 + hl];
							}
							const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
							const int bi = 2 * hl + k; // luma block index (raster) when tl_luma
							int r[16];
							bool any;
							{
								int vv[16];
								int dc_override = 0;
								const bool use_override = tl_luma && !bpred;
								if (use_override) dc_override = ws.res[0][bi];
								// quick outs on the raw words: nothing coded at all / only the DC slot
								const uint32_t ac_raw = (cw[0] & 0xffff0000u) | cw[1] | cw[2] | cw[3] | cw[4] | cw[5] | cw[6] | cw[7];
								if (ac_raw == 0) {
									const int v0 = use_override ? dc_override : s16(s16(cw[0]) * dcq);
									const int dc = s16((v0 + 4) >> 3);
									any = dc != 0;
#pragma unroll
									for (int i = 0; i < 16; i++) r[i] = dc;
								} else {
#pragma unroll
									for (int i = 0; i < 8; i++) {
										vv[2 * i] = s16(s16(cw[i]) * (i == 0 ? dcq : acq));
										vv[2 * i + 1] = s16(((int)cw[i] >> 16) * acq);
									}
									if (use_override) vv[0] = dc_override;
									idct4x4(vv, r);
									any = true;
								}
							}
							if (park) {
								uint32_t pk[8];
#pragma unroll
								for (int i = 0; i < 8; i++) pk[i] = (uint32_t)(r[2 * i] & 0xffff) | ((uint32_t)r[2 * i + 1] << 16);
								uint4* dst = reinterpret_cast<uint4*>(ws.res[bi]);
								dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
								dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
							} else {
								const int bx = tl_bx0 + 4 * k;
								uint32_t pw[4];
								if (mode == 1) { // V
									const uint32_t a = ld32(tile + 4 + bx);
									pw[0] = pw[1] = pw[2] = pw[3] = a;
								} else if (mode == 2) { // H
#pragma unroll
									for (int j = 0; j < 4; j++) pw[j] = ((lw >> (8 * j)) & 255u) * 0x01010101u;
								} else if (mode == 3) { // TM
									const uint32_t a = ld32(tile + 4 + bx);
#pragma unroll
									for (int j = 0; j < 4; j++) {
										const int base = (int)((lw >> (8 * j)) & 255u) - tm_p;
										uint32_t w = 0;
#pragma unroll
										for (int i = 0; i < 4; i++) w |= (uint32_t)add_clip255(base, (int)((a >> (8 * i)) & 255u)) << (8 * i);
										pw[j] = w;
									}
								} else {
									pw[0] = pw[1] = pw[2] = pw[3] = dcw;
								}
								uint8_t* dst = tile + (tl_by + 1) * stride + 4 + bx;
#pragma unroll
								for (int j = 0; j < 4; j++) {
									uint32_t w = pw[j];
									if (any) {
										uint32_t o = 0;
#pragma unroll
										for (int i = 0; i < 4; i++) o |= (uint32_t)add_clip255((int)((w >> (8 * i)) & 255u), r[4 * j + i]) << (8 * i);
										w = o;
									}
									st32(dst + j * stride, w);
								}
							}
						}
					}
					__syncwarp();

					// ---- staged coefficients are consumed: request the next macroblock of this half's row
					if (row_ok && x + 1 == x_pref && x_pref < cols) {
						if (G_COMPACT) {
							staged_nz = fetch_compact(ws, G_PACKED, G_MASK[mb_row0 + x_pref], G_FIRST[mb_row0 + x_pref], bit0, hl);
						} else if (hl < 13) {
							cp_async16(&ws.coef[0 * 13 + hl], cptr);
							cp_async16(&ws.coef[1 * 13 + hl], cptr + 8);
							if (hl < 12) {
								cp_async16(&ws.coef[2 * 13 + hl], cptr + 16);
								cp_async16(&ws.coef[3 * 13 + hl], cptr + 24);
							}
						}
						cptr += cstep;
						x_pref++;
					}
					cp_async_commit();

					// ---- B_PRED luma: 16 sub-blocks in raster order, one lane per pixel (both halves at once)
					const bool bp = v && bpred;
					if (__any_sync(FULL, bp)) {
						const uint32_t sp = __ballot_sync(FULL, bp && (bmode <= 1 || bmode == 10));
						const uint32_t spm = (sp | (sp >> 16)) & 0xffffu; // sub-blocks needing B_DC / B_TM / 128 in either half
						VP8P_UNROLL(VP8P_BPRED_UNROLL)
						for (int s = 0; s < 16; s++) {
							const int tile_c = ((s >> 2) * 4 + 1) * 24 + 4 + (s & 3) * 4; // pixel (0,0) of sub-block s
							const int mode = __shfl_sync(FULL, bmode, hbit | s);
							uint32_t tw = 0;
							int e = 0, rs = 0;
							if (bp) {
								tw = *reinterpret_cast<const uint32_t*>(bp_tab + mode * 64);
								e = bp_edge[tile_c];
								rs = bp_res[s * 16];
							}
							const int a = __shfl_sync(FULL, e, tw);
							const int b = __shfl_sync(FULL, e, tw >> 8);
							const int c = __shfl_sync(FULL, e, tw >> 16);
							int pv = (a + c + 2 + 2 * b) >> 2;
							if (spm & (1u << s)) {
								const uint32_t kind = tw >> 24;
								if (kind == 1) pv = clip255(a + b - c);
								// B_DC: (A0..A3 + L0..L3 + 4) >> 3; one reduction serves both halves (low / high 16 bits)
								const int mine = dc_tap ? (e << hbit) : 0;
								const int both = __reduce_add_sync(FULL, mine);
								if (kind == 2) pv = (((both >> hbit) & 0xffff) + 4) >> 3;
								if (kind == 3) pv = 128;
							}
							if (bp) bp_out[tile_c] = (uint8_t)add_clip255(pv, rs);
							__syncwarp();
						}
					}
				} else {
					// ================================================================== stand-alone m07: load the MB
					if (v) {
						const uint8_t* sy = sd->src_y + (size_t)(16 * y) * sd->src_stride_y + 16 * x;
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int i = hl + 16 * j, row = i >> 2, wc = i & 3;
							st32(ws.rt_y + (row + 1) * 24 + 4 + 4 * wc, ld32(sy + (size_t)row * sd->src_stride_y + 4 * wc));
						}
#pragma unroll
						for (int j = 0; j < 2; j++) {
							const int row = hl >> 1, wc = hl & 1;
							const uint8_t* sp = (j ? sd->src_v : sd->src_u) + (size_t)(8 * y + row) * sd->src_stride_uv + 8 * x + 4 * wc;
							st32((j ? ws.rt_v : ws.rt_u) + (row + 1) * 12 + 4 + 4 * wc, ld32(sp));
						}
					}
					__syncwarp();
				}

				// ---- snapshot of the unfiltered tile that the neighbours will need (read phase)
				uint32_t edge_y = 0, edge_c = 0, bot_w = 0, corner_b = 0;
				if (RECON && v) {
					edge_y = ws.rt_y[(hl + 1) * 24 + 4 + 15];
					edge_c = (hl < 8) ? ws.rt_u[(hl + 1) * 12 + 4 + 7] : ws.rt_v[(hl - 8 + 1) * 12 + 4 + 7];
					if (hl < 4) bot_w = ld32(ws.rt_y + 16 * 24 + 4 + 4 * hl);
					else if (hl < 6) bot_w = ld32(ws.rt_u + 8 * 12 + 4 + 4 * (hl - 4));
					else if (hl < 8) bot_w = ld32(ws.rt_v + 8 * 12 + 4 + 4 * (hl - 6));
					if (hl == 8) corner_b = ws.rt_y[4 + 15];
					if (hl == 9) corner_b = ws.rt_u[4 + 7];
					if (hl == 10) corner_b = ws.rt_v[4 + 7];
				}

				if (!FILTER) {
					// ================================================================== unfiltered output (-yuv)
					if (v) {
						const bool fast = words_ok && !last_col && !last_row;
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int i = hl + 16 * j, row = i >> 2, wc = i & 3;
							const uint32_t w = ld32(ws.rt_y + (row + 1) * 24 + 4 + 4 * wc);
							if (fast) st32(oy.p + (size_t)(16 * y + row) * oy.stride + 16 * x + 4 * wc, w);
							else put_word(oy, 16 * x + 4 * wc, 16 * y + row, w);
						}
#pragma unroll
						for (int j = 0; j < 2; j++) {
							const int row = hl >> 1, wc = hl & 1;
							const uint32_t w = ld32((j ? ws.rt_v : ws.rt_u) + (row + 1) * 12 + 4 + 4 * wc);
							if (fast) st32((j ? ov.p : ou.p) + (size_t)(8 * y + row) * ou.stride + 8 * x + 4 * wc, w);
							else put_word(j ? ov : ou, 8 * x + 4 * wc, 8 * y + row, w);
						}
					}
				} else {
					// ================================================================== m07: loop filter
					// ---- assemble the filter tile: left apron = previous tile's right 4 columns (all 20 / 12 rows, so the
					//      corner above-left travels along), top apron = the prefetched tf words, interior = the reconstruction
					{
						uint32_t la0, la1, la2 = 0, in[4], ic[2];
						la0 = ld32(ws.ft_y + (hl + 4) * 20 + 16);                                  // luma rows 0..15
						la1 = (hl < 4) ? ld32(ws.ft_y + hl * 20 + 16) : ld32(ws.ft_u + (hl - 4) * 12 + 8); // luma rows -4..-1, U rows -4..7
						if (hl < 12) la2 = ld32(ws.ft_v + hl * 12 + 8);                            // V rows -4..7
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int i = hl + 16 * j;
							in[j] = ld32(ws.rt_y + ((i >> 2) + 1) * 24 + 4 + 4 * (i & 3));
						}
						ic[0] = ld32(ws.rt_u + ((hl >> 1) + 1) * 12 + 4 + 4 * (hl & 1));
						ic[1] = ld32(ws.rt_v + ((hl >> 1) + 1) * 12 + 4 + 4 * (hl & 1));
						__syncwarp();
						st32(ws.ft_y + (hl + 4) * 20, la0);
						if (hl < 4) st32(ws.ft_y + hl * 20, la1);
						else st32(ws.ft_u + (hl - 4) * 12, la1);
						if (hl < 12) st32(ws.ft_v + hl * 12, la2);
						st32(ws.ft_y + (hl >> 2) * 20 + 4 + 4 * (hl & 3), ta_y);
						st32((hl < 8 ? ws.ft_u : ws.ft_v) + ((hl & 7) >> 1) * 12 + 4 + 4 * (hl & 1), ta_c);
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int i = hl + 16 * j;
							st32(ws.ft_y + ((i >> 2) + 4) * 20 + 4 + 4 * (i & 3), in[j]);
						}
						st32(ws.ft_u + ((hl >> 1) + 4) * 12 + 4 + 4 * (hl & 1), ic[0]);
						st32(ws.ft_v + ((hl >> 1) + 4) * 12 + 4 + 4 * (hl & 1), ic[1]);
						__syncwarp();
					}

					// ---- edge filters in the reference's order per plane (vp8_loopfilter.c:226-277); both halves in lockstep,
					//      16 luma lines, then 8 U + 8 V lines
					const uint8_t* lfp = sd->lf[seg][bpred ? 1 : 0];
					const int level = lfp[0], interior = lfp[1], hev_thr = lfp[2];
					const bool do_f = v && level > 0;
					const int lim_mb = 2 * (level + 2) + interior, lim_in = 2 * level + interior;
					const bool f_left = do_f && x > 0, f_top = do_f && y > 0, f_in = do_f && inner;
					uint8_t* const ty = ws.ft_y + 4 * 20 + 4;                                   // luma pixel (0,0)
					uint8_t* const tc = (hl < 8 ? ws.ft_u : ws.ft_v) + 4 * 12 + 4;              // this lane's chroma plane (0,0)
					const int cn = hl & 7;
#if VP8P_LF_COMPACT
					if (__any_sync(FULL, do_f)) {
						// 12 passes: (plane, direction, edge) packed 4 bits each: bit 0 chroma, bits 1-2 edge/4, bit 3 between rows.
						// Per plane the reference's order holds: left MB edge, inner columns, top MB edge, inner rows.
						const unsigned long long kPasses = 0xECBA98643210ull;
#pragma unroll 1
						for (int p = 0; p < 12; p++) {
							const int d = (int)(kPasses >> (4 * p)) & 15;
							const bool chroma = d & 1, rows_dir = d & 8;
							const int e = (d & 6) * 2;
							uint8_t* const base = chroma ? tc : ty;
							const int stride = chroma ? 12 : 20, n = chroma ? cn : hl;
							const bool act = (e == 0 ? (rows_dir ? f_top : f_left) : f_in) && !(lf_simple && chroma);
							if (__any_sync(FULL, act)) {
								if (act)
									lf_line(rows_dir ? base + e * stride + n : base + n * stride + e, rows_dir ? stride : 1,
									        lf_simple ? EDGE_SIMPLE : (e == 0 ? EDGE_MB : EDGE_INNER), e == 0 ? lim_mb : lim_in, interior, hev_thr);
								__syncwarp();
							}
						}
					}

#else
					if (__any_sync(FULL, do_f)) {
						if (!lf_simple) {
							if (__any_sync(FULL, f_left)) {
								if (f_left) lf_across_columns<EDGE_MB>(ty + hl * 20, lim_mb, interior, hev_thr);
								if (f_left) lf_across_columns<EDGE_MB>(tc + cn * 12, lim_mb, interior, hev_thr);
								__syncwarp();
							}
							const bool any_in = __any_sync(FULL, f_in);
#if VP8P_LF_LOOP
							if (any_in) {
								if (f_in) lf_across_columns<EDGE_INNER>(tc + cn * 12 + 4, lim_in, interior, hev_thr);
#pragma unroll 1
								for (int e = 4; e < 16; e += 4) {
									if (f_in) lf_across_columns<EDGE_INNER>(ty + hl * 20 + e, lim_in, interior, hev_thr);
									__syncwarp();
								}
							}
							if (__any_sync(FULL, f_top)) {
								if (f_top) lf_across_rows<EDGE_MB>(ty + hl, 20, lim_mb, interior, hev_thr);
								if (f_top) lf_across_rows<EDGE_MB>(tc + cn, 12, lim_mb, interior, hev_thr);
								__syncwarp();
							}
							if (any_in) {
								if (f_in) lf_across_rows<EDGE_INNER>(tc + 4 * 12 + cn, 12, lim_in, interior, hev_thr);
#pragma unroll 1
								for (int e = 4; e < 16; e += 4) {
									if (f_in) lf_across_rows<EDGE_INNER>(ty + e * 20 + hl, 20, lim_in, interior, hev_thr);
									__syncwarp();
								}
							}
#else
							if (any_in) {
								if (f_in) lf_across_columns<EDGE_INNER>(ty + hl * 20 + 4, lim_in, interior, hev_thr);
								if (f_in) lf_across_columns<EDGE_INNER>(tc + cn * 12 + 4, lim_in, interior, hev_thr);
								__syncwarp();
								if (f_in) lf_across_columns<EDGE_INNER>(ty + hl * 20 + 8, lim_in, interior, hev_thr);
								__syncwarp();
								if (f_in) lf_across_columns<EDGE_INNER>(ty + hl * 20 + 12, lim_in, interior, hev_thr);
								__syncwarp();
							}
							if (__any_sync(FULL, f_top)) {
								if (f_top) lf_across_rows<EDGE_MB>(ty + hl, 20, lim_mb, interior, hev_thr);
								if (f_top) lf_across_rows<EDGE_MB>(tc + cn, 12, lim_mb, interior, hev_thr);
								__syncwarp();
							}
							if (any_in) {
								if (f_in) lf_across_rows<EDGE_INNER>(ty + 4 * 20 + hl, 20, lim_in, interior, hev_thr);
								if (f_in) lf_across_rows<EDGE_INNER>(tc + 4 * 12 + cn, 12, lim_in, interior, hev_thr);
								__syncwarp();
								if (f_in) lf_across_rows<EDGE_INNER>(ty + 8 * 20 + hl, 20, lim_in, interior, hev_thr);
								__syncwarp();
								if (f_in) lf_across_rows<EDGE_INNER>(ty + 12 * 20 + hl, 20, lim_in, interior, hev_thr);
								__syncwarp();
							}
#endif
						} else {
							// simple filter: luma only (vp8_loopfilter.c:228-244)
							if (f_left) lf_across_columns<EDGE_SIMPLE>(ty + hl * 20, lim_mb, 0, 0);
							__syncwarp();
#pragma unroll 1
							for (int e = 4; e < 16; e += 4) {
								if (f_in) lf_across_columns<EDGE_SIMPLE>(ty + hl * 20 + e, lim_in, 0, 0);
								__syncwarp();
							}
							if (f_top) lf_across_rows<EDGE_SIMPLE>(ty + hl, 20, lim_mb, 0, 0);
							__syncwarp();
#pragma unroll 1
							for (int e = 4; e < 16; e += 4) {
								if (f_in) lf_across_rows<EDGE_SIMPLE>(ty + e * 20 + hl, 20, lim_in, 0, 0);
								__syncwarp();
							}
						}
					}

#endif
					// ---- store what can no longer change: the 16x16 (8x8) block whose origin is 4 pixels up and left of the
					//      macroblock; the last column / row of macroblocks also flush the strips nobody else will
					if (v) {
						const bool fast = words_ok && x > 0 && y > 0 && !last_col && !last_row;
#pragma unroll
						for (int j = 0; j < 4; j++) {
							const int i = hl + 16 * j, row = i >> 2, wc = i & 3; // tile rows -4..11, word columns -1..2
							const uint32_t w = ld32(ws.ft_y + row * 20 + 4 * wc);
							if (fast) st32(oy.p + (size_t)(16 * y - 4 + row) * oy.stride + (16 * x - 4 + 4 * wc), w);
							else put_word(oy, 16 * x - 4 + 4 * wc, 16 * y - 4 + row, w);
						}
#pragma unroll
						for (int j = 0; j < 2; j++) {
							const int row = hl >> 1, wc = hl & 1;
							const uint32_t w = ld32((j ? ws.ft_v : ws.ft_u) + row * 12 + 4 * wc);
							if (fast) st32((j ? ov.p : ou.p) + (size_t)(8 * y - 4 + row) * ou.stride + (8 * x - 4 + 4 * wc), w);
							else put_word(j ? ov : ou, 8 * x - 4 + 4 * wc, 8 * y - 4 + row, w);
						}
						if (last_col) { // right strip: columns 12..15 (4..7), rows -4..11 (-4..3)
							put_word(oy, 16 * x + 12, 16 * y - 4 + hl, ld32(ws.ft_y + hl * 20 + 16));
							put_word(hl < 8 ? ou : ov, 8 * x + 4, 8 * y - 4 + cn, ld32((hl < 8 ? ws.ft_u : ws.ft_v) + cn * 12 + 8));
						}
						if (last_row) { // bottom strip: rows 12..15 (4..7), columns -4..15 (-4..7)
							{
								const int rr = hl >> 2, ww = hl & 3;
								put_word(oy, 16 * x - 4 + 4 * ww, 16 * y + 12 + rr, ld32(ws.ft_y + (16 + rr) * 20 + 4 * ww));
								if (last_col && ww == 0) put_word(oy, 16 * x + 12, 16 * y + 12 + rr, ld32(ws.ft_y + (16 + rr) * 20 + 16));
							}
							{
								const int rr = (hl & 7) >> 1, ww = hl & 1;
								uint8_t* fc = hl < 8 ? ws.ft_u : ws.ft_v;
								put_word(hl < 8 ? ou : ov, 8 * x - 4 + 4 * ww, 8 * y + 4 + rr, ld32(fc + (8 + rr) * 12 + 4 * ww));
								if (last_col && ww == 0) put_word(hl < 8 ? ou : ov, 8 * x + 4, 8 * y + 4 + rr, ld32(fc + (8 + rr) * 12 + 8));
							}
						}
						if (!last_row) { // hand the bottom 4 filtered rows to the row below: columns -4..11 (+12..15 on the last column)
							if (x > 0 || (hl & 3)) stcg32(tf_y + (hl >> 2) * line_px + 16 * x - 4 + 4 * (hl & 3), ld32(ws.ft_y + ((hl >> 2) + 16) * 20 + 4 * (hl & 3)));
							if (x > 0 || (hl & 1))
								stcg32((hl < 8 ? tf_u : tf_v) + (cn >> 1) * line_c + 8 * x - 4 + 4 * (hl & 1),
								       ld32((hl < 8 ? ws.ft_u : ws.ft_v) + ((cn >> 1) + 8) * 12 + 4 * (hl & 1)));
							if (last_col) {
								if (hl < 4) stcg32(tf_y + hl * line_px + 16 * x + 12, ld32(ws.ft_y + (hl + 16) * 20 + 16));
								else if (hl < 12) {
									const int k = hl - 4;
									stcg32(((k >> 2) ? tf_v : tf_u) + (k & 3) * line_c + 8 * x + 4, ld32(((k >> 2) ? ws.ft_v : ws.ft_u) + ((k & 3) + 8) * 12 + 8));
								}
							}
						}
					}
				}

				// ---- hand the unfiltered borders on (write phase) and publish progress
				__syncwarp();
				if (RECON && v) {
					ws.lcol[hl] = (uint8_t)edge_y;
					ws.lcol[16 + hl] = (uint8_t)edge_c;
					ws.rt_y[(hl + 1) * 24 + 3] = (uint8_t)edge_y;
					if (hl < 8) ws.rt_u[(hl + 1) * 12 + 3] = (uint8_t)edge_c;
					else ws.rt_v[(hl - 8 + 1) * 12 + 3] = (uint8_t)edge_c;
					if (hl == 8) ws.rt_y[3] = (uint8_t)corner_b;
					if (hl == 9) ws.rt_u[3] = (uint8_t)corner_b;
					if (hl == 10) ws.rt_v[3] = (uint8_t)corner_b;
					if (!last_row) {
						if (hl < 4) st_line<CL>(tu_y + 16 * x + 4 * hl, bot_w);
						else if (hl < 6) st_line<CL>(tu_u + 8 * x + 4 * (hl - 4), bot_w);
						else if (hl < 8) st_line<CL>(tu_v + 8 * x + 4 * (hl - 6), bot_w);
					}
				}
				__syncwarp();
				if (hl == 0 && v) {
					if (CL) __threadfence();
					else __threadfence_block();
					prog[y & prog_mask] = (y + 1) * kStampRow + x + 1;
				}
			}
		}
	}
}

template <int NW, bool RECON, bool FILTER>
int launch_pairs_t(const Vp8ImgDesc* descs, int n, int line_px, int grid, size_t smem, uint8_t* scratch, int cluster, cudaStream_t st) {
	if (cluster <= 1) {
		auto k = vp8_mb_pairs<NW, RECON, FILTER, false>;
		cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return (int)e;
		k<<<grid, NW * 32, smem, st>>>(descs, n, line_px, scratch);
		return (int)cudaGetLastError();
	}
	if (NW != 16) return (int)cudaErrorInvalidValue; // the cluster flavour is only built for 16 warps per CTA
	auto k = vp8_mb_pairs<16, RECON, FILTER, true>;
	cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return (int)e;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(16 * 32);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = st;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)cluster;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	e = cudaLaunchKernelEx(&cfg, k, descs, n, line_px, scratch);
	if (e != cudaSuccess) return (int)e;
	return (int)cudaGetLastError();
}

template <int NW, bool RECON, bool FILTER>
int occupancy_pairs_t(size_t smem) {
	auto k = vp8_mb_pairs<NW, RECON, FILTER, false>;
	if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
	int nb = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, NW * 32, smem) != cudaSuccess) return 0;
	return nb;
}

} // namespace

int vp8_pairs_smem_bytes(int warps_per_image, int max_mb_cols) {
	return kSmemFixed + 2 * 16 * max_mb_cols + warps_per_image * 2 * (int)sizeof(HalfWs);
}

// one scratch slot per CTA (or per cluster): filtered line (8 B/column), unfiltered line (2 B/column, cluster mode), stamps
size_t vp8_pairs_scratch_bytes(int slots, int max_mb_cols) { return (size_t)slots * ((size_t)10 * 16 * max_mb_cols + kClusterProg * 4); }

#define VP8_PAIRS_DISPATCH(FN, ...)                                \
	switch (mode * 100 + warps_per_image) {                        \
		case 4: return FN<4, true, false>(__VA_ARGS__);             \
		case 8: return FN<8, true, false>(__VA_ARGS__);             \
		case 16: return FN<16, true, false>(__VA_ARGS__);           \
		case 104: return FN<4, true, true>(__VA_ARGS__);            \
		case 108: return FN<8, true, true>(__VA_ARGS__);            \
		case 116: return FN<16, true, true>(__VA_ARGS__);           \
		case 204: return FN<4, false, true>(__VA_ARGS__);           \
		case 208: return FN<8, false, true>(__VA_ARGS__);           \
		case 216: return FN<16, false, true>(__VA_ARGS__);          \
		default: return -1;                                         \
	}

int vp8_launch_pairs(int mode, int warps_per_image, const Vp8ImgDesc* descs_dev, int n_images, int max_mb_cols, int grid_ctas,
                     uint8_t* scratch, int cluster, void* stream) {
	const size_t smem = (size_t)vp8_pairs_smem_bytes(warps_per_image, max_mb_cols);
	const int line_px = 16 * max_mb_cols;
	cudaStream_t st = (cudaStream_t)stream;
	VP8_PAIRS_DISPATCH(launch_pairs_t, descs_dev, n_images, line_px, grid_ctas, smem, scratch, cluster, st)
}

int vp8_pairs_max_ctas_per_sm(int mode, int warps_per_image, int max_mb_cols) {
	const size_t smem = (size_t)vp8_pairs_smem_bytes(warps_per_image, max_mb_cols);
	VP8_PAIRS_DISPATCH(occupancy_pairs_t, smem)
}
