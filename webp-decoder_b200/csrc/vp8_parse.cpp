// vp8_parse.cpp - re-entrant host front end: .webp bytes -> Vp8DecodedFrame. See include/vp8_parse.h.
//
// Entropy decoding is a serial dependency chain per frame (RFC 6386 section 7), so it stays on host threads,
// one image per thread; everything after it runs on the GPU. Output arrays are laid out exactly as the
// reference's vp8_decode_decoded_frame produces them (vp8_tokens.c:673-1001), array for array, so either
// front end can feed the kernels.
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/vp8_parse.h"
#include "vp8_tables.h"

namespace {

// ------------------------------------------------------------------------------------------------ boolean decoder
// RFC 6386 section 7 arithmetic decoder with a 64-bit window. Bytes past the end of the partition read as zero,
// which is also what the reference does (bool_decoder.c:5-15).
struct BoolReader {
	const uint8_t* p;
	const uint8_t* end;
	uint64_t window = 0; // next bits, MSB-aligned
	int avail = 0;       // valid bits in window
	uint32_t range = 255;

	BoolReader(const uint8_t* data, size_t n) : p(data), end(data + n) {}

	inline void refill() {
		while (avail <= 56) {
			const uint64_t byte = (p < end) ? *p++ : 0;
			window |= byte << (56 - avail);
			avail += 8;
		}
	}
	inline int bit(uint32_t prob) {
		if (avail < 16) refill();
		const uint32_t split = 1 + (((range - 1) * prob) >> 8);
		const uint64_t big = (uint64_t)split << 56;
		int b;
		if (window >= big) {
			range -= split;
			window -= big;
			b = 1;
		} else {
			range = split;
			b = 0;
		}
		const int shift = __builtin_clz(range) - 24; // renormalise to [128, 255]
		range <<= shift;
		window <<= shift;
		avail -= shift;
		return b;
	}
	inline uint32_t literal(int bits) {
		uint32_t v = 0;
		while (bits--) v = (v << 1) | (uint32_t)bit(128);
		return v;
	}
	// magnitude then sign; like the reference (bool_decoder.c:79-84) no sign bit follows a zero magnitude
	inline int sint(int bits) {
		const int mag = (int)literal(bits);
		if (mag == 0) return 0;
		return bit(128) ? -mag : mag;
	}
};

inline int clamp_i8(int v) { return v < -128 ? -128 : (v > 127 ? 127 : v); }

// ------------------------------------------------------------------------------------------------ syntax trees
// RFC 6386 sections 8.1, 11.2-11.5 and 13.2 written out as decision chains.
inline int read_segment_id(BoolReader& br, const uint8_t p[3]) {
	return br.bit(p[0]) ? 2 + br.bit(p[2]) : br.bit(p[1]);
}
inline int read_kf_ymode(BoolReader& br) { // 0..4 = DC,V,H,TM,B_PRED
	if (!br.bit(145)) return 4;
	if (!br.bit(156)) return br.bit(163); // DC / V
	return 2 + br.bit(128);               // H / TM
}
inline int read_uv_mode(BoolReader& br) {
	if (!br.bit(142)) return 0;
	if (!br.bit(114)) return 1;
	return 2 + br.bit(183);
}
enum { B_DC, B_TM, B_VE, B_HE, B_LD, B_RD, B_VR, B_VL, B_HD, B_HU };
inline int read_bmode(BoolReader& br, const uint8_t* p) {
	if (!br.bit(p[0])) return B_DC;
	if (!br.bit(p[1])) return B_TM;
	if (!br.bit(p[2])) return B_VE;
	if (!br.bit(p[3])) {
		if (!br.bit(p[4])) return B_HE;
		return br.bit(p[5]) ? B_VR : B_RD;
	}
	if (!br.bit(p[6])) return B_LD;
	if (!br.bit(p[7])) return B_VL;
	return br.bit(p[8]) ? B_HU : B_HD;
}

const uint8_t kBands[16] = {0, 1, 2, 3, 6, 4, 5, 6, 6, 6, 6, 6, 6, 6, 6, 7};
const uint8_t kZigzag[16] = {0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15};
const uint8_t kCat3[] = {173, 148, 140}, kCat4[] = {176, 155, 140, 135}, kCat5[] = {180, 157, 141, 134, 130},
              kCat6[] = {254, 254, 243, 230, 196, 177, 153, 140, 133, 130, 129};

inline int read_extra(BoolReader& br, const uint8_t* p, int n) {
	int v = 0;
	for (int i = 0; i < n; i++) v = (v << 1) | br.bit(p[i]);
	return v;
}

// One 4x4 block of tokens (RFC 6386 section 13). probs = [8 bands][3 contexts][11] of the block's type.
// Returns 1 when any decoded coefficient is non-zero - that, not the end-of-block position, is what the reference
// propagates as the neighbour context (vp8_tokens.c:331-351).
inline int read_block(BoolReader& br, const uint8_t* probs, int first, int ctx, int16_t* out) {
	int nonzero = 0;
	bool prev_zero = false;
	for (int i = first; i < 16; i++) {
		const uint8_t* p = probs + (kBands[i] * 3 + ctx) * 11;
		if (!prev_zero && !br.bit(p[0])) break; // end of block
		if (!br.bit(p[1])) {
			ctx = 0;
			prev_zero = true;
			continue;
		}
		int v;
		if (!br.bit(p[2])) {
			v = 1;
		} else if (!br.bit(p[3])) {
			v = !br.bit(p[4]) ? 2 : 3 + br.bit(p[5]);
		} else if (!br.bit(p[6])) {
			if (!br.bit(p[7])) v = 5 + br.bit(159);
			else {
				v = 7 + 2 * br.bit(165);
				v += br.bit(145);
			}
		} else {
			const int hi = br.bit(p[8]);
			const int lo = br.bit(p[9 + hi]);
			switch (2 * hi + lo) {
				case 0: v = 11 + read_extra(br, kCat3, 3); break;
				case 1: v = 19 + read_extra(br, kCat4, 4); break;
				case 2: v = 35 + read_extra(br, kCat5, 5); break;
				default: v = 67 + read_extra(br, kCat6, 11); break;
			}
		}
		ctx = (v == 1) ? 1 : 2;
		prev_zero = false;
		out[kZigzag[i]] = (int16_t)(br.bit(128) ? -v : v);
		nonzero = 1;
	}
	return nonzero;
}

// ------------------------------------------------------------------------------------------------ container / header
inline uint32_t le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

int fail(int e) {
	errno = e;
	return -1;
}

// RIFF/WEBP with exactly one 'VP8 ' chunk and consistent sizes (reference webp_container.c:19-89).
int find_vp8_chunk(const uint8_t* f, size_t n, const uint8_t** payload, size_t* psize) {
	if (!f || n < 20 || memcmp(f, "RIFF", 4) || memcmp(f + 8, "WEBP", 4)) return fail(EINVAL);
	if ((size_t)le32(f + 4) + 8 != n) return fail(EINVAL);
	if (memcmp(f + 12, "VP8 ", 4)) return fail(EINVAL);
	const size_t csize = le32(f + 16);
	if (csize > n - 20) return fail(EINVAL);
	size_t end = 20 + csize;
	end += end & 1;
	if (end != n) return fail(EINVAL);
	*payload = f + 20;
	*psize = csize;
	return 0;
}

// Frame tag + start code + dimensions (reference vp8_header.c:13-65).
int read_frame_header(const uint8_t* d, size_t n, Vp8KeyFrameHeader* kf) {
	memset(kf, 0, sizeof(*kf));
	if (!d || n < 10) return fail(EINVAL);
	const uint32_t tag = d[0] | (d[1] << 8) | (d[2] << 16);
	kf->is_key_frame = !(tag & 1);
	kf->profile = (tag >> 1) & 7;
	kf->show_frame = (tag >> 4) & 1;
	kf->first_partition_len = (tag >> 5) & 0x7FFFF;
	if (!kf->is_key_frame) return fail(EINVAL);
	kf->start_code_ok = d[3] == 0x9d && d[4] == 0x01 && d[5] == 0x2a;
	if (!kf->start_code_ok) return fail(EINVAL);
	const uint32_t w = d[6] | (d[7] << 8), h = d[8] | (d[9] << 8);
	kf->width = w & 0x3FFF;
	kf->x_scale = (w >> 14) & 3;
	kf->height = h & 0x3FFF;
	kf->y_scale = (h >> 14) & 3;
	if (!kf->width || !kf->height) return fail(EINVAL);
	if (kf->first_partition_len > n - 10) return fail(EINVAL);
	return 0;
}

constexpr size_t kArenaAlign = 64;
inline size_t up(size_t v) { return (v + kArenaAlign - 1) / kArenaAlign * kArenaAlign; }

size_t arena_need(size_t mb) {
	return 5 * up(mb) + up(mb * 16) + up(mb * 32) + up(mb * 512) + 2 * up(mb * 128);
}

struct Carver {
	uint8_t* base;
	size_t at = 0;
	template <class T>
	T* take(size_t bytes) {
		T* r = reinterpret_cast<T*>(base + at);
		at += up(bytes);
		return r;
	}
};

int parse_payload(const uint8_t* d, size_t n, Vp8KeyFrameHeader* kf, Vp8DecodedFrame* out, void* arena, size_t arena_bytes) {
	if (!kf || !out) return fail(EINVAL);
	memset(out, 0, sizeof(*out));
	if (read_frame_header(d, n, kf)) return -1;
	const uint32_t cols = (kf->width + 15u) / 16u, rows = (kf->height + 15u) / 16u;
	const size_t mb = (size_t)cols * rows;
	if (mb > (1u << 20)) return fail(EINVAL);
	out->mb_cols = cols;
	out->mb_rows = rows;
	out->mb_total = (uint32_t)mb;

	// ---- array storage (zeroed: absent tokens are zero coefficients)
	const size_t need = arena_need(mb);
	bool own = false;
	if (!arena) {
		arena = aligned_alloc(kArenaAlign, need);
		if (!arena) return fail(ENOMEM);
		own = true;
	} else if (arena_bytes < need || ((uintptr_t)arena & (kArenaAlign - 1))) {
		return fail(EINVAL);
	}
	memset(arena, 0, need);
	Carver cv{(uint8_t*)arena};
	// coeff_y first: it is the address vp8_parse_free releases
	out->coeff_y = cv.take<int16_t>(mb * 512);
	out->coeff_u = cv.take<int16_t>(mb * 128);
	out->coeff_v = cv.take<int16_t>(mb * 128);
	out->coeff_y2 = cv.take<int16_t>(mb * 32);
	out->bmode = cv.take<uint8_t>(mb * 16);
	out->segment_id = cv.take<uint8_t>(mb);
	out->skip_coeff = cv.take<uint8_t>(mb);
	out->has_coeff = cv.take<uint8_t>(mb);
	out->ymode = cv.take<uint8_t>(mb);
	out->uv_mode = cv.take<uint8_t>(mb);
	out->stats_opaque[24] = own ? 0x6f776e6564ull : 0; // "owned" marker for vp8_parse_free
	// the arrays above are ONE block: lets vp8_gpu_upload move a frame with a single transfer (vp8_gpu.cu, kArenaMagic)
	out->stats_opaque[21] = 0x564138415245414eull;
	out->stats_opaque[22] = (uint64_t)(uintptr_t)arena;
	out->stats_opaque[23] = (uint64_t)need;
	auto bail = [&](int e) {
		if (own) free(arena);
		memset(out, 0, sizeof(*out));
		return fail(e);
	};

	// ---- first partition: frame header (RFC 6386 section 9.2-9.11, 19.2)
	BoolReader br(d + 10, kf->first_partition_len);
	br.bit(128); // color space
	br.bit(128); // clamping type
	uint8_t seg_probs[3] = {255, 255, 255};
	bool update_map = false;
	out->segmentation_enabled = (uint8_t)br.bit(128);
	if (out->segmentation_enabled) {
		update_map = br.bit(128);
		if (br.bit(128)) { // update segment feature data
			out->segmentation_abs = (uint8_t)br.bit(128);
			for (int i = 0; i < 4; i++)
				if (br.bit(128)) out->seg_quant_idx[i] = (int8_t)clamp_i8(br.sint(7));
			for (int i = 0; i < 4; i++)
				if (br.bit(128)) out->seg_lf_level[i] = (int8_t)clamp_i8(br.sint(6));
		}
		if (update_map)
			for (int i = 0; i < 3; i++)
				if (br.bit(128)) seg_probs[i] = (uint8_t)br.literal(8);
	}
	out->lf_use_simple = (uint8_t)br.bit(128);
	out->lf_level = (uint8_t)br.literal(6);
	out->lf_sharpness = (uint8_t)br.literal(3);
	out->lf_delta_enabled = (uint8_t)br.bit(128);
	if (out->lf_delta_enabled && br.bit(128)) {
		for (int i = 0; i < 4; i++)
			if (br.bit(128)) out->lf_ref_delta[i] = (int8_t)clamp_i8(br.sint(6));
		for (int i = 0; i < 4; i++)
			if (br.bit(128)) out->lf_mode_delta[i] = (int8_t)clamp_i8(br.sint(6));
	}
	const int partitions = 1 << br.literal(2);
	out->q_index = (uint8_t)br.literal(7);
	int8_t* dq[5] = {&out->y1_dc_delta_q, &out->y2_dc_delta_q, &out->y2_ac_delta_q, &out->uv_dc_delta_q, &out->uv_ac_delta_q};
	for (int i = 0; i < 5; i++) *dq[i] = br.bit(128) ? (int8_t)clamp_i8(br.sint(4)) : 0;
	br.bit(128); // refresh_entropy_probs

	// coefficient probabilities: defaults + updates, private to this call
	uint8_t probs[4 * 8 * 3 * 11];
	memcpy(probs, kDefaultCoeffProbs, sizeof(probs));
	for (int i = 0; i < 4 * 8 * 3 * 11; i++)
		if (br.bit(kCoeffUpdateProbs[i])) probs[i] = (uint8_t)br.literal(8);

	const bool has_skip = br.bit(128);
	const uint8_t skip_prob = has_skip ? (uint8_t)br.literal(8) : 0;

	// ---- first partition: per-macroblock modes (RFC 6386 sections 9.3, 11)
	{
		std::vector<uint8_t> above(cols * 4, B_DC);
		for (uint32_t my = 0; my < rows; my++) {
			uint8_t left[4] = {B_DC, B_DC, B_DC, B_DC};
			for (uint32_t mx = 0; mx < cols; mx++) {
				const size_t i = (size_t)my * cols + mx;
				if (update_map) out->segment_id[i] = (uint8_t)read_segment_id(br, seg_probs);
				if (has_skip) out->skip_coeff[i] = (uint8_t)br.bit(skip_prob);
				const int ym = read_kf_ymode(br);
				out->ymode[i] = (uint8_t)ym;
				uint8_t* bm = out->bmode + i * 16;
				uint8_t* ab = &above[mx * 4];
				if (ym == 4) {
					for (int r = 0; r < 4; r++)
						for (int c = 0; c < 4; c++) {
							const int A = r ? bm[4 * (r - 1) + c] : ab[c];
							const int L = c ? bm[4 * r + c - 1] : left[r];
							bm[4 * r + c] = (uint8_t)read_bmode(br, kKfBmodeProbs + (A * 10 + L) * 9);
						}
					for (int k = 0; k < 4; k++) {
						ab[k] = bm[12 + k];
						left[k] = bm[4 * k + 3];
					}
				} else {
					static const uint8_t implied[4] = {B_DC, B_VE, B_HE, B_TM};
					memset(bm, implied[ym], 16);
					memset(ab, implied[ym], 4);
					memset(left, implied[ym], 4);
				}
				out->uv_mode[i] = (uint8_t)read_uv_mode(br);
			}
		}
	}

	// ---- token partition (single; the reference rejects more, vp8_tokens.c:357-360)
	if (partitions != 1) return bail(ENOTSUP);
	const size_t tok_off = 10 + (size_t)kf->first_partition_len;
	BoolReader tr(d + tok_off, n - tok_off);
	{
		// non-zero contexts: per MB column 4 luma + 2 U + 2 V + 1 Y2 above flags; same set to the left
		std::vector<uint8_t> above(cols * 9, 0);
		for (uint32_t my = 0; my < rows; my++) {
			uint8_t left[9] = {0};
			for (uint32_t mx = 0; mx < cols; mx++) {
				const size_t i = (size_t)my * cols + mx;
				uint8_t* ab = &above[mx * 9];
				const bool has_y2 = out->ymode[i] != 4;
				if (out->skip_coeff[i]) {
					// no tokens: contexts clear, except Y2's which a B_PRED macroblock leaves alone
					const uint8_t a8 = ab[8], l8 = left[8];
					memset(ab, 0, 9);
					memset(left, 0, 9);
					if (!has_y2) {
						ab[8] = a8;
						left[8] = l8;
					}
					continue;
				}
				int any = 0;
				if (has_y2) {
					const int nz = read_block(tr, probs + 1 * 264, 0, ab[8] + left[8], out->coeff_y2 + i * 16);
					ab[8] = left[8] = (uint8_t)nz;
					any |= nz;
				}
				const uint8_t* py = probs + (has_y2 ? 0 : 3) * 264;
				for (int r = 0; r < 4; r++)
					for (int c = 0; c < 4; c++) {
						const int nz = read_block(tr, py, has_y2 ? 1 : 0, ab[c] + left[r], out->coeff_y + (i * 16 + 4 * r + c) * 16);
						ab[c] = left[r] = (uint8_t)nz;
						any |= nz;
					}
				for (int pl = 0; pl < 2; pl++) {
					int16_t* dst = (pl ? out->coeff_v : out->coeff_u) + i * 64;
					uint8_t* a = ab + 4 + 2 * pl;
					uint8_t* l = left + 4 + 2 * pl;
					for (int r = 0; r < 2; r++)
						for (int c = 0; c < 2; c++) {
							const int nz = read_block(tr, probs + 2 * 264, 0, a[c] + l[r], dst + (2 * r + c) * 16);
							a[c] = l[r] = (uint8_t)nz;
							any |= nz;
						}
				}
				out->has_coeff[i] = (uint8_t)any;
			}
		}
	}
	return 0;
}

} // namespace

extern "C" {

size_t vp8_parse_arena_bytes(uint32_t width, uint32_t height) {
	return arena_need((size_t)((width + 15) / 16) * ((height + 15) / 16));
}

int vp8_parse_webp_size(const uint8_t* file, size_t size, uint32_t* width, uint32_t* height) {
	const uint8_t* payload;
	size_t psize;
	Vp8KeyFrameHeader kf;
	if (find_vp8_chunk(file, size, &payload, &psize) || read_frame_header(payload, psize, &kf)) return -1;
	if (width) *width = kf.width;
	if (height) *height = kf.height;
	return 0;
}

int vp8_parse_vp8(const uint8_t* payload, size_t size, Vp8KeyFrameHeader* kf, Vp8DecodedFrame* out, void* arena,
                  size_t arena_bytes) {
	return parse_payload(payload, size, kf, out, arena, arena_bytes);
}

int vp8_parse_webp(const uint8_t* file, size_t size, Vp8KeyFrameHeader* kf, Vp8DecodedFrame* out, void* arena,
                   size_t arena_bytes) {
	const uint8_t* payload;
	size_t psize;
	if (find_vp8_chunk(file, size, &payload, &psize)) return -1;
	return parse_payload(payload, psize, kf, out, arena, arena_bytes);
}

void vp8_parse_free(Vp8DecodedFrame* f) {
	if (!f) return;
	if (f->stats_opaque[24] == 0x6f776e6564ull) free(f->coeff_y);
	memset(f, 0, sizeof(*f));
}

int vp8_parse_batch(const uint8_t* const* files, const size_t* sizes, int n, int threads, Vp8KeyFrameHeader* kf,
                    Vp8DecodedFrame* out, void* const* arenas, const size_t* arena_bytes, int* status) {
	if (!files || !sizes || !kf || !out || n < 0) {
		errno = EINVAL;
		return -1;
	}
	if (threads < 1) threads = 1;
	if (threads > n) threads = n;
	std::atomic<int> next{0}, failed{0};
	auto worker = [&]() {
		for (int i; (i = next.fetch_add(1)) < n;) {
			const int rc = vp8_parse_webp(files[i], sizes[i], &kf[i], &out[i], arenas ? arenas[i] : nullptr,
			                              arenas && arena_bytes ? arena_bytes[i] : 0);
			if (status) status[i] = rc ? errno : 0;
			if (rc) failed++;
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; t++) pool.emplace_back(worker);
	worker();
	for (auto& t : pool) t.join();
	return failed.load();
}

} // extern "C"
