// vp8_lf2.cuh - the normal loop filter on TWO edge positions at once, one per 16-bit half of a register.
//
// Every tap p3..q3 is a 32-bit word holding the pixel of position A in bits 0..15 and of position B in bits 16..31
// (values 0..255). The arithmetic of vp8_loopfilter.c:24-104 is restated with the packed-halfword instructions sm_100a
// has natively - VIADD.16x2, VIADDMNMX.S16x2(.RELU), VIMNMX(3).S16x2, VABSDIFF4.U8, plus PRMT for the sign masks - and
// with biases wherever the scalar code shifts a negative number (there is no packed arithmetic shift):
//     (x + 128) >> 3 == (x >> 3) + 16          (27 w' + 63) >> 7 == ((27 w + 63) >> 7) + 27   with w' = w + 128
// Branches of the scalar code (threshold not met, high edge variance or not) become byte masks and selects.
//
// The header also compiles for the host (each instruction emulated per half), so that tests/test_host.py can compare it
// with the oracle over every threshold and millions of tap combinations without a GPU. Test infrastructure may include
// it; nothing on the product path runs the host flavour.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VP8_LF2_FN __device__ __forceinline__
#else
#define VP8_LF2_FN static inline
#endif

namespace lf2 {

// ---- the instruction set, per 16-bit half
#if defined(__CUDA_ARCH__)
VP8_LF2_FN uint32_t add2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
VP8_LF2_FN uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
VP8_LF2_FN uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
VP8_LF2_FN uint32_t addmin(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }           // min(a+b, c)
VP8_LF2_FN uint32_t addmin_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); } // max(min(a+b, c), 0)
VP8_LF2_FN uint32_t absdiff_u8(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
VP8_LF2_FN uint32_t sign_mask(uint32_t a) { // 0xffff per half whose sign bit is set: PRMT with sign replication of bytes 1 and 3
	uint32_t r;
	asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(r) : "r"(a));
	return r;
}
#else
VP8_LF2_FN uint32_t lf2_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
VP8_LF2_FN int lf2_lo(uint32_t a) { return (int16_t)(a & 0xffffu); }
VP8_LF2_FN int lf2_hi(uint32_t a) { return (int16_t)(a >> 16); }
VP8_LF2_FN int lf2_min(int a, int b) { return a < b ? a : b; }
VP8_LF2_FN int lf2_max(int a, int b) { return a > b ? a : b; }
VP8_LF2_FN uint32_t add2(uint32_t a, uint32_t b) { return lf2_pack(lf2_lo(a) + lf2_lo(b), lf2_hi(a) + lf2_hi(b)); }
VP8_LF2_FN uint32_t max2(uint32_t a, uint32_t b) { return lf2_pack(lf2_max(lf2_lo(a), lf2_lo(b)), lf2_max(lf2_hi(a), lf2_hi(b))); }
VP8_LF2_FN uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return max2(max2(a, b), c); }
VP8_LF2_FN uint32_t addmin(uint32_t a, uint32_t b, uint32_t c) {
	return lf2_pack(lf2_min((int16_t)(lf2_lo(a) + lf2_lo(b)), lf2_lo(c)), lf2_min((int16_t)(lf2_hi(a) + lf2_hi(b)), lf2_hi(c)));
}
VP8_LF2_FN uint32_t addmin_relu(uint32_t a, uint32_t b, uint32_t c) {
	const uint32_t m = addmin(a, b, c);
	return lf2_pack(lf2_max(lf2_lo(m), 0), lf2_max(lf2_hi(m), 0));
}
VP8_LF2_FN uint32_t absdiff_u8(uint32_t a, uint32_t b) {
	uint32_t r = 0;
	for (int i = 0; i < 4; i++) {
		const int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255;
		r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
	}
	return r;
}
VP8_LF2_FN uint32_t sign_mask(uint32_t a) { return ((a & 0x8000u) ? 0xffffu : 0u) | ((a & 0x80000000u) ? 0xffff0000u : 0u); }
#endif

// Is there anything to filter at all? On the device the question is put to the whole warp (uniform branch: the threshold
// test fails for every position of a noisy macroblock, and then the filter arithmetic is skipped as in the scalar code).
#if defined(__CUDA_ARCH__)
VP8_LF2_FN bool anyone(uint32_t mask) { return __any_sync(0xffffffffu, mask != 0); }
#else
VP8_LF2_FN bool anyone(uint32_t mask) { return mask != 0; }
#endif

VP8_LF2_FN uint32_t both(int v) { return ((uint32_t)v & 0xffffu) * 0x00010001u; } // the same constant in both halves
VP8_LF2_FN uint32_t select(uint32_t mask, uint32_t yes, uint32_t no) { return (yes & mask) | (no & ~mask); }

// Per-position limits, already in the form the packed compare wants: c = -(limit + 1), so that x + c < 0 <=> x <= limit.
struct Limits {
	uint32_t edge;     // -(E + 1): 2|p0-q0| + |p1-q1|/2 <= E   (vp8_loopfilter.c:24-32)
	uint32_t interior; // -(I + 1): six neighbour differences <= I (:34-49)
	uint32_t hev;      // -(T + 1): max(|p1-p0|, |q1-q0|) <= T means NO high edge variance (:51-56)
};
VP8_LF2_FN Limits make_limits(int e_lo, int e_hi, int i_lo, int i_hi, int t_lo, int t_hi) {
	Limits l;
	l.edge = ((uint32_t)(-(e_lo + 1)) & 0xffffu) | ((uint32_t)(-(e_hi + 1)) << 16);
	l.interior = ((uint32_t)(-(i_lo + 1)) & 0xffffu) | ((uint32_t)(-(i_hi + 1)) << 16);
	l.hev = ((uint32_t)(-(t_lo + 1)) & 0xffffu) | ((uint32_t)(-(t_hi + 1)) << 16);
	return l;
}

// Masks: `ok` = normal_threshold holds, `calm` = no high edge variance. 0xffff / 0 per half.
VP8_LF2_FN void thresholds(uint32_t p3, uint32_t p2, uint32_t p1, uint32_t p0, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3,
                           const Limits& lim, uint32_t& ok, uint32_t& calm) {
	const uint32_t d0 = absdiff_u8(p0, q0), d1 = absdiff_u8(p1, q1);
	const uint32_t e = add2(add2(d0, d0), (d1 >> 1) & 0x007f007fu);
	const uint32_t a1 = absdiff_u8(p1, p0), b1 = absdiff_u8(q1, q0);
	uint32_t m = max3(absdiff_u8(p3, p2), absdiff_u8(p2, p1), a1);
	m = max3(m, absdiff_u8(q3, q2), absdiff_u8(q2, q1));
	m = max2(m, b1);
	ok = sign_mask(add2(e, lim.edge) & add2(m, lim.interior));
	calm = sign_mask(add2(max2(a1, b1), lim.hev));
}

// a = clamp8(3 (q0 - p0) + (clamp8(p1 - q1) & use_p1q1))   (vp8_loopfilter.c:58-63, 81-86)
VP8_LF2_FN uint32_t filter_value(uint32_t p1, uint32_t p0, uint32_t q0, uint32_t q1, uint32_t use_p1q1) {
	uint32_t t = addmin(p1, ~q1, both(126)); // min(p1 - q1 - 1, 126)
	t = add2(max2(t, both(-129)), both(1));  // clamp8(p1 - q1)
	t &= use_p1q1;
	const uint32_t d = add2(q0, ~p0); // q0 - p0 - 1
	const uint32_t a = addmin(add2(add2(add2(d, d), d), t), both(3), both(127));
	return max2(a, both(-128));
}

// filter_common (vp8_loopfilter.c:58-79) on p0/q0, and on p1/q1 where `inner_too` is set. a = filter_value().
VP8_LF2_FN void apply_common(uint32_t a, uint32_t inner_too, uint32_t& p1, uint32_t& p0, uint32_t& q0, uint32_t& q1) {
	const uint32_t u1 = (addmin(a, both(4 + 128), both(255)) >> 3) & 0x001f001fu; // f1 + 16
	const uint32_t u2 = (addmin(a, both(3 + 128), both(255)) >> 3) & 0x001f001fu; // f2 + 16
	const uint32_t h = (add2(u1, both(1)) >> 1) & 0x001f001fu;                    // ((f1 + 1) >> 1) + 8
	q0 = addmin_relu(add2(q0, ~u1), both(17), both(255));
	p0 = addmin_relu(add2(p0, u2), both(-16), both(255));
	const uint32_t q1n = addmin_relu(add2(q1, ~h), both(9), both(255));
	const uint32_t p1n = addmin_relu(add2(p1, h), both(-8), both(255));
	q1 = select(inner_too, q1n, q1);
	p1 = select(inner_too, p1n, p1);
}

// filter_mb_edge (vp8_loopfilter.c:81-104) for w = filter_value(): returns the new taps, leaves the selection to the caller.
VP8_LF2_FN void apply_mb(uint32_t w, uint32_t p2, uint32_t p1, uint32_t p0, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t& np2,
                         uint32_t& np1, uint32_t& np0, uint32_t& nq0, uint32_t& nq1, uint32_t& nq2) {
	const uint32_t wb = add2(w, both(128)); // 0..255 per half: the 32-bit products below cannot carry between halves
	const uint32_t a27 = ((wb * 27u + both(63)) >> 7) & 0x01ff01ffu; // ((27 w + 63) >> 7) + 27
	const uint32_t a18 = ((wb * 18u + both(63)) >> 7) & 0x01ff01ffu; // ... + 18
	const uint32_t a9 = ((wb * 9u + both(63)) >> 7) & 0x01ff01ffu;   // ... + 9
	np0 = addmin_relu(add2(p0, a27), both(-27), both(255));
	nq0 = addmin_relu(add2(q0, ~a27), both(28), both(255));
	np1 = addmin_relu(add2(p1, a18), both(-18), both(255));
	nq1 = addmin_relu(add2(q1, ~a18), both(19), both(255));
	np2 = addmin_relu(add2(p2, a9), both(-9), both(255));
	nq2 = addmin_relu(add2(q2, ~a9), both(10), both(255));
}

// Sub-block (inner) edge: filter_subblock_*_edge, vp8_loopfilter.c:123-150. `on` = 0xffff per half that is to be filtered
// at all (the macroblock has a filter level, inner edges enabled ...). Returns the mask of halves that changed.
VP8_LF2_FN uint32_t inner_edge(uint32_t p3, uint32_t p2, uint32_t& p1, uint32_t& p0, uint32_t& q0, uint32_t& q1, uint32_t q2, uint32_t q3,
                               const Limits& lim, uint32_t on) {
	uint32_t ok, calm;
	thresholds(p3, p2, p1, p0, q0, q1, q2, q3, lim, ok, calm);
	ok &= on;
	if (!anyone(ok)) return 0;
	const uint32_t a = filter_value(p1, p0, q0, q1, ~calm); // the p1-q1 term only with high edge variance
	uint32_t n1 = p1, n0 = p0, m0 = q0, m1 = q1;
	apply_common(a, calm, n1, n0, m0, m1); // p1/q1 move only without it
	p1 = select(ok, n1, p1);
	p0 = select(ok, n0, p0);
	q0 = select(ok, m0, q0);
	q1 = select(ok, m1, q1);
	return ok;
}

// Macroblock edge: filter_mb_*_edge, vp8_loopfilter.c:106-121.
VP8_LF2_FN uint32_t mb_edge(uint32_t p3, uint32_t& p2, uint32_t& p1, uint32_t& p0, uint32_t& q0, uint32_t& q1, uint32_t& q2, uint32_t q3,
                            const Limits& lim, uint32_t on) {
	uint32_t ok, calm;
	thresholds(p3, p2, p1, p0, q0, q1, q2, q3, lim, ok, calm);
	ok &= on;
	if (!anyone(ok)) return 0;
	const uint32_t a = filter_value(p1, p0, q0, q1, 0xffffffffu);
	// high edge variance: filter_common with the outer taps, p0/q0 only
	uint32_t h1 = p1, h0 = p0, g0 = q0, g1 = q1;
	apply_common(a, 0u, h1, h0, g0, g1);
	// otherwise the wide filter
	uint32_t n2, n1, n0, m0, m1, m2;
	apply_mb(a, p2, p1, p0, q0, q1, q2, n2, n1, n0, m0, m1, m2);
	const uint32_t wide = ok & calm;
	p0 = select(ok, select(calm, n0, h0), p0);
	q0 = select(ok, select(calm, m0, g0), q0);
	p1 = select(wide, n1, p1);
	q1 = select(wide, m1, q1);
	p2 = select(wide, n2, p2);
	q2 = select(wide, m2, q2);
	return ok;
}

} // namespace lf2

#if defined(__CUDACC__)
// ---- drivers over a filter tile in shared memory (row stride 20 bytes for every plane, see HalfWs in vp8_pairs.cu)

// Words wa / wb = four consecutive pixels of rows A / B  <->  four taps [a, 0, b, 0]
__device__ __forceinline__ void lf2_unpack(uint32_t wa, uint32_t wb, uint32_t (&t)[4]) {
	const uint32_t x = __byte_perm(wa, wb, 0x5410), y = __byte_perm(wa, wb, 0x7632); // [a0 a1 b0 b1], [a2 a3 b2 b3]
	t[0] = x & 0x00ff00ffu;
	t[1] = (x >> 8) & 0x00ff00ffu;
	t[2] = y & 0x00ff00ffu;
	t[3] = (y >> 8) & 0x00ff00ffu;
}
__device__ __forceinline__ void lf2_pack(const uint32_t (&t)[4], uint32_t& wa, uint32_t& wb) {
	const uint32_t x = t[0] | (t[1] << 8), y = t[2] | (t[3] << 8);
	wa = __byte_perm(x, y, 0x5410);
	wb = __byte_perm(x, y, 0x7632);
}

// The four vertical edges (x = 0, 4, 8, 12) of two pixel rows. row_a / row_b point at pixel x = 0 of the rows; the tile has
// four more pixels to the left. Chroma rows (luma == false) are 8 pixels wide: edges 0 and 4 only.
__device__ __forceinline__ void lf2_vertical_edges(uint8_t* row_a, uint8_t* row_b, bool luma, bool any_mb, bool any_inner, const lf2::Limits& l_mb,
                                                   const lf2::Limits& l_in, uint32_t on_mb, uint32_t on_inner, uint32_t on_inner_luma) {
	uint32_t t[5][4];
#pragma unroll
	for (int w = 0; w < 5; w++) {
		uint32_t wa = 0, wb = 0;
		if (w < 3 || luma) {
			wa = *reinterpret_cast<const uint32_t*>(row_a + 4 * w - 4);
			wb = *reinterpret_cast<const uint32_t*>(row_b + 4 * w - 4);
		}
		lf2_unpack(wa, wb, t[w]);
	}
	if (any_mb) lf2::mb_edge(t[0][0], t[0][1], t[0][2], t[0][3], t[1][0], t[1][1], t[1][2], t[1][3], l_mb, on_mb);
	if (any_inner) {
		lf2::inner_edge(t[1][0], t[1][1], t[1][2], t[1][3], t[2][0], t[2][1], t[2][2], t[2][3], l_in, on_inner);
		lf2::inner_edge(t[2][0], t[2][1], t[2][2], t[2][3], t[3][0], t[3][1], t[3][2], t[3][3], l_in, on_inner_luma);
		lf2::inner_edge(t[3][0], t[3][1], t[3][2], t[3][3], t[4][0], t[4][1], t[4][2], t[4][3], l_in, on_inner_luma);
	}
#pragma unroll
	for (int w = 0; w < 5; w++) {
		uint32_t wa, wb;
		lf2_pack(t[w], wa, wb);
		if (w < 3 || luma) {
			*reinterpret_cast<uint32_t*>(row_a + 4 * w - 4) = wa;
			*reinterpret_cast<uint32_t*>(row_b + 4 * w - 4) = wb;
		}
	}
}

// The four horizontal edges (y = 0, 4, 8, 12) of two adjacent pixel columns. col points at pixel y = 0 of the left column
// (an even x); the tile has four more rows above. Chroma columns are 8 pixels tall: edges 0 and 4 only.
__device__ __forceinline__ void lf2_horizontal_edges(uint8_t* col, bool luma, bool any_mb, bool any_inner, const lf2::Limits& l_mb,
                                                     const lf2::Limits& l_in, uint32_t on_mb, uint32_t on_inner, uint32_t on_inner_luma) {
	constexpr int S = 20;
	uint32_t r[20]; // rows -4 .. 15 as [left, 0, right, 0]
#pragma unroll
	for (int i = 0; i < 20; i++) {
		uint32_t v = 0;
		if (i < 12 || luma) v = *reinterpret_cast<const uint16_t*>(col + (i - 4) * S);
		r[i] = __byte_perm(v, 0, 0x4140);
	}
	if (any_mb) lf2::mb_edge(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], l_mb, on_mb);
	if (any_inner) {
		lf2::inner_edge(r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], l_in, on_inner);
		lf2::inner_edge(r[8], r[9], r[10], r[11], r[12], r[13], r[14], r[15], l_in, on_inner_luma);
		lf2::inner_edge(r[12], r[13], r[14], r[15], r[16], r[17], r[18], r[19], l_in, on_inner_luma);
	}
#pragma unroll
	for (int i = 1; i < 19; i++) // rows -3 .. 14 can have changed
		if (i < 11 || luma) *reinterpret_cast<uint16_t*>(col + (i - 4) * S) = (uint16_t)__byte_perm(r[i], 0, 0x4420);
}
#endif

