// lf2_check.cpp - host check of the packed two-position loop filter (webp-decoder_b200/csrc/vp8_lf2.cuh, host flavour)
// against a scalar restatement of the reference's filter (vp8_loopfilter.c:24-121). Built and run by tests/test_host.py.
// Prints "ok <cases>" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../webp-decoder_b200/csrc/vp8_lf2.cuh"

static int clamp8(int v) { return v < -128 ? -128 : v > 127 ? 127 : v; }
static int u8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static int ad(int a, int b) { return a > b ? a - b : b - a; }

struct Taps { int p3, p2, p1, p0, q0, q1, q2, q3; };

static bool normal_threshold(const Taps& t, int E, int I) {
	if (2 * ad(t.p0, t.q0) + (ad(t.p1, t.q1) >> 1) > E) return false;
	return ad(t.p3, t.p2) <= I && ad(t.p2, t.p1) <= I && ad(t.p1, t.p0) <= I && ad(t.q3, t.q2) <= I && ad(t.q2, t.q1) <= I && ad(t.q1, t.q0) <= I;
}
static bool hev(const Taps& t, int T) { return ad(t.p1, t.p0) > T || ad(t.q1, t.q0) > T; }
static void common(Taps& t, bool outer) {
	int a = 3 * (t.q0 - t.p0);
	if (outer) a += clamp8(t.p1 - t.q1);
	a = clamp8(a);
	const int f1 = clamp8(a + 4) >> 3, f2 = clamp8(a + 3) >> 3;
	t.q0 = u8(t.q0 - f1);
	t.p0 = u8(t.p0 + f2);
	if (!outer) {
		const int h = (f1 + 1) >> 1;
		t.q1 = u8(t.q1 - h);
		t.p1 = u8(t.p1 + h);
	}
}
static void wide(Taps& t) {
	const int w = clamp8(clamp8(t.p1 - t.q1) + 3 * (t.q0 - t.p0));
	int a = (27 * w + 63) >> 7;
	t.q0 = u8(t.q0 - a); t.p0 = u8(t.p0 + a);
	a = (18 * w + 63) >> 7;
	t.q1 = u8(t.q1 - a); t.p1 = u8(t.p1 + a);
	a = (9 * w + 63) >> 7;
	t.q2 = u8(t.q2 - a); t.p2 = u8(t.p2 + a);
}
static void ref_inner(Taps& t, int E, int I, int T, bool on) {
	if (on && normal_threshold(t, E, I)) common(t, hev(t, T));
}
static void ref_mb(Taps& t, int E, int I, int T, bool on) {
	if (on && normal_threshold(t, E, I)) {
		if (hev(t, T)) common(t, true);
		else wide(t);
	}
}

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }
static Taps random_taps(int spread) {
	Taps t;
	int base = rnd() % 256;
	int* f[8] = {&t.p3, &t.p2, &t.p1, &t.p0, &t.q0, &t.q1, &t.q2, &t.q3};
	for (int i = 0; i < 8; i++) {
		int v = spread >= 256 ? (int)(rnd() % 256) : base + (int)(rnd() % (2 * spread + 1)) - spread;
		*f[i] = v < 0 ? 0 : v > 255 ? 255 : v;
	}
	return t;
}
static uint32_t pk(int a, int b) { return (uint32_t)a | ((uint32_t)b << 16); }

int main() {
	long cases = 0;
	const int spreads[] = {0, 1, 2, 4, 8, 16, 40, 256};
	for (int iter = 0; iter < 400000; iter++) {
		const int sp = spreads[iter % 8];
		Taps a = random_taps(sp), b = random_taps(spreads[(iter / 8) % 8]);
		if (iter % 97 == 0) { a.p0 = 0; a.q0 = 255; a.p1 = 255; a.q1 = 0; }
		if (iter % 89 == 0) { b.p0 = 255; b.q0 = 0; b.p1 = 0; b.q1 = 255; }
		// levels 0..63, sharpness 0..7 -> interior 1..63 (0 unreachable but harmless), hev 0..2, E up to 2*(63+2)+63
		const int Ea = rnd() % 194, Eb = rnd() % 194, Ia = rnd() % 64, Ib = rnd() % 64, Ta = rnd() % 3, Tb = rnd() % 3;
		const bool ona = rnd() % 8 != 0, onb = rnd() % 8 != 0;
		const lf2::Limits lim = lf2::make_limits(Ea, Eb, Ia, Ib, Ta, Tb);
		const uint32_t on = (ona ? 0xffffu : 0u) | (onb ? 0xffff0000u : 0u);
		for (int kind = 0; kind < 2; kind++) {
			Taps ra = a, rb = b;
			uint32_t p3 = pk(a.p3, b.p3), p2 = pk(a.p2, b.p2), p1 = pk(a.p1, b.p1), p0 = pk(a.p0, b.p0);
			uint32_t q0 = pk(a.q0, b.q0), q1 = pk(a.q1, b.q1), q2 = pk(a.q2, b.q2), q3 = pk(a.q3, b.q3);
			if (kind == 0) {
				ref_inner(ra, Ea, Ia, Ta, ona);
				ref_inner(rb, Eb, Ib, Tb, onb);
				lf2::inner_edge(p3, p2, p1, p0, q0, q1, q2, q3, lim, on);
			} else {
				ref_mb(ra, Ea, Ia, Ta, ona);
				ref_mb(rb, Eb, Ib, Tb, onb);
				lf2::mb_edge(p3, p2, p1, p0, q0, q1, q2, q3, lim, on);
			}
			const uint32_t want[8] = {pk(ra.p3, rb.p3), pk(ra.p2, rb.p2), pk(ra.p1, rb.p1), pk(ra.p0, rb.p0),
			                          pk(ra.q0, rb.q0), pk(ra.q1, rb.q1), pk(ra.q2, rb.q2), pk(ra.q3, rb.q3)};
			const uint32_t got[8] = {p3, p2, p1, p0, q0, q1, q2, q3};
			for (int i = 0; i < 8; i++)
				if (want[i] != got[i]) {
					printf("MISMATCH kind %d tap %d: got %08x want %08x (iter %d; A %d %d %d %d | %d %d %d %d  E %d I %d T %d on %d)\n", kind, i,
					       got[i], want[i], iter, a.p3, a.p2, a.p1, a.p0, a.q0, a.q1, a.q2, a.q3, Ea, Ia, Ta, (int)ona);
					return 1;
				}
			cases += 2;
		}
	}
	printf("ok %ld\n", cases);
	return 0;
}
