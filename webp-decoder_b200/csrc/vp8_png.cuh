// vp8_png.cuh - m09 on the device: the arithmetic of the PNG framing kernels (vp8_png.cu), written so that the host can run
// it too (tests/native/png_check.cpp replays the kernels' grid on the CPU against an independent writer).
//
// The file the reference emits (yuv2rgb_png.c:208-364) is a pure function of the RGB bytes:
//   0  signature (8) | 8 IHDR chunk (25) | 33 IDAT length (4) | 37 "IDAT" | 41 zlib stream (zsize) | CRC (4) | IEND chunk (12)
//   zlib stream = 78 01, then stored blocks of at most 65535 bytes (5-byte header each) over the scanlines (filter byte 0 +
//   3*w RGB bytes per line), then the Adler-32 of the scanline bytes.
// Every thread produces 16 consecutive FILE bytes per round (whole 16-byte stores at 16-byte aligned file offsets; the file
// starts on a 256-byte boundary of the device buffer), so the kernel is a copy with a byte shift that changes at line and
// block boundaries. Both checksums are linear in the data and are formed in pieces:
//   Adler-32: A = 1 + sum v_j, B = n + sum v_j * (n - j) (j = position in the scanline stream, n = its length); pieces are
//             added up in 64-bit atomics without any ordering, the finish kernel reduces mod 65521.
//   CRC-32:   the register is a polynomial over GF(2) (reflected, bit 31 = x^0); running over n more bytes multiplies it by
//             x^(8n) mod P. A thread runs a table CRC over its 16-byte segments, moving from one segment to the next of the
//             same CTA (4080 bytes on) by one table multiplication; at the end it is multiplied up to the end of the CTA's
//             span, the CTA's XOR is multiplied up to the end of the region the main kernel covers and XORed into the
//             image's accumulator. The finish kernel adds the initial value's term, the last < 16 bytes and the Adler bytes.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PNG_HD __host__ __device__ __forceinline__
#else
#define PNG_HD inline
#endif

namespace pngk {

constexpr uint32_t kPoly = 0xEDB88320u;
constexpr uint32_t kOne = 0x80000000u;        // x^0
constexpr uint32_t kSeg = 16;                 // file bytes per thread per round
constexpr uint32_t kThreads = 256;
constexpr uint32_t kRound = kSeg * kThreads;  // 4096 file bytes per CTA round
constexpr uint32_t kRounds = 16;
constexpr uint32_t kSpan = kRound * kRounds;  // 65536 file bytes per CTA
constexpr uint32_t kHead = 43;                // signature .. 78 01: bytes that do not depend on the pixels
constexpr uint32_t kCrcBegin = 37;            // "IDAT": first byte the IDAT CRC covers
constexpr uint32_t kStored = 65535;           // scanline bytes per stored block
constexpr uint32_t kStoredZ = kStored + 5;    // ... and what a full block takes in the zlib stream
constexpr uint32_t kAdlerMod = 65521;

// Geometry of one image's file. Everything fits 32 bits: the reference refuses scanline streams beyond 0x7FFFFFFF bytes.
struct Geom {
	uint32_t w, h;
	uint32_t row;      // 3 * w
	uint32_t line;     // row + 1
	uint32_t raw;      // line * h, the scanline stream
	uint32_t blocks;   // stored blocks
	uint32_t zsize;    // 2 + raw + 5 * blocks + 4
	uint32_t adler_at; // file offset of the Adler-32 = 37 + zsize
	uint32_t crc_end;  // adler_at rounded down to 16: the main kernel's CRC covers [37, crc_end)
	uint32_t file_len; // 57 + zsize
	uint32_t line_inv; // ceil(2^32 / line): scanline position -> line by one multiplication (div_line)
};

PNG_HD Geom geom(uint32_t w, uint32_t h) {
	Geom g;
	g.w = w;
	g.h = h;
	g.row = 3 * w;
	g.line = g.row + 1;
	g.raw = g.line * h;
	g.blocks = (g.raw + kStored - 1) / kStored;
	g.zsize = 2 + g.raw + 5 * g.blocks + 4;
	g.adler_at = 37 + g.zsize;
	g.crc_end = g.adler_at & ~15u;
	g.file_len = 57 + g.zsize;
	g.line_inv = 0xFFFFFFFFu / g.line + 1;
	return g;
}
// r / line and r % line for r < 2^31: the estimate umulhi(r, ceil(2^32 / line)) is the quotient or one more.
PNG_HD uint32_t div_line(const Geom& g, uint32_t r, uint32_t& rem) {
#if defined(__CUDA_ARCH__)
	uint32_t q = __umulhi(r, g.line_inv);
#else
	uint32_t q = (uint32_t)(((uint64_t)r * g.line_inv) >> 32);
#endif
	int32_t m = (int32_t)(r - q * g.line);
	if (m < 0) m += (int32_t)g.line, q--;
	rem = (uint32_t)m;
	return q;
}
PNG_HD bool geom_ok(uint64_t w, uint64_t h) { return w && h && (3 * w + 1) * h <= 0x7FFFFFFFull; }
PNG_HD uint32_t ctas_of(const Geom& g) { return (g.file_len + kSpan - 1) / kSpan; }

// a * b mod P (zlib's multmodp, branch-free)
PNG_HD uint32_t mulmod(uint32_t a, uint32_t b) {
	uint32_t p = 0;
	for (int i = 0; i < 32; i++) {
		p ^= b & (0u - (a >> 31));
		a <<= 1;
		b = (b >> 1) ^ (kPoly & (0u - (b & 1)));
	}
	return p;
}

// Tables, built once on the host (build_tables) and kept in device memory.
struct Tables {
	uint32_t z4[4][256];  // slice-by-4: register ^ word -> register 4 bytes on. z4[t][i] = byte i followed by t zero bytes
	uint32_t zr[4][256];  // register -> register (kRound - kSeg) zero bytes on, by register byte
	uint32_t xp16[256];   // x^(8 * 16 * k)
	uint32_t x2n[32];     // x^(8 * 16 * 2^k)
};

inline void build_tables(Tables& t) {
	for (uint32_t i = 0; i < 256; i++) {
		uint32_t v = i;
		for (int k = 0; k < 8; k++) v = (v & 1) ? kPoly ^ (v >> 1) : v >> 1;
		t.z4[0][i] = v;
	}
	for (uint32_t i = 0; i < 256; i++)
		for (int k = 1; k < 4; k++) t.z4[k][i] = t.z4[0][t.z4[k - 1][i] & 255] ^ (t.z4[k - 1][i] >> 8);
	uint32_t x8 = kOne >> 8; // x^8: one zero byte
	uint32_t x128 = kOne;
	for (int k = 0; k < 16; k++) x128 = mulmod(x128, x8);
	t.x2n[0] = x128;
	for (int k = 1; k < 32; k++) t.x2n[k] = mulmod(t.x2n[k - 1], t.x2n[k - 1]);
	t.xp16[0] = kOne;
	for (int k = 1; k < 256; k++) t.xp16[k] = mulmod(t.xp16[k - 1], x128);
	const uint32_t adv = t.xp16[(kRound - kSeg) / 16]; // 4080 bytes = 255 segments
	for (int j = 0; j < 4; j++)
		for (uint32_t b = 0; b < 256; b++) t.zr[j][b] = mulmod(b << (8 * j), adv);
}

PNG_HD uint32_t crc_word(const uint32_t (*z4)[256], uint32_t reg, uint32_t word) {
	const uint32_t x = reg ^ word;
	return z4[3][x & 255] ^ z4[2][(x >> 8) & 255] ^ z4[1][(x >> 16) & 255] ^ z4[0][x >> 24];
}
PNG_HD uint32_t crc_advance_round(const uint32_t (*zr)[256], uint32_t reg) {
	return zr[0][reg & 255] ^ zr[1][(reg >> 8) & 255] ^ zr[2][(reg >> 16) & 255] ^ zr[3][reg >> 24];
}
PNG_HD uint32_t crc_byte(uint32_t reg, uint32_t byte) { // table-free, for the finish kernel's few bytes
	reg ^= byte;
	for (int k = 0; k < 8; k++) reg = (reg >> 1) ^ (kPoly & (0u - (reg & 1)));
	return reg;
}
// x^(8 * 16 * n16), serially (the kernel's first warp forms the same product as a tree)
PNG_HD uint32_t xpow16(const Tables& t, uint32_t n16) {
	uint32_t p = kOne;
	for (int k = 0; n16; k++, n16 >>= 1)
		if (n16 & 1) p = mulmod(p, t.x2n[k]);
	return p;
}
// x^(8 * n) for any byte count (host: the initial value's term)
inline uint32_t xpow_bytes(uint32_t n) {
	uint32_t p = kOne, base = kOne >> 8;
	for (; n; n >>= 1, base = mulmod(base, base))
		if (n & 1) p = mulmod(p, base);
	return p;
}
// The register's initial value carried over the bytes the main kernel covers.
inline uint32_t crc_init_term(const Geom& g) { return mulmod(0xFFFFFFFFu, xpow_bytes(g.crc_end - kCrcBegin)); }

// The 43 bytes in front of the first stored block: signature, IHDR chunk, IDAT length, "IDAT", 78 01.
inline void build_head(const Geom& g, const Tables& t, uint8_t head[44]) {
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
	auto be32 = [](uint8_t* p, uint32_t v) {
		p[0] = (uint8_t)(v >> 24);
		p[1] = (uint8_t)(v >> 16);
		p[2] = (uint8_t)(v >> 8);
		p[3] = (uint8_t)v;
	};
	for (int i = 0; i < 8; i++) head[i] = sig[i];
	be32(head + 8, 13);
	head[12] = 'I', head[13] = 'H', head[14] = 'D', head[15] = 'R';
	be32(head + 16, g.w);
	be32(head + 20, g.h);
	head[24] = 8, head[25] = 2, head[26] = head[27] = head[28] = 0;
	uint32_t crc = 0xFFFFFFFFu;
	for (int i = 12; i < 29; i++) crc = t.z4[0][(crc ^ head[i]) & 255] ^ (crc >> 8);
	be32(head + 29, crc ^ 0xFFFFFFFFu);
	be32(head + 33, g.zsize);
	head[37] = 'I', head[38] = 'D', head[39] = 'A', head[40] = 'T';
	head[41] = 0x78, head[42] = 0x01;
	head[43] = 0;
}

// Byte-at-a-time walk through the file from any offset (line ends, block headers, the head, the tail): the slow path of a
// segment. Also feeds the Adler sums: a += v, b += v * ((raw - r) mod 65521) for every scanline byte.
struct Walker {
	uint32_t f;        // file offset
	uint32_t k, o;     // stored block and offset inside its kStoredZ bytes (valid once f >= kHead)
	uint32_t rowi, col;
	uint32_t wt;       // (raw - r) mod 65521 of the next scanline byte, ~0u = not formed yet
};
PNG_HD Walker walker_at(const Geom& g, uint32_t f) {
	Walker wk;
	wk.f = f;
	wk.k = wk.o = wk.rowi = wk.col = 0;
	wk.wt = ~0u;
	if (f >= kHead) {
		const uint32_t z = f - kHead;
		wk.k = z / kStoredZ;
		wk.o = z - wk.k * kStoredZ;
		const uint32_t r = wk.k * kStored + (wk.o > 5 ? wk.o - 5 : 0); // scanline position of the next scanline byte
		wk.rowi = div_line(g, r, wk.col);
	}
	return wk;
}
PNG_HD uint32_t walker_next(const Geom& g, const uint8_t* head, const uint8_t* rgb, Walker& wk, uint32_t& a, uint32_t& b) {
	uint32_t v = 0;
	if (wk.f < kHead) {
		v = head[wk.f];
	} else if (wk.f >= g.adler_at) {
		v = 0; // Adler, CRC, IEND: the finish kernel writes them
	} else if (wk.o < 5) {
		const uint32_t left = g.raw - wk.k * kStored, len = left < kStored ? left : kStored;
		v = wk.o == 0 ? (wk.k + 1 == g.blocks ? 1u : 0u) : wk.o == 1 ? (len & 255) : wk.o == 2 ? (len >> 8) : wk.o == 3 ? (~len & 255) : ((~len >> 8) & 255);
		wk.o++;
	} else {
		if (wk.col) v = rgb[(size_t)wk.rowi * g.row + wk.col - 1];
		if (wk.wt == ~0u) wk.wt = (g.raw - (wk.k * kStored + wk.o - 5)) % kAdlerMod;
		a += v;
		b += v * wk.wt;
		wk.wt = wk.wt ? wk.wt - 1 : kAdlerMod - 1;
		if (++wk.col == g.line) wk.col = 0, wk.rowi++;
		if (++wk.o == kStoredZ) wk.o = 0, wk.k++;
	}
	wk.f++;
	return v;
}

// Is the 16-byte segment at file offset f (a multiple of 16) made of 16 scanline bytes of one stored block, at most one of them
// a filter byte? The RGB image is tight, so those are consecutive RGB bytes with (perhaps) one zero put in between: src = offset
// of the first RGB byte in the image, zero_at = where the filter byte goes (16 = nowhere; bytes behind it come from one RGB byte
// earlier), r0 = scanline position of the segment's first byte.
PNG_HD bool seg_is_copy(const Geom& g, uint32_t f, size_t& src, uint32_t& r0, uint32_t& zero_at) {
	if (f < 48 || g.line < kSeg) return false;
	const uint32_t z = f - kHead, k = z / kStoredZ, o = z - k * kStoredZ;
	if (o < 5 || o + kSeg > kStoredZ) return false;
	r0 = k * kStored + o - 5;
	if (r0 + kSeg > g.raw) return false;
	uint32_t col;
	const uint32_t rowi = div_line(g, r0, col);
	const uint32_t s0 = rowi * g.row + (col ? col - 1 : 0);
	if (s0 + kSeg > g.row * g.h) return false; // the last bytes of the image go the slow way (16 RGB bytes are read regardless)
	src = s0;
	zero_at = col ? g.line - col : 0;
	if (zero_at > kSeg) zero_at = kSeg;
	return true;
}
// 16 bytes s[] (little-endian words) with a zero byte put in at position at (0..15): bytes behind it move up by one, the last
// one drops out.
PNG_HD void insert_zero(uint32_t s[4], uint32_t at) {
	const uint32_t t0 = s[0] << 8, t1 = (s[1] << 8) | (s[0] >> 24), t2 = (s[2] << 8) | (s[1] >> 24), t3 = (s[3] << 8) | (s[2] >> 24);
	const uint32_t t[4] = {t0, t1, t2, t3};
	for (int i = 0; i < 4; i++) {
		const int d = (int)at - 4 * i; // position of the zero relative to this word
		const uint32_t lo = d <= 0 ? 0u : d >= 4 ? 0xFFFFFFFFu : (1u << (8 * d)) - 1u;         // bytes in front of it: as they are
		const uint32_t hi = d < 0 ? 0xFFFFFFFFu : d >= 3 ? 0u : ~((1u << (8 * d + 8)) - 1u);   // bytes behind it: moved up
		s[i] = (s[i] & lo) | (t[i] & hi);
	}
}

PNG_HD uint32_t dot4(uint32_t bytes, uint32_t weights, uint32_t acc) {
#if defined(__CUDA_ARCH__)
	return __dp4a(bytes, weights, acc);
#else
	for (int k = 0; k < 4; k++) acc += ((bytes >> (8 * k)) & 255) * ((weights >> (8 * k)) & 255);
	return acc;
#endif
}
// Adler pieces of a plain segment: a += sum v_j, b += sum v_j * (raw - r0 - j)  (mod 65521 where it matters)
PNG_HD void seg_adler(const Geom& g, const uint32_t o[4], uint32_t r0, uint32_t& a, uint64_t& b) {
	uint32_t sa = 0, sb = 0;
	sa = dot4(o[0], 0x01010101u, sa), sb = dot4(o[0], 0x0D0E0F10u, sb);
	sa = dot4(o[1], 0x01010101u, sa), sb = dot4(o[1], 0x090A0B0Cu, sb);
	sa = dot4(o[2], 0x01010101u, sa), sb = dot4(o[2], 0x05060708u, sb);
	sa = dot4(o[3], 0x01010101u, sa), sb = dot4(o[3], 0x01020304u, sb);
	a += sa;
	b += sb + (uint64_t)sa * ((g.raw - r0 - kSeg) % kAdlerMod); // sa * q <= 4080 * 65520
}

// What one thread carries through its rounds.
struct ThreadAcc {
	uint32_t crc;   // register, aligned to the end of the last segment it covered
	uint32_t a;     // sum of scanline bytes
	uint64_t b;     // weighted sum
	int last_seg;   // index (inside the CTA's span) of the last segment the CRC covered, -1 = none
};

// One segment: 16 file bytes at offset f into o[4] (little-endian words), checksums updated. rgb must be 4-byte aligned.
PNG_HD void segment(const Geom& g, const Tables* t, const uint8_t* head, const uint8_t* rgb, uint32_t f, int seg_index, uint32_t o[4],
                    ThreadAcc& acc) {
	size_t src;
	uint32_t r0, zero_at;
	if (seg_is_copy(g, f, src, r0, zero_at)) {
		const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(rgb + (src & ~(size_t)3));
		const uint32_t sh = (uint32_t)(src & 3) * 8;
		const uint32_t w0 = wsrc[0], w1 = wsrc[1], w2 = wsrc[2], w3 = wsrc[3];
		if (sh) {
			const uint32_t w4 = wsrc[4];
			o[0] = (w0 >> sh) | (w1 << (32 - sh));
			o[1] = (w1 >> sh) | (w2 << (32 - sh));
			o[2] = (w2 >> sh) | (w3 << (32 - sh));
			o[3] = (w3 >> sh) | (w4 << (32 - sh));
		} else {
			o[0] = w0, o[1] = w1, o[2] = w2, o[3] = w3;
		}
		if (zero_at < kSeg) insert_zero(o, zero_at); // a scanline ends inside the segment
		seg_adler(g, o, r0, acc.a, acc.b);
	} else {
		Walker wk = walker_at(g, f);
		uint32_t sa = 0, sb = 0;
		for (int i = 0; i < 4; i++) {
			uint32_t word = 0;
			for (int k = 0; k < 4; k++) word |= walker_next(g, head, rgb, wk, sa, sb) << (8 * k);
			o[i] = word;
		}
		acc.a += sa;
		acc.b += sb;
	}
	if (f + kSeg <= g.crc_end) {
		uint32_t c0 = o[0], c1 = o[1], c2 = o[2], c3 = o[3];
		if (f < 48) { // bytes in front of "IDAT" are not covered: leading zeros leave a zero register alone
			if (f < 32) c0 = c1 = c2 = c3 = 0;
			else c0 = 0, c1 &= 0xFFFFFF00u; // file bytes 32..36
		}
		uint32_t r = crc_advance_round(t->zr, acc.crc); // from the end of this thread's previous segment (0 stays 0)
		r = crc_word(t->z4, r, c0);
		r = crc_word(t->z4, r, c1);
		r = crc_word(t->z4, r, c2);
		r = crc_word(t->z4, r, c3);
		acc.crc = r;
		acc.last_seg = seg_index;
	}
}

// Segments of the CTA's span that the CRC covers.
PNG_HD uint32_t span_crc_segments(const Geom& g, uint32_t span_base) {
	if (g.crc_end <= span_base) return 0;
	const uint32_t n = (g.crc_end - span_base) / kSeg;
	return n < kSpan / kSeg ? n : kSpan / kSeg;
}

// Per-image accumulators (zeroed before the main kernel).
struct Accum {
	uint32_t crc;
	uint32_t pad_;
	unsigned long long a, b;
};

// Finish: everything behind crc_end. crc_init = crc_init_term(g).
PNG_HD void finish(const Geom& g, uint32_t crc_init, const Accum& acc, uint8_t* out) {
	uint32_t crc = crc_init ^ acc.crc;
	for (uint32_t f = g.crc_end; f < g.adler_at; f++) crc = crc_byte(crc, out[f]);
	const uint32_t a = (uint32_t)((1 + acc.a) % kAdlerMod);
	const uint32_t b = (uint32_t)((g.raw % kAdlerMod + acc.b % kAdlerMod) % kAdlerMod);
	const uint32_t adler = (b << 16) | a;
	uint8_t* p = out + g.adler_at;
	for (int k = 0; k < 4; k++) {
		p[k] = (uint8_t)(adler >> (24 - 8 * k));
		crc = crc_byte(crc, p[k]);
	}
	crc ^= 0xFFFFFFFFu;
	for (int k = 0; k < 4; k++) p[4 + k] = (uint8_t)(crc >> (24 - 8 * k));
	const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
	for (int k = 0; k < 12; k++) p[8 + k] = iend[k];
}

} // namespace pngk
