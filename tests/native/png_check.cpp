// png_check.cpp - host replay of the m09 kernels' grid (webp-decoder_b200/csrc/vp8_png.cuh: every CTA, every thread, the
// reductions and the finish step, in plain loops) against an independent, byte-at-a-time PNG writer that follows the
// reference's framing (yuv2rgb_png.c:208-364) with a bitwise CRC-32 and a scalar Adler-32. Built and run by
// tests/test_host.py. Prints "ok <images> <bytes>" or the first mismatch; `png_check dump <w> <h> <file>` writes the
// replayed file of one random image so that the test can take it apart with zlib.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../webp-decoder_b200/csrc/vp8_png.cuh"

using namespace pngk;

static uint32_t rng_state = 2463534242u;
static uint32_t rnd() { return rng_state = rng_state * 1664525u + 1013904223u; }

static uint32_t crc_bitwise(uint32_t crc, const uint8_t* p, size_t n) {
	for (size_t i = 0; i < n; i++) {
		crc ^= p[i];
		for (int k = 0; k < 8; k++) crc = (crc & 1) ? 0xEDB88320u ^ (crc >> 1) : crc >> 1;
	}
	return crc;
}
static void put32(std::vector<uint8_t>& o, uint32_t v) {
	for (int k = 3; k >= 0; k--) o.push_back((uint8_t)(v >> (8 * k)));
}
static void chunk(std::vector<uint8_t>& o, const char* type, const std::vector<uint8_t>& data) {
	put32(o, (uint32_t)data.size());
	const size_t at = o.size();
	o.insert(o.end(), type, type + 4);
	o.insert(o.end(), data.begin(), data.end());
	put32(o, crc_bitwise(0xFFFFFFFFu, o.data() + at, 4 + data.size()) ^ 0xFFFFFFFFu);
}
static std::vector<uint8_t> plain_png(const uint8_t* rgb, uint32_t w, uint32_t h) {
	std::vector<uint8_t> o = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A}, ihdr, raw, z;
	put32(ihdr, w);
	put32(ihdr, h);
	ihdr.insert(ihdr.end(), {8, 2, 0, 0, 0});
	chunk(o, "IHDR", ihdr);
	for (uint32_t y = 0; y < h; y++) {
		raw.push_back(0);
		raw.insert(raw.end(), rgb + (size_t)y * w * 3, rgb + (size_t)(y + 1) * w * 3);
	}
	z = {0x78, 0x01};
	uint32_t a = 1, b = 0;
	for (size_t pos = 0; pos < raw.size(); pos += 65535) {
		const size_t len = raw.size() - pos < 65535 ? raw.size() - pos : 65535;
		z.push_back(pos + len == raw.size() ? 1 : 0);
		z.push_back((uint8_t)len), z.push_back((uint8_t)(len >> 8)), z.push_back((uint8_t)~len), z.push_back((uint8_t)(~len >> 8));
		z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
	}
	for (uint8_t v : raw) a = (a + v) % 65521, b = (b + a) % 65521;
	put32(z, (b << 16) | a);
	chunk(o, "IDAT", z);
	chunk(o, "IEND", {});
	return o;
}

// the grid of vp8_png_frame + vp8_png_finish for one image
static std::vector<uint8_t> replay(const Tables& tb, const uint8_t* rgb, uint32_t w, uint32_t h) {
	const Geom g = geom(w, h);
	uint8_t head[44];
	build_head(g, tb, head);
	const uint32_t write_end = (g.file_len + 15) & ~15u;
	std::vector<uint8_t> out(write_end + 16, 0xA5);
	Accum accum{0, 0, 0, 0};
	for (uint32_t cta = 0; cta < ctas_of(g); cta++) {
		const uint32_t span_base = cta * kSpan, nseg = span_crc_segments(g, span_base);
		uint32_t crc = 0;
		unsigned long long a = 0, b = 0;
		for (uint32_t t = 0; t < kThreads; t++) {
			ThreadAcc acc{0, 0, 0, -1};
			for (uint32_t it = 0; it < kRounds; it++) {
				const uint32_t s = it * kThreads + t, f = span_base + s * kSeg;
				if (f >= write_end) break;
				uint32_t o[4];
				segment(g, &tb, head, rgb, f, (int)s, o, acc);
				memcpy(out.data() + f, o, 16);
			}
			if (acc.last_seg >= 0) crc ^= mulmod(acc.crc, tb.xp16[nseg - 1 - (uint32_t)acc.last_seg]);
			a += acc.a;
			b += acc.b;
		}
		const uint32_t n16 = nseg ? (g.crc_end - (span_base + nseg * kSeg)) / kSeg : 0;
		if (n16) crc = mulmod(crc, xpow16(tb, n16));
		accum.crc ^= crc;
		accum.a += a;
		accum.b += b;
	}
	finish(g, crc_init_term(g), accum, out.data());
	out.resize(g.file_len);
	return out;
}

static std::vector<uint32_t> random_rgb(uint32_t w, uint32_t h, int flavour) {
	std::vector<uint32_t> words(((size_t)w * h * 3 + 3) / 4 + 2, 0);
	uint8_t* p = (uint8_t*)words.data();
	for (size_t i = 0; i < (size_t)w * h * 3; i++) p[i] = flavour == 0 ? (uint8_t)(rnd() >> 24) : flavour == 1 ? 255 : (uint8_t)((rnd() >> 30) ? 0 : 255);
	return words;
}

int main(int argc, char** argv) {
	Tables tb;
	build_tables(tb);
	if (argc == 5 && !strcmp(argv[1], "dump")) {
		const uint32_t w = (uint32_t)atoi(argv[2]), h = (uint32_t)atoi(argv[3]);
		const auto rgb = random_rgb(w, h, 0);
		const auto png = replay(tb, (const uint8_t*)rgb.data(), w, h);
		FILE* f = fopen(argv[4], "wb");
		if (!f || fwrite(png.data(), 1, png.size(), f) != png.size()) return 2;
		fclose(f);
		return 0;
	}
	// the reciprocal division behind the scanline index, at both ends of its range
	for (uint32_t w : {1u, 2u, 5u, 21u, 85u, 341u, 1365u, 1920u, 5461u, 16383u}) {
		Geom g = geom(w, 1);
		for (int i = 0; i < 200000; i++) {
			const uint32_t r = i < 1000 ? (uint32_t)i : i < 2000 ? 0x7FFFFFFFu - (uint32_t)(i - 1000) : i < 3000 ? g.line * (uint32_t)(i - 2000) : (rnd() >> 1);
			uint32_t rem;
			const uint32_t q = div_line(g, r, rem);
			if (q != r / g.line || rem != r % g.line) {
				printf("div_line(%u, line %u) = %u rem %u\n", r, g.line, q, rem);
				return 1;
			}
		}
	}
	// widths around the segment / line / block geometry, the smallest images, a file of many spans
	static const uint32_t sizes[][2] = {{1, 1},    {1, 2},    {2, 1},     {3, 3},    {5, 1},     {4, 7},     {5, 5},   {16, 16},  {21, 13},
	                                    {85, 3},   {129, 129}, {256, 255}, {341, 64}, {1000, 7},  {21845, 1}, {21845, 2}, {21846, 3}, {7, 9000},
	                                    {1365, 48}, {1366, 49}, {640, 480}, {1000, 700}, {1920, 1080}, {3840, 2160}, {16383, 5}, {3, 16383}};
	long images = 0, bytes = 0;
	for (auto& wh : sizes)
		for (int flavour = 0; flavour < 3; flavour++) {
			const uint32_t w = wh[0], h = wh[1];
			if ((uint64_t)w * h > 700000 && flavour == 2) continue;
			const auto rgb = random_rgb(w, h, flavour);
			const auto want = plain_png((const uint8_t*)rgb.data(), w, h);
			const auto got = replay(tb, (const uint8_t*)rgb.data(), w, h);
			if (got.size() != want.size()) {
				printf("size mismatch %ux%u: %zu vs %zu\n", w, h, got.size(), want.size());
				return 1;
			}
			for (size_t i = 0; i < want.size(); i++)
				if (got[i] != want[i]) {
					printf("mismatch %ux%u flavour %d at byte %zu of %zu: %02x vs %02x\n", w, h, flavour, i, want.size(), got[i], want[i]);
					return 1;
				}
			images++;
			bytes += (long)want.size();
		}
	printf("ok %ld %ld\n", images, bytes);
	return 0;
}
