// rgb_check.cpp - host check of the m08 tile arithmetic (webp-decoder_b200/csrc/vp8_rgb.cuh, host flavour) against the
// oracle's orc_i420_to_rgb (oracle/vp8_oracle.c, pinned against the reference's yuv2rgb_ppm.c). Built and run by
// tests/test_host.py; links oracle/liboracle.so. Prints "ok <images> <pixels>" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include "../../webp-decoder_b200/csrc/vp8_rgb.cuh"

extern "C" void orc_i420_to_rgb(const uint8_t* y, const uint8_t* u, const uint8_t* v, uint32_t width, uint32_t height,
                                uint32_t stride_y, uint32_t stride_uv, uint8_t* rgb);

static uint32_t rng_state = 12345;
static uint32_t rnd() { return rng_state = rng_state * 1664525u + 1013904223u; }

// the byte-access loader of the kernel (vp8_rgb.cu, !in_vec branch), restated for the host
static void run_image(uint32_t w, uint32_t h, int flavour, long& pixels) {
	const uint32_t cw = (w + 1) / 2, ch = (h + 1) / 2;
	std::vector<uint8_t> y(w * h), u(cw * ch), v(cw * ch), want(w * h * 3), got(w * h * 3, 0xA5);
	auto fill = [&](std::vector<uint8_t>& p) {
		for (auto& b : p) {
			const uint32_t r = rnd();
			b = flavour == 0 ? (uint8_t)(r >> 24) : flavour == 1 ? (uint8_t)((r >> 24) < 128 ? 0 : 255) : (uint8_t)(((r >> 28) & 1) ? (r >> 16) : 128 + ((r >> 20) & 7));
		}
	};
	fill(y); fill(u); fill(v);
	orc_i420_to_rgb(y.data(), u.data(), v.data(), w, h, w, cw, want.data());
	const uint32_t tiles_x = (w + 15) / 16, pairs = h / 2 + 1;
	for (uint32_t p = 0; p < pairs; p++)
		for (uint32_t tx = 0; tx < tiles_x; tx++) {
			const uint32_t px0 = tx * 16, j0 = px0 / 2;
			const bool has_top = p > 0, has_bot = 2 * p < h;
			const uint32_t ra = p > 0 ? p - 1 : 0, rb = std::min(p, ch - 1);
			const uint8_t* yt = y.data() + (size_t)(has_top ? 2 * p - 1 : 0) * w;
			const uint8_t* yb = y.data() + (size_t)(has_bot ? 2 * p : h - 1) * w;
			rgbk::Tile t;
			for (int q = 0; q < 4; q++) {
				uint32_t a = 0, b = 0;
				for (int k = 0; k < 4; k++) {
					const uint32_t x = std::min(px0 + 4 * q + k, w - 1);
					a |= (uint32_t)yt[x] << (8 * k);
					b |= (uint32_t)yb[x] << (8 * k);
				}
				t.ya[q] = a; t.yb[q] = b;
			}
			for (int i = 0; i < rgbk::kTileCols; i++) {
				const uint32_t c = (uint32_t)std::min(std::max((int)j0 - 1 + i, 0), (int)cw - 1);
				t.ca[i] = u[ra * cw + c] | ((uint32_t)v[ra * cw + c] << 16);
				t.cb[i] = u[rb * cw + c] | ((uint32_t)v[rb * cw + c] << 16);
			}
			uint32_t top[12], bot[12];
			rgbk::tile_rgb(t, top, bot);
			for (int row = 0; row < 2; row++) {
				if (!(row ? has_bot : has_top)) continue;
				const uint32_t* o = row ? bot : top;
				uint8_t* dst = got.data() + ((size_t)(2 * p - 1 + row) * w + px0) * 3;
				const uint32_t n = std::min(16u, w - px0) * 3;
				for (uint32_t k = 0; k < n; k++) dst[k] = (uint8_t)(o[k / 4] >> (8 * (k % 4)));
			}
		}
	for (size_t i = 0; i < want.size(); i++)
		if (want[i] != got[i]) {
			printf("mismatch %ux%u flavour %d at pixel (%zu, %zu) channel %zu: want %u got %u\n", w, h, flavour, (i / 3) % w, (i / 3) / w, i % 3, want[i], got[i]);
			exit(1);
		}
	pixels += (long)w * h;
}

int main() {
	// the clamped-index form of the reference's edge cases: (3a + b + 2) >> 2 == (a + ((a + b + 2) >> 1)) >> 1
	for (int a = 0; a < 256; a++)
		for (int b = 0; b < 256; b++)
			if (((3 * a + b + 2) >> 2) != ((a + ((a + b + 2) >> 1)) >> 1)) {
				printf("edge identity fails for %d %d\n", a, b);
				return 1;
			}
	long pixels = 0;
	int images = 0;
	const uint32_t sizes[][2] = {{1, 1}, {2, 2}, {1, 7}, {7, 1}, {3, 5}, {16, 16}, {17, 9}, {18, 10}, {31, 32}, {32, 31}, {33, 33}, {64, 7},
	                             {15, 2}, {16, 2}, {47, 48}, {48, 47}, {129, 129}, {200, 120}, {255, 3}, {1000, 70}, {1920, 34}};
	for (auto& s : sizes)
		for (int f = 0; f < 3; f++) {
			run_image(s[0], s[1], f, pixels);
			images++;
		}
	for (int i = 0; i < 200; i++) {
		run_image(1 + rnd() % 97, 1 + (rnd() >> 8) % 61, (int)(rnd() >> 12) % 3, pixels);
		images++;
	}
	printf("ok %d %ld\n", images, pixels);
	return 0;
}
