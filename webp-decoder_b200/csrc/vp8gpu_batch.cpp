// vp8gpu_batch - batch front end over libvp8gpu.so: the reference CLI's -yuv / -yuvf / -ppm / -png sub-commands
// (reference src/main.c:556-842) for MANY files at once. Container + header + token parsing run on host threads (one
// image per thread), every pixel on the GPU (one pipelined batch), outputs are written as <out_dir>/<stem>.<ext>.
//
//   vp8gpu_batch -yuv|-yuvf|-ppm|-png <out_dir> [--device N] [--threads T] [--chunk C] file.webp ...
//
// Exit status 0 when every file decoded, 1 otherwise (one line per failed file on stderr, like the reference CLI).
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

#include "../../include/vp8_gpu.h"
#include "../../include/vp8_parse.h"

static bool read_file(const char* path, std::vector<uint8_t>& out) {
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	out.resize(n > 0 ? (size_t)n : 0);
	const bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
	fclose(f);
	return ok;
}

static bool write_file(const std::string& path, const uint8_t* p, size_t n) {
	int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
	if (fd < 0) return false;
	bool ok = true;
	while (n && ok) {
		ssize_t k = write(fd, p, n);
		if (k < 0) {
			if (errno == EINTR) continue;
			ok = false;
		} else {
			p += k;
			n -= (size_t)k;
		}
	}
	close(fd);
	return ok;
}

static std::string stem_of(const char* path) {
	std::string s(path);
	size_t slash = s.find_last_of('/');
	if (slash != std::string::npos) s = s.substr(slash + 1);
	size_t dot = s.find_last_of('.');
	if (dot != std::string::npos) s = s.substr(0, dot);
	return s;
}

int main(int argc, char** argv) {
	if (argc < 4) {
		fprintf(stderr, "usage: %s -yuv|-yuvf|-ppm|-png <out_dir> [--device N] [--threads T] [--chunk C] file.webp ...\n", argv[0]);
		return 2;
	}
	const std::string mode = argv[1], out_dir = argv[2];
	if (mode != "-yuv" && mode != "-yuvf" && mode != "-ppm" && mode != "-png") {
		fprintf(stderr, "error: unknown sub-command %s\n", mode.c_str());
		return 2;
	}
	int device = 0, threads = (int)std::thread::hardware_concurrency(), chunk = 0;
	std::vector<const char*> paths;
	for (int i = 3; i < argc; i++) {
		if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--chunk") && i + 1 < argc) chunk = atoi(argv[++i]);
		else paths.push_back(argv[i]);
	}
	if (paths.empty()) return 2;
	mkdir(out_dir.c_str(), 0755);

	// ---- read + parse on host threads (m01 + m02 + m05), arrays straight into pinned memory
	const int n_all = (int)paths.size();
	std::vector<std::vector<uint8_t>> bytes(n_all);
	std::vector<int> ok_idx;
	int failed = 0;
	for (int i = 0; i < n_all; i++) {
		uint32_t w, h;
		if (!read_file(paths[i], bytes[i]) || vp8_parse_webp_size(bytes[i].data(), bytes[i].size(), &w, &h)) {
			fprintf(stderr, "error: %s: not a supported simple lossy WebP (RIFF/WEBP + single VP8 chunk)\n", paths[i]);
			failed++;
		} else {
			ok_idx.push_back(i);
		}
	}
	const int n = (int)ok_idx.size();
	if (n == 0) return 1;
	std::vector<size_t> arena_bytes(n), arena_off(n + 1, 0);
	for (int k = 0; k < n; k++) {
		uint32_t w, h;
		vp8_parse_webp_size(bytes[ok_idx[k]].data(), bytes[ok_idx[k]].size(), &w, &h);
		arena_bytes[k] = vp8_parse_arena_bytes(w, h);
		arena_off[k + 1] = arena_off[k] + arena_bytes[k];
	}
	uint8_t* pinned = (uint8_t*)vp8_gpu_host_alloc(arena_off[n]);
	if (!pinned) {
		fprintf(stderr, "error: cannot allocate pinned staging memory (no CUDA device? this tool has no CPU fallback)\n");
		return 1;
	}
	std::vector<const uint8_t*> fptr(n);
	std::vector<size_t> fsize(n);
	std::vector<void*> arenas(n);
	for (int k = 0; k < n; k++) {
		fptr[k] = bytes[ok_idx[k]].data();
		fsize[k] = bytes[ok_idx[k]].size();
		arenas[k] = pinned + arena_off[k];
	}
	std::vector<Vp8KeyFrameHeader> kf(n);
	std::vector<Vp8DecodedFrame> fr(n);
	std::vector<int> status(n, 0);
	vp8_parse_batch(fptr.data(), fsize.data(), n, threads, kf.data(), fr.data(), arenas.data(), arena_bytes.data(), status.data());
	std::vector<const Vp8KeyFrameHeader*> kfp;
	std::vector<const Vp8DecodedFrame*> frp;
	std::vector<int> dec_idx;
	for (int k = 0; k < n; k++) {
		if (status[k]) {
			fprintf(stderr, "error: %s: VP8 macroblock/token decode failed (%s)\n", paths[ok_idx[k]], strerror(status[k]));
			failed++;
			continue;
		}
		kfp.push_back(&kf[k]);
		frp.push_back(&fr[k]);
		dec_idx.push_back(ok_idx[k]);
	}
	const int m = (int)kfp.size();
	if (m == 0) return 1;

	// ---- one pipelined GPU batch
	vp8_gpu_ctx* ctx = nullptr;
	if (vp8_gpu_init(device, nullptr, &ctx)) {
		fprintf(stderr, "error: %s\n", vp8_gpu_last_error());
		return 1;
	}
	const bool rgb = mode == "-ppm" || mode == "-png";
	const size_t cap = vp8_gpu_decode_bytes(kfp.data(), m, rgb);
	uint8_t* out = (uint8_t*)vp8_gpu_host_alloc(cap);
	std::vector<size_t> offs(m), sizes(m);
	int rc = out ? 0 : -1;
	if (!rc) {
		rc = rgb ? vp8_gpu_decode_ppm(ctx, kfp.data(), frp.data(), m, out, cap, offs.data(), sizes.data(), chunk)
		         : vp8_gpu_decode_i420(ctx, kfp.data(), frp.data(), m, mode == "-yuvf", out, cap, offs.data(), sizes.data(), chunk);
	}
	if (rc) {
		fprintf(stderr, "error: VP8 reconstruction failed: %s\n", vp8_gpu_last_error());
		return 1;
	}

	// ---- write results
	for (int k = 0; k < m; k++) {
		const std::string base = out_dir + "/" + stem_of(paths[dec_idx[k]]);
		bool ok;
		if (mode == "-png") {
			// the PNG container is framed on the host from the RGB bytes (the PPM payload), via the reference-shaped
			// writer fed with ... the RGB already computed: reuse the library's framing through a pipe-free path
			const uint32_t w = kfp[k]->width, h = kfp[k]->height;
			const uint8_t* rgbp = out + offs[k] + (sizes[k] - (size_t)w * h * 3);
			std::vector<uint8_t> png(vp8_gpu_png_bound(w, h));
			const size_t len = vp8_gpu_png_frame(rgbp, w, h, png.data());
			ok = len && write_file(base + ".png", png.data(), len);
		} else {
			ok = write_file(base + (mode == "-ppm" ? ".ppm" : ".i420"), out + offs[k], sizes[k]);
		}
		if (!ok) {
			fprintf(stderr, "error: %s: write failed\n", base.c_str());
			failed++;
		}
	}
	vp8_gpu_host_free(out);
	vp8_gpu_host_free(pinned);
	vp8_gpu_destroy(ctx);
	return failed ? 1 : 0;
}
