// vp8_png.cu - m09 on the device: RGB24 images in HBM -> the PNG files the reference's yuv420_write_png_fd emits
// (yuv2rgb_png.c:208-364: stored deflate, filter 0), checksums included, so that the host only moves bytes.
// Arithmetic in vp8_png.cuh; two launches per batch:
//   vp8_png_frame   grid = sum over images of ceil(file_len / 65536) CTAs of 256 threads; a CTA covers 64 KB of one file in
//                   16 rounds of 4 KB, thread t takes the t-th 16-byte segment of every round (coalesced 16-byte stores,
//                   RGB read as aligned words and funnel-shifted). HBM-bound copy: file_len bytes read + written per image.
//   vp8_png_finish  one thread per image: folds the accumulators into the Adler-32 and the IDAT CRC, writes them and IEND.
#include <cuda_runtime.h>
#include <stdint.h>
#include "vp8_dev.h"
#include "vp8_png.cuh"

namespace {

using namespace pngk;

__device__ __forceinline__ int image_of_cta(const Vp8PngDesc* __restrict__ descs, int n, uint32_t cta) {
	int lo = 0, hi = n - 1; // last image whose first_cta <= cta
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (descs[mid].first_cta <= cta) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(kThreads) vp8_png_frame(const Vp8PngDesc* __restrict__ descs, int n_images, const Tables* __restrict__ tables,
                                                          Accum* __restrict__ accum) {
	__shared__ Tables st;
	__shared__ uint32_t red_crc[kThreads / 32];
	__shared__ uint32_t red_a[kThreads / 32];
	__shared__ unsigned long long red_b[kThreads / 32];
	__shared__ uint8_t head[44];
	const uint32_t t = threadIdx.x;
	const int img = image_of_cta(descs, n_images, blockIdx.x);
	const Vp8PngDesc& d = descs[img];
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
		uint32_t* dst = reinterpret_cast<uint32_t*>(&st);
		for (uint32_t i = t; i < sizeof(Tables) / 4; i += kThreads) dst[i] = src[i];
		if (t < 44) head[t] = d.head[t];
	}
	__syncthreads();
	const Geom g = geom(d.width, d.height);
	const uint32_t span_base = (blockIdx.x - d.first_cta) * kSpan;
	const uint32_t write_end = (g.file_len + 15) & ~15u;
	ThreadAcc acc{0, 0, 0, -1};
#pragma unroll 1
	for (uint32_t it = 0; it < kRounds; it++) {
		const uint32_t s = it * kThreads + t, f = span_base + s * kSeg;
		if (f >= write_end) break;
		uint32_t o[4];
		segment(g, &st, head, d.rgb, f, (int)s, o, acc);
		__stcs(reinterpret_cast<uint4*>(d.out + f), make_uint4(o[0], o[1], o[2], o[3])); // read by nobody on the device
	}
	// thread -> end of the span's covered part
	const uint32_t nseg = span_crc_segments(g, span_base);
	uint32_t crc = 0;
	if (acc.last_seg >= 0) crc = mulmod(acc.crc, st.xp16[nseg - 1 - (uint32_t)acc.last_seg]);
	crc = __reduce_xor_sync(0xFFFFFFFFu, crc);
	uint32_t a = __reduce_add_sync(0xFFFFFFFFu, acc.a);
	unsigned long long b = acc.b;
	for (int off = 16; off; off >>= 1) b += __shfl_xor_sync(0xFFFFFFFFu, b, off);
	if ((t & 31) == 0) red_crc[t >> 5] = crc, red_a[t >> 5] = a, red_b[t >> 5] = b;
	__syncthreads();
	if (t < 32) {
		crc = t < kThreads / 32 ? red_crc[t] : 0;
		a = t < kThreads / 32 ? red_a[t] : 0;
		b = t < kThreads / 32 ? red_b[t] : 0;
		crc = __reduce_xor_sync(0xFFFFFFFFu, crc);
		a = __reduce_add_sync(0xFFFFFFFFu, a);
		for (int off = 4; off; off >>= 1) b += __shfl_xor_sync(0xFFFFFFFFu, b, off);
		// span -> end of the main kernel's region: x^(8 * 16 * n16) as a product tree over the lanes
		const uint32_t n16 = nseg ? (g.crc_end - (span_base + nseg * kSeg)) / kSeg : 0;
		if (n16) {
			uint32_t m = ((n16 >> t) & 1) ? st.x2n[t] : kOne;
			for (int off = 16; off; off >>= 1) m = mulmod(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
			crc = mulmod(crc, m);
		}
		if (t == 0) {
			if (crc) atomicXor(&accum[img].crc, crc);
			if (a) atomicAdd(&accum[img].a, (unsigned long long)a);
			if (b) atomicAdd(&accum[img].b, b);
		}
	}
}

__global__ void vp8_png_finish(const Vp8PngDesc* __restrict__ descs, int n_images, const Accum* __restrict__ accum) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_images) return;
	const Vp8PngDesc& d = descs[i];
	finish(geom(d.width, d.height), d.crc_init, accum[i], d.out);
}

} // namespace

// descs_dev[i].first_cta must be the running sum of vp8_png_ctas over the images before i; total_ctas their sum.
// accum_dev: n_images * vp8_png_accum_bytes() bytes, zeroed here. tables_dev: vp8_png_tables() copied to the device.
int vp8_launch_png(const Vp8PngDesc* descs_dev, int n_images, uint32_t total_ctas, const void* tables_dev, void* accum_dev, void* stream) {
	cudaStream_t st = (cudaStream_t)stream;
	cudaError_t e = cudaMemsetAsync(accum_dev, 0, sizeof(Accum) * (size_t)n_images, st);
	if (e != cudaSuccess) return (int)e;
	vp8_png_frame<<<total_ctas, kThreads, 0, st>>>(descs_dev, n_images, (const Tables*)tables_dev, (Accum*)accum_dev);
	if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
	vp8_png_finish<<<(n_images + 127) / 128, 128, 0, st>>>(descs_dev, n_images, (const Accum*)accum_dev);
	return (int)cudaGetLastError();
}

uint32_t vp8_png_ctas(uint32_t width, uint32_t height) { return ctas_of(geom(width, height)); }
size_t vp8_png_file_bytes(uint32_t width, uint32_t height) { return geom(width, height).file_len; }
size_t vp8_png_accum_bytes(void) { return sizeof(Accum); }
size_t vp8_png_tables_bytes(void) { return sizeof(Tables); }
void vp8_png_tables(void* dst) { build_tables(*(Tables*)dst); }
void vp8_png_fill_desc(Vp8PngDesc* d, const void* tables) {
	const Geom g = geom(d->width, d->height);
	build_head(g, *(const Tables*)tables, d->head);
	d->crc_init = crc_init_term(g);
}
