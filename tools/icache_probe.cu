// icache_probe.cu - how much integer issue rate does a B200 SM keep when the loop body outgrows the instruction caches?
//
// Development probe behind DESIGN.md's "why the wavefront kernel sits at ~0.6 IPC per scheduler" section. Every warp runs
// the same straight-line integer body (4 independent chains, so a single warp could issue every cycle or two) in a loop;
// the body is 0.5 K ... 6 K SASS instructions (8 ... 96 KB). Two arrangements per size:
//   in-phase : all warps of an SM start together, so they walk the body side by side and share every fetched line
//   skewed   : each warp first burns a different delay, so the 28 warps of an SM sit at 28 different places in the body
//              (this is what independent macroblock engines do)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/icache_probe tools/icache_probe.cu && tools/icache_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int BLOCKS>
__global__ void __launch_bounds__(128) body_kernel(unsigned* out, int iters, int skew) {
	unsigned a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7, c = a ^ 0x9e3779b9u, d = b + 0x7f4a7c15u;
	const unsigned warp = (blockIdx.x * 4 + (threadIdx.x >> 5));
	if (skew) {
		// a dependent chain of a per-warp length: 0 ... ~8 K instructions of delay
		const int n = (warp * 977u) % 2048u;
		for (int i = 0; i < n; i++) a = a * 1664525u + 1013904223u;
	}
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int k = 0; k < BLOCKS; k++) {
			// 8 instructions, 4 independent chains, constants differ per block so nothing is merged
			a = a * (2 * k + 3) + b;
			c = c * (2 * k + 5) + d;
			b = (b >> 3) ^ a;
			d = (d >> 5) ^ c;
			a += c & (k + 1);
			c += a | (k + 2);
		}
	}
	if ((a ^ b ^ c ^ d) == 0x12345678u) out[threadIdx.x] = a; // keep the chains alive
}

// Same total body, but cut into PHASES pieces with a CTA-wide barrier after each piece: the 28 warps of a 896-thread CTA
// walk one piece (BLOCKS / PHASES blocks) together before any of them enters the next.
template <int BLOCKS, int PHASES>
__global__ void __launch_bounds__(896) phased_kernel(unsigned* out, int iters) {
	unsigned a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 7, c = a ^ 0x9e3779b9u, d = b + 0x7f4a7c15u;
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int k = 0; k < BLOCKS; k++) {
			a = a * (2 * k + 3) + b;
			c = c * (2 * k + 5) + d;
			b = (b >> 3) ^ a;
			d = (d >> 5) ^ c;
			a += c & (k + 1);
			c += a | (k + 2);
			if ((k + 1) % (BLOCKS / PHASES) == 0) __syncthreads();
		}
	}
	if ((a ^ b ^ c ^ d) == 0x12345678u) out[threadIdx.x] = a;
}

template <int BLOCKS, int PHASES>
static void run_phased(unsigned* out, int sms, double khz) {
	const int iters = 64 * 4096 / BLOCKS;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	phased_kernel<BLOCKS, PHASES><<<sms, 896>>>(out, iters / 8);
	cudaEventRecord(e0);
	phased_kernel<BLOCKS, PHASES><<<sms, 896>>>(out, iters);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	const double blocks = (double)iters * BLOCKS * 28;
	printf("{\"blocks\": %d, \"approx_body_bytes\": %d, \"ctas_per_sm\": 1, \"arrangement\": \"one 28-warp CTA, barrier every %d bytes\", "
	       "\"ms\": %.3f, \"blocks_per_us_per_sm\": %.2f, \"blocks_per_kcycle_per_scheduler_at_max_clock\": %.2f}\n",
	       BLOCKS, 10 * BLOCKS * 16, 10 * BLOCKS * 16 / PHASES, ms, blocks / (ms * 1e3), blocks / 4 / (ms * khz) * 1e3);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
}

template <int BLOCKS>
static void run(unsigned* out, int ctas_per_sm, int sms, double khz) {
	const int iters = 64 * 4096 / BLOCKS; // same number of executed blocks for every size
	for (int skew = 0; skew < 2; skew++) {
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0);
		cudaEventCreate(&e1);
		body_kernel<BLOCKS><<<sms * ctas_per_sm, 128>>>(out, iters / 8, skew);
		cudaEventRecord(e0);
		body_kernel<BLOCKS><<<sms * ctas_per_sm, 128>>>(out, iters, skew);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		const double blocks = (double)iters * BLOCKS * 4 * ctas_per_sm; // warp-level blocks per SM
		printf("{\"blocks\": %d, \"approx_body_bytes\": %d, \"ctas_per_sm\": %d, \"arrangement\": \"%s\", \"ms\": %.3f, "
		       "\"blocks_per_us_per_sm\": %.2f, \"blocks_per_kcycle_per_scheduler_at_max_clock\": %.2f}\n",
		       BLOCKS, (10 * BLOCKS) * 16, ctas_per_sm, skew ? "skewed" : "in-phase", ms, blocks / (ms * 1e3),
		       blocks / 4 / (ms * khz) * 1e3);
		cudaEventDestroy(e0);
		cudaEventDestroy(e1);
	}
}

int main(int argc, char** argv) {
	int dev = 0;
	cudaSetDevice(dev);
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, dev);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	unsigned* out;
	cudaMalloc(&out, 4096);
	printf("# %s, %d SMs, max clock %d kHz; a block is 6 statements = 10 SASS instructions of 16 bytes (cuobjdump)\n", p.name,
	       p.multiProcessorCount, khz);
	const int sms = p.multiProcessorCount;
	for (int cps : {1, 4, 7}) {
		run<32>(out, cps, sms, khz);
		run<128>(out, cps, sms, khz);
		run<256>(out, cps, sms, khz);
		run<384>(out, cps, sms, khz);
		run<512>(out, cps, sms, khz);
		run<768>(out, cps, sms, khz);
	}
	run_phased<384, 1>(out, sms, khz);
	run_phased<384, 3>(out, sms, khz);
	run_phased<384, 6>(out, sms, khz);
	run_phased<384, 12>(out, sms, khz);
	run_phased<768, 6>(out, sms, khz);
	run_phased<768, 12>(out, sms, khz);
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) {
		fprintf(stderr, "cuda error: %s\n", cudaGetErrorString(e));
		return 1;
	}
	return 0;
}
