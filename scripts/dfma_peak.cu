// Measures the FP64 FMA peak of the device (the roofline denominator of the double kernel layers): every thread runs
// 16 independent DFMA chains from registers.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_peak dfma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_chain(double* out, int iters, double a, double b) {
	double v[16];
	#pragma unroll
	for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3 + i;
	for (int it = 0; it < iters; ++it) {
		#pragma unroll
		for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
	}
	double s = 0;
	#pragma unroll
	for (int i = 0; i < 16; ++i) s += v[i];
	out[blockIdx.x * 256 + threadIdx.x] = s;
}

int main() {
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	const int blocks = p.multiProcessorCount * 8, iters = 20000;
	double* out;
	cudaMalloc(&out, sizeof(double) * blocks * 256);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	dfma_chain<<<blocks, 256>>>(out, 1000, 0.999999, 1e-9);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		cudaEventRecord(e0);
		dfma_chain<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	const double flops = 2.0 * 16 * iters * (double) blocks * 256;
	printf("{\"dfma_peak_tflops\": %.2f, \"sms\": %d, \"ms\": %.3f, \"dfma_per_clk_per_sm_at_1965mhz\": %.1f}\n", flops / best * 1e-9,
			p.multiProcessorCount, best, flops / 2 / (best * 1e-3) / p.multiProcessorCount / 1.965e9);
	return 0;
}
