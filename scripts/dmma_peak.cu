// Measures the FP64 tensor-core (mma.sync m8n8k4 f64) rate of the device next to the DFMA rate (dfma_peak.cu).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dmma_chain(double* out, int iters, double a, double b) {
	double c[8][2];
	#pragma unroll
	for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; }
	for (int it = 0; it < iters; ++it) {
		#pragma unroll
		for (int i = 0; i < 8; ++i)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
					: "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
	}
	double s = 0;
	#pragma unroll
	for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
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
	dmma_chain<<<blocks, 256>>>(out, 1000, 1e-3, 1e-3);
	cudaDeviceSynchronize();
	float best = 1e30f;
	for (int rep = 0; rep < 5; ++rep) {
		cudaEventRecord(e0);
		dmma_chain<<<blocks, 256>>>(out, iters, 1e-3, 1e-3);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		if (ms < best) best = ms;
	}
	// one mma.m8n8k4 per warp = 8*8*4 FMA = 512 FLOP
	const double flops = 512.0 * 8 * iters * (double) blocks * 8;
	printf("{\"dmma_m8n8k4_tflops\": %.2f, \"ms\": %.3f}\n", flops / best * 1e-9, best);
	return 0;
}
