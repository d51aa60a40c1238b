// tf32_peak.cu -- the roofline denominator of the float kernels: the dense tcgen05.mma kind::tf32 rate of this
// device, measured (MEASURED_PEAKS.json holds bf16 and HBM figures only).  Four issue modes, each first VERIFIED on
// exactly representable data against a host product and then timed on uniform(-1, 1) data:
//   1sm_ss : cta_group::1, M = 128, N = 256, K = 8, A and B from 128B-swizzled shared memory
//   1sm_ts : cta_group::1, A from tensor memory (what conv_tc.cu's kernels issue)
//   2sm_ss : cta_group::2, M = 256 (128 rows per CTA of the pair), N = 256 (128 rows of B per CTA)
//   2sm_ts : cta_group::2, A from each CTA's tensor memory
// One elected thread per CTA (per pair) issues `iters` k-blocks of four MMAs on one resident tile and commits once;
// one CTA per SM, all SMs.  burst = best of 7 launches of ~1 ms after 3 s of idling, sustained = 1.5 s of back-to-back
// launches (the power cap pulls the SM clock down under a continuous tensor load).  Prints one JSON line.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {   // K-major, 128 B rows, SWIZZLE_128B, 8-row atoms 1024 B apart
	return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7) << 4); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
			:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// exactly representable test data (integers of magnitude <= 4): any error is a layout error, not rounding
__host__ __device__ inline float a_val(int m, int k, int verify, uint32_t salt) {
	if (verify) return (float) ((m * 3 + k) % 7 - 3);
	uint32_t h = (uint32_t) (m * 131 + k * 7919) * 2654435761u + salt; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
	return (float) (h & 0xFFFFFF) * (2.0f / 16777216.0f) - 1.0f;
}
__host__ __device__ inline float b_val(int n, int k, int verify, uint32_t salt) {
	if (verify) return (float) ((n * 5 + k * 2) % 9 - 4);
	return a_val(n + 1000, k, 0, salt ^ 0x9e3779b9u);
}

constexpr int KBLK = 32;   // one resident k-block: 32 fp32 = one 128-byte swizzled row, four K = 8 MMAs

template<int CTAS, bool TS>
__global__ void __launch_bounds__(128, 1) tf32_mma_kernel(int n_cols, int iters, int verify, float* out) {
	extern __shared__ __align__(1024) uint8_t raw[];
	uint8_t* smem = (uint8_t*) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
	uint8_t* sa = smem;                       // 128 rows x 128 B (SS modes)
	uint8_t* sb = smem + 16384;               // n_cols / CTAS rows x 128 B
	uint64_t* bar = (uint64_t*) (smem + 16384 + 32768);
	uint32_t* slot = (uint32_t*) (bar + 1);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t rank = CTAS == 2 ? cluster_rank() : 0u;
	const int b_rows = n_cols / CTAS;
	const int m0 = (int) rank * 128, n0 = (int) rank * b_rows;   // this CTA's rows of A (and D) and of B
	for (int i = tid; i < 128 * KBLK; i += 128) {
		const int m = i / KBLK, k = i % KBLK;
		*(float*) (sa + swz(m * 128 + k * 4)) = a_val(m0 + m, k, verify, blockIdx.x);
	}
	for (int i = tid; i < b_rows * KBLK; i += 128) {
		const int n = i / KBLK, k = i % KBLK;
		*(float*) (sb + swz(n * 128 + k * 4)) = b_val(n0 + n, k, verify, blockIdx.x);
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core's async proxy
	if (warp == 0) {
		if (CTAS == 2) {
			asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(512u) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
		} else {
			asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(512u) : "memory");
			asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tbase = *slot;
	const uint32_t a_col = 256;   // TS modes: A in TMEM columns [256, 288), lane = row
	if (TS) {
		uint32_t r[KBLK];
		#pragma unroll
		for (int k = 0; k < KBLK; ++k) r[k] = __float_as_uint(a_val(m0 + tid, k, verify, blockIdx.x));
		const uint32_t taddr = tbase + ((uint32_t) (32 * warp) << 16) + a_col;
		#pragma unroll
		for (int h = 0; h < 2; ++h)
			asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
					:: "r"(taddr + 16 * h), "r"(r[16 * h + 0]), "r"(r[16 * h + 1]), "r"(r[16 * h + 2]), "r"(r[16 * h + 3]), "r"(r[16 * h + 4]),
					   "r"(r[16 * h + 5]), "r"(r[16 * h + 6]), "r"(r[16 * h + 7]), "r"(r[16 * h + 8]), "r"(r[16 * h + 9]), "r"(r[16 * h + 10]),
					   "r"(r[16 * h + 11]), "r"(r[16 * h + 12]), "r"(r[16 * h + 13]), "r"(r[16 * h + 14]), "r"(r[16 * h + 15]) : "memory");
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	if (CTAS == 2) cluster_sync(); else __syncthreads();   // both CTAs' operands (and barriers) are in place
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	if (tid == 0 && rank == 0) {
		const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (n_cols >> 3) << 17) | ((uint32_t) ((128 * CTAS) >> 4) << 24);
		const uint64_t da = make_desc(smem_u32(sa)), db = make_desc(smem_u32(sb));
		for (int it = 0; it < iters; ++it) {
			#pragma unroll
			for (int ks = 0; ks < KBLK / 8; ++ks) {
				const uint32_t acc = (it | ks) != 0 ? 1u : 0u;
				if (CTAS == 2) {
					if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
							:: "r"(tbase), "r"(tbase + a_col + 8 * ks), "l"(db + 2 * ks), "r"(idesc), "r"(acc) : "memory");
					else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
							:: "r"(tbase), "l"(da + 2 * ks), "l"(db + 2 * ks), "r"(idesc), "r"(acc) : "memory");
				} else {
					if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
							:: "r"(tbase), "r"(tbase + a_col + 8 * ks), "l"(db + 2 * ks), "r"(idesc), "r"(acc) : "memory");
					else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
							:: "r"(tbase), "l"(da + 2 * ks), "l"(db + 2 * ks), "r"(idesc), "r"(acc) : "memory");
				}
			}
		}
		if (CTAS == 2)
			asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
					:: "r"(smem_u32(bar)), "h"((uint16_t) 3) : "memory");
		else
			asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
	}
	mbar_wait(bar, 0);
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (out) {
		// D: lane = row of this CTA's 128, column = n; out[(m0 + row) * n_cols + n]
		for (int c0 = 0; c0 < n_cols; c0 += 16) {
			uint32_t r[16];
			const uint32_t taddr = tbase + ((uint32_t) (32 * warp) << 16) + c0;
			asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
					: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
					  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			if (blockIdx.x < CTAS)
				for (int i = 0; i < 16; ++i) out[(long long) (m0 + 32 * warp + lane) * n_cols + c0 + i] = __uint_as_float(r[i]);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	if (CTAS == 2) cluster_sync(); else __syncthreads();
	if (warp == 0) {
		if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512u) : "memory");
		else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512u) : "memory");
	}
}

template<int CTAS, bool TS>
static int launch(int grid, int n_cols, int iters, int verify, float* out) {
	auto kern = tf32_mma_kernel<CTAS, TS>;
	const int smem = 16384 + 32768 + 1024 + 1024 + 68 * 1024;   // more than half an SM's shared memory: one CTA per SM
	cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return (int) cudaLaunchKernelEx(&cfg, kern, n_cols, iters, verify, out);
}

typedef int (*LaunchFn)(int, int, int, int, float*);

int main(int argc, char** argv) {
	// usage: tf32_peak [sustained seconds = 1.5] [mode filter, e.g. 1sm_ts] [N = 256]
	const double sustain_s = argc > 1 ? atof(argv[1]) : 1.5;
	const char* only = argc > 2 ? argv[2] : nullptr;
	const int n_arg = argc > 3 ? atoi(argv[3]) : 256;   // accumulator columns (a multiple of 32 up to 256)
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	const int sms = prop.multiProcessorCount;
	const int N = n_arg;
	float* d_out; cudaMalloc(&d_out, 256 * 256 * 4);
	static float h[256 * 256];
	struct Mode { const char* name; LaunchFn fn; int ctas; } modes[] = {
		{"1sm_ss", launch<1, false>, 1}, {"1sm_ts", launch<1, true>, 1}, {"2sm_ss", launch<2, false>, 2}, {"2sm_ts", launch<2, true>, 2} };
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	printf("{\"device\": \"%s\", \"sms\": %d, \"shape\": \"M=128 per CTA, N=%d, K=8, kind::tf32, fp32 accumulate in TMEM\"", prop.name, sms, N);
	for (Mode& md : modes) {
		if (only && !strstr(md.name, only)) continue;
		// ---- verification: one k-block of exactly representable integers
		cudaMemset(d_out, 0xff, 256 * N * 4);
		int rc = md.fn(md.ctas, N, 1, 1, d_out);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(h, d_out, 128 * md.ctas * N * 4, cudaMemcpyDeviceToHost);
		double maxerr = 0;
		for (int m = 0; m < 128 * md.ctas; ++m) for (int n = 0; n < N; ++n) {
			double ref = 0;
			for (int k = 0; k < KBLK; ++k) ref += (double) a_val(m, k, 1, 0) * b_val(n, k, 1, 0);
			const double d = fabs(ref - h[m * N + n]);
			if (!(d <= maxerr)) maxerr = d;
		}
		printf(", \"%s\": {\"launch_rc\": %d, \"sync\": \"%s\", \"verify_max_abs_err\": %g", md.name, rc, cudaGetErrorString(e), maxerr);
		if (rc != 0 || e != cudaSuccess) { printf("}"); continue; }
		// ---- timing on uniform(-1, 1) data; every mode starts from an idle, cooled-down device
		sleep(3);
		const int grid = sms / md.ctas * md.ctas;
		const int iters = 4000;   // x 4 MMAs of 128 x 256 x 8 per SM: about 1 ms
		md.fn(grid, N, iters, 0, nullptr);
		cudaDeviceSynchronize();
		float best = 1e30f;
		for (int rep = 0; rep < 7; ++rep) {
			cudaEventRecord(e0);
			md.fn(grid, N, iters, 0, nullptr);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			if (ms < best) best = ms;
		}
		const double flop = 2.0 * 128 * N * 8 * 4 * (double) iters * grid;
		const int reps = (int) (sustain_s * 1e3 / best) + 1;
		cudaEventRecord(e0);
		for (int rep = 0; rep < reps; ++rep) md.fn(grid, N, iters, 0, nullptr);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms_long; cudaEventElapsedTime(&ms_long, e0, e1);
		printf(", \"burst_tflops\": %.1f, \"sustained_tflops\": %.1f, \"ctas\": %d, \"burst_ms\": %.3f, \"sustained_s\": %.2f}",
				flop / best * 1e-9, flop * reps / ms_long * 1e-9, grid, best, ms_long * 1e-3);
	}
	printf("}\n");
	return 0;
}
