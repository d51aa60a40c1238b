// mma_stress.cu -- how the dense tcgen05.mma kind::tf32 rate (scripts/tf32_peak.cu: one resident tile, one commit) reacts to what the
// real kernels add around the MMAs.  cta_group::1, TS issue (A from tensor memory), M = 128, N = 256, K = 8, all SMs:
//   bit 0: a tcgen05.commit (to a barrier nobody waits on) after every 12 MMAs, as the kernels release a stage per k-block
//   bit 1: B rotates over four 32 KB tiles and A over four tensor-memory slots (addresses change every k-block)
//   bit 2: four other warps stream shared memory beside the MMAs (each thread: 8 x ld.shared.v4 + 8 x st.shared.v4 per round on a
//          private 64 KB region: what converters and TMA writes add)
//   bit 3: the same four warps also write tensor memory (tcgen05.st of 64 columns per round, the converters' stores)
// Prints TFLOP/s per mode.  Development probe, not product code.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
	return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(160, 1) stress(int iters, int mode, float* sink) {
	extern __shared__ __align__(1024) uint8_t raw[];
	uint8_t* smem = (uint8_t*) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
	uint8_t* sb = smem;                       // 4 x 32 KB B tiles
	uint8_t* traffic = smem + 131072;         // 64 KB for the background warps
	uint64_t* bar = (uint64_t*) (smem + 131072 + 65536);
	uint64_t* dummy = bar + 1;
	uint32_t* slot = (uint32_t*) (bar + 2);
	volatile int* stop = (volatile int*) (bar + 3);
	const int tid = threadIdx.x, warp = tid >> 5;
	for (int i = tid; i < 131072 / 4; i += blockDim.x) ((float*) sb)[i] = (float) ((i * 2654435761u) >> 8 & 0xFFFF) * (1.0f / 65536.0f) - 0.5f;
	for (int i = tid; i < 65536 / 4; i += blockDim.x) ((float*) traffic)[i] = 1.0f;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(dummy)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		*stop = 0;
	}
	asm volatile("fence.proxy.async;" ::: "memory");
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tbase = *slot;
	if (warp >= 1) {
		// fill the A slots (columns 256..511) with finite data
		uint32_t r[16];
		for (int k = 0; k < 16; ++k) r[k] = __float_as_uint(0.001f * (float) (tid + k));
		const uint32_t taddr = tbase + ((uint32_t) (32 * ((warp - 1) & 3)) << 16) + 256;
		for (int c = 0; c < 256; c += 16)
			asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
					:: "r"(taddr + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
					   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (tid == 0) {
		const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (256 >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
		for (int it = 0; it < iters; ++it) {
			const int rot = (mode & 2) ? (it & 3) : 0;
			const uint64_t db = make_desc(smem_u32(sb + rot * 32768));
			const uint32_t ta = tbase + 256 + (uint32_t) (rot * 64);
			#pragma unroll
			for (int i = 0; i < 12; ++i) {
				const uint32_t acc = (it | i) != 0 ? 1u : 0u;
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
						:: "r"(tbase), "r"(ta + 8 * (i & 3)), "l"(db + 2 * (i & 3) + ((i >> 2) == 1 ? 1024 : 0)), "r"(idesc), "r"(acc) : "memory");
			}
			if (mode & 1)
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(dummy)) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
		asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
				:: "r"(smem_u32(bar)), "r"(0u) : "memory");
		*stop = 1;
	} else if (warp >= 1 && (mode & 12)) {
		// background traffic until the MMA thread is done
		const uint32_t base = smem_u32(traffic) + (uint32_t) (tid - 32) * 16u;
		float acc = 0.f;
		uint32_t r[16];
		for (int k = 0; k < 16; ++k) r[k] = tid + k;
		const uint32_t taddr = tbase + ((uint32_t) (32 * ((warp - 1) & 3)) << 16) + 448;   // a slot the MMAs do not read
		while (!*stop) {
			if (mode & 4) {
				#pragma unroll
				for (int u = 0; u < 8; ++u) {
					float4 v;
					asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(base + u * 2048u) : "memory");
					acc += v.x;
					asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(base + 16384u + u * 2048u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
				}
			}
			if (mode & 8) {
				for (int c = 0; c < 64; c += 16)
					asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
							:: "r"(taddr + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
							   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
				asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
			}
		}
		if (acc == 123.456f) sink[tid] = acc;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512u) : "memory");
}

int main() {
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
	const int sms = prop.multiProcessorCount, iters = 1400;
	float* sink; cudaMalloc(&sink, 4096);
	const int smem = 131072 + 65536 + 2048;
	cudaFuncSetAttribute(stress, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	printf("{\"shape\": \"M=128 N=256 K=8 TS, 12 MMAs per round, %d rounds per SM\"", iters);
	for (int mode = 0; mode < 16; ++mode) {
		if (mode == 3 || mode == 5 || mode == 9 || mode == 10 || mode == 11 || mode == 13) continue;
		stress<<<sms, 160, smem>>>(iters, mode, sink);
		cudaDeviceSynchronize();
		float best = 1e30f;
		for (int rep = 0; rep < 5; ++rep) {
			cudaEventRecord(e0);
			stress<<<sms, 160, smem>>>(iters, mode, sink);
			cudaEventRecord(e1);
			cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			if (ms < best) best = ms;
		}
		printf(", \"mode_%d\": %.1f", mode, 2.0 * 128 * 256 * 8 * 12 * (double) iters * sms / best * 1e-9);
		fflush(stdout);
	}
	printf(", \"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
