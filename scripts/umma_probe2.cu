// Development probe 2: K = 16 (two kind::tf32 MMAs of K = 8) with
//   A: K-major SW64 (64 B rows)  or  MN-major "128B base-32B" with LBO = 2048 (16 k-rows per 32-row group)
//   B: K-major SW64 (64 B rows, 8-row atoms 512 B apart)
// to pin the descriptors used by the KB = 16 pipeline.  Not product code.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
	return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFF) << 32) |
			(1ull << 46) | ((uint64_t) lt << 61);
}
__device__ __forceinline__ uint32_t sw64(uint32_t off) { return off ^ (((off >> 7) & 3) << 4); }

constexpr int NN = 64;
// amode 0: A K-major SW64; 1: A MN-major base32B LBO=2048.  A[m][k] = (m%7) + 0.5k, B[n][k] = (n%5) - 0.25k
__global__ void probe(int amode, float* out) {
	extern __shared__ __align__(1024) uint8_t raw[];
	uint8_t* smem = (uint8_t*) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
	uint8_t* sa = smem;             // 8 KB
	uint8_t* sb = smem + 8192;      // NN * 64 B
	uint64_t* bar = (uint64_t*) (smem + 8192 + 4096);
	uint32_t* slot = (uint32_t*) (bar + 1);
	const int tid = threadIdx.x;
	for (int i = tid; i < (8192 + 4096) / 4; i += blockDim.x) ((float*) smem)[i] = 0.f;
	__syncthreads();
	for (int i = tid; i < 128 * 16; i += blockDim.x) {
		int m = i / 16, k = i % 16;
		float v = (float) (m % 7) + 0.5f * k;
		uint32_t off;
		if (amode == 0) off = sw64(m * 64 + k * 4);
		else { off = (m / 32) * 2048 + k * 128 + (m % 32) * 4; off ^= ((off >> 7) & 3) << 5; }
		*(float*) (sa + off) = v;
	}
	for (int i = tid; i < NN * 16; i += blockDim.x) {
		int n = i / 16, k = i % 16;
		*(float*) (sb + sw64(n * 64 + k * 4)) = (float) (n % 5) - 0.25f * k;
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (tid < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(128u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tbase = *slot;
	if (amode == 2 && tid < 128) {
		// A in TMEM: lane = m, columns 64..79 = k; each thread writes the 16 k values of its row
		const int m = tid;
		uint32_t r[16];
		for (int k = 0; k < 16; ++k) r[k] = __float_as_uint((float) (m % 7) + 0.5f * k);
		uint32_t taddr = tbase + ((uint32_t) (32 * (tid >> 5)) << 16) + 64;
		asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
				"{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
				:: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
				   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (tid == 0 && amode == 2) {
		uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (NN >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
		for (int ks = 0; ks < 2; ++ks) {
			uint64_t db = make_desc(smem_u32(sb) + ks * 32, 16, 512, 4);
			asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
					"tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
					:: "r"(tbase), "r"(tbase + 64 + 8 * ks), "l"(db), "r"(idesc), "r"((uint32_t) ks) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
	}
	if (tid == 0 && amode != 2) {
		uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (NN >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
		if (amode == 1) idesc |= (1u << 15);
		for (int ks = 0; ks < 2; ++ks) {
			uint64_t da = amode == 0 ? make_desc(smem_u32(sa) + ks * 32, 16, 512, 4) : make_desc(smem_u32(sa) + ks * 1024, 2048, 512, 1);
			uint64_t db = make_desc(smem_u32(sb) + ks * 32, 16, 512, 4);
			asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
					"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
					:: "r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t) ks) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
	}
	asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
			:: "r"(smem_u32(bar)), "r"(0u) : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (tid < 128) {
		const int warp = tid >> 5, lane = tid & 31;
		for (int c0 = 0; c0 < NN; c0 += 16) {
			uint32_t r[16];
			uint32_t taddr = tbase + ((uint32_t) (32 * warp) << 16) + c0;
			asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
					"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
					: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
					  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
					: "r"(taddr) : "memory");
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
			for (int j = 0; j < 16; ++j) out[(32 * warp + lane) * NN + c0 + j] = __uint_as_float(r[j]);
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(128u) : "memory");
}

int main() {
	float* d; cudaMalloc(&d, 128 * NN * 4);
	static float h[128 * NN];
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
	const char* names[] = { "A K-major SW64 + B K-major SW64, K=16", "A MN-major base32B LBO=2048 + B K-major SW64, K=16", "A in TMEM (lane=m, col=k) + B K-major SW64, K=16" };
	for (int amode = 0; amode < 3; ++amode) {
		cudaMemset(d, 0xff, 128 * NN * 4);
		probe<<<1, 128, 32 * 1024, 0>>>(amode, d);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
		double maxerr = 0, maxabs = 0;
		for (int m = 0; m < 128; ++m) for (int n = 0; n < NN; ++n) {
			double ref = 0;
			for (int k = 0; k < 16; ++k) ref += ((m % 7) + 0.5 * k) * ((n % 5) - 0.25 * k);
			maxerr = fmax(maxerr, fabs(ref - h[m * NN + n])); maxabs = fmax(maxabs, fabs(h[m * NN + n]));
		}
		printf("%-55s err=%s maxerr=%g maxabs=%g D[0][0..3]=%g %g %g %g D[33][1]=%g\n", names[amode], cudaGetErrorString(e),
				maxerr, maxabs, h[0], h[1], h[2], h[3], h[33 * NN + 1]);
	}
	return 0;
}
