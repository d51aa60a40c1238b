// Development probe: one tcgen05.mma kind::tf32 (M=128, N=32, K=8) from hand-filled shared memory,
// to pin down the smem-descriptor semantics for K-major and MN-major operands.  Not product code.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt = 2) {
	return (uint64_t) ((addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t) ((sbo >> 4) & 0x3FFF) << 32) |
			(1ull << 46) | ((uint64_t) lt << 61);
}
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7) << 4); }

// mode 0: A K-major; mode 1: A MN-major.  B always K-major.  A[m][k] = (m % 7) + 0.5 * k ; B[n][k] = (n % 5) - k
__global__ void probe(int mode, int with_idesc_major, float* out, int variant, uint32_t lt, uint32_t lbo, uint32_t sbo) {
	extern __shared__ __align__(1024) uint8_t raw[];
	uint8_t* smem = (uint8_t*) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
	uint8_t* sa = smem;             // 16 KB region
	uint8_t* sb = smem + 16384;     // 4 KB region
	uint64_t* bar = (uint64_t*) (smem + 16384 + 4096);
	uint32_t* slot = (uint32_t*) (bar + 1);
	const int tid = threadIdx.x;
	for (int i = tid; i < (16384 + 4096) / 4; i += blockDim.x) ((float*) smem)[i] = 0.f;
	__syncthreads();
	for (int i = tid; i < 128 * 8; i += blockDim.x) {
		int m = i / 8, k = i % 8;
		float v = (float) (m % 7) + 0.5f * k;
		uint32_t off;
		if (mode == 0) off = m * 128 + k * 4;                         // K-major rows of 128 B
		else if (variant == 0) { off = swz((m / 32) * 4096 + k * 128 + (m % 32) * 4); }   // SW128, 8 k-rows of 128 B
		else if (variant == 1) {  // SW128 base-32B: atoms of 4 k-rows x 128 B (512 B); k-atoms 512 apart inside a 4096 group
			off = (m / 32) * 4096 + (k / 4) * 512 + (k % 4) * 128 + (m % 32) * 4;
			off ^= ((off >> 7) & 3) << 5;
		} else if (variant == 2) { // SW64: 16 floats per row (64 B), 8 k-rows = 512 B atom; 8 MN groups of 16
			off = (m / 16) * 2048 + k * 64 + (m % 16) * 4;
			off ^= ((off >> 7) & 3) << 4;
		} else if (variant == 3) { // SW32: 8 floats per row (32 B), 8 k-rows = 256 B atom
			off = (m / 8) * 1024 + k * 32 + (m % 8) * 4;
			off ^= ((off >> 7) & 1) << 4;
		} else { // no swizzle: core matrix 8 k x 16 B
			off = (m / 4) * 128 + k * 16 + (m % 4) * 4;
		}
		if (mode == 0) off = swz(off);
		*(float*) (sa + off) = v;
	}
	for (int i = tid; i < 32 * 8; i += blockDim.x) {
		int n = i / 8, k = i % 8;
		*(float*) (sb + swz(n * 128 + k * 4)) = (float) (n % 5) - (float) k;
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if (tid < 32) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(32u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tbase = *slot;
	if (tid == 0) {
		uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (32 >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
		if (mode == 1 && with_idesc_major) idesc |= (1u << 15);
		uint64_t da = mode == 0 ? make_desc(smem_u32(sa), 16, 1024) : make_desc(smem_u32(sa), lbo, sbo, lt);
		uint64_t db = make_desc(smem_u32(sb), 16, 1024);
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
				:: "r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
	}
	asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
			:: "r"(smem_u32(bar)), "r"(0u) : "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (tid < 128) {
		const int warp = tid >> 5, lane = tid & 31;
		uint32_t r[32];
		uint32_t taddr = tbase + ((uint32_t) (32 * warp) << 16);
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
				"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
				"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
				: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
				  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
				  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
				  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
				: "r"(taddr) : "memory");
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		for (int j = 0; j < 32; ++j) out[(32 * warp + lane) * 32 + j] = __uint_as_float(r[j]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(32u) : "memory");
}

int main() {
	float* d; cudaMalloc(&d, 128 * 32 * 4);
	static float h[128 * 32];
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
	struct Cfg { int mode, maj, variant; uint32_t lt, lbo, sbo; const char* name; };
	Cfg cfgs[] = {
		{0, 0, 0, 2, 16, 1024, "K-major SW128"},
		{1, 1, 0, 2, 4096, 1024, "MN SW128 lbo=4096 sbo=1024"},
		{1, 1, 0, 2, 1024, 4096, "MN SW128 lbo=1024 sbo=4096 (swapped)"},
		{1, 1, 1, 1, 4096, 512, "MN SW128_BASE32B lbo=4096 sbo=512"},
		{1, 1, 1, 1, 512, 4096, "MN SW128_BASE32B lbo=512 sbo=4096"},
		{1, 1, 2, 4, 2048, 512, "MN SW64 lbo=2048 sbo=512"},
		{1, 1, 2, 4, 512, 2048, "MN SW64 lbo=512 sbo=2048"},
		{1, 1, 3, 6, 1024, 256, "MN SW32 lbo=1024 sbo=256"},
		{1, 1, 3, 6, 256, 1024, "MN SW32 lbo=256 sbo=1024"},
		{1, 1, 4, 0, 128, 128, "MN NOSWIZZLE lbo=128 sbo=128"},
		{1, 1, 4, 0, 1024, 128, "MN NOSWIZZLE lbo=1024 sbo=128"},
		{1, 1, 4, 0, 128, 1024, "MN NOSWIZZLE lbo=128 sbo=1024"},
	};
	for (Cfg& c : cfgs) {
		int mode = c.mode, maj = c.maj;
		cudaMemset(d, 0xff, 128 * 32 * 4);
		probe<<<1, 128, 32 * 1024, 0>>>(mode, maj, d, c.variant, c.lt, c.lbo, c.sbo);
		cudaError_t e = cudaDeviceSynchronize();
		cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
		double maxerr = 0, maxabs = 0;
		for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
			double ref = 0;
			for (int k = 0; k < 8; ++k) ref += ((m % 7) + 0.5 * k) * ((n % 5) - (double) k);
			maxerr = fmax(maxerr, fabs(ref - h[m * 32 + n])); maxabs = fmax(maxabs, fabs(h[m * 32 + n]));
		}
		printf("%-40s err=%s maxerr=%g maxabs=%g  D[0][0..3]=%g %g %g %g D[1][0]=%g D[33][1]=%g\n", c.name,
				cudaGetErrorString(e), maxerr, maxabs, h[0], h[1], h[2], h[3], h[32], h[33 * 32 + 1]);
	}
	return 0;
}
