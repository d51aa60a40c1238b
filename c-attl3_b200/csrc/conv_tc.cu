// conv_tc.cu -- the float kernel-layer hot path on 5th-generation tensor cores:
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma kind::tf32 -> TMEM
// accumulator -> tcgen05.ld epilogue, as a persistent warp-specialised implicit GEMM with NO im2col
// buffer in HBM (the reference materialises one, C-ATTL3/layer/kernel/ConvKernelLayer.hpp:127).
//
// Precision: 3xTF32 split.  a = a_hi + a_lo with a_hi = the top 19 bits of a (what kind::tf32
// reads from an fp32 word) and a_lo = a - a_hi; D += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in an FP32
// TMEM accumulator keeps the result at FP32-GEMM accuracy ("3xTF32").
//
// Layout.  Every tensor of the reference is N-fastest (C-ATTL3/core/EigenProxy.hpp:56-57), so the
// GEMM row index m = n + N*(oh + OH*ow) is contiguous along n.  A 4-D tiled TMA box
// [32 n][1][1][KB channels] at coordinate (n0, ih, iw, c0) therefore lands as KB rows of 128 bytes:
//   * gather GEMM (forward / input gradient): that is an MN-major (M contiguous) A operand; four
//     such boxes (LBO apart) form the 128-row tile; padding comes from TMA out-of-bounds zero fill;
//   * weight gradient: the same box is a K-major operand whose reduction dimension is m.
// The output tile sits in TMEM with lane = m and column = filter, so the epilogue stores each
// column as 32 consecutive floats per warp: y(m + M*f) is written fully coalesced.
#include <cuda.h>

#include "common.cuh"

namespace cattl3 {

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra WAIT_DONE;\n\t"
		"bra WAIT_LOOP;\n\t"
		"WAIT_DONE:\n\t}"
		:: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
	uint32_t pred = 0;
	asm volatile(
		"{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
		"elect.sync rx|px, %1;\n\t"
		"@px mov.s32 %0, 1;\n\t}"
		: "+r"(pred) : "r"(0xffffffffu));
	return pred != 0;
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
	asm volatile(
		"cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
		:: "r"(smem_u32(dst)), "l"((uint64_t) tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
		:: "r"(smem_u32(dst)), "l"((uint64_t) tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
	asm volatile(
		"cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
		:: "r"(smem_u32(dst)), "l"((uint64_t) tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
	asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t) tm) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
		:: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane = thread).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
	uint32_t* r = reinterpret_cast<uint32_t*>(v);
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		: "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, float* v) {
	uint32_t* r = reinterpret_cast<uint32_t*>(v);
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x16.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		: "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address, leading /
// stride byte offsets (all >> 4), version = 1 (Blackwell), layout type (2 = SWIZZLE_128B,
// 1 = SWIZZLE_128B_BASE32B).
//
// Measured on B200 (scripts/umma_probe.cu): an MN-major kind::tf32 operand is only honoured with
// layout type 1 (128-byte rows, 32-byte swizzle atom, Swizzle<2,5,2>): K atoms of 4 rows (512 B,
// SBO apart), 32-element MN groups LBO apart; every other swizzle mode yields zeros.  TMA writes
// that layout with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  K-major operands use the ordinary
// SWIZZLE_128B (8-row atoms, SBO = 1024).
constexpr uint32_t LT_SW128 = 2, LT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
	return (uint64_t) ((smem_addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16) |
			((uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t) layout_type << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, M = 128, N runtime.
__host__ __device__ inline uint32_t make_idesc_tf32(int n, bool a_mn_major, bool b_mn_major) {
	return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
			((uint32_t) (n >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
}

// ---- operand preparation ---------------------------------------------------------------------------
// lo = rn_tf32(x - trunc_tf32(x)): the part of x the tensor core does not see when it reads the raw
// fp32 word (kind::tf32 uses the top 19 bits), rounded to nearest-even at TF32 precision so that the
// hardware's own truncation of lo is exact and the split error is unbiased (~2^-22 relative).
__device__ __forceinline__ float tf32_lo(float v) {
	const float r = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
	const uint32_t u = __float_as_uint(r);
	return __uint_as_float((u + 0xFFFu + ((u >> 13) & 1u)) & 0xFFFFE000u);
}
__global__ void __launch_bounds__(256) split_lo_kernel(long long count, const float* __restrict__ x, float* __restrict__ lo) {
	const long long nvec = count >> 2;
	const long long stride = (long long) gridDim.x * 256;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += stride) {
		float4 v = reinterpret_cast<const float4*>(x)[i];
		v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
		reinterpret_cast<float4*>(lo)[i] = v;
	}
	for (long long i = (nvec << 2) + blockIdx.x * 256ll + threadIdx.x; i < count; i += stride)
		lo[i] = tf32_lo(x[i]);
}

// Packs the weights of one gather-GEMM pass into K-major tiles [tap][j_pad][r_pad] (r contiguous),
// split into hi (truncated to TF32) and lo, zero padded.
__global__ void __launch_bounds__(256) pack_weights_kernel(GatherGeom gg, int r_pad, int j_pad, const float* __restrict__ w,
		float* __restrict__ hi, float* __restrict__ lo) {
	const int T = gg.RH * gg.RW;
	const long long total = (long long) T * j_pad * r_pad;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long) gridDim.x * 256) {
		const int r = (int) (i % r_pad);
		const int j = (int) ((i / r_pad) % j_pad);
		const int tap = (int) (i / ((long long) r_pad * j_pad));
		float v = 0.f;
		if (r < gg.SC && j < gg.J) v = w[tap * gg.w_stap + r * gg.w_sr + j * gg.w_sj];
		hi[i] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
		lo[i] = tf32_lo(v);
	}
}

// ---- the gather-GEMM kernel --------------------------------------------------------------------------
struct TcGemmParams {
	int N, SH, SW, OH, OW, J, RH, RW;
	int ah, bh, ch, aw, bw, cw;
	long long M, P;
	int m_tiles, j_tiles;
	int BN;          // filter tile (multiple of 16, <= 256)
	int kc;          // channel chunks of 32 per tap
	int stages;
	int tmem_cols;   // power of two >= 2 * BN
	int bias_mode;
	const float* bias;
	float* out;
};

constexpr int TC_BM = 128, TC_KB = 32;
constexpr int TC_A_BYTES = TC_BM * TC_KB * 4;  // 16 KB per split part

__global__ void __launch_bounds__(192, 1) tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
		const __grid_constant__ CUtensorMap tm_a_lo, const __grid_constant__ CUtensorMap tm_b_hi,
		const __grid_constant__ CUtensorMap tm_b_lo, const TcGemmParams p) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
	const int b_bytes = p.BN * TC_KB * 4;
	const int stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) p.stages * stage_bytes);
	uint64_t* full = bars;
	uint64_t* empty = bars + p.stages;
	uint64_t* acc_full = bars + 2 * p.stages;
	uint64_t* acc_empty = acc_full + 2;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int T = p.RH * p.RW;
	const int kblocks = T * p.kc;
	const int tiles = p.m_tiles * p.j_tiles;

	if (warp == 0 && elect_one()) {
		tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
		tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
		for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
		fence_barrier_init();
	}
	if (warp == 1) tmem_alloc(tmem_slot, (uint32_t) p.tmem_cols);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		// ===== TMA producer =====
		if (elect_one()) {
			int s = 0; uint32_t ph = 0;
			for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
				const int mt = tile % p.m_tiles, jt = tile / p.m_tiles;
				// the four 32-row groups of this tile: (n0, oh, ow) each, or out of range
				int gn[4], goh[4], gow[4];
				#pragma unroll
				for (int g = 0; g < 4; ++g) {
					const long long m = (long long) mt * TC_BM + 32 * g;
					if (m < p.M) {
						gn[g] = (int) (m % p.N);
						const long long pix = m / p.N;
						goh[g] = (int) (pix % p.OH);
						gow[g] = (int) (pix / p.OH);
					} else {
						gn[g] = 0; goh[g] = -0x40000; gow[g] = -0x40000;  // far out of bounds: zero fill
					}
				}
				for (int kb = 0; kb < kblocks; ++kb) {
					const int tap = kb / p.kc, c0 = (kb % p.kc) * TC_KB;
					const int rh = tap % p.RH, rw = tap / p.RH;
					mbar_wait(&empty[s], ph ^ 1);
					uint8_t* st = smem + (size_t) s * stage_bytes;
					mbar_expect_tx(&full[s], (uint32_t) stage_bytes);
					#pragma unroll
					for (int g = 0; g < 4; ++g) {
						int ih = goh[g] * p.ah + rh * p.bh + p.ch;
						int iw = gow[g] * p.aw + rw * p.bw + p.cw;
						if (goh[g] < -0x10000) { ih = -0x40000; iw = -0x40000; }
						tma_load_4d(st + g * (TC_KB * 128), &tm_a_hi, &full[s], gn[g], ih, iw, c0);
						tma_load_4d(st + TC_A_BYTES + g * (TC_KB * 128), &tm_a_lo, &full[s], gn[g], ih, iw, c0);
					}
					tma_load_3d(st + 2 * TC_A_BYTES, &tm_b_hi, &full[s], c0, jt * p.BN, tap);
					tma_load_3d(st + 2 * TC_A_BYTES + b_bytes, &tm_b_lo, &full[s], c0, jt * p.BN, tap);
					if (++s == p.stages) { s = 0; ph ^= 1; }
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer (one elected thread) =====
		if (elect_one()) {
			const uint32_t idesc = make_idesc_tf32(p.BN, true, false);
			int s = 0; uint32_t ph = 0;
			int acc = 0; uint32_t acc_ph = 0;
			for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
				mbar_wait(&acc_empty[acc], acc_ph ^ 1);
				tc_fence_after();
				const uint32_t d = tmem_base + (uint32_t) (acc * p.BN);
				for (int kb = 0; kb < kblocks; ++kb) {
					mbar_wait(&full[s], ph);
					tc_fence_after();
					const uint32_t a_hi = smem_u32(smem + (size_t) s * stage_bytes);
					const uint32_t a_lo = a_hi + TC_A_BYTES;
					const uint32_t b_hi = a_hi + 2 * TC_A_BYTES;
					const uint32_t b_lo = b_hi + b_bytes;
					#pragma unroll
					for (int pass = 0; pass < 3; ++pass) {
						const uint32_t a = pass == 0 ? a_lo : a_hi;
						const uint32_t b = pass == 1 ? b_lo : b_hi;
						#pragma unroll
						for (int ks = 0; ks < TC_KB / 8; ++ks) {
							// A: MN-major, 128 B rows of 32 consecutive m; one K step = 8 rows = two 4-row atoms
							// (SBO = 512); the four 32-row M groups are LBO = KB*128 apart
							const uint64_t da = make_smem_desc(a + ks * 1024, TC_KB * 128, 512, LT_SW128_BASE32B);
							// B: K-major, 128 B rows, 8-row groups SBO = 1024 apart, K step = 32 B inside the row
							const uint64_t db = make_smem_desc(b + ks * 32, 16, 1024, LT_SW128);
							umma_tf32(d, da, db, idesc, (kb | pass | ks) != 0 ? 1u : 0u);
						}
					}
					umma_commit(&empty[s]);   // frees the smem stage once these MMAs have read it
					if (++s == p.stages) { s = 0; ph ^= 1; }
				}
				umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
				if (++acc == 2) { acc = 0; acc_ph ^= 1; }
			}
		}
	} else {
		// ===== epilogue warps 2..5: TMEM -> registers -> (+bias) -> coalesced global stores =====
		const int q = warp & 3;  // TMEM lane quarter this warp may access
		int acc = 0; uint32_t acc_ph = 0;
		for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
			const int mt = tile % p.m_tiles, jt = tile / p.m_tiles;
			mbar_wait(&acc_full[acc], acc_ph);
			tc_fence_after();
			const long long m = (long long) mt * TC_BM + 32 * q + lane;
			const bool m_ok = m < p.M;
			const long long pix = m / p.N;
			const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + (uint32_t) (acc * p.BN);
			for (int c0 = 0; c0 < p.BN; c0 += 16) {
				float v[16];
				tmem_ld_16(taddr + c0, v);
				tmem_ld_wait();
				#pragma unroll
				for (int i = 0; i < 16; ++i) {
					const int j = jt * p.BN + c0 + i;
					if (m_ok && j < p.J) {
						float r = v[i];
						if (p.bias_mode == 1) r += __ldg(p.bias + j);
						else if (p.bias_mode == 2) r += __ldg(p.bias + pix + p.P * j);
						p.out[m + p.M * j] = r;
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&acc_empty[acc]);
			if (++acc == 2) { acc = 0; acc_ph ^= 1; }
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) tmem_dealloc(tmem_base, (uint32_t) p.tmem_cols);
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
		const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
		CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
	static EncodeTiledFn fn = nullptr;
	if (!fn) {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
				q == cudaDriverEntryPointSuccess)
			fn = (EncodeTiledFn) p;
	}
	return fn;
}

static int encode_map(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
		const cuuint32_t* box, CUtensorMapSwizzle swizzle) {
	EncodeTiledFn enc = get_encode();
	if (!enc) {
		set_error("cuTensorMapEncodeTiled entry point unavailable");
		return CATTL3_ERR_CUDA;
	}
	cuuint32_t estr[5] = { 1, 1, 1, 1, 1 };
	CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t) rank, const_cast<void*>(base), dims, strides_bytes, box,
			estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int) r, rank);
		return CATTL3_ERR_CUDA;
	}
	return CATTL3_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Low-order split of a tensor into one of the context's two scratch slots; a split made earlier in
// the same API call (ctx->lo_src[slot] == ptr) is reused -- conv_backward needs lo(dY) twice.
static int get_lo_split(cattl3_ctx* ctx, const float* ptr, long long elems, int slot, const float** lo_out) {
	for (int s = 0; s < 2; ++s)
		if (ctx->lo_src[s] == ptr && ctx->lo_elems[s] == elems) { *lo_out = (const float*) ctx->lo_buf[s]; return CATTL3_OK; }
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->lo_buf[slot], &ctx->lo_bytes[slot], (size_t) elems * 4));
	split_lo_kernel<<<ew_grid(ctx, elems / 4 + 1, 256), 256, 0, ctx->stream>>>(elems, ptr, (float*) ctx->lo_buf[slot]);
	CATTL3_LAUNCHED(ctx);
	ctx->lo_src[slot] = ptr;
	ctx->lo_elems[slot] = elems;
	*lo_out = (const float*) ctx->lo_buf[slot];
	return CATTL3_OK;
}

bool tc_gather_gemm_supported(const cattl3_ctx*, const GatherGeom& gg) {
	// 32-row TMA boxes along n; no per-pixel divisibility tests (strided transposed gathers go to SIMT)
	if (gg.N % 32 != 0 || gg.denh != 1 || gg.denw != 1) return false;
	// tiny reduce / filter counts cannot fill a tensor-core tile: those layers are HBM / latency
	// bound and stay on the SIMT kernel ("Tiny-channel configs")
	if (gg.SC < 16 || gg.J < 16) return false;
	return get_encode() != nullptr;
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

int tc_gather_gemm_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* w, const float* bias,
		int bias_mode, float* out) {
	CATTL3_REQUIRE(aligned16(src) && aligned16(out), "tcgen05 path needs 16-byte aligned tensors");
	const int T = gg.RH * gg.RW;
	const int r_pad = round_up(gg.SC, TC_KB);
	const int BN = gg.J >= 256 ? 256 : round_up(gg.J, 16);
	const int j_tiles = (gg.J + BN - 1) / BN;
	const int j_pad = j_tiles * BN;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long src_elems = (long long) gg.N * gg.SH * gg.SW * gg.SC;
	const long long w_elems = (long long) T * j_pad * r_pad;

	// operand preparation: low-order split of the activations, packed + split weights
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->tc_w, &ctx->tc_w_bytes, (size_t) w_elems * 8));
	const float* a_lo = nullptr;
	CATTL3_CHECK(get_lo_split(ctx, src, src_elems, 0, &a_lo));
	float* w_hi = (float*) ctx->tc_w;
	float* w_lo = w_hi + w_elems;
	pack_weights_kernel<<<ew_grid(ctx, w_elems, 256), 256, 0, ctx->stream>>>(gg, r_pad, j_pad, w, w_hi, w_lo);
	CATTL3_LAUNCHED(ctx);

	CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
	{
		cuuint64_t dims[4] = { (cuuint64_t) gg.N, (cuuint64_t) gg.SH, (cuuint64_t) gg.SW, (cuuint64_t) gg.SC };
		cuuint64_t str[3] = { (cuuint64_t) gg.N * 4, (cuuint64_t) gg.N * gg.SH * 4, (cuuint64_t) gg.N * gg.SH * gg.SW * 4 };
		cuuint32_t box[4] = { 32, 1, 1, (cuuint32_t) TC_KB };
		CATTL3_CHECK(encode_map(&tm_a_hi, src, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
		CATTL3_CHECK(encode_map(&tm_a_lo, a_lo, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
	}
	{
		cuuint64_t dims[3] = { (cuuint64_t) r_pad, (cuuint64_t) j_pad, (cuuint64_t) T };
		cuuint64_t str[2] = { (cuuint64_t) r_pad * 4, (cuuint64_t) r_pad * j_pad * 4 };
		cuuint32_t box[3] = { 32, (cuuint32_t) BN, 1 };
		CATTL3_CHECK(encode_map(&tm_b_hi, w_hi, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
		CATTL3_CHECK(encode_map(&tm_b_lo, w_lo, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
	}

	TcGemmParams p;
	p.N = gg.N; p.SH = gg.SH; p.SW = gg.SW; p.OH = gg.OH; p.OW = gg.OW; p.J = gg.J; p.RH = gg.RH; p.RW = gg.RW;
	p.ah = gg.ah; p.bh = gg.bh; p.ch = gg.ch; p.aw = gg.aw; p.bw = gg.bw; p.cw = gg.cw;
	p.M = M; p.P = (long long) gg.OH * gg.OW;
	p.m_tiles = (int) ceil_div(M, TC_BM); p.j_tiles = j_tiles;
	p.BN = BN; p.kc = r_pad / TC_KB;
	const int stage_bytes = 2 * TC_A_BYTES + 2 * BN * TC_KB * 4;
	int stages = (227 * 1024 - 2048) / stage_bytes;
	if (stages > 6) stages = 6;
	p.stages = stages;
	int cols = 32;
	while (cols < 2 * BN) cols <<= 1;
	p.tmem_cols = cols;
	p.bias_mode = bias_mode; p.bias = bias; p.out = out;
	const size_t smem_bytes = (size_t) stages * stage_bytes + 1024 + 256;
	CATTL3_CUDA(cudaFuncSetAttribute(tc_gather_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
	const int tiles = p.m_tiles * p.j_tiles;
	const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
	tc_gather_gemm_kernel<<<grid, 192, smem_bytes, ctx->stream>>>(tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- the weight-gradient kernel ---------------------------------------------------------------------
// dw(tap, r, j) += sum_m src(m, tap, r) * plain(m, j): a GEMM whose reduction runs over
// m = N*OH*OW.  Both operands are K-major here (m is contiguous in HBM for both):
//   A tile: 128 rows = (tap, channel) pairs, built from 128/RB TMA boxes [32 m][1][1][RB channels]
//           of the gathered tensor (one box per tap, at that tap's spatial coordinate);
//   B tile: BN rows = output channels, one 2-D box [32 m][BN] of the plain tensor.
// One k-block = one 32-row m-group (32 batch entries of one pixel).  Each CTA owns one
// (row tile, column tile) and one contiguous range of m-groups (split-K); the fp32 partial tile is
// written to scratch and reduced into dw in split order by wgrad_reduce_tc_kernel (deterministic,
// accumulating: Parameters::accumulate_grad, C-ATTL3/parameters/StandardParameters.hpp:115-123).
struct TcWgradParams {
	int N, OH, OW, R, J, RH, RW;
	int ah, bh, ch, aw, bw, cw;
	int RB;            // channel rows per TMA box (32 / 64 / 128)
	int rchunks;       // r_pad / RB
	int row_blocks;    // T * rchunks
	int row_tiles, j_tiles, splits;
	long long mgroups, mg_per_split;
	int BN, stages, tmem_cols;
	int flush;         // k-blocks accumulated in TMEM before the partial tile is folded into fp32 scratch
	long long w_stap, w_sr, w_sj, dw_elems;
	float* partial;    // [split][tile][BN columns][128 rows]
};

// Tensor-core accumulation truncates (round-toward-zero) at every MMA, so a reduction of n MMA steps
// carries a bias of ~n * 2^-24 relative (measured: 2e-4 on the 10^4-step config-2 weight gradient).
// The kernel therefore closes a TMEM accumulator every `flush` k-blocks; the epilogue warps fold it
// into a CTA-private fp32 tile with round-to-nearest adds while the MMA warp fills the other
// accumulator.
__global__ void __launch_bounds__(192, 1) tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
		const __grid_constant__ CUtensorMap tm_a_lo, const __grid_constant__ CUtensorMap tm_b_hi,
		const __grid_constant__ CUtensorMap tm_b_lo, const TcWgradParams p) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
	const int b_bytes = p.BN * 128;
	const int stage_bytes = 2 * TC_A_BYTES + 2 * b_bytes;
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) p.stages * stage_bytes);
	uint64_t* full = bars;
	uint64_t* empty = bars + p.stages;
	uint64_t* acc_full = bars + 2 * p.stages;
	uint64_t* acc_empty = acc_full + 2;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int tiles = p.row_tiles * p.j_tiles;
	const int tile = blockIdx.x % tiles, z = blockIdx.x / tiles;
	const int rt = tile % p.row_tiles, jt = tile / p.row_tiles;
	const int boxes_per_tile = 128 / p.RB;
	const int rb0 = rt * boxes_per_tile;
	int nboxes = p.row_blocks - rb0;
	if (nboxes > boxes_per_tile) nboxes = boxes_per_tile;
	const long long mg0 = (long long) z * p.mg_per_split;
	long long mg1 = mg0 + p.mg_per_split;
	if (mg1 > p.mgroups) mg1 = p.mgroups;
	const long long kblocks = mg1 > mg0 ? mg1 - mg0 : 0;
	const long long chunks = (kblocks + p.flush - 1) / p.flush;

	if (warp == 0 && elect_one()) {
		tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
		tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
		for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
		fence_barrier_init();
	}
	if (warp == 1) tmem_alloc(tmem_slot, (uint32_t) p.tmem_cols);
	tc_fence_before();
	__syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		if (elect_one()) {
			int s = 0; uint32_t ph = 0;
			const uint32_t tx = (uint32_t) (2 * nboxes * p.RB * 128 + 2 * b_bytes);
			for (long long kb = 0; kb < kblocks; ++kb) {
				const long long m = (mg0 + kb) * 32;
				const int n0 = (int) (m % p.N);
				const long long pix = m / p.N;
				const int oh = (int) (pix % p.OH), ow = (int) (pix / p.OH);
				mbar_wait(&empty[s], ph ^ 1);
				uint8_t* st = smem + (size_t) s * stage_bytes;
				mbar_expect_tx(&full[s], tx);
				for (int bx = 0; bx < nboxes; ++bx) {
					const int rb = rb0 + bx;
					const int tap = rb / p.rchunks, c0 = (rb % p.rchunks) * p.RB;
					const int rh = tap % p.RH, rw = tap / p.RH;
					const int ih = oh * p.ah + rh * p.bh + p.ch, iw = ow * p.aw + rw * p.bw + p.cw;
					tma_load_4d(st + bx * p.RB * 128, &tm_a_hi, &full[s], n0, ih, iw, c0);
					tma_load_4d(st + TC_A_BYTES + bx * p.RB * 128, &tm_a_lo, &full[s], n0, ih, iw, c0);
				}
				tma_load_2d(st + 2 * TC_A_BYTES, &tm_b_hi, &full[s], (int) m, jt * p.BN);
				tma_load_2d(st + 2 * TC_A_BYTES + b_bytes, &tm_b_lo, &full[s], (int) m, jt * p.BN);
				if (++s == p.stages) { s = 0; ph ^= 1; }
			}
		}
	} else if (warp == 1) {
		if (elect_one()) {
			const uint32_t idesc = make_idesc_tf32(p.BN, false, false);
			int s = 0; uint32_t ph = 0;
			int acc = 0; uint32_t acc_ph = 0;
			long long kb = 0;
			for (long long c = 0; c < chunks; ++c) {
				mbar_wait(&acc_empty[acc], acc_ph ^ 1);
				tc_fence_after();
				const uint32_t d = tmem_base + (uint32_t) (acc * p.BN);
				const long long kend = kb + p.flush < kblocks ? kb + p.flush : kblocks;
				for (bool first = true; kb < kend; ++kb) {
					mbar_wait(&full[s], ph);
					tc_fence_after();
					const uint32_t a_hi = smem_u32(smem + (size_t) s * stage_bytes);
					const uint32_t a_lo = a_hi + TC_A_BYTES;
					const uint32_t b_hi = a_hi + 2 * TC_A_BYTES;
					const uint32_t b_lo = b_hi + b_bytes;
					#pragma unroll
					for (int pass = 0; pass < 3; ++pass) {
						const uint32_t a = pass == 0 ? a_lo : a_hi;
						const uint32_t b = pass == 1 ? b_lo : b_hi;
						#pragma unroll
						for (int ks = 0; ks < 4; ++ks) {
							const uint64_t da = make_smem_desc(a + ks * 32, 16, 1024, LT_SW128);
							const uint64_t db = make_smem_desc(b + ks * 32, 16, 1024, LT_SW128);
							umma_tf32(d, da, db, idesc, (first && pass == 0 && ks == 0) ? 0u : 1u);
						}
					}
					first = false;
					umma_commit(&empty[s]);
					if (++s == p.stages) { s = 0; ph ^= 1; }
				}
				umma_commit(&acc_full[acc]);
				if (++acc == 2) { acc = 0; acc_ph ^= 1; }
			}
		}
	} else {
		const int q = warp & 3;
		const int row = 32 * q + lane;  // accumulator lane = row of the tile
		float* dst = p.partial + ((long long) z * tiles + tile) * p.BN * 128 + row;
		if (chunks == 0) {
			for (int c0 = 0; c0 < p.BN; ++c0) dst[(long long) c0 * 128] = 0.f;
		}
		int acc = 0; uint32_t acc_ph = 0;
		for (long long c = 0; c < chunks; ++c) {
			mbar_wait(&acc_full[acc], acc_ph);
			tc_fence_after();
			const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + (uint32_t) (acc * p.BN);
			for (int c0 = 0; c0 < p.BN; c0 += 16) {
				float v[16];
				tmem_ld_16(taddr + c0, v);
				tmem_ld_wait();
				if (c > 0) {
					#pragma unroll
					for (int i = 0; i < 16; ++i) v[i] += dst[(long long) (c0 + i) * 128];
				}
				#pragma unroll
				for (int i = 0; i < 16; ++i) dst[(long long) (c0 + i) * 128] = v[i];
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(&acc_empty[acc]);
			if (++acc == 2) { acc = 0; acc_ph ^= 1; }
		}
	}
	tc_fence_before();
	__syncthreads();
	if (warp == 1) tmem_dealloc(tmem_base, (uint32_t) p.tmem_cols);
}

// dw(tap, r, j) += sum over splits of the scratch tiles, in split order (deterministic).
__global__ void __launch_bounds__(256) wgrad_reduce_tc_kernel(const TcWgradParams p, int T, float* __restrict__ dw) {
	const long long total = (long long) T * p.R * p.J;
	const int tiles = p.row_tiles * p.j_tiles;
	const int boxes_per_tile = 128 / p.RB;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long) gridDim.x * 256) {
		const int r = (int) (i % p.R);
		const int tap = (int) ((i / p.R) % T);
		const int j = (int) (i / ((long long) p.R * T));
		const int rb = tap * p.rchunks + r / p.RB;
		const int rt = rb / boxes_per_tile;
		const int row = (rb % boxes_per_tile) * p.RB + r % p.RB;
		const int jt = j / p.BN, col = j % p.BN;
		const long long off = ((long long) (rt + p.row_tiles * jt) * p.BN + col) * 128 + row;
		float s = 0.f;
		for (int z = 0; z < p.splits; ++z) s += p.partial[(long long) z * tiles * p.BN * 128 + off];
		dw[tap * p.w_stap + r * p.w_sr + j * p.w_sj] += s;
	}
}

bool tc_wgrad_supported(const cattl3_ctx*, const GatherGeom& gg) {
	if (gg.N % 32 != 0 || gg.denh != 1 || gg.denw != 1) return false;
	if (gg.SC < 16 || gg.J < 16) return false;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	if (M >= (1ll << 31)) return false;  // TMA coordinates are 32-bit
	return get_encode() != nullptr;
}

int tc_wgrad_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* plain, float* dw) {
	CATTL3_REQUIRE(aligned16(src) && aligned16(plain), "tcgen05 path needs 16-byte aligned tensors");
	const int T = gg.RH * gg.RW;
	const int r_pad = round_up(gg.SC, 32);
	const int RB = r_pad % 128 == 0 ? 128 : (r_pad % 64 == 0 ? 64 : 32);
	const int BN = gg.J >= 256 ? 256 : round_up(gg.J, 16);
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long src_elems = (long long) gg.N * gg.SH * gg.SW * gg.SC;
	const long long plain_elems = M * gg.J;

	const float *a_lo = nullptr, *b_lo = nullptr;
	CATTL3_CHECK(get_lo_split(ctx, src, src_elems, 0, &a_lo));
	CATTL3_CHECK(get_lo_split(ctx, plain, plain_elems, 1, &b_lo));

	CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
	{
		cuuint64_t dims[4] = { (cuuint64_t) gg.N, (cuuint64_t) gg.SH, (cuuint64_t) gg.SW, (cuuint64_t) gg.SC };
		cuuint64_t str[3] = { (cuuint64_t) gg.N * 4, (cuuint64_t) gg.N * gg.SH * 4, (cuuint64_t) gg.N * gg.SH * gg.SW * 4 };
		cuuint32_t box[4] = { 32, 1, 1, (cuuint32_t) RB };
		CATTL3_CHECK(encode_map(&tm_a_hi, src, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
		CATTL3_CHECK(encode_map(&tm_a_lo, a_lo, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
	}
	{
		cuuint64_t dims[2] = { (cuuint64_t) M, (cuuint64_t) gg.J };
		cuuint64_t str[1] = { (cuuint64_t) M * 4 };
		cuuint32_t box[2] = { 32, (cuuint32_t) BN };
		CATTL3_CHECK(encode_map(&tm_b_hi, plain, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
		CATTL3_CHECK(encode_map(&tm_b_lo, b_lo, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
	}

	TcWgradParams p;
	p.N = gg.N; p.OH = gg.OH; p.OW = gg.OW; p.R = gg.SC; p.J = gg.J; p.RH = gg.RH; p.RW = gg.RW;
	p.ah = gg.ah; p.bh = gg.bh; p.ch = gg.ch; p.aw = gg.aw; p.bw = gg.bw; p.cw = gg.cw;
	p.RB = RB; p.rchunks = r_pad / RB; p.row_blocks = T * p.rchunks;
	p.row_tiles = (p.row_blocks * RB + 127) / 128;
	p.j_tiles = (gg.J + BN - 1) / BN;
	p.mgroups = M / 32;
	const int tiles = p.row_tiles * p.j_tiles;
	long long splits = ctx->sm_count / tiles;
	if (splits < 1) splits = 1;
	if (splits > p.mgroups) splits = p.mgroups;
	p.mg_per_split = ceil_div(p.mgroups, splits);
	p.splits = (int) ceil_div(p.mgroups, p.mg_per_split);
	p.BN = BN;
	const int stage_bytes = 2 * TC_A_BYTES + 2 * BN * 128;
	int stages = (227 * 1024 - 2048) / stage_bytes;
	if (stages > 6) stages = 6;
	p.stages = stages;
	int cols = 32;
	while (cols < 2 * BN) cols <<= 1;
	p.tmem_cols = cols;
	p.flush = 32;
	p.w_stap = gg.w_stap; p.w_sr = gg.w_sr; p.w_sj = gg.w_sj;
	p.dw_elems = (long long) T * gg.SC * gg.J;
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) p.splits * tiles * BN * 128 * 4));
	p.partial = (float*) ctx->ws;
	const size_t smem_bytes = (size_t) stages * stage_bytes + 1024 + 256;
	CATTL3_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
	tc_wgrad_kernel<<<tiles * p.splits, 192, smem_bytes, ctx->stream>>>(tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo, p);
	CATTL3_LAUNCHED(ctx);
	wgrad_reduce_tc_kernel<<<ew_grid(ctx, p.dw_elems, 256), 256, 0, ctx->stream>>>(p, T, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

} // namespace cattl3
