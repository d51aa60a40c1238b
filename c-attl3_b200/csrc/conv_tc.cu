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
#include <stdlib.h>
#include <string.h>

#include "activations.cuh"

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
// 16-byte asynchronous global -> shared copy (LDGSTS); src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async_16(void* dst, const void* src, uint32_t src_bytes) {
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ float4 lds_128(uint32_t addr) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts_128(uint32_t addr, float4 v) {
	asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
	asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t) tm) : "memory");
}
// ---- CTA pairs (cta_group::2) ------------------------------------------------------------------------------
// Two CTAs of one cluster (the two SMs of a TPC) issue ONE tcgen05.mma of M = 256: each CTA holds its own 128 rows of A
// (and of the accumulator) in its own tensor memory and HALF of the rows of B in its own shared memory, so every byte
// of B in shared memory feeds 256 accumulator rows instead of 128 (measured: scripts/tf32_peak.cu verifies the
// operand placement).  CTA 0 (the leader) issues the MMAs and the commits; barriers that both CTAs arrive on live in
// the leader and are reached through mapa; commits are multicast to the same barrier offset in both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at this offset in CTA `cta` of the cluster (release at cluster scope)
// Measured (profiles/README.md, r2): cluster-scope release / acquire on these per-k-block handshakes compile to
// MEMBAR.ALL.GPU and cost more than the pair saves; the plain forms (what CUTLASS' ClusterBarrier uses) order the
// tcgen05 and TMA traffic they guard through tcgen05.fence / the mbarrier itself.
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
	asm volatile(
		"{\n\t.reg .b32 ra;\n\t"
		"mapa.shared::cluster.u32 ra, %0, %1;\n\t"
		"mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
		:: "r"(smem_u32(bar)), "r"(cta) : "memory");
}
template<int CTAS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
	if (CTAS == 2) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(cols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	} else {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(cols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
}
template<int CTAS> __device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
	if (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
	else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all MMAs issued so far by this thread -> one arrival on `bar` when they have completed; for a CTA pair the arrival
// goes to the barrier at the same offset in both CTAs
template<int CTAS> __device__ __forceinline__ void umma_commit(uint64_t* bar) {
	if (CTAS == 2)
		asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
				:: "r"(smem_u32(bar)), "h"((uint16_t) 3) : "memory");
	else
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
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 rows = TMEM lanes, K elements = consecutive columns)
// is read from tensor memory, so only B costs shared-memory bandwidth.  CTAS = 2: M = 256, the pair's MMA.
template<int CTAS> __device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
		uint32_t accumulate) {
	if (CTAS == 2)
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"setp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
			:: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
	else
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"setp.ne.b32 p, %4, 0;\n\t"
			"tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
			:: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> 32 lanes x 16 consecutive columns of TMEM (lane = thread of the warp's lane quarter)
__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const uint32_t* r) {
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
		"{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
		:: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
		   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// stride byte offsets (all >> 4), version = 1 (Blackwell), layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
// K-major operand, measured on B200 (scripts/umma_probe.cu, scripts/umma_probe2.cu): rows of 32 fp32
// (128 B) use SWIZZLE_128B with 8-row atoms SBO = 1024 apart, rows of 16 fp32 (64 B) use SWIZZLE_64B with
// 8-row atoms SBO = 512 apart; one K step of 8 elements = +32 B on the start address.
constexpr uint32_t LT_SW128 = 2, LT_SW64 = 4;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
	return (uint64_t) ((smem_addr >> 4) & 0x3FFF) | ((uint64_t) ((lbo_bytes >> 4) & 0x3FFF) << 16) |
			((uint64_t) ((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t) layout_type << 61);
}
template<int KB>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile_addr, int kstep) {
	return KB == 32 ? make_smem_desc(tile_addr + kstep * 32, 16, 1024, LT_SW128)
			: make_smem_desc(tile_addr + kstep * 32, 16, 512, LT_SW64);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N runtime.
__host__ __device__ inline uint32_t make_idesc_tf32(int n, int m = 128) {
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (n >> 3) << 17) | ((uint32_t) (m >> 4) << 24);
}

// ---- 3xTF32 operand split -------------------------------------------------------------------------------
// kind::tf32 uses the top 19 bits of an fp32 word.  x = hi + lo with hi = x truncated to TF32 and
// lo = x - hi (exact in fp32), D += lo_a*hi_b + hi_a*lo_b + hi_a*hi_b in an FP32 accumulator.  Adding half a
// TF32 ulp to lo's magnitude turns the hardware's truncation of lo into a round-to-nearest: unbiased,
// ~2^-22 relative to x; the carry cannot leave a finite exponent.
__device__ __forceinline__ uint32_t tf32_hi_bits(float v) { return __float_as_uint(v) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo_bits(float v) {
	const float r = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
	return __float_as_uint(r) + 0x1000u;
}

// The weights of the gather GEMM -- its shared-memory operand -- are split once in HBM while being repacked
// K-major as [part][tap][j_pad][r_pad] (r contiguous, zero padded; part 0 = hi, 1 = lo).
// tapbox: the k-blocks are tap ROWS; element r of k-block rh is tap column r % RW of channel r / RW (the row order of
// the TMA box [n][1][RW][C] the activations arrive in).
__global__ void __launch_bounds__(256) pack_weights_kernel(GatherGeom gg, int r_pad, int j_pad, int tapbox, const float* __restrict__ w,
		float* __restrict__ packed) {
	const int T = tapbox ? gg.RH : gg.RH * gg.RW;
	const long long total = (long long) T * j_pad * r_pad;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long) gridDim.x * 256) {
		const int r = (int) (i % r_pad);
		const int j = (int) ((i / r_pad) % j_pad);
		const int tap = (int) (i / ((long long) r_pad * j_pad));
		float v = 0.f;
		if (tapbox) {
			if (r < gg.RW * gg.SC && j < gg.J)
				v = w[gg.w_off + tap * gg.w_stap + (r % gg.RW) * (gg.w_srw ? gg.w_srw : gg.w_stap * gg.RH) + (r / gg.RW) * gg.w_sr +
						j * gg.w_sj];
		} else if (r < gg.SC && j < gg.J) {
			const int rh = tap % gg.RH, rw = tap / gg.RH;
			v = w[gg.w_off + rh * gg.w_stap + rw * (gg.w_srw ? gg.w_srw : gg.w_stap * gg.RH) + r * gg.w_sr + j * gg.w_sj];
		}
		packed[i] = __uint_as_float(tf32_hi_bits(v));
		packed[total + i] = __uint_as_float(tf32_lo_bits(v));
	}
}
// ---- pipeline shared by the two kernels ------------------------------------------------------------------
// Shared memory is the scarce resource: measured on B200, the TMA writes, the converters' reads and writes
// and the tensor core's operand reads of one SM together sustain ~95 B/clk (profiles/README.md), while one
// 128 x 256 x 8 kind::tf32 MMA alone wants 96 B/clk when both operands come from shared memory.  So the
// 128-row operand -- always the big activation tensor -- goes to TENSOR MEMORY: TMA lands the raw fp32 tile
// in shared memory once, the converter warps read it once, split it and store hi / lo into TMEM
// (tcgen05.st), and tcgen05.mma takes A from TMEM; only the B operand is read from shared memory.
//
// 320 threads: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2-5 = epilogue,
// warps 6-9 = converters (a warp may only touch TMEM lanes 32*(warp % 4) ...).  Per stage:
//   TMA --full--> converters --ready--> MMA --empty (tcgen05.commit)--> TMA.
// TMEM (512 columns): accumulator(s) first, then per stage 2*KB columns holding A hi | A lo.
constexpr int TC_BM = 128, TC_THREADS = 320, TC_MAX_STAGES = 8;

struct TcGemmParams {
	int N, OH, OW, J, RH, RW;
	int ah, bh, ch, aw, bw, cw;
	long long M, P;
	int m_tiles, j_tiles;
	int BN;          // filter tile (multiple of 16, <= 256)
	int r_pad;       // reduce channels, padded to a multiple of KB
	int nb;          // batch entries per A box (32 / 64 / 128): a tile is 128 / nb boxes of [nb n][KB channels]
	int a_rows;      // rows a TMA box of A delivers (KB; fewer for a tap box [nb n][1][RW][C]: the rest of the tile stays zero)
	int stages, nacc;
	int bias_mode;
	const float* bias;
	float* out;
	// where row m = n + N*(i + OH*j) of the GEMM lives in the output tensor: pixel (h0 + hs*i, w0 + ws*j) of an
	// oH x oW image (dense: h0 = w0 = 0, hs = ws = 1, oH = OH, oW = OW); out_cs = N*oH*oW is the channel stride
	int h0, hs, oH, w0, ws, oW;
	long long out_cs;
	// fused epilogue (cattl3_epilogue): out may be null when act_out is given; stat_partial = per-CTA column sums
	int act_kind;
	float act_param;
	float* act_out;
	double* stat_partial;   // [gridDim.x][2][j_tiles * BN] or null
};

// Gather GEMM (conv / dense forward, stride-1 input gradient):
//   out[m + M*j] = bias + sum_{tap, r} src(m, tap, r) * w(tap, r, j),  m = n + N*(oh + OH*ow).
// A: 128 / nb TMA boxes [nb n][1][1][KB channels] of the source tensor at the tap's coordinate (nb = 32, 64 or
// 128 consecutive batch entries of one pixel), unswizzled; padding = TMA out-of-bounds zero fill.
// B: one TMA box [KB r][BN j] of the packed weights, hi and lo.
// The reduction walks channel chunks OUTER and taps INNER: the 148 CTAs then work on the same few channels
// of neighbouring pixels at the same time, so every source byte is fetched from HBM once and the tap-shifted
// re-reads hit L2.
// Warps: 0 = TMA producer, 1 = MMA issuer, then NEPI epilogue warps, then NCONV converter warps (4 or 8 each: a warp
// reaches the TMEM lane quarter warp % 4 only, so eight warps are two per quarter -- epilogue warps then split the
// tile's columns, converter warps alternate k-blocks).  Short k-blocks (few filters: the input gradient) need the
// second converter set, wide tiles the second epilogue set.
template<int KB, int CTAS, int NEPI, int NCONV>
__global__ void __launch_bounds__((2 + NEPI + NCONV) * 32, 1) tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap tm_a,
		const __grid_constant__ CUtensorMap tm_b, const TcGemmParams p) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
	constexpr int A_BYTES = TC_BM * KB * 4;
	const int b_bytes = p.BN / CTAS * KB * 4;   // a CTA of a pair holds half of the rows of B
	const int stage_bytes = A_BYTES + 2 * b_bytes;
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) p.stages * stage_bytes);
	uint64_t* full = bars;
	uint64_t* ready = bars + TC_MAX_STAGES;
	uint64_t* empty = bars + 2 * TC_MAX_STAGES;
	uint64_t* acc_full = bars + 3 * TC_MAX_STAGES;
	uint64_t* acc_empty = acc_full + 2;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
	float* sbias = reinterpret_cast<float*>(bars + 32);  // 256 floats behind the barriers
	double* sstat = reinterpret_cast<double*>(sbias + 256);  // [4 quarters][2][BN] column sums (only with stat_partial)
	float* spatch = reinterpret_cast<float*>(sstat + 8 * p.BN);  // [NEPI warps][32][17] transpose patches (likewise)
	constexpr int EPI_THREADS = NEPI * 32;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int T = p.RH * p.RW;
	const int kblocks = T * (p.r_pad / KB);
	const int tiles = p.m_tiles * p.j_tiles;   // m_tiles counts tiles of 128 * CTAS rows: a pair works on one together
	const uint32_t a_col0 = (uint32_t) (p.nacc * p.BN);
	const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
	const int unit = blockIdx.x / CTAS, units = gridDim.x / CTAS;   // CTA (pair) index and count
	constexpr int TILE_M = TC_BM * CTAS;

	if (warp == 0 && elect_one()) {
		tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
		// full: this CTA's TMA bytes; ready: one arrival per converter warp of the pair (in the leader); empty / acc_full:
		// the leader's commit, multicast; acc_empty: one arrival per epilogue warp of the pair (in the leader)
		for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 4 * CTAS); mbar_init(&empty[s], 1); }
		for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], NEPI * CTAS); }
		fence_barrier_init();
	}
	if (p.a_rows < KB) {
		// tap boxes fill only the first a_rows k-rows of a tile: the others are zero for the whole kernel
		for (int s = 0; s < p.stages; ++s) {
			float4* a = reinterpret_cast<float4*>(smem + (size_t) s * stage_bytes);
			for (int i = threadIdx.x; i < A_BYTES / 16; i += blockDim.x) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		}
	}
	if (warp == 1) tmem_alloc<CTAS>(tmem_slot, 512u);
	tc_fence_before();
	if (CTAS == 2) cluster_sync(); else __syncthreads();   // the peer's barriers exist before anything arrives on them
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		// ===== TMA producer =====
		if (elect_one()) {
			int s = 0; uint32_t ph = 0;
			for (int tile = unit; tile < tiles; tile += units) {
				const int mt = tile % p.m_tiles, jt = tile / p.m_tiles;
				int gn[4], goh[4], gow[4];  // the 128 / nb row groups of this tile: (n0, oh, ow) each, or out of range
				const int ngroups = TC_BM / p.nb;
				#pragma unroll
				for (int g = 0; g < 4; ++g) {
					const long long m = (long long) mt * TILE_M + rank * TC_BM + p.nb * g;
					if (g < ngroups && m < p.M) {
						gn[g] = (int) (m % p.N);
						const long long pix = m / p.N;
						goh[g] = (int) (pix % p.OH);
						gow[g] = (int) (pix / p.OH);
					} else {
						gn[g] = 0; goh[g] = -0x40000; gow[g] = -0x40000;
					}
				}
				int rh = 0, rw = 0, tap = 0, c0 = 0;
				for (int kb = 0; kb < kblocks; ++kb) {
					mbar_wait(&empty[s], ph ^ 1);
					uint8_t* st = smem + (size_t) s * stage_bytes;
					mbar_expect_tx(&full[s], (uint32_t) (TC_BM * p.a_rows * 4 + 2 * b_bytes));
					#pragma unroll
					for (int g = 0; g < 4; ++g) {
						if (g < ngroups) {
							int ih = goh[g] * p.ah + rh * p.bh + p.ch;
							int iw = gow[g] * p.aw + rw * p.bw + p.cw;
							if (goh[g] < -0x10000) { ih = -0x40000; iw = -0x40000; }  // far out of bounds: zero fill
							tma_load_4d(st + g * (KB * p.nb * 4), &tm_a, &full[s], gn[g], ih, iw, c0);
						}
					}
					const int j0 = jt * p.BN + (int) rank * (p.BN / CTAS);   // this CTA's rows of B
					tma_load_4d(st + A_BYTES, &tm_b, &full[s], c0, j0, tap, 0);
					tma_load_4d(st + A_BYTES + b_bytes, &tm_b, &full[s], c0, j0, tap, 1);
					if (++s == p.stages) { s = 0; ph ^= 1; }
					// next k-block: taps inner, channel chunks outer
					++tap;
					if (++rh == p.RH) { rh = 0; if (++rw == p.RW) { rw = 0; tap = 0; c0 += KB; } }
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer (one elected thread; of a pair, the leader's) =====
		if (rank == 0 && elect_one()) {
			const uint32_t idesc = make_idesc_tf32(p.BN, TILE_M);
			int s = 0; uint32_t ph = 0;
			int acc = 0; uint32_t acc_ph = 0;
			for (int tile = unit; tile < tiles; tile += units) {
				mbar_wait(&acc_empty[acc], acc_ph ^ 1);
				tc_fence_after();
				const uint32_t d = tmem_base + (uint32_t) (acc * p.BN);
				for (int kb = 0; kb < kblocks; ++kb) {
					mbar_wait(&ready[s], ph);
					tc_fence_after();
					const uint32_t b_hi = smem_u32(smem + (size_t) s * stage_bytes + A_BYTES);
					const uint32_t b_lo = b_hi + b_bytes;
					const uint32_t a_hi = tmem_base + a_col0 + (uint32_t) (s * 2 * KB);
					const uint32_t a_lo = a_hi + KB;
					#pragma unroll
					for (int pass = 0; pass < 3; ++pass) {
						const uint32_t a = pass == 0 ? a_lo : a_hi;
						const uint32_t b = pass == 1 ? b_lo : b_hi;
						#pragma unroll
						for (int ks = 0; ks < KB / 8; ++ks)
							umma_tf32_ts<CTAS>(d, a + 8 * ks, kmajor_desc<KB>(b, ks), idesc, (kb | pass | ks) != 0 ? 1u : 0u);
					}
					umma_commit<CTAS>(&empty[s]);   // frees the smem stage and its TMEM columns once these MMAs have read them
					if (++s == p.stages) { s = 0; ph ^= 1; }
				}
				umma_commit<CTAS>(&acc_full[acc]);  // accumulator complete -> epilogue
				if (++acc == p.nacc) { acc = 0; acc_ph ^= 1; }
			}
		}
	} else if (warp < 2 + NEPI) {
		// ===== epilogue warps: TMEM -> registers -> (+bias, activation, column statistics) -> coalesced stores =====
		// The per-filter bias of the tile is staged in shared memory while the MMAs of the tile still run,
		// so the drain itself is tcgen05.ld + add + store with no dependent global loads.
		const int q = warp & 3;  // TMEM lane quarter this warp may access
		const int et = threadIdx.x - 64;  // index among the epilogue threads
		const int ew = warp - 2;          // index among the epilogue warps
		// two warps per lane quarter split the tile's columns (BN % 32 == 0 then)
		const int c_begin = (ew >> 2) * (p.BN / (NEPI / 4)), c_end = c_begin + p.BN / (NEPI / 4);
		int acc = 0; uint32_t acc_ph = 0;
		int staged_jt = -1;
		const bool stats = p.stat_partial != nullptr;
		// c_end - c_begin is a multiple of 32 whenever the tile is (every pair tile is)
		const bool fast = !stats && p.bias_mode == 1 && p.out != nullptr && p.act_out == nullptr && (c_end - c_begin) % 32 == 0;
		const int j_pad = p.j_tiles * p.BN;
		double* my_stat = sstat + q * 2 * p.BN;
		if (stats) {
			for (int c = et; c < 8 * p.BN; c += EPI_THREADS) sstat[c] = 0.0;
			asm volatile("bar.sync 1, %0;" :: "n"(EPI_THREADS) : "memory");
		}
		// Column sums of this CTA's tiles of filter tile `jt` -> its slot of the partial buffer (plain stores:
		// the tile order of a CTA visits every jt at most once, in increasing order).
		auto flush_stats = [&](int jt) {
			asm volatile("bar.sync 1, %0;" :: "n"(EPI_THREADS) : "memory");
			for (int c = et; c < 2 * p.BN; c += EPI_THREADS) {
				const int k = c / p.BN, col = c % p.BN;
				const double t = ((sstat[(0 * 2 + k) * p.BN + col] + sstat[(1 * 2 + k) * p.BN + col]) +
						sstat[(2 * 2 + k) * p.BN + col]) + sstat[(3 * 2 + k) * p.BN + col];
				p.stat_partial[((long long) blockIdx.x * 2 + k) * j_pad + jt * p.BN + col] = t;
				sstat[(0 * 2 + k) * p.BN + col] = 0.0; sstat[(1 * 2 + k) * p.BN + col] = 0.0;
				sstat[(2 * 2 + k) * p.BN + col] = 0.0; sstat[(3 * 2 + k) * p.BN + col] = 0.0;
			}
			asm volatile("bar.sync 1, %0;" :: "n"(EPI_THREADS) : "memory");
		};
		for (int tile = unit; tile < tiles; tile += units) {
			const int mt = tile % p.m_tiles, jt = tile / p.m_tiles;
			if (jt != staged_jt) {
				if (stats && staged_jt >= 0) flush_stats(staged_jt);
				if (p.bias_mode == 1) {
					asm volatile("bar.sync 1, %0;" :: "n"(EPI_THREADS) : "memory");  // nobody still reads the previous tile's bias
					for (int c = et; c < p.BN; c += EPI_THREADS) {
						const int j = jt * p.BN + c;
						sbias[c] = j < p.J ? __ldg(p.bias + j) : 0.f;
					}
					asm volatile("bar.sync 1, %0;" :: "n"(EPI_THREADS) : "memory");
				}
				staged_jt = jt;
			}
			mbar_wait(&acc_full[acc], acc_ph);
			tc_fence_after();
			const long long m = (long long) mt * TILE_M + rank * TC_BM + 32 * q + lane;
			const bool m_ok = m < p.M;
			// M is a multiple of 32 (N % 32 == 0), so the 32 rows of a warp are valid or invalid together
			const bool warp_ok = (long long) mt * TILE_M + rank * TC_BM + 32 * q < p.M;
			const long long gpix = m / p.N;
			const int gi = (int) (gpix % p.OH), gj = (int) (gpix / p.OH);
			const long long pix = (p.h0 + p.hs * gi) + (long long) p.oH * (p.w0 + p.ws * gj);  // pixel in the output tensor
			const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + (uint32_t) (acc * p.BN);
			const long long out_off = (m % p.N) + p.N * pix + p.out_cs * (long long) (jt * p.BN);
			float* out_m = p.out ? p.out + out_off : nullptr;
			float* act_m = p.act_out ? p.act_out + out_off : nullptr;
			const int jn = p.J - jt * p.BN < p.BN ? p.J - jt * p.BN : p.BN;  // valid columns of this tile
			if (fast && m_ok && jn == p.BN) {
				// the plain layer (per-filter bias, nothing fused, whole tile): two 16-column loads in flight, the stores of
				// one overlapping the tensor-memory latency of the next
				float va[16], vb[16];
				tmem_ld_16(taddr + c_begin, va);
				for (int c0 = c_begin; c0 < c_end; c0 += 32) {
					tmem_ld_wait();
					tmem_ld_16(taddr + c0 + 16, vb);
					#pragma unroll
					for (int i = 0; i < 16; ++i) out_m[p.out_cs * (long long) (c0 + i)] = va[i] + sbias[c0 + i];
					tmem_ld_wait();
					if (c0 + 32 < c_end) tmem_ld_16(taddr + c0 + 32, va);
					#pragma unroll
					for (int i = 0; i < 16; ++i) out_m[p.out_cs * (long long) (c0 + 16 + i)] = vb[i] + sbias[c0 + 16 + i];
				}
			} else
			for (int c0 = c_begin; c0 < c_end; c0 += 16) {
				float v[16], bv[16];
				tmem_ld_16(taddr + c0, v);
				if (p.bias_mode == 1) {
					#pragma unroll
					for (int i = 0; i < 16; ++i) bv[i] = sbias[c0 + i];
				} else if (p.bias_mode == 2) {
					// one bias per output element (TransConvKernelLayer): independent loads, issued together
					#pragma unroll
					for (int i = 0; i < 16; ++i) {
						const int c = c0 + i < jn ? c0 + i : jn - 1;
						bv[i] = m_ok ? __ldg(p.bias + pix + p.P * (long long) (jt * p.BN + c)) : 0.f;
					}
				} else {
					#pragma unroll
					for (int i = 0; i < 16; ++i) bv[i] = 0.f;
				}
				tmem_ld_wait();
				if (stats && warp_ok) {
					// Per column: sum and sum of squares of the 32 accumulators (= y - bias) of this warp's rows.  The
					// 32 x 16 chunk is transposed through a padded shared-memory patch (conflict free both ways):
					// lane L then owns column L % 16 over rows 16 * (L / 16) ..., sums d = value - first value and
					// d^2 in fp32 (magnitudes ~ sigma), re-bases to shift 0 in double, and the two half-columns
					// meet in lanes 0..15, which accumulate in double.
					float* patch = spatch + ew * (32 * 17);
					#pragma unroll
					for (int i = 0; i < 16; ++i) patch[lane * 17 + i] = v[i];
					__syncwarp();
					const float* colp = patch + (lane >> 4) * (16 * 17) + (lane & 15);
					const float k0 = colp[0];
					float s1 = 0.f, s2 = 0.f;
					#pragma unroll
					for (int r = 1; r < 16; ++r) {
						const float d = colp[r * 17] - k0;
						s1 += d;
						s2 = fmaf(d, d, s2);
					}
					const double K = (double) k0, d1 = (double) s1;
					double S1 = d1 + 16.0 * K;
					double S2 = (double) s2 + 2.0 * K * d1 + 16.0 * K * K;
					S1 += __shfl_down_sync(0xffffffffu, S1, 16);
					S2 += __shfl_down_sync(0xffffffffu, S2, 16);
					if (lane < 16) {
						my_stat[c0 + lane] += S1;
						my_stat[p.BN + c0 + lane] += S2;
					}
					__syncwarp();   // the patch is rewritten by the next chunk
				}
				if (m_ok) {
					#pragma unroll
					for (int i = 0; i < 16; ++i) v[i] += bv[i];
					if (out_m) {
						if (c0 + 16 <= jn) {
							#pragma unroll
							for (int i = 0; i < 16; ++i) out_m[p.out_cs * (long long) (c0 + i)] = v[i];
						} else {
							#pragma unroll
							for (int i = 0; i < 16; ++i)
								if (c0 + i < jn) out_m[p.out_cs * (long long) (c0 + i)] = v[i];
						}
					}
					if (act_m) {
						act_fwd_rt_n<float, 16>(p.act_kind, v, p.act_param);
						if (c0 + 16 <= jn) {
							#pragma unroll
							for (int i = 0; i < 16; ++i) act_m[p.out_cs * (long long) (c0 + i)] = v[i];
						} else {
							#pragma unroll
							for (int i = 0; i < 16; ++i)
								if (c0 + i < jn) act_m[p.out_cs * (long long) (c0 + i)] = v[i];
						}
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { if (CTAS == 2) mbar_arrive_cta(&acc_empty[acc], 0); else mbar_arrive(&acc_empty[acc]); }
			if (++acc == p.nacc) { acc = 0; acc_ph ^= 1; }
		}
		if (stats && staged_jt >= 0) flush_stats(staged_jt);
	} else {
		// ===== converter warps 6..9: raw A tile (shared memory) -> hi | lo (tensor memory) =====
		// The tile lies in shared memory as 128 / nb unswizzled boxes of KB rows x nb batch entries; row m of the
		// tile is entry m % nb of box m / nb.  Lane = row within the quarter, so a warp reads 32 consecutive
		// floats per k (conflict free) and owns TMEM lanes 32q .. 32q+31.
		const int q = warp & 3;
		const int cset = (warp - 2 - NEPI) >> 2;   // with eight converter warps, set 0 takes the even k-blocks, set 1 the odd ones
		const int row = 32 * q + lane;
		const uint32_t row_off = (uint32_t) ((row / p.nb) * (KB * p.nb * 4) + (row % p.nb) * 4);
		const uint32_t k_stride = (uint32_t) (p.nb * 4);
		// the k-blocks of all of this unit's tiles form one sequence; this set takes every (NCONV / 4)-th of them
		constexpr int NSETS = NCONV / 4;
		const long long my_tiles = tiles > unit ? (tiles - unit + units - 1) / units : 0;
		const long long total = my_tiles * kblocks;
		int s = cset % p.stages; uint32_t ph = (uint32_t) ((cset / p.stages) & 1);
		{
			for (long long g = cset; g < total; g += NSETS) {
				mbar_wait(&full[s], ph);
				const uint8_t* grp = smem + (size_t) s * stage_bytes + row_off;
				const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + a_col0 + (uint32_t) (s * 2 * KB);
				#pragma unroll
				for (int k0 = 0; k0 < KB; k0 += 16) {
					uint32_t hi[16], lo[16];
					#pragma unroll
					for (int k = 0; k < 16; ++k) {
						const float v = *reinterpret_cast<const float*>(grp + (uint32_t) (k0 + k) * k_stride);
						hi[k] = tf32_hi_bits(v);
						lo[k] = tf32_lo_bits(v);
					}
					tmem_st_16(taddr + k0, hi);
					tmem_st_16(taddr + KB + k0, lo);
				}
				tmem_st_wait();
				tc_fence_before();
				__syncwarp();
				if (lane == 0) { if (CTAS == 2) mbar_arrive_cta(&ready[s], 0); else mbar_arrive(&ready[s]); }
				s += NSETS;
				if (s >= p.stages) { s -= p.stages; ph ^= 1; }   // (NSETS <= stages)
			}
		}
	}
	tc_fence_before();
	if (CTAS == 2) cluster_sync(); else __syncthreads();   // nothing of the pair is in flight towards this CTA any more
	if (warp == 1) tmem_dealloc<CTAS>(tmem_base, 512u);
}

// ---- host side -----------------------------------------------------------------------------------------
// The packed (K-major, hi | lo) weights of a gather GEMM: repacked into the context's scratch, or -- inside a
// cattl3_weights_stable scope -- found in / added to the scope's slots.
static int packed_weights(cattl3_ctx* ctx, const GatherGeom& gg, int r_pad, int j_pad, int tapbox, const float* w, long long w_elems,
		float** w_packed) {
	const size_t bytes = (size_t) w_elems * 8;
	float* dst = nullptr;
	if (ctx->pack_stable) {
		cattl3_ctx::PackKey key;
		memset(&key, 0, sizeof(key));
		key.w = w; key.RH = gg.RH; key.RW = gg.RW; key.SC = gg.SC; key.J = gg.J; key.r_pad = r_pad; key.j_pad = j_pad; key.tapbox = tapbox;
		key.w_off = gg.w_off; key.w_stap = gg.w_stap; key.w_srw = gg.w_srw; key.w_sr = gg.w_sr; key.w_sj = gg.w_sj;
		for (int i = 0; i < ctx->pack_used; ++i) {
			if (memcmp(&ctx->pack_slots[i].key, &key, sizeof(key)) == 0) {
				*w_packed = (float*) ctx->pack_slots[i].buf;
				return CATTL3_OK;
			}
		}
		if (ctx->pack_used < cattl3_ctx::MAX_PACK_SLOTS) {
			cattl3_ctx::PackSlot& slot = ctx->pack_slots[ctx->pack_used];
			if (slot.bytes < bytes) {
				if (ctx->capturing) {
					set_error("packed-weight slot %d would have to grow during a graph capture", ctx->pack_used);
					return CATTL3_ERR_UNSUPPORTED;
				}
				if (slot.buf) CATTL3_CUDA(cudaFree(slot.buf));   // (synchronises: nothing in flight still reads it)
				slot.buf = nullptr; slot.bytes = 0;
				CATTL3_CUDA(cudaMalloc(&slot.buf, bytes));
				slot.bytes = bytes;
			}
			slot.key = key;
			++ctx->pack_used;
			dst = (float*) slot.buf;
		}
	}
	if (!dst) {
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->tc_w, &ctx->tc_w_bytes, bytes));
		dst = (float*) ctx->tc_w;
	}
	pack_weights_kernel<<<ew_grid(ctx, w_elems, 256), 256, 0, ctx->stream>>>(gg, r_pad, j_pad, tapbox, w, dst);
	CATTL3_LAUNCHED(ctx);
	*w_packed = dst;
	return CATTL3_OK;
}

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

static int round_up(int v, int m) { return (v + m - 1) / m * m; }
// CATTL3_TC_PAIRS: bit 0 = CTA pairs in the gather GEMM, bit 1 = in the weight gradient (default: both)
static int pair_mask() {
	static const int mask = getenv("CATTL3_TC_PAIRS") ? atoi(getenv("CATTL3_TC_PAIRS")) : 3;
	return mask;
}

// <<<grid, threads, smem, stream>>> with clusters of `ctas` consecutive CTAs (a CTA pair must be the two SMs of one TPC)
template<typename... P, typename... A>
static cudaError_t launch_clustered(void (*kern)(P...), int grid, int threads, size_t smem, cudaStream_t stream, int ctas, A... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned) grid); cfg.blockDim = dim3((unsigned) threads);
	cfg.dynamicSmemBytes = smem; cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned) ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kern, args...);
}
constexpr int TC_SMEM_LIMIT = 227 * 1024 - 2560;

bool tc_gather_gemm_supported(const cattl3_ctx*, const GatherGeom& gg) {
	// 32-row TMA boxes along n; no per-pixel divisibility tests (strided transposed gathers go to SIMT)
	if (gg.N % 32 != 0 || gg.denh != 1 || gg.denw != 1) return false;
	// tiny reduce / filter counts cannot fill a tensor-core tile: those layers are HBM / latency
	// bound and stay on the SIMT kernel ("Tiny-channel configs")
	if (gg.J < 16) return false;
	// few reduce channels are zero-padded to a 16-channel k-block (TMA fills out-of-bounds channels with zeros): worth it
	// when the padding is cheaper than leaving the tensor cores -- a many-tap stem convolution (7x7 over 3 channels:
	// 5.3x padded MMAs still beat the FFMA kernel), not a 1x1 over 3 channels
	if (gg.SC < 16 && (gg.SC < 3 || gg.RH * gg.RW < 9 || getenv("CATTL3_NO_TC_STEM"))) return false;
	return get_encode() != nullptr;
}

// col_stats[k * J + j] = sum over the CTAs' partials, in CTA order (deterministic).
__global__ void __launch_bounds__(256) colstats_reduce_tc_kernel(const double* __restrict__ partial, int ctas, int j_pad, int J,
		double* __restrict__ col_stats) {
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= 2 * J) return;
	const int k = i / J, j = i % J;
	double s = 0.0;
	for (int c = 0; c < ctas; ++c) s += partial[((long long) c * 2 + k) * j_pad + j];
	col_stats[i] = s;
}

static bool tc_rows_gemm_applies(const GatherGeom& gg, int bias_mode, const EpilogueArgs* ep);
static int tc_rows_gemm_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* w, const float* bias,
		int bias_mode, float* out);

int tc_gather_gemm_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* w, const float* bias,
		int bias_mode, float* out, const EpilogueArgs* ep) {
	CATTL3_REQUIRE(aligned16(src) && (!out || aligned16(out)), "tcgen05 path needs 16-byte aligned tensors");
	if (out && tc_rows_gemm_applies(gg, bias_mode, ep))
		return tc_rows_gemm_f32(ctx, gg, src, w, bias, bias_mode, out);
	const bool want_stats = ep && ep->col_stats;
	const bool want_act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || want_act, "gather GEMM: no output tensor");
	CATTL3_REQUIRE(!want_stats || bias_mode == 1, "column statistics need a per-column bias");
	// 256-wide tiles leave room for ONE accumulator only, so the epilogue sits on the critical path: with column
	// statistics (the longest epilogue) take 128-wide tiles, whose two accumulators let it hide behind the next
	// tile's MMAs (profiles/README.md, r1e)
	const bool want_stats_early = ep && ep->col_stats;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	// CTA pairs (cta_group::2, M = 256): whenever there are two 128-row tiles to pair and the filter tile splits into two
	// halves of whole 8-row swizzle atoms with N % 32 == 0 (CATTL3_TC_1CTA=1 keeps single CTAs, for A/B measurements)
	static const bool no_pairs = (pair_mask() & 1) == 0;
	const int ctas = (!no_pairs && M > TC_BM && (gg.J >= 256 || round_up(gg.J, 16) % 32 == 0)) ? 2 : 1;
	const int BN = gg.J >= 256 ? (want_stats_early ? 128 : 256) : round_up(gg.J, 16);
	// 32-element k-blocks (128 B weight rows) where shared and tensor memory allow four stages of them
	// few channels with unit tap steps along W: one TMA box [nb n][1][RW][C] per tap ROW instead of a zero-padded 16-channel
	// box per tap (a 7 x 7 x 3 stem: 7 k-blocks of 21 useful rows instead of 49 of 3)
	const bool tapbox = gg.SC < 16 && gg.bw == 1 && gg.RW > 1 && gg.RW * gg.SC <= 32 && !getenv("CATTL3_NO_TAPBOX");
	const int KB = tapbox ? (gg.RW * gg.SC > 16 ? 32 : 16) : (BN / ctas <= 128 && gg.SC > 16) ? 32 : 16;
	const int r_pad = tapbox ? KB : round_up(gg.SC, KB);
	// the A operand goes through registers into tensor memory, so its shared-memory image needs no MMA layout:
	// take the longest contiguous run of batch entries TMA can deliver per row (up to 128 = 512 B)
	const int nb = gg.N % 128 == 0 ? 128 : (gg.N % 64 == 0 ? 64 : 32);
	const int j_tiles = (gg.J + BN - 1) / BN;
	const int j_pad = j_tiles * BN;
	const int T = tapbox ? gg.RH : gg.RH * gg.RW;   // k-blocks per channel chunk
	const long long w_elems = (long long) T * j_pad * r_pad;

	float* w_packed = nullptr;
	CATTL3_CHECK(packed_weights(ctx, gg, r_pad, j_pad, tapbox ? 1 : 0, w, w_elems, &w_packed));

	CUtensorMap tm_a, tm_b;
	{
		cuuint64_t dims[4] = { (cuuint64_t) gg.N, (cuuint64_t) gg.SH, (cuuint64_t) gg.SW, (cuuint64_t) gg.SC };
		cuuint64_t str[3] = { (cuuint64_t) gg.N * 4, (cuuint64_t) gg.N * gg.SH * 4, (cuuint64_t) gg.N * gg.SH * gg.SW * 4 };
		cuuint32_t box[4] = { (cuuint32_t) nb, 1, 1, (cuuint32_t) KB };
		if (tapbox) { box[2] = (cuuint32_t) gg.RW; box[3] = (cuuint32_t) gg.SC; }
		CATTL3_CHECK(encode_map(&tm_a, src, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));
	}
	{
		cuuint64_t dims[4] = { (cuuint64_t) r_pad, (cuuint64_t) j_pad, (cuuint64_t) T, 2 };
		cuuint64_t str[3] = { (cuuint64_t) r_pad * 4, (cuuint64_t) r_pad * j_pad * 4, (cuuint64_t) w_elems * 4 };
		cuuint32_t box[4] = { (cuuint32_t) KB, (cuuint32_t) (BN / ctas), 1, 1 };
		CATTL3_CHECK(encode_map(&tm_b, w_packed, 4, dims, str, box,
				KB == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B));
	}

	TcGemmParams p;
	p.N = gg.N; p.OH = gg.OH; p.OW = gg.OW; p.J = gg.J; p.RH = gg.RH; p.RW = tapbox ? 1 : gg.RW;
	p.a_rows = tapbox ? gg.RW * gg.SC : KB;
	p.ah = gg.ah; p.bh = gg.bh; p.ch = gg.ch; p.aw = gg.aw; p.bw = gg.bw; p.cw = gg.cw;
	p.M = M;
	p.h0 = gg.out_h0; p.hs = gg.out_hs; p.w0 = gg.out_w0; p.ws = gg.out_ws;
	p.oH = gg.out_H ? gg.out_H : gg.OH; p.oW = gg.out_H ? gg.out_W : gg.OW;
	p.P = (long long) p.oH * p.oW;
	p.out_cs = (long long) gg.N * p.P;
	p.m_tiles = (int) ceil_div(M, TC_BM * ctas); p.j_tiles = j_tiles;
	p.BN = BN; p.r_pad = r_pad; p.nb = nb;
	// one accumulator of 256 columns leaves room for the A stages; narrower tiles double-buffer it so that the
	// epilogue of one tile overlaps the MMAs of the next
	p.nacc = 2 * BN + 4 * 2 * KB <= 512 ? 2 : 1;
	const int stage_bytes = TC_BM * KB * 4 + 2 * (BN / ctas) * KB * 4;
	// converter / epilogue warp sets: a k-block's MMAs take 3 * (KB / 8) * BN / 2 clocks per SM; below ~800 the four
	// converter warps (~350 clocks per k-block each) cannot keep up, so a second set alternates with them; otherwise the
	// second set of four goes to the epilogue of wide tiles (one accumulator: the drain is exposed)
	const bool short_kblocks = 3 * (KB / 8) * BN / 2 < 800;
	const int nconv = short_kblocks ? 8 : 4;
	const int nepi = (!short_kblocks && BN % 32 == 0 && BN >= 128) ? 8 : 4;
	const int stat_bytes = want_stats ? 4 * 2 * BN * 8 + nepi * 32 * 17 * 4 : 0;
	int stages = (TC_SMEM_LIMIT - stat_bytes) / stage_bytes;
	const int tmem_stages = (512 - p.nacc * BN) / (2 * KB);
	if (stages > tmem_stages) stages = tmem_stages;
	if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
	p.stages = stages;
	p.bias_mode = bias_mode; p.bias = bias; p.out = out;
	// at least half of the shared memory: two co-resident CTAs would deadlock allocating 512 TMEM columns each
	size_t smem_bytes = (size_t) stages * stage_bytes + 1024 + 256 + 1024 + stat_bytes;  // alignment slack, barriers, bias, sums
	if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
	const int tiles = p.m_tiles * p.j_tiles;
	const int grid = ctas * (tiles < ctx->sm_count / ctas ? tiles : ctx->sm_count / ctas);
	p.act_kind = want_act ? ep->act_kind : CATTL3_ACT_NONE;
	p.act_param = want_act ? (float) ep->act_param : 0.f;
	p.act_out = want_act ? (float*) ep->act_out : nullptr;
	p.stat_partial = nullptr;
	if (want_stats) {
		const size_t bytes = (size_t) grid * 2 * j_pad * sizeof(double);
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->stat_ws, &ctx->stat_ws_bytes, bytes));
		CATTL3_CUDA(cudaMemsetAsync(ctx->stat_ws, 0, bytes, ctx->stream));
		p.stat_partial = (double*) ctx->stat_ws;
	}
	typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const TcGemmParams);
	GemmKernel kern;
	if (nconv == 8)
		kern = ctas == 2 ? (KB == 32 ? (GemmKernel) tc_gather_gemm_kernel<32, 2, 4, 8> : tc_gather_gemm_kernel<16, 2, 4, 8>)
				: (KB == 32 ? (GemmKernel) tc_gather_gemm_kernel<32, 1, 4, 8> : tc_gather_gemm_kernel<16, 1, 4, 8>);
	else if (nepi == 8)
		kern = ctas == 2 ? (KB == 32 ? (GemmKernel) tc_gather_gemm_kernel<32, 2, 8, 4> : tc_gather_gemm_kernel<16, 2, 8, 4>)
				: (KB == 32 ? (GemmKernel) tc_gather_gemm_kernel<32, 1, 8, 4> : tc_gather_gemm_kernel<16, 1, 8, 4>);
	else
		kern = ctas == 2 ? (KB == 32 ? (GemmKernel) tc_gather_gemm_kernel<32, 2, 4, 4> : tc_gather_gemm_kernel<16, 2, 4, 4>)
				: (KB == 32 ? (GemmKernel) tc_gather_gemm_kernel<32, 1, 4, 4> : tc_gather_gemm_kernel<16, 1, 4, 4>);
	CATTL3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
	CATTL3_CUDA(launch_clustered(kern, grid, (2 + nepi + nconv) * 32, smem_bytes, ctx->stream, ctas, tm_a, tm_b, p));
	CATTL3_LAUNCHED(ctx);
	if (want_stats) {
		colstats_reduce_tc_kernel<<<(unsigned) ceil_div(2 * gg.J, 256), 256, 0, ctx->stream>>>(p.stat_partial, grid, j_pad,
				gg.J, ep->col_stats);
		CATTL3_LAUNCHED(ctx);
	}
	return CATTL3_OK;
}

// ---- the row-sharing gather GEMM ------------------------------------------------------------------------------------
// A stride-1 gather with few output channels (the input gradient of config 2: 256 -> 64 channels; the 64 -> 64 layers of
// config 4) is bound by L2 -> SM bandwidth, not by the tensor pipe: every 128 x 32 tile of the big operand feeds only
// 3 * 64 columns of MMAs, it is fetched once per tap (nine times), and the weights are streamed again for every tile --
// measured 37 B/clk/SM against the ~42 B/clk/SM the L2 delivers (profiles/README.md, r2).  Here a tile is R = 3
// vertically adjacent output pixels (x 256 batch entries for the pair) with one accumulator each: the R + RH - 1 source
// rows they need are fetched ONCE per tap column and each source tile feeds the accumulator of every output row it
// belongs to (15 fetches instead of 27), and the RH weight tiles of a tap column stay in shared memory for those
// R + RH - 1 k-blocks (24 fetches instead of 72): 2.1x fewer bytes per MMA.
// Always a CTA pair; warps: 0 = TMA, 1 = MMA (leader), 2-5 = epilogue, 6-9 = converters.
struct TcRowsParams {
	int N, OH, OW, J, RH, RW;
	int bh, ch, aw, bw, cw;   // source row = oh + rh * bh + ch (bh = +-1), source column = ow * aw + rw * bw + cw
	int BN, r_pad, R, IR, ohg, nblocks, tiles;
	int bias_mode;
	const float* bias;
	float* out;
	long long out_cs;
};
constexpr int ROWS_KB = 32, ROWS_ASTAGES = 5, ROWS_BSLOTS = 2, ROWS_R = 3;

__global__ void __launch_bounds__(320, 1) tc_rows_gemm_kernel(const __grid_constant__ CUtensorMap tm_a,
		const __grid_constant__ CUtensorMap tm_b, const TcRowsParams p) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
	constexpr int A_BYTES = TC_BM * ROWS_KB * 4;
	const int b_bytes = p.BN / 2 * ROWS_KB * 4;         // one weight tile (hi or lo) of this CTA's half of the columns
	const int bslot_bytes = p.RH * 2 * b_bytes;         // [tap row][hi | lo]
	uint8_t* a_smem = smem;
	uint8_t* b_smem = smem + ROWS_ASTAGES * A_BYTES;
	uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + ROWS_BSLOTS * bslot_bytes);
	uint64_t* a_full = bars;
	uint64_t* a_ready = bars + 8;
	uint64_t* a_empty = bars + 16;
	uint64_t* b_full = bars + 24;
	uint64_t* b_empty = bars + 26;
	uint64_t* acc_full = bars + 28;
	uint64_t* acc_empty = bars + 29;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
	float* sbias = reinterpret_cast<float*>(bars + 32);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t rank = cluster_ctarank();
	const int unit = blockIdx.x / 2, units = gridDim.x / 2;
	const int chunks = p.r_pad / ROWS_KB;
	const int base_off = p.bh > 0 ? 0 : -(p.RH - 1);
	const uint32_t a_col0 = (uint32_t) (p.R * p.BN);

	if (warp == 0 && elect_one()) {
		tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b);
		for (int s = 0; s < ROWS_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_ready[s], 8); mbar_init(&a_empty[s], 1); }
		for (int s = 0; s < ROWS_BSLOTS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
		mbar_init(acc_full, 1); mbar_init(acc_empty, 8);
		fence_barrier_init();
	}
	if (p.bias_mode == 1)
		for (int c = threadIdx.x; c < p.BN; c += blockDim.x) sbias[c] = c < p.J ? __ldg(p.bias + c) : 0.f;
	if (warp == 1) tmem_alloc<2>(tmem_slot, 512u);
	tc_fence_before();
	cluster_sync();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		if (elect_one()) {
			int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
			for (int tile = unit; tile < p.tiles; tile += units) {
				const int nb_i = tile % p.nblocks, rest = tile / p.nblocks;
				const int oh0 = (rest % p.ohg) * p.R, ow = rest / p.ohg;
				const int n0 = nb_i * (2 * TC_BM) + (int) rank * TC_BM;
				const int ih_base = oh0 + p.ch + base_off;
				for (int c0 = 0; c0 < p.r_pad; c0 += ROWS_KB) {
					for (int rw = 0; rw < p.RW; ++rw) {
						const int iw = ow * p.aw + rw * p.bw + p.cw;
						mbar_wait(&b_empty[bs], bph ^ 1);
						mbar_expect_tx(&b_full[bs], (uint32_t) bslot_bytes);
						for (int rh = 0; rh < p.RH; ++rh) {
							uint8_t* dst = b_smem + bs * bslot_bytes + rh * 2 * b_bytes;
							tma_load_4d(dst, &tm_b, &b_full[bs], c0, (int) rank * (p.BN / 2), rh + p.RH * rw, 0);
							tma_load_4d(dst + b_bytes, &tm_b, &b_full[bs], c0, (int) rank * (p.BN / 2), rh + p.RH * rw, 1);
						}
						if (++bs == ROWS_BSLOTS) { bs = 0; bph ^= 1; }
						for (int t = 0; t < p.IR; ++t) {
							mbar_wait(&a_empty[as], aph ^ 1);
							mbar_expect_tx(&a_full[as], (uint32_t) A_BYTES);
							tma_load_4d(a_smem + as * A_BYTES, &tm_a, &a_full[as], n0, ih_base + t, iw, c0);
							if (++as == ROWS_ASTAGES) { as = 0; aph ^= 1; }
						}
					}
				}
			}
		}
	} else if (warp == 1) {
		if (rank == 0 && elect_one()) {
			const uint32_t idesc = make_idesc_tf32(p.BN, 2 * TC_BM);
			int as = 0; uint32_t aph = 0; int bs = 0;
			uint32_t acc_ph = 0;
			for (int tile = unit; tile < p.tiles; tile += units) {
				const int oh0 = ((tile / p.nblocks) % p.ohg) * p.R;
				mbar_wait(acc_empty, acc_ph ^ 1);
				tc_fence_after();
				uint32_t started = 0;
				for (int c = 0; c < chunks; ++c) {
					for (int rw = 0; rw < p.RW; ++rw) {
						const uint32_t bslot = smem_u32(b_smem + bs * bslot_bytes);
						for (int t = 0; t < p.IR; ++t) {
							mbar_wait(&a_ready[as], aph);
							tc_fence_after();
							const uint32_t a_hi = tmem_base + a_col0 + (uint32_t) (as * 2 * ROWS_KB);
							const uint32_t a_lo = a_hi + ROWS_KB;
							for (int i = 0; i < p.R; ++i) {
								const int d = base_off + t - i;
								const int rh = p.bh > 0 ? d : -d;
								if (oh0 + i >= p.OH || rh < 0 || rh >= p.RH) continue;
								const uint32_t acc = tmem_base + (uint32_t) (i * p.BN);
								const uint32_t b_hi = bslot + (uint32_t) (rh * 2 * b_bytes);
								const uint32_t b_lo = b_hi + (uint32_t) b_bytes;
								uint32_t accumulate = (started >> i) & 1u;
								#pragma unroll
								for (int pass = 0; pass < 3; ++pass) {
									const uint32_t a = pass == 0 ? a_lo : a_hi;
									const uint32_t b = pass == 1 ? b_lo : b_hi;
									#pragma unroll
									for (int ks = 0; ks < ROWS_KB / 8; ++ks) {
										umma_tf32_ts<2>(acc, a + 8 * ks, kmajor_desc<ROWS_KB>(b, ks), idesc, accumulate);
										accumulate = 1u;
									}
								}
								started |= 1u << i;
							}
							umma_commit<2>(&a_empty[as]);
							if (++as == ROWS_ASTAGES) { as = 0; aph ^= 1; }
						}
						umma_commit<2>(&b_empty[bs]);   // the tap column's weight tiles have been read
						if (++bs == ROWS_BSLOTS) bs = 0;
					}
				}
				umma_commit<2>(acc_full);
				acc_ph ^= 1;
			}
		}
	} else if (warp < 6) {
		const int q = warp & 3;
		uint32_t acc_ph = 0;
		for (int tile = unit; tile < p.tiles; tile += units) {
			const int nb_i = tile % p.nblocks, rest = tile / p.nblocks;
			const int oh0 = (rest % p.ohg) * p.R, ow = rest / p.ohg;
			const int n = nb_i * (2 * TC_BM) + (int) rank * TC_BM + 32 * q + lane;
			mbar_wait(acc_full, acc_ph);
			tc_fence_after();
			for (int i = 0; i < p.R; ++i) {
				if (oh0 + i >= p.OH) break;
				float* out_m = p.out + n + (long long) p.N * ((oh0 + i) + (long long) p.OH * ow);
				const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + (uint32_t) (i * p.BN);
				for (int c0 = 0; c0 < p.BN; c0 += 16) {
					float v[16];
					tmem_ld_16(taddr + c0, v);
					tmem_ld_wait();
					#pragma unroll
					for (int k = 0; k < 16; ++k) {
						if (c0 + k < p.J)
							out_m[p.out_cs * (long long) (c0 + k)] = p.bias_mode == 1 ? v[k] + sbias[c0 + k] : v[k];
					}
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive_cta(acc_empty, 0);
			acc_ph ^= 1;
		}
	} else {
		const int q = warp & 3;
		const int row = 32 * q + lane;
		const uint32_t row_off = (uint32_t) (row * 4);            // the box is [32 k][128 n]: row m of the tile = entry m of each k-row
		constexpr uint32_t k_stride = TC_BM * 4;
		int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
		for (int tile = unit; tile < p.tiles; tile += units) {
			for (int c = 0; c < chunks; ++c) {
				for (int rw = 0; rw < p.RW; ++rw) {
					for (int t = 0; t < p.IR; ++t) {
						mbar_wait(&a_full[as], aph);
						if (t == 0) mbar_wait(&b_full[bs], bph);   // this CTA's half of the tap column's weights has landed too
						const uint8_t* grp = a_smem + as * A_BYTES + row_off;
						const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + a_col0 + (uint32_t) (as * 2 * ROWS_KB);
						#pragma unroll
						for (int k0 = 0; k0 < ROWS_KB; k0 += 16) {
							uint32_t hi[16], lo[16];
							#pragma unroll
							for (int k = 0; k < 16; ++k) {
								const float v = *reinterpret_cast<const float*>(grp + (uint32_t) (k0 + k) * k_stride);
								hi[k] = tf32_hi_bits(v);
								lo[k] = tf32_lo_bits(v);
							}
							tmem_st_16(taddr + k0, hi);
							tmem_st_16(taddr + ROWS_KB + k0, lo);
						}
						tmem_st_wait();
						tc_fence_before();
						__syncwarp();
						if (lane == 0) mbar_arrive_cta(&a_ready[as], 0);
						if (++as == ROWS_ASTAGES) { as = 0; aph ^= 1; }
					}
					if (++bs == ROWS_BSLOTS) { bs = 0; bph ^= 1; }
				}
			}
		}
	}
	tc_fence_before();
	cluster_sync();
	if (warp == 1) tmem_dealloc<2>(tmem_base, 512u);
}

static bool tc_rows_gemm_applies(const GatherGeom& gg, int bias_mode, const EpilogueArgs* ep) {
	static const bool off = getenv("CATTL3_NO_ROWS") != nullptr || (pair_mask() & 1) == 0;
	if (off || ep || bias_mode > 1) return false;
	if (gg.ah != 1 || gg.aw != 1 || (gg.bh != 1 && gg.bh != -1) || gg.denh != 1 || gg.denw != 1) return false;
	if (gg.N % (2 * TC_BM) != 0 || gg.J > 64 || gg.J < 16 || gg.SC < 32 || gg.RH < 2 || gg.RH > 4 || gg.OH < ROWS_R) return false;
	if (gg.out_H != 0 || gg.w_off != 0 || gg.w_srw != 0) return false;
	return true;
}

static int tc_rows_gemm_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* w, const float* bias,
		int bias_mode, float* out) {
	const int T = gg.RH * gg.RW;
	const int BN = round_up(gg.J, 32);
	const int r_pad = round_up(gg.SC, ROWS_KB);
	const long long w_elems = (long long) T * BN * r_pad;
	float* w_packed = nullptr;
	CATTL3_CHECK(packed_weights(ctx, gg, r_pad, BN, 0, w, w_elems, &w_packed));
	CUtensorMap tm_a, tm_b;
	{
		cuuint64_t dims[4] = { (cuuint64_t) gg.N, (cuuint64_t) gg.SH, (cuuint64_t) gg.SW, (cuuint64_t) gg.SC };
		cuuint64_t str[3] = { (cuuint64_t) gg.N * 4, (cuuint64_t) gg.N * gg.SH * 4, (cuuint64_t) gg.N * gg.SH * gg.SW * 4 };
		cuuint32_t box[4] = { (cuuint32_t) TC_BM, 1, 1, (cuuint32_t) ROWS_KB };
		CATTL3_CHECK(encode_map(&tm_a, src, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));
	}
	{
		cuuint64_t dims[4] = { (cuuint64_t) r_pad, (cuuint64_t) BN, (cuuint64_t) T, 2 };
		cuuint64_t str[3] = { (cuuint64_t) r_pad * 4, (cuuint64_t) r_pad * BN * 4, (cuuint64_t) w_elems * 4 };
		cuuint32_t box[4] = { (cuuint32_t) ROWS_KB, (cuuint32_t) (BN / 2), 1, 1 };
		CATTL3_CHECK(encode_map(&tm_b, w_packed, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
	}
	TcRowsParams p;
	p.N = gg.N; p.OH = gg.OH; p.OW = gg.OW; p.J = gg.J; p.RH = gg.RH; p.RW = gg.RW;
	p.bh = gg.bh; p.ch = gg.ch; p.aw = gg.aw; p.bw = gg.bw; p.cw = gg.cw;
	p.BN = BN; p.r_pad = r_pad; p.R = ROWS_R; p.IR = ROWS_R + gg.RH - 1;
	p.ohg = (gg.OH + ROWS_R - 1) / ROWS_R;
	p.nblocks = gg.N / (2 * TC_BM);
	p.tiles = p.nblocks * p.ohg * gg.OW;
	p.bias_mode = bias_mode; p.bias = bias; p.out = out;
	p.out_cs = (long long) gg.N * gg.OH * gg.OW;
	const int pairs = p.tiles < ctx->sm_count / 2 ? p.tiles : ctx->sm_count / 2;
	size_t smem_bytes = (size_t) ROWS_ASTAGES * TC_BM * ROWS_KB * 4 + (size_t) ROWS_BSLOTS * gg.RH * 2 * (BN / 2) * ROWS_KB * 4 +
			1024 + 256 + 1024;
	if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
	CATTL3_REQUIRE(smem_bytes <= 227 * 1024, "rows GEMM: shared memory");
	CATTL3_CUDA(cudaFuncSetAttribute(tc_rows_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
	CATTL3_CUDA(launch_clustered(tc_rows_gemm_kernel, 2 * pairs, 320, smem_bytes, ctx->stream, 2, tm_a, tm_b, p));
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- the weight-gradient kernel ---------------------------------------------------------------------
// dw(tap, r, j) += sum_m src(m, tap, r) * plain(m, j): a GEMM whose reduction runs over m = N*OH*OW, so m
// is the K dimension and both operands are K-major where they lie (m is contiguous in HBM for both).
//   A (tensor memory, 128 rows = output channels j): one 2-D TMA box [32 m][128 j] of the plain tensor
//     (dY), split hi / lo into TMEM by the converter warps;
//   B (shared memory, up to 192 rows = (tap, channel) pairs): per tap one box [32 m][RB channels] of the
//     gathered tensor at that tap's coordinate.  The tensor core truncates the raw tile itself (= hi); the
//     converters write the lo tile next to it (element-wise, so the swizzled layout is preserved).
// One k-block = 32 consecutive batch entries of one pixel.  What limits this kernel is the NUMBER OF TMA ROWS:
// every row of either operand lies in a different 2 MB page (the channel stride is N*H*W*4 bytes), and the
// TMA unit sustains only ~1 such row per 6-8 clocks (measured; profiles/README.md).  Hence raw-only loads
// (no second, pre-split copy of an operand) and 128-byte rows.
// Each CTA owns one (j tile, column tile) and one contiguous range of k-blocks (split-K); the fp32 partial
// tile goes to scratch and wgrad_reduce_tc_kernel adds the partials to dw in split order (deterministic,
// accumulating: Parameters::accumulate_grad, C-ATTL3/parameters/StandardParameters.hpp:115-123).
struct TcWgradParams {
	int N, OH, OW, R, J, RH, RW;
	int ah, bh, ch, aw, bw, cw;
	int tapbox;        // few channels: ONE box [32 m][RH][RWp][R] per k-block holds every (tap, channel) row; RWp = tap_wp
	int tap_wp;
	int RB;            // channel rows per TMA box (16 / 32 / 64)
	int rchunks;       // r_pad / RB
	int boxes;         // T * rchunks
	int boxes_per_tile, col_tiles, j_tiles, splits;
	int BNW;           // boxes_per_tile * RB: columns of the accumulator (<= 192)
	long long mgroups, mg_per_split, mg_base;   // the k-blocks [mg_base, mg_base + mgroups) of 32 rows are reduced
	int stages;
	int flush;         // k-blocks accumulated in TMEM before the partial tile is folded into fp32 scratch
	long long w_stap, w_sr, w_sj, dw_elems;
	float* partial;    // [split][tile][BNW columns][128 rows]
	float* db_partial; // [split][j_tiles * 128] column sums of the plain tensor (bias gradient), or null
	const float* plain;  // the A operand: M x J, m contiguous
	long long M;
};

// Tensor-core accumulation truncates (round-toward-zero) at every MMA, so a reduction of n MMA steps
// carries a bias of ~n * 2^-24 relative (measured: 2e-4 on the 10^4-step config-2 weight gradient).
// The kernel therefore closes the TMEM accumulator every `flush` k-blocks and the epilogue warps fold it
// into a CTA-private fp32 tile with round-to-nearest adds (TMA and the converters keep filling the stages
// meanwhile; the MMA warp idles for the few thousand clocks of the drain, ~4 % of a flush period).
constexpr int WG_KB = 32;
constexpr uint32_t WG_A_COL0 = 192;   // TMEM: accumulator in columns [0, 192), A stages (64 columns each) from column 192: five fit
constexpr int WG_RING = 4;             // k-blocks of dY in flight per converter warp (cp.async ring of 4 KB patches)
template<int CTAS>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ CUtensorMap tm_b,
		const TcWgradParams p) {
	extern __shared__ __align__(1024) uint8_t smem_raw[];
	uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
	const int b_bytes = p.BNW / CTAS * WG_KB * 4;   // a CTA of a pair holds half of the accumulator's columns (= rows of B)
	const int stage_bytes = 2 * b_bytes;   // [B raw][B lo]; the A operand has its own cp.async ring
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t) p.stages * stage_bytes);
	uint64_t* full = bars;
	uint64_t* ready = bars + TC_MAX_STAGES;
	uint64_t* empty = bars + 2 * TC_MAX_STAGES;
	uint64_t* acc_full = bars + 3 * TC_MAX_STAGES;
	uint64_t* acc_empty = acc_full + 2;
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	// A unit = one CTA, or one pair working on 256 output channels (the leader on the first 128, its peer on the next)
	// and each loading half of the unit's boxes.  p.j_tiles counts 128-row tiles (a multiple of CTAS).
	const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
	const int tiles = p.col_tiles * p.j_tiles;
	const int unit_tiles = tiles / CTAS;
	const int unit = blockIdx.x / CTAS;
	const int utile = unit % unit_tiles, z = unit / unit_tiles;
	const int ct = utile % p.col_tiles, jt = (utile / p.col_tiles) * CTAS + (int) rank;  // units sharing a dY tile are neighbours
	const int tile = ct + p.col_tiles * jt;
	const int my_boxes = p.boxes_per_tile / CTAS;
	const int box0 = ct * p.boxes_per_tile + (int) rank * my_boxes;
	int nboxes = p.boxes - box0;
	if (nboxes > my_boxes) nboxes = my_boxes;
	if (nboxes < 0) nboxes = 0;
	const long long mg0 = p.mg_base + (long long) z * p.mg_per_split;
	long long mg1 = mg0 + p.mg_per_split;
	if (mg1 > p.mg_base + p.mgroups) mg1 = p.mg_base + p.mgroups;
	const long long kblocks = mg1 > mg0 ? mg1 - mg0 : 0;
	const long long chunks = (kblocks + p.flush - 1) / p.flush;

	if (warp == 0 && elect_one()) {
		tma_prefetch_desc(&tm_b);
		for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], 4 * CTAS); mbar_init(&empty[s], 1); }
		mbar_init(&acc_full[0], 1); mbar_init(&acc_empty[0], 4 * CTAS);
		fence_barrier_init();
	}
	if (warp == 1) tmem_alloc<CTAS>(tmem_slot, 512u);
	tc_fence_before();
	if (CTAS == 2) cluster_sync(); else __syncthreads();
	tc_fence_after();
	const uint32_t tmem_base = *tmem_slot;

	if (warp == 0) {
		if (elect_one()) {
			// One thread feeds the whole pipeline, so its loop is kept to a handful of instructions per box: the boxes'
			// tap offsets and channel origins are tabulated once, and the k-block's (n0, oh, ow) advances incrementally
			// (the 64-bit divisions this loop used to do per k-block made it as slow as the MMAs it feeds).
			int s = 0; uint32_t ph = 0;
			const uint32_t tx = p.tapbox ? (uint32_t) (p.RH * p.tap_wp * p.R * WG_KB * 4) : (uint32_t) (nboxes * p.RB * WG_KB * 4);
			int* box_tab = reinterpret_cast<int*>(bars + 40);   // [12][3]: dh, dw, c0 (behind the barriers and the TMEM slot)
			for (int bx = 0; bx < nboxes; ++bx) {
				const int box = box0 + bx;
				const int tap = box / p.rchunks;
				const int rh = tap % p.RH, rw = tap / p.RH;
				box_tab[3 * bx + 0] = rh * p.bh + p.ch;
				box_tab[3 * bx + 1] = rw * p.bw + p.cw;
				box_tab[3 * bx + 2] = (box % p.rchunks) * p.RB;
			}
			if (p.tapbox) { box_tab[0] = p.ch; box_tab[1] = p.cw; box_tab[2] = 0; }   // the one box starts at tap (0, 0), channel 0
			const long long m_first = mg0 * WG_KB;
			int n0 = (int) (m_first % p.N);
			const long long pix0 = m_first / p.N;
			int oh = (int) (pix0 % p.OH), ow = (int) (pix0 / p.OH);
			const uint32_t box_bytes = (uint32_t) (p.RB * WG_KB * 4);
			for (long long kb = 0; kb < kblocks; ++kb) {
				mbar_wait(&empty[s], ph ^ 1);
				uint8_t* st = smem + (size_t) s * stage_bytes;
				if (tx) mbar_expect_tx(&full[s], tx); else mbar_arrive(&full[s]);
				const int ih0 = oh * p.ah, iw0 = ow * p.aw;
				for (int bx = 0; bx < nboxes; ++bx)
					tma_load_4d(st + bx * box_bytes, &tm_b, &full[s], n0, ih0 + box_tab[3 * bx], iw0 + box_tab[3 * bx + 1], box_tab[3 * bx + 2]);
				if (++s == p.stages) { s = 0; ph ^= 1; }
				n0 += WG_KB;
				if (n0 >= p.N) { n0 = 0; if (++oh == p.OH) { oh = 0; ++ow; } }
			}
		}
	} else if (warp == 1) {
		if (rank == 0 && elect_one()) {
			const uint32_t idesc = make_idesc_tf32(p.BNW, TC_BM * CTAS);
			int s = 0; uint32_t ph = 0;
			uint32_t acc_ph = 0;
			long long kb = 0;
			for (long long c = 0; c < chunks; ++c) {
				mbar_wait(&acc_empty[0], acc_ph ^ 1);
				tc_fence_after();
				const uint32_t d = tmem_base;
				const long long kend = kb + p.flush < kblocks ? kb + p.flush : kblocks;
				for (bool first = true; kb < kend; ++kb) {
					mbar_wait(&ready[s], ph);
					tc_fence_after();
					const uint32_t b_hi = smem_u32(smem + (size_t) s * stage_bytes);
					const uint32_t b_lo = b_hi + b_bytes;
					const uint32_t a_hi = tmem_base + WG_A_COL0 + (uint32_t) (s * 2 * WG_KB);
					const uint32_t a_lo = a_hi + WG_KB;
					#pragma unroll
					for (int pass = 0; pass < 3; ++pass) {
						const uint32_t a = pass == 0 ? a_lo : a_hi;
						const uint32_t b = pass == 1 ? b_lo : b_hi;
						#pragma unroll
						for (int ks = 0; ks < WG_KB / 8; ++ks)
							umma_tf32_ts<CTAS>(d, a + 8 * ks, kmajor_desc<WG_KB>(b, ks), idesc, (first && pass == 0 && ks == 0) ? 0u : 1u);
					}
					first = false;
					umma_commit<CTAS>(&empty[s]);
					if (++s == p.stages) { s = 0; ph ^= 1; }
				}
				umma_commit<CTAS>(&acc_full[0]);
				acc_ph ^= 1;
			}
		}
	} else if (warp < 6) {
		const int q = warp & 3;
		const int row = 32 * q + lane;  // accumulator lane = output channel within the j tile
		float* dst = p.partial + ((long long) z * tiles + tile) * p.BNW * 128 + row;
		if (chunks == 0) {
			for (int c0 = 0; c0 < p.BNW; ++c0) dst[(long long) c0 * 128] = 0.f;
		}
		uint32_t acc_ph = 0;
		for (long long c = 0; c < chunks; ++c) {
			mbar_wait(&acc_full[0], acc_ph);
			tc_fence_after();
			const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16);
			for (int c0 = 0; c0 < p.BNW; c0 += 16) {
				float v[16];
				tmem_ld_16(taddr + c0, v);
				tmem_ld_wait();
				// fold into the CTA's fp32 partial tile: the first chunk stores, the later ones add with fire-and-forget
				// reductions (RED.ADD.F32, round to nearest; one writer per address, program order = flush order, so the
				// sum is the same sequence of additions as a read-modify-write, without its L2 round trip per 16 columns,
				// which cost 12 % of the kernel)
				if (c > 0) {
					#pragma unroll
					for (int i = 0; i < 16; ++i) atomicAdd(&dst[(long long) (c0 + i) * 128], v[i]);
				} else {
					#pragma unroll
					for (int i = 0; i < 16; ++i) dst[(long long) (c0 + i) * 128] = v[i];
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { if (CTAS == 2) mbar_arrive_cta(&acc_empty[0], 0); else mbar_arrive(&acc_empty[0]); }
			acc_ph ^= 1;
		}
	} else {
		// converters.  A (dY): each warp owns 32 rows (= output channels) of the tile.  A k-block of a row is 128
		// contiguous bytes in HBM.  TMA pays 6-9 clocks for every such row (each lies in another 2 MB page), so the
		// warp fetches its rows itself with 16-byte cp.async copies -- eight lanes per row, four whole 128-byte
		// lines per instruction, WG_RING k-blocks in flight -- into a private ring of 4 KB patches whose 16-byte
		// chunks are XOR-swizzled with the row so that the later one-row-per-lane read is conflict free.  The values
		// are split and stored to TMEM with tcgen05.st.
		// B: lo tile of the gathered operand, element-wise in shared memory.
		const int q = warp & 3;
		const int row = 32 * q + lane;
		const int tid = threadIdx.x - 192;
		uint8_t* ring = smem + (size_t) p.stages * stage_bytes + 512 + q * (WG_RING * 4096);
		const int lrow = lane >> 3, lchunk = lane & 7;   // this lane copies chunk lchunk of rows lrow, lrow + 4, ...
		const float* a_src[8];
		uint32_t a_dst[8], a_bytes[8];
		#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const int r = 4 * i + lrow;
			const int j = jt * TC_BM + 32 * q + r;
			a_bytes[i] = j < p.J ? 16u : 0u;   // rows past the last output channel are zero filled
			a_src[i] = p.plain + (long long) (j < p.J ? j : 0) * p.M + mg0 * WG_KB + 4 * lchunk;
			a_dst[i] = (uint32_t) (r * 128 + ((lchunk ^ (r & 7)) << 4));
		}
		#pragma unroll
		for (int d = 0; d < WG_RING - 1; ++d) {
			if (d < kblocks) {
				#pragma unroll
				for (int i = 0; i < 8; ++i) cp_async_16(ring + d * 4096 + a_dst[i], a_src[i] + (long long) d * WG_KB, a_bytes[i]);
			}
			cp_async_commit();
		}
		// the bias gradient db(j) = sum_m dY(m, j) falls out of the same read: every dY value passes through this
		// thread's registers exactly once (in the CTAs of column tile 0)
		const bool want_db = p.db_partial != nullptr && ct == 0;
		float db_sum = 0.f;
		int s = 0; uint32_t ph = 0;
		int slot = 0;
		for (long long kb = 0; kb < kblocks; ++kb) {
			{
				// refill the slot read in the previous iteration with k-block kb + WG_RING - 1
				const long long nk = kb + WG_RING - 1;
				const int nslot = slot == 0 ? WG_RING - 1 : slot - 1;
				if (nk < kblocks) {
					#pragma unroll
					for (int i = 0; i < 8; ++i) cp_async_16(ring + nslot * 4096 + a_dst[i], a_src[i] + nk * WG_KB, a_bytes[i]);
				}
				cp_async_commit();
			}
			cp_async_wait<WG_RING - 1>();   // this thread's copies of k-block kb have landed ...
			__syncwarp();                   // ... and so have the other lanes'
			const uint8_t* patch = ring + slot * 4096;
			float4 mine[8];
			#pragma unroll
			for (int c = 0; c < 8; ++c)
				mine[c] = *reinterpret_cast<const float4*>(patch + lane * 128 + ((c ^ (lane & 7)) << 4));
			__syncwarp();
			if (++slot == WG_RING) slot = 0;
			mbar_wait(&full[s], ph);   // B landed; the producer saw empty[s], so TMEM slot s is free as well
			uint8_t* st = smem + (size_t) s * stage_bytes;
			const uint32_t taddr = tmem_base + ((uint32_t) (32 * q) << 16) + WG_A_COL0 + (uint32_t) (s * 2 * WG_KB);
			float blk_sum = 0.f;
			#pragma unroll
			for (int h = 0; h < 2; ++h) {
				uint32_t hi[16], lo[16];
				#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const float4 v = mine[4 * h + c];
					blk_sum += (v.x + v.y) + (v.z + v.w);
					hi[4 * c + 0] = tf32_hi_bits(v.x); lo[4 * c + 0] = tf32_lo_bits(v.x);
					hi[4 * c + 1] = tf32_hi_bits(v.y); lo[4 * c + 1] = tf32_lo_bits(v.y);
					hi[4 * c + 2] = tf32_hi_bits(v.z); lo[4 * c + 2] = tf32_lo_bits(v.z);
					hi[4 * c + 3] = tf32_hi_bits(v.w); lo[4 * c + 3] = tf32_lo_bits(v.w);
				}
				tmem_st_16(taddr + 16 * h, hi);
				tmem_st_16(taddr + WG_KB + 16 * h, lo);
			}
			{
				// lo tile of B: 16-byte vectors, this thread's are 2 KB apart; all loads of a batch are issued before the first
				// is used (one shared-memory latency per batch of six instead of one per vector)
				const uint32_t src = smem_u32(st) + (uint32_t) tid * 16u;
				const int vecs = b_bytes >> 11;   // vectors per thread: b_bytes / (128 threads * 16 B); b_bytes is a multiple of 2 KB
				for (int v0 = 0; v0 < vecs; v0 += 6) {
					float4 v[6];
					#pragma unroll
					for (int u = 0; u < 6; ++u)
						if (v0 + u < vecs) v[u] = lds_128(src + (uint32_t) (v0 + u) * 2048u);
					#pragma unroll
					for (int u = 0; u < 6; ++u) {
						if (v0 + u < vecs) {
							float4 t = v[u];
							t.x = __uint_as_float(tf32_lo_bits(t.x)); t.y = __uint_as_float(tf32_lo_bits(t.y));
							t.z = __uint_as_float(tf32_lo_bits(t.z)); t.w = __uint_as_float(tf32_lo_bits(t.w));
							sts_128(src + (uint32_t) b_bytes + (uint32_t) (v0 + u) * 2048u, t);
						}
					}
				}
			}
			db_sum += blk_sum;
			// the lo tile was written through the generic proxy; the MMA (of either CTA of a pair) reads it through the async proxy
			// (.shared::cta: the unqualified fence compiles to MEMBAR.ALL.GPU and took 47 % of the converters' time)
			fence_proxy_async();
			tmem_st_wait();
			tc_fence_before();
			__syncwarp();
			if (lane == 0) { if (CTAS == 2) mbar_arrive_cta(&ready[s], 0); else mbar_arrive(&ready[s]); }
			if (++s == p.stages) { s = 0; ph ^= 1; }
		}
		if (want_db) p.db_partial[(long long) z * p.j_tiles * TC_BM + jt * TC_BM + row] = db_sum;
	}
	tc_fence_before();
	if (CTAS == 2) cluster_sync(); else __syncthreads();
	if (warp == 1) tmem_dealloc<CTAS>(tmem_base, 512u);
}

// dw(tap, r, j) += sum over splits of the scratch tiles; db(j) += sum over splits of the column sums.  Small layers
// (an LSTM gate kernel: 2 304 weights, 128 splits) made the one-thread-per-weight loop over the splits a chain of
// dependent strided loads (20 us for 1 MB); here eight warps of a block share 32 outputs, each summing every eighth
// split, and the eight partial sums are added in a fixed order: deterministic, and a few loads deep.
constexpr int WR_ZL = 8;
__global__ void __launch_bounds__(32 * WR_ZL) wgrad_reduce_tc_kernel(const TcWgradParams p, int T, float* __restrict__ dw,
		float* __restrict__ db) {
	__shared__ float red[WR_ZL][32];
	const long long n_dw = (long long) T * p.R * p.J;
	const long long total = n_dw + (db != nullptr ? p.J : 0);   // the bias gradient's columns follow the weights
	const int tiles = p.col_tiles * p.j_tiles;
	const int lane = threadIdx.x & 31, zl = threadIdx.x >> 5;
	for (long long base = blockIdx.x * 32ll; base < total; base += (long long) gridDim.x * 32) {
		const long long i = base + lane;
		const float* src = nullptr;
		long long z_stride = 0;
		float* dst = nullptr;
		if (i < n_dw) {
			const int j = (int) (i % p.J);
			const int r = (int) ((i / p.J) % p.R);
			const int tap = (int) (i / ((long long) p.J * p.R));
			const int box = tap * p.rchunks + r / p.RB;
			int ct = box / p.boxes_per_tile;
			int col = (box % p.boxes_per_tile) * p.RB + r % p.RB;
			if (p.tapbox) {   // rows of the tap box: tap row fastest, then tap column, then channel (the tensor's own order)
				ct = 0;
				col = tap % p.RH + p.RH * (tap / p.RH + p.tap_wp * r);
			}
			const int jt = j / TC_BM, row = j % TC_BM;
			src = p.partial + ((long long) (ct + p.col_tiles * jt) * p.BNW + col) * 128 + row;
			z_stride = (long long) tiles * p.BNW * 128;
			dst = dw + tap * p.w_stap + r * p.w_sr + j * p.w_sj;
		} else if (i < total) {
			const int j = (int) (i - n_dw);
			src = p.db_partial + j;
			z_stride = (long long) p.j_tiles * TC_BM;
			dst = db + j;
		}
		float s = 0.f;
		if (src) {
			#pragma unroll 4
			for (int z = zl; z < p.splits; z += WR_ZL) s += src[z * z_stride];
		}
		red[zl][lane] = s;
		__syncthreads();
		if (zl == 0 && dst) {
			float t = red[0][lane];
			#pragma unroll
			for (int k = 1; k < WR_ZL; ++k) t += red[k][lane];
			*dst += t;
		}
		__syncthreads();
	}
}

// Few channels (the 3-channel stem of config 4): with undilated taps the RH x RW x C patch of a pixel is a BOX of the
// source tensor, so one TMA box [32 m][RH][RWp][C] delivers every (tap, channel) row of a k-block -- no zero-padded
// 16-channel boxes per tap (5.3x the MMAs and 16 boxes for a 7 x 7 x 3 stem), one accumulator tile for the whole dW.
// RWp >= RW makes the row count a multiple of 8 (whole swizzle atoms); the extra columns map to no weight.
static int tapbox_wp(const GatherGeom& gg) {
	if (gg.SC >= 16 || gg.bh != 1 || gg.bw != 1 || getenv("CATTL3_NO_TC_STEM")) return 0;
	for (int wp = gg.RW; wp <= gg.RW + 8; ++wp) {
		const int rows = gg.RH * wp * gg.SC;
		if (rows % 8 == 0 && rows <= 192 && wp <= gg.SW + 8) return wp;
	}
	return 0;
}

bool tc_wgrad_supported(const cattl3_ctx*, const GatherGeom& gg) {
	if (gg.N % 32 != 0 || gg.denh != 1 || gg.denw != 1) return false;
	if (gg.J < 16) return false;
	if (gg.SC < 16 && (gg.RH * gg.RW < 9 || tapbox_wp(gg) == 0)) return false;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	if (M >= (1ll << 31)) return false;  // TMA coordinates are 32-bit
	return get_encode() != nullptr;
}

int tc_wgrad_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* plain, float* dw, float* db) {
	CATTL3_REQUIRE(aligned16(src) && aligned16(plain), "tcgen05 path needs 16-byte aligned tensors");
	const int T = gg.RH * gg.RW;
	const int r_pad = round_up(gg.SC, 16);
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	// CTA pairs (cta_group::2): 256 output channels per unit, each CTA loading and splitting half of the gathered tile --
	// half the shared-memory traffic per MMA.  Needs a second 128-row tile of output channels to pair with.
	static const bool no_pairs = (pair_mask() & 2) == 0;
	const int tap_wp = gg.SC < 16 ? tapbox_wp(gg) : 0;
	const int ctas = (!no_pairs && gg.J > TC_BM && !tap_wp) ? 2 : 1;   // (a tap box is one TMA box: not split over a pair)
	// boxes of 32 channel rows for a pair (a tile of 192 columns = 6 boxes, 3 per CTA)
	const int RB = (ctas == 1 && r_pad % 64 == 0) ? 64 : (r_pad % 32 == 0 ? 32 : 16);

	CUtensorMap tm_b;
	{
		cuuint64_t dims[4] = { (cuuint64_t) gg.N, (cuuint64_t) gg.SH, (cuuint64_t) gg.SW, (cuuint64_t) gg.SC };
		cuuint64_t str[3] = { (cuuint64_t) gg.N * 4, (cuuint64_t) gg.N * gg.SH * 4, (cuuint64_t) gg.N * gg.SH * gg.SW * 4 };
		cuuint32_t box[4] = { (cuuint32_t) WG_KB, 1, 1, (cuuint32_t) RB };
		if (tap_wp) { box[1] = (cuuint32_t) gg.RH; box[2] = (cuuint32_t) tap_wp; box[3] = (cuuint32_t) gg.SC; }
		CATTL3_CHECK(encode_map(&tm_b, src, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B));
	}

	TcWgradParams p;
	p.N = gg.N; p.OH = gg.OH; p.OW = gg.OW; p.R = gg.SC; p.J = gg.J; p.RH = gg.RH; p.RW = gg.RW;
	p.ah = gg.ah; p.bh = gg.bh; p.ch = gg.ch; p.aw = gg.aw; p.bw = gg.bw; p.cw = gg.cw;
	p.RB = RB; p.rchunks = r_pad / RB; p.boxes = T * p.rchunks;
	const int max_boxes = 192 / RB;
	p.col_tiles = (p.boxes + max_boxes - 1) / max_boxes;
	p.boxes_per_tile = (p.boxes + p.col_tiles - 1) / p.col_tiles;  // balanced: 9 boxes of 64 -> 3 tiles of 192 columns
	if (ctas == 2 && p.boxes_per_tile % 2) ++p.boxes_per_tile;     // each CTA of a pair takes half of a tile's boxes
	p.col_tiles = (p.boxes + p.boxes_per_tile - 1) / p.boxes_per_tile;
	p.BNW = p.boxes_per_tile * RB;
	p.tapbox = tap_wp ? 1 : 0; p.tap_wp = tap_wp;
	if (tap_wp) {
		// one box, one column tile: the accumulator holds the whole (padded) dW of this j tile
		p.boxes = 1; p.boxes_per_tile = 1; p.col_tiles = 1;
		p.BNW = round_up(gg.RH * tap_wp * gg.SC, 16);
	}
	if (ctas == 2 && p.BNW % 32) { set_error("wgrad: pair tile of %d columns", p.BNW); return CATTL3_ERR_UNSUPPORTED; }
	p.j_tiles = round_up((gg.J + TC_BM - 1) / TC_BM, ctas);   // 128-row tiles, whole pairs
	p.mgroups = (gg.m_count ? gg.m_count : M) / WG_KB;
	p.mg_base = gg.m_first / WG_KB;
	const int tiles = p.col_tiles * p.j_tiles;
	long long splits = ctx->sm_count / tiles;
	if (splits < 1) splits = 1;
	if (splits > p.mgroups) splits = p.mgroups;
	p.mg_per_split = ceil_div(p.mgroups, splits);
	p.splits = (int) ceil_div(p.mgroups, p.mg_per_split);
	const int ring_bytes = 4 * WG_RING * 4096;   // the converters' dY rings
	const int stage_bytes = 2 * (p.BNW / ctas) * WG_KB * 4;
	int stages = (TC_SMEM_LIMIT - ring_bytes) / stage_bytes;
	const int tmem_stages = (512 - (int) WG_A_COL0) / (2 * WG_KB);
	if (stages > tmem_stages) stages = tmem_stages;
	if (getenv("CATTL3_WG_STAGES") && atoi(getenv("CATTL3_WG_STAGES")) < stages) stages = atoi(getenv("CATTL3_WG_STAGES"));
	p.stages = stages;
	p.flush = getenv("CATTL3_WG_FLUSH") ? atoi(getenv("CATTL3_WG_FLUSH")) : 32;
	p.w_stap = gg.w_stap; p.w_sr = gg.w_sr; p.w_sj = gg.w_sj;
	p.dw_elems = (long long) T * gg.SC * gg.J;
	const size_t partial_elems = (size_t) p.splits * tiles * p.BNW * 128;
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (partial_elems + (size_t) p.splits * p.j_tiles * TC_BM) * 4));
	p.partial = (float*) ctx->ws;
	p.db_partial = db ? p.partial + partial_elems : nullptr;
	p.plain = plain; p.M = M;
	size_t smem_bytes = (size_t) stages * stage_bytes + 1024 + 512 + ring_bytes;
	if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;
	void (*kern)(const CUtensorMap, const TcWgradParams) = ctas == 2 ? tc_wgrad_kernel<2> : tc_wgrad_kernel<1>;
	CATTL3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
	CATTL3_CUDA(launch_clustered(kern, tiles * p.splits, TC_THREADS, smem_bytes, ctx->stream, ctas, tm_b, p));
	CATTL3_LAUNCHED(ctx);
	wgrad_reduce_tc_kernel<<<ew_grid(ctx, ceil_div(p.dw_elems + (db ? gg.J : 0), 32), 1), 32 * WR_ZL, 0, ctx->stream>>>(p, T, dw, db);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

} // namespace cattl3
