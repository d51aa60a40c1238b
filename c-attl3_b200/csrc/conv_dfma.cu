// conv_dfma.cu -- big-tile FMA implicit-GEMM kernels for GEMM-sized shapes the tensor-core path does not take:
//   * double: the product path of the double instantiation, built around the FP64 FMA pipe (64 DFMA / clk / SM on
//     B200: half the FP32 rate; this design has no tensor-core path for fp64);
//   * float: layers the tcgen05 kernels refuse -- fewer than 16 input channels (the 3-channel stem convolution of
//     config 4, whose reduction is walked as one flattened (tap, channel) index so that 3 channels do not become 8)
//     or a batch that is not a multiple of 32.
// An FMA kernel is bound by its pipe only if little else competes for issue slots and shared memory, so:
//   * 128 x 128 (or 128 x 64) block tiles, 256 threads, 8 x 8 (8 x 4) accumulators per thread: 64 DFMA per
//     8 shared-memory loads per k;
//   * a warp covers 64 rows x 32 columns as 8 x 4 lanes, so every 16-byte shared-memory read of a warp touches
//     128 contiguous bytes (A) or 64 (B, broadcast): one wavefront each;
//   * global -> register prefetch of the next k-block while the current one is multiplied, two shared-memory
//     buffers, one __syncthreads per k-block;
//   * no im2col buffer: the gather coordinates are recomputed per load from the flattened reduction index
//     k = tap * R + r (C-ATTL3/layer/kernel/ConvKernelLayer.hpp:117-142 materialises that matrix instead).
// Shapes that cannot fill such tiles (tiny filter counts: configs 1 and 3) stay on conv_simt.cu.
#include "activations.cuh"

namespace cattl3 {

namespace {

constexpr int DF_BM = 128, DF_BK = 8, DF_THREADS = 256;

// One element global -> shared without passing through a register (LDGSTS); !pred zero-fills the destination.
template<typename S>
__device__ __forceinline__ void cp_async_elem(S* dst, const S* src, bool pred) {
	const uint32_t d = (uint32_t) __cvta_generic_to_shared(dst);
	const int bytes = pred ? (int) sizeof(S) : 0;
	if (sizeof(S) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(src), "r"(bytes) : "memory");
	else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_all() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template<typename S> struct Pair;
template<> struct Pair<double> { typedef double2 type; };
template<> struct Pair<float> { typedef float2 type; };
template<typename S> __device__ __forceinline__ typename Pair<S>::type make_pair2(S a, S b);
template<> __device__ __forceinline__ double2 make_pair2<double>(double a, double b) { return make_double2(a, b); }
template<> __device__ __forceinline__ float2 make_pair2<float>(float a, float b) { return make_float2(a, b); }

// FLAT: the reduction index is the flattened k = tap * R + r (k-blocks may span taps: few reduce channels);
// otherwise taps outer, channel blocks inner.
template<typename S, int TN, bool FLAT>
__global__ void __launch_bounds__(DF_THREADS, (TN == 4 || sizeof(S) == 4) ? 2 : 1) fma_gather_gemm_kernel(GatherGeom gg,
		const S* __restrict__ src, const S* __restrict__ w, const S* __restrict__ bias, int bias_mode, S* __restrict__ out,
		int act_kind, S act_param, S* __restrict__ act_out, int vec_ok) {
	typedef typename Pair<S>::type P2;
	constexpr int BN = 16 * TN;
	constexpr int B_ITERS = DF_BK * BN / DF_THREADS;   // 4 (BN = 128) or 2 (BN = 64)
	constexpr int B_KSTEP = DF_THREADS / BN;           // 2 or 4
	__shared__ __align__(16) S As[2][DF_BK][DF_BM];
	__shared__ __align__(16) S Bs[2][DF_BK][BN];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = lane & 7, ty = lane >> 3;
	const int rm = (warp & 1) * 64 + tx * 2;            // + 16 * i, i < 4 (two consecutive rows each)
	const int cn = (warp >> 1) * (BN / 4) + ty * 2;     // + 8 * jq, jq < TN / 2 (two consecutive columns each)
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long m0 = (long long) blockIdx.x * DF_BM;
	const int j0 = blockIdx.y * BN;
	const int R = gg.SC, J = gg.J;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;

	// loader coordinates: this thread always fetches row a_ml of the A tile and column b_j of the B tile
	const int a_ml = tid & (DF_BM - 1), a_k0 = tid >> 7;
	const long long am = m0 + a_ml;
	const bool m_ok = am < M;
	const int an = (int) (am % gg.N);
	const long long apix = am / gg.N;
	const int aoh = (int) (apix % gg.OH), aow = (int) (apix / gg.OH);
	const int b_j = tid & (BN - 1), b_k0 = tid / BN;
	const bool j_ok = j0 + b_j < J;
	const long long wj = (long long) (j0 + b_j) * gg.w_sj;

	S acc[8][TN];
	#pragma unroll
	for (int i = 0; i < 8; ++i)
		#pragma unroll
		for (int j = 0; j < TN; ++j) acc[i][j] = (S) 0;

	const int T = gg.RH * gg.RW;
	const bool unit_den = gg.denh == 1 && gg.denw == 1;
	// Source address of (this thread's row, tap (rh, rw)), or null outside the tensor / off the stride lattice.
	auto tap_src = [&](int rh, int rw) -> const S* {
		const int th = aoh * gg.ah + rh * gg.bh + gg.ch;
		const int tw = aow * gg.aw + rw * gg.bw + gg.cw;
		if (!m_ok || th < 0 || tw < 0) return nullptr;
		int ih = th, iw = tw;
		if (!unit_den) {   // strided transposed gathers only: the divisions stay out of every other layer's loop
			if (th % gg.denh != 0 || tw % gg.denw != 0) return nullptr;
			ih = th / gg.denh; iw = tw / gg.denw;
		}
		if (ih >= gg.SH || iw >= gg.SW) return nullptr;
		return src + an + (long long) gg.N * (ih + (long long) gg.SH * iw);
	};
	// The tiles go global -> shared with cp.async (no staging registers: they are all needed by the accumulators and
	// the fragments): the copies of k-block ks + 1 are issued before the multiply of k-block ks and awaited after it.
	auto multiply = [&](int buf) {
		#pragma unroll
		for (int kk = 0; kk < DF_BK; ++kk) {
			S a[8], b[TN];
			#pragma unroll
			for (int i = 0; i < 4; ++i)
				*reinterpret_cast<P2*>(&a[2 * i]) = *reinterpret_cast<const P2*>(&As[buf][kk][rm + 16 * i]);
			#pragma unroll
			for (int jq = 0; jq < TN / 2; ++jq)
				*reinterpret_cast<P2*>(&b[2 * jq]) = *reinterpret_cast<const P2*>(&Bs[buf][kk][cn + 8 * jq]);
			#pragma unroll
			for (int i = 0; i < 8; ++i)
				#pragma unroll
				for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
		}
	};

	if (FLAT) {
		// Every load slot walks its own flattened index k = k_slot + 8 * step: (r, rh, rw) advance incrementally
		// (no divisions in the loop); past the last tap a slot loads zeros.
		const int ksteps = (T * R + DF_BK - 1) / DF_BK;
		int a_r[4], a_rh[4], a_rw[4], b_r[B_ITERS], b_tap[B_ITERS];
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int k = a_k0 + 2 * i, tap = k / R;
			a_r[i] = k - tap * R; a_rw[i] = tap / gg.RH; a_rh[i] = tap - a_rw[i] * gg.RH;
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int k = b_k0 + B_KSTEP * i;
			b_tap[i] = k / R; b_r[i] = k - b_tap[i] * R;
		}
		auto fetch = [&](int buf) {
			#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const S* ps = a_rw[i] < gg.RW ? tap_src(a_rh[i], a_rw[i]) : nullptr;
				cp_async_elem(&As[buf][a_k0 + 2 * i][a_ml], ps ? ps + a_r[i] * plane : src, ps != nullptr);
				a_r[i] += DF_BK;
				while (a_r[i] >= R) {
					a_r[i] -= R;
					if (++a_rh[i] == gg.RH) { a_rh[i] = 0; ++a_rw[i]; }
				}
			}
			#pragma unroll
			for (int i = 0; i < B_ITERS; ++i) {
				const bool ok = j_ok && b_tap[i] < T;
				cp_async_elem(&Bs[buf][b_k0 + B_KSTEP * i][b_j], ok ? w + wj + b_tap[i] * gg.w_stap + b_r[i] * gg.w_sr : w, ok);
				b_r[i] += DF_BK;
				while (b_r[i] >= R) { b_r[i] -= R; ++b_tap[i]; }
			}
			cp_async_commit_all();
		};
		fetch(0);
		cp_async_wait_all();
		__syncthreads();
		for (int ks = 0; ks < ksteps; ++ks) {
			const int buf = ks & 1;
			if (ks + 1 < ksteps) fetch(buf ^ 1);
			multiply(buf);
			cp_async_wait_all();
			__syncthreads();
		}
	} else {
		// Taps outer, channel blocks of DF_BK inner (a partial last block is zero padded): the gather coordinates --
		// divisions, bounds and divisibility tests -- are evaluated once per tap; inside a tap a load is
		// base + r * plane.  The fetch stream runs one k-block ahead of the multiply stream.
		const int rblocks = (R + DF_BK - 1) / DF_BK;
		const int ksteps = T * rblocks;
		int f_tap = 0, f_r0 = 0;
		const S* f_src = tap_src(0, 0);
		const S* f_w = w + wj;
		auto fetch = [&](int buf) {
			#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int r = f_r0 + a_k0 + 2 * i;
				const bool ok = f_src && r < R;
				cp_async_elem(&As[buf][a_k0 + 2 * i][a_ml], ok ? f_src + r * plane : src, ok);
			}
			#pragma unroll
			for (int i = 0; i < B_ITERS; ++i) {
				const int r = f_r0 + b_k0 + B_KSTEP * i;
				const bool ok = j_ok && r < R;
				cp_async_elem(&Bs[buf][b_k0 + B_KSTEP * i][b_j], ok ? f_w + r * gg.w_sr : w, ok);
			}
			cp_async_commit_all();
			f_r0 += DF_BK;
			if (f_r0 >= R) {
				f_r0 = 0;
				if (++f_tap < T) {
					const int rw = f_tap / gg.RH;
					f_src = tap_src(f_tap - rw * gg.RH, rw);
					f_w = w + wj + f_tap * gg.w_stap;
				}
			}
		};
		fetch(0);
		cp_async_wait_all();
		__syncthreads();
		for (int ks = 0; ks < ksteps; ++ks) {
			const int buf = ks & 1;
			if (ks + 1 < ksteps) fetch(buf ^ 1);
			multiply(buf);
			cp_async_wait_all();
			__syncthreads();
		}
	}

	// epilogue: bias, fused activation, stores (pairs along m where the tensor allows it)
	const long long P = (long long) gg.OH * gg.OW;
	#pragma unroll
	for (int jq = 0; jq < TN / 2; ++jq) {
		#pragma unroll
		for (int f = 0; f < 2; ++f) {
			const int j = j0 + cn + 8 * jq + f;
			if (j >= J) continue;
			const S bj = bias_mode == 1 ? __ldg(bias + j) : (S) 0;
			#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const long long m = m0 + rm + 16 * i;
				if (m >= M) continue;
				S v0 = acc[2 * i][2 * jq + f] + bj, v1 = acc[2 * i + 1][2 * jq + f] + bj;
				const bool two = m + 1 < M;
				if (bias_mode == 2) {
					v0 += __ldg(bias + m / gg.N + P * j);
					if (two) v1 += __ldg(bias + (m + 1) / gg.N + P * j);
				}
				const long long o = m + M * j;
				if (out) {
					if (vec_ok) *reinterpret_cast<P2*>(out + o) = make_pair2<S>(v0, v1);
					else { out[o] = v0; if (two) out[o + 1] = v1; }
				}
				if (act_out) {
					v0 = act_fwd_rt<S>(act_kind, v0, act_param);
					v1 = act_fwd_rt<S>(act_kind, v1, act_param);
					if (vec_ok) *reinterpret_cast<P2*>(act_out + o) = make_pair2<S>(v0, v1);
					else { act_out[o] = v0; if (two) act_out[o + 1] = v1; }
				}
			}
		}
	}
}

// Weight gradient: dw(tap, r, j) += sum_m src(m, tap, r) * plain[m + M*j].  The reduction runs over m, which is
// contiguous in HBM for BOTH operands, so the tiles are fetched 8 consecutive m at a time (two full sectors) and
// transposed on the way into shared memory (row pitch +2 doubles: conflict-free both ways).  Output tile: 128
// flattened (tap, r) rows x BN filters; split over m (grid z) into per-split partials that wgrad_reduce_kernel adds
// to dw in split order.
constexpr int DW_BKM = 8;

template<typename S, int TN>
__global__ void __launch_bounds__(DF_THREADS, (TN == 4 || sizeof(S) == 4) ? 2 : 1) fma_wgrad_kernel(GatherGeom gg,
		const S* __restrict__ src, const S* __restrict__ plain, S* __restrict__ partial, long long m_per_split,
		long long dw_elems) {
	typedef typename Pair<S>::type P2;
	constexpr int BN = 16 * TN;
	constexpr int PAD = 16 / sizeof(S);   // +16 bytes per row: rows stay 16-byte aligned, transposing stores spread over the banks
	constexpr int PITCH_A = DF_BM + PAD, PITCH_B = BN + PAD;
	constexpr int B_ITERS = BN / 32;
	__shared__ __align__(16) S As[2][DW_BKM][PITCH_A];
	__shared__ __align__(16) S Bs[2][DW_BKM][PITCH_B];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tx = lane & 7, ty = lane >> 3;
	const int rm = (warp & 1) * 64 + tx * 2;
	const int cn = (warp >> 1) * (BN / 4) + ty * 2;
	const int R = gg.SC, J = gg.J;
	const int Ktot = gg.RH * gg.RW * R;
	const int k0 = blockIdx.x * DF_BM, j0 = blockIdx.y * BN;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long ms = (long long) blockIdx.z * m_per_split;
	const long long me = ms + m_per_split < M ? ms + m_per_split : M;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;

	// loader: 8 consecutive m per (tap, r) row / filter column; this thread owns m offset l_mm and rows l_r0 + 32 i
	const int l_mm = tid & 7, l_r0 = tid >> 3;
	int row_dh[4], row_dw[4], row_r[4];   // per owned row: rh*bh + ch, rw*bw + cw, reduce channel (or -1: padding row)
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int k = k0 + l_r0 + 32 * i;
		if (k < Ktot) {
			const int tap = k / R, r = k - tap * R;
			const int rw = tap / gg.RH, rh = tap - rw * gg.RH;
			row_dh[i] = rh * gg.bh + gg.ch; row_dw[i] = rw * gg.bw + gg.cw; row_r[i] = r;
		} else {
			row_dh[i] = 0; row_dw[i] = 0; row_r[i] = -1;
		}
	}

	S acc[8][TN];
	#pragma unroll
	for (int i = 0; i < 8; ++i)
		#pragma unroll
		for (int j = 0; j < TN; ++j) acc[i][j] = (S) 0;

	auto fetch = [&](long long mc, int buf) {
		const long long m = mc + l_mm;
		const bool ok = m < me;
		// M < 2^31 (dfma_wgrad_supported): 32-bit divisions
		const unsigned mu = (unsigned) m, pixu = mu / (unsigned) gg.N;
		const int n = (int) (mu - pixu * (unsigned) gg.N);
		const int ow = (int) (pixu / (unsigned) gg.OH), oh = (int) (pixu - (unsigned) ow * (unsigned) gg.OH);
		const int bh = oh * gg.ah, bw = ow * gg.aw;
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int th = bh + row_dh[i], tw = bw + row_dw[i];
			// weight gradients only ever gather forward-style (denh = denw = 1, fma_wgrad_supported)
			const bool aok = ok && row_r[i] >= 0 && th >= 0 && tw >= 0 && th < gg.SH && tw < gg.SW;
			cp_async_elem(&As[buf][l_mm][l_r0 + 32 * i],
					aok ? src + n + (long long) gg.N * (th + (long long) gg.SH * tw) + row_r[i] * plane : src, aok);
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int j = j0 + l_r0 + 32 * i;
			const bool bok = ok && j < J;
			cp_async_elem(&Bs[buf][l_mm][l_r0 + 32 * i], bok ? plain + m + M * j : plain, bok);
		}
		cp_async_commit_all();
	};

	const long long steps = me > ms ? (me - ms + DW_BKM - 1) / DW_BKM : 0;
	if (steps > 0) {
		fetch(ms, 0);
		cp_async_wait_all();
	}
	__syncthreads();
	for (long long st = 0; st < steps; ++st) {
		const int buf = (int) (st & 1);
		if (st + 1 < steps) fetch(ms + (st + 1) * DW_BKM, buf ^ 1);
		#pragma unroll
		for (int mk = 0; mk < DW_BKM; ++mk) {
			S a[8], b[TN];
			#pragma unroll
			for (int i = 0; i < 4; ++i)
				*reinterpret_cast<P2*>(&a[2 * i]) = *reinterpret_cast<const P2*>(&As[buf][mk][rm + 16 * i]);
			#pragma unroll
			for (int jq = 0; jq < TN / 2; ++jq)
				*reinterpret_cast<P2*>(&b[2 * jq]) = *reinterpret_cast<const P2*>(&Bs[buf][mk][cn + 8 * jq]);
			#pragma unroll
			for (int i = 0; i < 8; ++i)
				#pragma unroll
				for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
		}
		cp_async_wait_all();
		__syncthreads();
	}

	S* dst = partial + (long long) blockIdx.z * dw_elems;
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		#pragma unroll
		for (int e = 0; e < 2; ++e) {
			const int k = k0 + rm + 16 * i + e;
			if (k >= Ktot) continue;
			const int tap = k / R, r = k - tap * R;
			const long long base = tap * gg.w_stap + r * gg.w_sr;
			#pragma unroll
			for (int jq = 0; jq < TN / 2; ++jq) {
				#pragma unroll
				for (int f = 0; f < 2; ++f) {
					const int j = j0 + cn + 8 * jq + f;
					if (j < J) dst[base + j * gg.w_sj] = acc[2 * i + e][2 * jq + f];
				}
			}
		}
	}
}

template<typename S>
__global__ void __launch_bounds__(256) fma_wgrad_reduce_kernel(const S* __restrict__ partial, int splits, long long elems,
		S* __restrict__ dw) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < elems; i += (long long) gridDim.x * 256) {
		S s = 0;
		for (int z = 0; z < splits; ++z) s += partial[(long long) z * elems + i];
		dw[i] += s;
	}
}

} // namespace

// The big-tile kernels pay off once a tile is mostly real work: enough filters for a 64-wide tile and a reduction
// longer than a couple of k-blocks.
template<typename S>
bool fma_gather_gemm_supported(const GatherGeom& gg) {
	const long long K = (long long) gg.RH * gg.RW * gg.SC;
	// double walks reduce channels in blocks of 8 per tap (tiny channel counts would mostly multiply padding);
	// float has the flattened walk for them
	return gg.J > 32 && K >= 32 && (sizeof(S) == 4 || gg.SC >= 6);
}
template bool fma_gather_gemm_supported<float>(const GatherGeom&);
template bool fma_gather_gemm_supported<double>(const GatherGeom&);

template<typename S>
int fma_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias, int bias_mode, S* out,
		const EpilogueArgs* ep) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	S* act_out = act ? (S*) ep->act_out : nullptr;
	const size_t pair = 2 * sizeof(S);
	const int vec_ok = M % 2 == 0 && (!out || (uintptr_t) out % pair == 0) && (!act_out || (uintptr_t) act_out % pair == 0);
	const int act_kind = act ? ep->act_kind : CATTL3_ACT_NONE;
	const S act_param = act ? (S) ep->act_param : (S) 0;
	// measured at config 2, double (profiles/README.md, r1e): 64 filters -> 128 x 64 tiles, two CTAs per SM (20 TFLOP/s
	// against 12 with half-empty 128-wide tiles); 256 filters -> 128 x 128 tiles (18.8 against 15.7)
	const bool wide = gg.J > 64;
	const bool flat = sizeof(S) == 4 && gg.SC < DF_BK;
	dim3 grid((unsigned) ceil_div(M, DF_BM), (unsigned) ceil_div(gg.J, wide ? 128 : 64));
#define LAUNCH(TN, FLAT) fma_gather_gemm_kernel<S, TN, FLAT><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, w, bias, \
		bias_mode, out, act_kind, act_param, act_out, vec_ok)
	if (flat) { if (wide) LAUNCH(8, true); else LAUNCH(4, true); }
	else { if (wide) LAUNCH(8, false); else LAUNCH(4, false); }
#undef LAUNCH
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int fma_gather_gemm<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, const float*, int, float*,
		const EpilogueArgs*);
template int fma_gather_gemm<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, const double*, int, double*,
		const EpilogueArgs*);

template<typename S>
bool fma_wgrad_supported(const GatherGeom& gg) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	return gg.J > 32 && (long long) gg.RH * gg.RW * gg.SC >= 64 && M >= 1024 && M < (1ll << 31) && gg.denh == 1 && gg.denw == 1;
}
template bool fma_wgrad_supported<float>(const GatherGeom&);
template bool fma_wgrad_supported<double>(const GatherGeom&);

template<typename S>
int fma_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw) {
	const int Ktot = gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) Ktot * gg.J;
	const int BN = gg.J > 64 ? 128 : 64;
	const long long gx = ceil_div(Ktot, DF_BM), gy = ceil_div(gg.J, BN);
	// one (double, BN = 128) or two CTAs per SM: as many m-splits as fill the machine once
	const int per_sm = (BN == 128 && sizeof(S) == 8) ? 1 : 2;
	long long splits = (long long) per_sm * ctx->sm_count / (gx * gy);
	if (splits < 1) splits = 1;
	const long long max_splits = ceil_div(M, 1024);
	if (splits > max_splits) splits = max_splits;
	long long m_per_split = ceil_div(ceil_div(M, splits), DW_BKM) * DW_BKM;
	splits = ceil_div(M, m_per_split);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (splits * elems) * sizeof(S)));
	// rows of a tile beyond Ktot and columns beyond J are never written: the reduce reads only real elements
	dim3 grid((unsigned) gx, (unsigned) gy, (unsigned) splits);
	if (BN == 128)
		fma_wgrad_kernel<S, 8><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, plain, (S*) ctx->ws, m_per_split, elems);
	else
		fma_wgrad_kernel<S, 4><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, plain, (S*) ctx->ws, m_per_split, elems);
	CATTL3_LAUNCHED(ctx);
	fma_wgrad_reduce_kernel<S><<<ew_grid(ctx, elems, 256), 256, 0, ctx->stream>>>((const S*) ctx->ws, (int) splits, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int fma_wgrad<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, float*);
template int fma_wgrad<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, double*);

} // namespace cattl3
