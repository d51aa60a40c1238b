// conv_simt.cu -- FFMA / DFMA implicit-GEMM kernels: the any-shape path of the kernel layers
// (float and double) and the product path of the double instantiation.
//
// One "gather GEMM" covers ConvKernelLayer forward and input-gradient, TransConvKernelLayer
// forward and input-gradient and DenseKernelLayer forward / input-gradient: the A operand is
// never materialised (no im2col buffer in HBM, unlike C-ATTL3/layer/kernel/ConvKernelLayer.hpp:127);
// its element (m, tap, r) is read straight from the source tensor at the coordinate the
// GatherGeom describes.  The batch index n is the fastest dimension of every tensor
// (C-ATTL3/core/EigenProxy.hpp:56-57), so a run of consecutive m for a fixed tap / channel is
// contiguous in HBM: global loads and the output stores are coalesced along m.
//
// The weight gradient is a split-K reduction over m = N*OH*OW with a deterministic second stage
// that ACCUMULATES into the gradient (Parameters::accumulate_grad semantics,
// C-ATTL3/parameters/StandardParameters.hpp:115-123).
#include <algorithm>
#include "activations.cuh"

namespace cattl3 {

template<typename S> struct Vec;
template<> struct Vec<float> { typedef float4 type; static constexpr int G = 4; };
template<> struct Vec<double> { typedef double2 type; static constexpr int G = 2; };

// ------------------------------------------------------------------------------------------------
// gather GEMM: out[m + M*j] = bias + sum_{tap, r} src(m, tap, r) * w(tap, r, j)
// block tile BM x 64, BK = 16, 256 threads, each thread 2 groups of G rows x 4 columns.
// ------------------------------------------------------------------------------------------------
template<typename S>
__global__ void __launch_bounds__(256) gather_gemm_kernel(GatherGeom gg, const S* __restrict__ src,
		const S* __restrict__ w, const S* __restrict__ bias, int bias_mode, S* __restrict__ out, int act_kind, S act_param,
		S* __restrict__ act_out) {
	constexpr int G = Vec<S>::G;
	constexpr int BM = 32 * G, BN = 64, BK = 16;
	constexpr int A_ROWS = 256 / BM, A_ITERS = BK / A_ROWS;
	typedef typename Vec<S>::type V;
	__shared__ __align__(16) S As[BK][BM];
	__shared__ __align__(16) S Bs[BK][BN];

	const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long m0 = (long long) blockIdx.x * BM;
	const int j0 = blockIdx.y * BN;
	const int T = gg.RH * gg.RW, R = gg.SC, J = gg.J;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;

	// A-loader coordinates: this thread always loads row a_ml of the tile
	const int a_ml = tid % BM, a_k0 = tid / BM;
	const long long am = m0 + a_ml;
	const bool m_ok = am < M;
	const int an = (int) (am % gg.N);
	const long long apix = am / gg.N;
	const int aoh = (int) (apix % gg.OH), aow = (int) (apix / gg.OH);
	// B-loader coordinates
	const int b_jj = tid & 63, b_k0 = tid >> 6;

	S acc[2 * G][4];
	#pragma unroll
	for (int i = 0; i < 2 * G; ++i)
		#pragma unroll
		for (int j = 0; j < 4; ++j) acc[i][j] = 0;

	for (int tap = 0; tap < T; ++tap) {
		const int rh = tap % gg.RH, rw = tap / gg.RH;
		const int th = aoh * gg.ah + rh * gg.bh + gg.ch;
		const int tw = aow * gg.aw + rw * gg.bw + gg.cw;
		bool ok = m_ok && th >= 0 && tw >= 0 && (th % gg.denh) == 0 && (tw % gg.denw) == 0;
		const int ih = th / gg.denh, iw = tw / gg.denw;
		ok = ok && ih < gg.SH && iw < gg.SW;
		const long long base = an + (long long) gg.N * (ih + (long long) gg.SH * iw);
		const long long wtap = tap * gg.w_stap;
		for (int r0 = 0; r0 < R; r0 += BK) {
			#pragma unroll
			for (int i = 0; i < A_ITERS; ++i) {
				const int kk = a_k0 + A_ROWS * i, r = r0 + kk;
				As[kk][a_ml] = (ok && r < R) ? src[base + r * plane] : (S) 0;
			}
			#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const int kk = b_k0 + 4 * i, r = r0 + kk, j = j0 + b_jj;
				Bs[kk][b_jj] = (r < R && j < J) ? w[wtap + r * gg.w_sr + j * gg.w_sj] : (S) 0;
			}
			__syncthreads();
			#pragma unroll
			for (int kk = 0; kk < BK; ++kk) {
				S a[2 * G], b[4];
				*reinterpret_cast<V*>(&a[0]) = *reinterpret_cast<const V*>(&As[kk][tm * G]);
				*reinterpret_cast<V*>(&a[G]) = *reinterpret_cast<const V*>(&As[kk][BM / 2 + tm * G]);
				if (G == 4) {
					*reinterpret_cast<V*>(&b[0]) = *reinterpret_cast<const V*>(&Bs[kk][tn * 4]);
				} else {
					*reinterpret_cast<V*>(&b[0]) = *reinterpret_cast<const V*>(&Bs[kk][tn * 4]);
					*reinterpret_cast<V*>(&b[2]) = *reinterpret_cast<const V*>(&Bs[kk][tn * 4 + 2]);
				}
				#pragma unroll
				for (int i = 0; i < 2 * G; ++i)
					#pragma unroll
					for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
			}
			__syncthreads();
		}
	}
	const long long P = (long long) gg.OH * gg.OW;
	#pragma unroll
	for (int g = 0; g < 2; ++g) {
		#pragma unroll
		for (int jj = 0; jj < 4; ++jj) {
			const int j = j0 + tn * 4 + jj;
			if (j >= J) continue;
			#pragma unroll
			for (int e = 0; e < G; ++e) {
				const long long m = m0 + g * (BM / 2) + tm * G + e;
				if (m >= M) continue;
				S v = acc[g * G + e][jj];
				if (bias_mode == 1) v += bias[j];
				else if (bias_mode == 2) v += bias[m / gg.N + P * j];
				if (out) out[m + M * j] = v;
				if (act_out) act_out[m + M * j] = act_fwd_rt<S>(act_kind, v, act_param);
			}
		}
	}
}

template<typename S>
int simt_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias,
		int bias_mode, S* out, const EpilogueArgs* ep) {
	constexpr int BM = 32 * Vec<S>::G;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	dim3 grid((unsigned) ceil_div(M, BM), (unsigned) ceil_div(gg.J, 64));
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	gather_gemm_kernel<S><<<grid, 256, 0, ctx->stream>>>(gg, src, w, bias, bias_mode, out,
			act ? ep->act_kind : CATTL3_ACT_NONE, act ? (S) ep->act_param : (S) 0, act ? (S*) ep->act_out : (S*) nullptr);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int simt_gather_gemm<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, const float*, int, float*,
		const EpilogueArgs*);
template int simt_gather_gemm<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, const double*, int, double*,
		const EpilogueArgs*);

// ------------------------------------------------------------------------------------------------
// weight gradient: dw(tap, r, j) += sum_m src(m, tap, r) * plain[m + M*j]
// grid (T * ceil(R/64), ceil(J/64), splits); each block reduces its m range into a 64 x 64 tile
// of a per-split partial buffer; wgrad_reduce_kernel then adds the partials to dw in split order.
// ------------------------------------------------------------------------------------------------
template<typename S>
__global__ void __launch_bounds__(256) wgrad_kernel(GatherGeom gg, const S* __restrict__ src,
		const S* __restrict__ plain, S* __restrict__ partial, long long m_per_split, long long dw_elems) {
	constexpr int BKM = 32, BT = 64;
	typedef typename Vec<S>::type V;
	constexpr int G = Vec<S>::G;
	// +4 padding: the loaders write a column at a time (consecutive threads = consecutive m)
	__shared__ __align__(16) S As[BKM][BT + 4];
	__shared__ __align__(16) S Bs[BKM][BT + 4];
	const int tid = threadIdx.x, tr = tid & 15, tj = tid >> 4;
	const int T = gg.RH * gg.RW, R = gg.SC, J = gg.J;
	const int tap = blockIdx.x % T, r0 = (blockIdx.x / T) * BT, j0 = blockIdx.y * BT;
	const int rh = tap % gg.RH, rw = tap / gg.RH;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long ms = (long long) blockIdx.z * m_per_split;
	const long long me = (ms + m_per_split < M) ? ms + m_per_split : M;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;
	const int l_mm = tid & 31, l_c0 = tid >> 5;

	S acc[4][4];
	#pragma unroll
	for (int i = 0; i < 4; ++i)
		#pragma unroll
		for (int j = 0; j < 4; ++j) acc[i][j] = 0;

	for (long long mc = ms; mc < me; mc += BKM) {
		const long long m = mc + l_mm;
		bool ok = m < me;
		const int n = (int) (m % gg.N);
		const long long pix = m / gg.N;
		const int oh = (int) (pix % gg.OH), ow = (int) (pix / gg.OH);
		const int th = oh * gg.ah + rh * gg.bh + gg.ch;
		const int tw = ow * gg.aw + rw * gg.bw + gg.cw;
		bool aok = ok && th >= 0 && tw >= 0 && (th % gg.denh) == 0 && (tw % gg.denw) == 0;
		const int ih = th / gg.denh, iw = tw / gg.denw;
		aok = aok && ih < gg.SH && iw < gg.SW;
		const long long base = n + (long long) gg.N * (ih + (long long) gg.SH * iw);
		#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const int col = l_c0 + 8 * i;
			const int r = r0 + col, j = j0 + col;
			As[l_mm][col] = (aok && r < R) ? src[base + r * plane] : (S) 0;
			Bs[l_mm][col] = (ok && j < J) ? plain[m + M * j] : (S) 0;
		}
		__syncthreads();
		#pragma unroll
		for (int mk = 0; mk < BKM; ++mk) {
			S a[4], b[4];
			if (G == 4) {
				*reinterpret_cast<V*>(&a[0]) = *reinterpret_cast<const V*>(&As[mk][tr * 4]);
				*reinterpret_cast<V*>(&b[0]) = *reinterpret_cast<const V*>(&Bs[mk][tj * 4]);
			} else {
				*reinterpret_cast<V*>(&a[0]) = *reinterpret_cast<const V*>(&As[mk][tr * 4]);
				*reinterpret_cast<V*>(&a[2]) = *reinterpret_cast<const V*>(&As[mk][tr * 4 + 2]);
				*reinterpret_cast<V*>(&b[0]) = *reinterpret_cast<const V*>(&Bs[mk][tj * 4]);
				*reinterpret_cast<V*>(&b[2]) = *reinterpret_cast<const V*>(&Bs[mk][tj * 4 + 2]);
			}
			#pragma unroll
			for (int i = 0; i < 4; ++i)
				#pragma unroll
				for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
	}
	S* dst = partial + (long long) blockIdx.z * dw_elems;
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int r = r0 + tr * 4 + i;
		if (r >= R) continue;
		#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int jj = j0 + tj * 4 + j;
			if (jj < J) dst[tap * gg.w_stap + r * gg.w_sr + jj * gg.w_sj] = acc[i][j];
		}
	}
}

// dw[i] += sum over splits of partial[split][i].  A small gradient with many splits (432 weights, 592 CTAs' partials)
// is a long chain of dependent loads for one thread per weight: eight warps of a block share 32 weights instead, each
// summing every eighth split, and the eight sums are added in a fixed order (deterministic).
template<typename S>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const S* __restrict__ partial, int splits,
		long long elems, S* __restrict__ dw) {
	__shared__ S red[8][32];
	const int lane = threadIdx.x & 31, zl = threadIdx.x >> 5;
	for (long long base = blockIdx.x * 32ll; base < elems; base += (long long) gridDim.x * 32) {
		const long long i = base + lane;
		S s0 = 0, s1 = 0;
		if (i < elems) {
			int z = zl;
			for (; z + 8 < splits; z += 16) { s0 += partial[(long long) z * elems + i]; s1 += partial[(long long) (z + 8) * elems + i]; }
			if (z < splits) s0 += partial[(long long) z * elems + i];
		}
		red[zl][lane] = s0 + s1;
		__syncthreads();
		if (zl == 0 && i < elems) {
			S t = red[0][lane];
			#pragma unroll
			for (int k = 1; k < 8; ++k) t += red[k][lane];
			dw[i] += t;
		}
		__syncthreads();
	}
}

template<typename S>
int simt_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw) {
	const int T = gg.RH * gg.RW, R = gg.SC, J = gg.J;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) T * R * J;
	const long long gx = (long long) T * ceil_div(R, 64), gy = ceil_div(J, 64);
	long long splits = ceil_div(4ll * ctx->sm_count, gx * gy);
	const long long max_splits = ceil_div(M, 256);
	if (splits > max_splits) splits = max_splits;
	if (splits < 1) splits = 1;
	long long m_per_split = ceil_div(ceil_div(M, splits), 32) * 32;
	splits = ceil_div(M, m_per_split);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (splits * elems) * sizeof(S)));
	dim3 grid((unsigned) gx, (unsigned) gy, (unsigned) splits);
	wgrad_kernel<S><<<grid, 256, 0, ctx->stream>>>(gg, src, plain, (S*) ctx->ws, m_per_split, elems);
	CATTL3_LAUNCHED(ctx);
	wgrad_reduce_kernel<S><<<ew_grid(ctx, ceil_div(elems, 32), 1), 256, 0, ctx->stream>>>((const S*) ctx->ws,
			(int) splits, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int simt_wgrad<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, float*);
template int simt_wgrad<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, double*);

// ------------------------------------------------------------------------------------------------
// Tiny-channel layers (configs 1 and 3: 1-8 filters or input channels): a GEMM tile would be almost
// empty, the work is a few FMAs per byte, so these are written as streaming kernels.
//
// tiny_gather_gemm_kernel: one thread per output row m keeps all J <= JT outputs in registers; the
// weights sit in shared memory as [k][JT] (every lane reads the same word: broadcast), the source is
// read once per (tap, channel), coalesced along m.  Forward of few-filter convolutions, input
// gradient of few-channel ones, transposed convolutions with few output channels.
// ------------------------------------------------------------------------------------------------
template<typename S, int JT>
__global__ void __launch_bounds__(256) tiny_gather_gemm_kernel(GatherGeom gg, const S* __restrict__ src,
		const S* __restrict__ w, const S* __restrict__ bias, int bias_mode, S* __restrict__ out, int act_kind, S act_param,
		S* __restrict__ act_out, int ks) {
	// ks (1, 2 or 4) threads share an output row, each walking a contiguous ks-th of the reduction index k = tap * R + r;
	// their partial sums meet in shared memory and are added in slice order.  Small layers (fewer rows than the GPU has
	// thread slots) get ks times the threads in flight -- the thread is a chain of load latencies -- at the price of
	// one barrier.
	extern __shared__ __align__(16) unsigned char tiny_smem[];
	S* ws = reinterpret_cast<S*>(tiny_smem);
	const int T = gg.RH * gg.RW, R = gg.SC, J = gg.J, K = T * R;
	S* red = ws + K * JT;   // [ks][rows][JT], only if ks > 1
	for (int i = threadIdx.x; i < K * JT; i += 256) {
		const int j = i % JT, k = i / JT, tap = k / R, r = k - tap * R;
		ws[i] = j < J ? w[tap * gg.w_stap + r * gg.w_sr + j * gg.w_sj] : (S) 0;
	}
	__syncthreads();
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const int rows = 256 / ks;
	const int row = threadIdx.x % rows, slice = threadIdx.x / rows;
	const long long m = (long long) blockIdx.x * rows + row;
	const bool live = m < M;
	const int n = (int) (m % gg.N);
	const long long pix = m / gg.N;
	const int oh = (int) (pix % gg.OH), ow = (int) (pix / gg.OH);
	const long long plane = (long long) gg.N * gg.SH * gg.SW;
	S acc[JT];
	#pragma unroll
	for (int j = 0; j < JT; ++j) acc[j] = (S) 0;
	if (live) {
		const int kb = (int) ((long long) K * slice / ks), ke = (int) ((long long) K * (slice + 1) / ks);
		for (int tap = kb / R; tap * R < ke; ++tap) {
			const int rw = tap / gg.RH, rh = tap - rw * gg.RH;
			const int th = oh * gg.ah + rh * gg.bh + gg.ch, tw = ow * gg.aw + rw * gg.bw + gg.cw;
			if (th < 0 || tw < 0 || th % gg.denh != 0 || tw % gg.denw != 0) continue;
			const int ih = th / gg.denh, iw = tw / gg.denw;
			if (ih >= gg.SH || iw >= gg.SW) continue;
			const S* ps = src + n + (long long) gg.N * (ih + (long long) gg.SH * iw);
			const S* pw = ws + tap * R * JT;
			const int r0 = kb > tap * R ? kb - tap * R : 0, r1 = ke - tap * R < R ? ke - tap * R : R;
			for (int r = r0; r < r1; ++r) {
				const S v = __ldg(ps + r * plane);
				#pragma unroll
				for (int j = 0; j < JT; ++j) acc[j] = fma(v, pw[r * JT + j], acc[j]);
			}
		}
	}
	if (ks > 1) {
		#pragma unroll
		for (int j = 0; j < JT; ++j) red[(slice * rows + row) * JT + j] = acc[j];
		__syncthreads();
		if (slice != 0) return;
		for (int z = 1; z < ks; ++z) {
			#pragma unroll
			for (int j = 0; j < JT; ++j) acc[j] += red[(z * rows + row) * JT + j];
		}
	}
	if (!live) return;
	const long long P = (long long) gg.OH * gg.OW;
	#pragma unroll
	for (int j = 0; j < JT; ++j) {
		if (j >= J) break;
		S v = acc[j];
		if (bias_mode == 1) v += __ldg(bias + j);
		else if (bias_mode == 2) v += __ldg(bias + pix + P * j);
		if (out) out[m + M * j] = v;
		if (act_out) act_out[m + M * j] = act_fwd_rt<S>(act_kind, v, act_param);
	}
}

bool tiny_gather_gemm_supported(const GatherGeom& gg, size_t scalar_bytes) {
	const long long K = (long long) gg.RH * gg.RW * gg.SC;
	return gg.J <= 8 && K * 8 * (long long) scalar_bytes <= 40 * 1024;
}

template<typename S>
int tiny_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias, int bias_mode, S* out,
		const EpilogueArgs* ep) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	const int K = gg.RH * gg.RW * gg.SC;
	const int JT = gg.J <= 1 ? 1 : (gg.J <= 2 ? 2 : (gg.J <= 4 ? 4 : 8));
	// fewer CTAs than two per SM: let 4 (fewer than four per SM: 2) threads share a row (see the kernel)
	const long long full_ctas = ceil_div(M, 256);
	int ks = K < 8 ? 1 : (full_ctas < 2ll * ctx->sm_count ? 4 : (full_ctas < 4ll * ctx->sm_count ? 2 : 1));
	if ((size_t) (K * JT + 256 * JT) * sizeof(S) > 48 * 1024) ks = 1;   // the exchange buffer must fit the default limit
	const unsigned grid = (unsigned) ceil_div(M, 256 / ks);
	const size_t smem = (size_t) (K * JT + (ks > 1 ? 256 * JT : 0)) * sizeof(S);
	const int kind = act ? ep->act_kind : CATTL3_ACT_NONE;
	const S ap = act ? (S) ep->act_param : (S) 0;
	S* ao = act ? (S*) ep->act_out : nullptr;
#define LAUNCH(JTV) tiny_gather_gemm_kernel<S, JTV><<<grid, 256, smem, ctx->stream>>>(gg, src, w, bias, bias_mode, out, kind, ap, ao, ks)
	if (JT == 1) LAUNCH(1); else if (JT == 2) LAUNCH(2); else if (JT == 4) LAUNCH(4); else LAUNCH(8);
#undef LAUNCH
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int tiny_gather_gemm<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, const float*, int, float*,
		const EpilogueArgs*);
template int tiny_gather_gemm<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, const double*, int, double*,
		const EpilogueArgs*);

// skinny_gather_gemm_kernel: a dense layer over a small batch (M = N rows, a long reduction, J <= 16 outputs: the
// classifier heads of configs 1 and 4) would be a single CTA walking K serially.  Here the reduction is split over
// grid.y; every CTA keeps its slice of the weights in shared memory, a thread owns one row and JT partial outputs,
// and skinny_reduce_kernel adds the slices in order (deterministic) and applies bias and activation.  T = 1 only.
template<typename S, int JT>
__global__ void __launch_bounds__(256) skinny_gather_gemm_kernel(GatherGeom gg, int k_per_slice, const S* __restrict__ src,
		const S* __restrict__ w, S* __restrict__ partial) {
	extern __shared__ __align__(16) unsigned char tiny_smem[];
	S* ws = reinterpret_cast<S*>(tiny_smem);
	const int R = gg.SC, J = gg.J;
	const int k0 = blockIdx.y * k_per_slice, k1 = k0 + k_per_slice < R ? k0 + k_per_slice : R;
	for (int i = threadIdx.x; i < (k1 - k0) * JT; i += 256) {
		const int j = i % JT, r = k0 + i / JT;
		ws[i] = j < J ? w[r * gg.w_sr + j * gg.w_sj] : (S) 0;
	}
	__syncthreads();
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long m = (long long) blockIdx.x * 256 + threadIdx.x;
	if (m >= M) return;
	// T = 1, stride 1, no padding (dense / 1x1): row m of the source is m itself
	const long long plane = (long long) gg.N * gg.SH * gg.SW;
	S acc[JT];
	#pragma unroll
	for (int j = 0; j < JT; ++j) acc[j] = (S) 0;
	const S* ps = src + m + (long long) k0 * plane;
	for (int r = 0; r < k1 - k0; ++r) {
		const S v = __ldg(ps + r * plane);
		#pragma unroll
		for (int j = 0; j < JT; ++j) acc[j] = fma(v, ws[r * JT + j], acc[j]);
	}
	S* dst = partial + ((long long) blockIdx.y * J) * M + m;
	#pragma unroll
	for (int j = 0; j < JT; ++j)
		if (j < J) dst[(long long) j * M] = acc[j];
}

template<typename S>
__global__ void __launch_bounds__(256) skinny_reduce_kernel(long long M, int J, int slices, const S* __restrict__ partial,
		const S* __restrict__ bias, int bias_mode, long long N, S* __restrict__ out, int act_kind, S act_param,
		S* __restrict__ act_out) {
	const long long total = M * J;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long) gridDim.x * 256) {
		S v = 0;
		for (int z = 0; z < slices; ++z) v += partial[(long long) z * total + i];
		const long long j = i / M, m = i - j * M;
		if (bias_mode == 1) v += __ldg(bias + j);
		else if (bias_mode == 2) v += __ldg(bias + m / N + (M / N) * j);
		if (out) out[i] = v;
		if (act_out) act_out[i] = act_fwd_rt<S>(act_kind, v, act_param);
	}
}

bool skinny_gather_gemm_supported(const cattl3_ctx* ctx, const GatherGeom& gg) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	return gg.RH == 1 && gg.RW == 1 && gg.J <= 16 && gg.SC >= 1024 && ceil_div(M, 256) * 4 <= ctx->sm_count &&
			gg.SH == gg.OH && gg.SW == gg.OW && gg.ah == 1 && gg.aw == 1 && gg.ch == 0 && gg.cw == 0 && gg.denh == 1 && gg.denw == 1;
}

template<typename S>
int skinny_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias, int bias_mode, S* out,
		const EpilogueArgs* ep) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	const int JT = gg.J <= 4 ? 4 : (gg.J <= 8 ? 8 : 16);
	const long long row_blocks = ceil_div(M, 256);
	long long slices = 2ll * ctx->sm_count / row_blocks;
	const long long max_slices = ceil_div(gg.SC, 128);
	if (slices > max_slices) slices = max_slices;
	int k_per_slice = (int) ceil_div(gg.SC, slices);
	if ((size_t) k_per_slice * JT * sizeof(S) > 40 * 1024) k_per_slice = (int) (40 * 1024 / (JT * sizeof(S)));
	slices = ceil_div(gg.SC, k_per_slice);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (slices * M * gg.J) * sizeof(S)));
	dim3 grid((unsigned) row_blocks, (unsigned) slices);
	const size_t smem = (size_t) k_per_slice * JT * sizeof(S);
#define LAUNCH(JTV) skinny_gather_gemm_kernel<S, JTV><<<grid, 256, smem, ctx->stream>>>(gg, k_per_slice, src, w, (S*) ctx->ws)
	if (JT == 4) LAUNCH(4); else if (JT == 8) LAUNCH(8); else LAUNCH(16);
#undef LAUNCH
	CATTL3_LAUNCHED(ctx);
	skinny_reduce_kernel<S><<<ew_grid(ctx, M * gg.J, 256), 256, 0, ctx->stream>>>(M, gg.J, (int) slices, (const S*) ctx->ws, bias,
			bias_mode, gg.N, out, act ? ep->act_kind : CATTL3_ACT_NONE, act ? (S) ep->act_param : (S) 0,
			act ? (S*) ep->act_out : (S*) nullptr);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int skinny_gather_gemm<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, const float*, int, float*,
		const EpilogueArgs*);
template int skinny_gather_gemm<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, const double*, int, double*,
		const EpilogueArgs*);

// tiny_wgrad_kernel: the whole K x J gradient (K = taps * channels <= 256, K * J <= 2048) is one CTA's
// output; the CTAs split the reduction over m.  Per chunk of MC rows the gathered source rows [MC][K]
// and the plain rows [MC][J] go to shared memory (coalesced along m), then every thread adds the chunk
// into its <= 8 outputs: <= 8 / JT items of one reduction index k and JT consecutive filters, so that one
// read of the source value and one (vector) read of the JT plain values feed JT FMAs (with one output per
// item the loop was two shared-memory reads per FMA and bound by them).  Partials per CTA, deterministic
// reduce (wgrad_reduce_kernel).
template<typename S, int JT>
__global__ void __launch_bounds__(256) tiny_wgrad_kernel(GatherGeom gg, const S* __restrict__ src,
		const S* __restrict__ plain, S* __restrict__ partial, long long m_per_cta, long long dw_elems, int Jp) {
	// Jp: the filter count rounded up to a multiple of JT; the plain rows are zero padded to it in shared memory
	constexpr int MC = 32, ITEMS = 8 / JT;
	extern __shared__ __align__(16) unsigned char tiny_smem[];
	const int T = gg.RH * gg.RW, R = gg.SC, J = gg.J, K = T * R, JG = Jp / JT, KG = K * JG;
	const int Kp = K | 1;   // odd pitch: the transposing stores spread over the banks
	S* Bs = reinterpret_cast<S*>(tiny_smem);       // [MC][Jp] (first: 16-byte aligned rows when Jp % 4 == 0)
	S* As = Bs + MC * Jp;                           // [MC][Kp]
	int* ktab = reinterpret_cast<int*>(As + MC * Kp); // per k: rh*bh + ch (16 bits) | rw*bw + cw (16 bits), and the channel
	for (int k = threadIdx.x; k < K; k += 256) {
		const int tap = k / R, r = k - tap * R, rw = tap / gg.RH, rh = tap - rw * gg.RH;
		ktab[2 * k] = ((rh * gg.bh + gg.ch) << 16) | ((rw * gg.bw + gg.cw) & 0xFFFF);
		ktab[2 * k + 1] = r;
	}
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long ms = (long long) blockIdx.x * m_per_cta;
	const long long me = ms + m_per_cta < M ? ms + m_per_cta : M;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;
	const int mm = threadIdx.x % MC, l0 = threadIdx.x / MC;   // loader: row mm of the chunk, items l0 + 8 * i
	S acc[ITEMS][JT];
	int item_k[ITEMS], item_j[ITEMS];
	#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		const int o = threadIdx.x + 256 * i;
		item_j[i] = o < KG ? (o / K) * JT : -1;
		item_k[i] = o < KG ? o % K : 0;
		#pragma unroll
		for (int t = 0; t < JT; ++t) acc[i][t] = (S) 0;
	}
	__syncthreads();
	for (long long mc = ms; mc < me; mc += MC) {
		const long long m = mc + mm;
		const bool ok = m < me;
		const int n = (int) (m % gg.N);
		const long long pix = m / gg.N;
		const int oh = (int) (pix % gg.OH), ow = (int) (pix / gg.OH);
		for (int k = l0; k < K; k += 256 / MC) {
			const int packed = ktab[2 * k];
			const int th = oh * gg.ah + (packed >> 16), tw = ow * gg.aw + (int) (short) (packed & 0xFFFF);
			S v = (S) 0;
			if (ok && th >= 0 && tw >= 0 && th < gg.SH && tw < gg.SW)   // forward-style gathers only (denh = denw = 1)
				v = __ldg(src + n + (long long) gg.N * (th + (long long) gg.SH * tw) + ktab[2 * k + 1] * plane);
			As[mm * Kp + k] = v;
		}
		for (int j = l0; j < Jp; j += 256 / MC) Bs[mm * Jp + j] = ok && j < J ? __ldg(plain + m + M * j) : (S) 0;
		__syncthreads();
		#pragma unroll
		for (int i = 0; i < ITEMS; ++i) {
			if (item_j[i] >= 0) {
				const S* ap = As + item_k[i];
				const S* bp = Bs + item_j[i];
				#pragma unroll 8
				for (int r = 0; r < MC; ++r) {
					const S a = ap[r * Kp];
					S bv[JT];
					if constexpr (JT == 4 && sizeof(S) == 4) {
						const float4 v = *reinterpret_cast<const float4*>(bp + r * Jp);
						bv[0] = v.x; bv[1] = v.y; bv[2] = v.z; bv[3] = v.w;
					} else {
						#pragma unroll
						for (int t = 0; t < JT; ++t) bv[t] = bp[r * Jp + t];
					}
					#pragma unroll
					for (int t = 0; t < JT; ++t) acc[i][t] = fma(a, bv[t], acc[i][t]);
				}
			}
		}
		__syncthreads();
	}
	S* dst = partial + (long long) blockIdx.x * dw_elems;
	#pragma unroll
	for (int i = 0; i < ITEMS; ++i) {
		if (item_j[i] >= 0) {
			const int k = item_k[i], tap = k / R, r = k - tap * R;
			#pragma unroll
			for (int t = 0; t < JT; ++t)
				if (item_j[i] + t < J) dst[tap * gg.w_stap + r * gg.w_sr + (item_j[i] + t) * gg.w_sj] = acc[i][t];
		}
	}
}

bool tiny_wgrad_supported(const GatherGeom& gg, size_t scalar_bytes) {
	const long long K = (long long) gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long Jp = (gg.J + 3) / 4 * 4;   // the plain rows are padded to four filters per thread
	const long long smem = (32 * ((K | 1) + Jp)) * (long long) scalar_bytes + 8 * K;
	// |rh*bh + ch| and |rw*bw + cw| are packed into 16 bits each
	const long long reach = (long long) (gg.RH > gg.RW ? gg.RH : gg.RW) * (gg.bh > gg.bw ? gg.bh : gg.bw) + gg.SH + gg.SW;
	return gg.J <= 32 && K <= 256 && K * Jp <= 2048 && smem <= 44 * 1024 && gg.denh == 1 && gg.denw == 1 && M >= 512 &&
			reach < 30000 && gg.bh > 0 && gg.bw > 0;
}

template<typename S>
int tiny_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw) {
	const int K = gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) K * gg.J;
	// a CTA's chunks of 32 rows run one after the other (load, barrier, accumulate, barrier), so short row ranges -- four
	// chunks -- and many CTAs are what keeps the SMs busy on the small layers this kernel exists for
	long long ctas = ceil_div(M, 128);
	if (ctas > 8ll * ctx->sm_count) ctas = 8ll * ctx->sm_count;
	const long long m_per_cta = ceil_div(ceil_div(M, ctas), 32) * 32;
	ctas = ceil_div(M, m_per_cta);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (ctas * elems) * sizeof(S)));
	// four filters per thread whenever there is more than one (three filters: one lane of the four idles, still three FMAs
	// per two shared-memory reads instead of one); a single filter keeps one output per item
	const int Jp = gg.J == 1 ? 1 : (gg.J + 3) / 4 * 4;
	const size_t smem = (size_t) (32 * ((K | 1) + Jp)) * sizeof(S) + 8 * (size_t) K;
	if (Jp > 1)
		tiny_wgrad_kernel<S, 4><<<(unsigned) ctas, 256, smem, ctx->stream>>>(gg, src, plain, (S*) ctx->ws, m_per_cta, elems, Jp);
	else
		tiny_wgrad_kernel<S, 1><<<(unsigned) ctas, 256, smem, ctx->stream>>>(gg, src, plain, (S*) ctx->ws, m_per_cta, elems, Jp);
	CATTL3_LAUNCHED(ctx);
	wgrad_reduce_kernel<S><<<ew_grid(ctx, ceil_div(elems, 32), 1), 256, 0, ctx->stream>>>((const S*) ctx->ws, (int) ctas, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int tiny_wgrad<float>(cattl3_ctx*, const GatherGeom&, const float*, const float*, float*);
template int tiny_wgrad<double>(cattl3_ctx*, const GatherGeom&, const double*, const double*, double*);

// ------------------------------------------------------------------------------------------------
// out[col] += sum_{row} a[row + rows*col]   (bias gradients: ConvKernelLayer.hpp:155,
// TransConvKernelLayer.hpp:160-161, DenseKernelLayer.hpp:109).  Fixed reduction tree => deterministic.
// ------------------------------------------------------------------------------------------------
template<typename S>
__global__ void __launch_bounds__(256) colsum_block_kernel(long long rows, const S* __restrict__ a,
		S* __restrict__ out) {
	__shared__ S red[256];
	const S* col = a + rows * (long long) blockIdx.x;
	S s = 0;
	for (long long i = threadIdx.x; i < rows; i += 256) s += col[i];
	red[threadIdx.x] = s;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int) threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) out[blockIdx.x] += red[0];
}

template<typename S>
__global__ void __launch_bounds__(256) colsum_thread_kernel(long long rows, long long cols,
		const S* __restrict__ a, S* __restrict__ out) {
	for (long long c = blockIdx.x * 256ll + threadIdx.x; c < cols; c += (long long) gridDim.x * 256) {
		S s = 0;
		for (long long i = 0; i < rows; ++i) s += a[i + rows * c];
		out[c] += s;
	}
}

// Few long columns (the bias gradient of a layer with a handful of filters over many pixels: 16 columns of 524 288
// rows left 132 SMs idle with one block per column): grid (columns, chunks) of per-chunk sums into scratch, then one
// thread per column adds the chunks in order.  Deterministic.
template<typename S>
__global__ void __launch_bounds__(256) colsum_chunk_kernel(long long rows, long long chunk, const S* __restrict__ a,
		S* __restrict__ partial) {
	__shared__ S red[256];
	const S* col = a + rows * (long long) blockIdx.x;
	const long long lo = (long long) blockIdx.y * chunk, hi = lo + chunk < rows ? lo + chunk : rows;
	S s0 = 0, s1 = 0, s2 = 0, s3 = 0;
	long long i = lo + threadIdx.x;
	for (; i + 768 < hi; i += 1024) { s0 += col[i]; s1 += col[i + 256]; s2 += col[i + 512]; s3 += col[i + 768]; }
	for (; i < hi; i += 256) s0 += col[i];
	red[threadIdx.x] = (s0 + s1) + (s2 + s3);
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int) threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[(long long) blockIdx.x * gridDim.y + blockIdx.y] = red[0];
}
template<typename S>
__global__ void __launch_bounds__(256) colsum_final_kernel(long long cols, int chunks, const S* __restrict__ partial,
		S* __restrict__ out) {
	const long long c = blockIdx.x * 256ll + threadIdx.x;
	if (c >= cols) return;
	S s = 0;
	for (int k = 0; k < chunks; ++k) s += partial[c * chunks + k];
	out[c] += s;
}

template<typename S>
int colsum_accumulate(cattl3_ctx* ctx, int64_t rows, int64_t cols, const S* a, S* out) {
	const int64_t chunks = std::min<int64_t>(ceil_div((int64_t) 4 * ctx->sm_count, cols), rows / 4096);
	if (chunks > 1 && cols <= 65535) {
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (cols * chunks) * sizeof(S)));
		const int64_t chunk = ceil_div(ceil_div(rows, chunks), 256) * 256;
		const int64_t used = ceil_div(rows, chunk);
		colsum_chunk_kernel<S><<<dim3((unsigned) cols, (unsigned) used), 256, 0, ctx->stream>>>(rows, chunk, a, (S*) ctx->ws);
		CATTL3_LAUNCHED(ctx);
		colsum_final_kernel<S><<<(unsigned) ceil_div(cols, 256), 256, 0, ctx->stream>>>(cols, (int) used, (const S*) ctx->ws, out);
	} else if (rows >= 128) {
		colsum_block_kernel<S><<<(unsigned) cols, 256, 0, ctx->stream>>>(rows, a, out);
	} else {
		colsum_thread_kernel<S><<<ew_grid(ctx, cols, 256), 256, 0, ctx->stream>>>(rows, cols, a, out);
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
// Shifted column statistics for a following BatchNormLayer where the GEMM epilogue did not produce them (SIMT path):
// grid (cols, chunks); per-chunk double partials, then a fixed-order sum.
template<typename S>
__global__ void __launch_bounds__(256) colstats_partial_kernel(long long rows, long long chunk, const S* __restrict__ a,
		const S* __restrict__ shift, double* __restrict__ partial) {
	__shared__ double r1[256], r2[256];
	const long long j = blockIdx.x;
	const S* col = a + rows * j;
	const double K = (double) shift[j];
	const long long lo = (long long) blockIdx.y * chunk, hi = lo + chunk < rows ? lo + chunk : rows;
	double s1 = 0, s2 = 0;
	for (long long i = lo + threadIdx.x; i < hi; i += 256) { const double d = (double) col[i] - K; s1 += d; s2 += d * d; }
	r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int) threadIdx.x < o) { r1[threadIdx.x] += r1[threadIdx.x + o]; r2[threadIdx.x] += r2[threadIdx.x + o]; }
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		partial[(j * gridDim.y + blockIdx.y) * 2] = r1[0];
		partial[(j * gridDim.y + blockIdx.y) * 2 + 1] = r2[0];
	}
}
__global__ void __launch_bounds__(256) colstats_final_kernel(long long cols, int chunks, const double* __restrict__ partial,
		double* __restrict__ col_stats) {
	const long long j = blockIdx.x * 256ll + threadIdx.x;
	if (j >= cols) return;
	double s1 = 0, s2 = 0;
	for (int c = 0; c < chunks; ++c) { s1 += partial[(j * chunks + c) * 2]; s2 += partial[(j * chunks + c) * 2 + 1]; }
	col_stats[j] = s1;
	col_stats[cols + j] = s2;
}

template<typename S>
int colstats_shifted(cattl3_ctx* ctx, int64_t rows, int64_t cols, const S* a, const S* shift, double* col_stats) {
	long long chunks = ceil_div(8ll * ctx->sm_count, cols);
	const long long maxc = ceil_div(rows, 2048);
	if (chunks > maxc) chunks = maxc;
	if (chunks < 1) chunks = 1;
	const long long chunk = ceil_div(rows, chunks);
	chunks = ceil_div(rows, chunk);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->stat_ws, &ctx->stat_ws_bytes, (size_t) (cols * chunks * 2) * sizeof(double)));
	dim3 grid((unsigned) cols, (unsigned) chunks);
	colstats_partial_kernel<S><<<grid, 256, 0, ctx->stream>>>(rows, chunk, a, shift, (double*) ctx->stat_ws);
	CATTL3_LAUNCHED(ctx);
	colstats_final_kernel<<<(unsigned) ceil_div(cols, 256), 256, 0, ctx->stream>>>(cols, (int) chunks,
			(const double*) ctx->stat_ws, col_stats);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template int colstats_shifted<float>(cattl3_ctx*, int64_t, int64_t, const float*, const float*, double*);
template int colstats_shifted<double>(cattl3_ctx*, int64_t, int64_t, const double*, const double*, double*);

template int colsum_accumulate<float>(cattl3_ctx*, int64_t, int64_t, const float*, float*);
template int colsum_accumulate<double>(cattl3_ctx*, int64_t, int64_t, const double*, double*);

} // namespace cattl3
