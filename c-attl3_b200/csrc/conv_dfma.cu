// conv_dfma.cu -- the double instantiation of the kernel layers at GEMM-sized shapes: implicit-GEMM kernels built
// around the FP64 FMA pipe (64 DFMA / clk / SM on B200: half the FP32 rate, no tensor-core path for fp64 in this
// design).  A DFMA kernel is bound by that pipe only if little else competes for issue slots and shared memory, so:
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

template<int TN>
__global__ void __launch_bounds__(DF_THREADS, TN == 4 ? 2 : 1) dfma_gather_gemm_kernel(GatherGeom gg, const double* __restrict__ src,
		const double* __restrict__ w, const double* __restrict__ bias, int bias_mode, double* __restrict__ out,
		int act_kind, double act_param, double* __restrict__ act_out, int vec_ok) {
	constexpr int BN = 16 * TN;
	constexpr int B_ITERS = DF_BK * BN / DF_THREADS;   // 4 (BN = 128) or 2 (BN = 64)
	constexpr int B_KSTEP = DF_THREADS / BN;           // 2 or 4
	__shared__ __align__(16) double As[2][DF_BK][DF_BM];
	__shared__ __align__(16) double Bs[2][DF_BK][BN];

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

	double acc[8][TN];
	#pragma unroll
	for (int i = 0; i < 8; ++i)
		#pragma unroll
		for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;

	// The reduction walks taps outer, channel blocks of DF_BK inner (a partial last block is zero padded): the gather
	// coordinates -- divisions, bounds and divisibility tests -- are evaluated once per tap; inside a tap a load is
	// base + r * plane.  The fetch stream runs one k-block ahead of the multiply stream.
	const int T = gg.RH * gg.RW;
	const int rblocks = (R + DF_BK - 1) / DF_BK;
	const int ksteps = T * rblocks;
	int f_tap = 0, f_r0 = 0;
	bool f_ok = false;
	const double* f_src = src;
	const double* f_w = w + wj;
	auto enter_tap = [&]() {
		const int rw = f_tap / gg.RH, rh = f_tap - rw * gg.RH;
		const int th = aoh * gg.ah + rh * gg.bh + gg.ch;
		const int tw = aow * gg.aw + rw * gg.bw + gg.cw;
		f_ok = false;
		if (m_ok && th >= 0 && tw >= 0 && th % gg.denh == 0 && tw % gg.denw == 0) {
			const int ih = th / gg.denh, iw = tw / gg.denw;
			if (ih < gg.SH && iw < gg.SW) {
				f_ok = true;
				f_src = src + an + (long long) gg.N * (ih + (long long) gg.SH * iw);
			}
		}
		f_w = w + wj + f_tap * gg.w_stap;
	};
	double pa[4], pb[B_ITERS];
	auto fetch = [&]() {
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int r = f_r0 + a_k0 + 2 * i;
			pa[i] = (f_ok && r < R) ? __ldg(f_src + r * plane) : 0.0;
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int r = f_r0 + b_k0 + B_KSTEP * i;
			pb[i] = (j_ok && r < R) ? __ldg(f_w + r * gg.w_sr) : 0.0;
		}
		f_r0 += DF_BK;
		if (f_r0 >= R) {
			f_r0 = 0;
			if (++f_tap < T) enter_tap();
		}
	};
	auto stash = [&](int buf) {
		#pragma unroll
		for (int i = 0; i < 4; ++i) As[buf][a_k0 + 2 * i][a_ml] = pa[i];
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) Bs[buf][b_k0 + B_KSTEP * i][b_j] = pb[i];
	};

	enter_tap();
	fetch();
	stash(0);
	__syncthreads();
	for (int ks = 0; ks < ksteps; ++ks) {
		const int buf = ks & 1;
		if (ks + 1 < ksteps) fetch();
		#pragma unroll
		for (int kk = 0; kk < DF_BK; ++kk) {
			double a[8], b[TN];
			#pragma unroll
			for (int i = 0; i < 4; ++i)
				*reinterpret_cast<double2*>(&a[2 * i]) = *reinterpret_cast<const double2*>(&As[buf][kk][rm + 16 * i]);
			#pragma unroll
			for (int jq = 0; jq < TN / 2; ++jq)
				*reinterpret_cast<double2*>(&b[2 * jq]) = *reinterpret_cast<const double2*>(&Bs[buf][kk][cn + 8 * jq]);
			#pragma unroll
			for (int i = 0; i < 8; ++i)
				#pragma unroll
				for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
		}
		if (ks + 1 < ksteps) stash(buf ^ 1);
		__syncthreads();
	}

	// epilogue: bias, fused activation, stores (16-byte pairs along m where the tensor allows it)
	const long long P = (long long) gg.OH * gg.OW;
	#pragma unroll
	for (int jq = 0; jq < TN / 2; ++jq) {
		#pragma unroll
		for (int f = 0; f < 2; ++f) {
			const int j = j0 + cn + 8 * jq + f;
			if (j >= J) continue;
			const double bj = bias_mode == 1 ? __ldg(bias + j) : 0.0;
			#pragma unroll
			for (int i = 0; i < 4; ++i) {
				const long long m = m0 + rm + 16 * i;
				if (m >= M) continue;
				double v0 = acc[2 * i][2 * jq + f] + bj, v1 = acc[2 * i + 1][2 * jq + f] + bj;
				const bool two = m + 1 < M;
				if (bias_mode == 2) {
					v0 += __ldg(bias + m / gg.N + P * j);
					if (two) v1 += __ldg(bias + (m + 1) / gg.N + P * j);
				}
				const long long o = m + M * j;
				if (out) {
					if (vec_ok) *reinterpret_cast<double2*>(out + o) = make_double2(v0, v1);
					else { out[o] = v0; if (two) out[o + 1] = v1; }
				}
				if (act_out) {
					v0 = act_fwd_rt<double>(act_kind, v0, act_param);
					v1 = act_fwd_rt<double>(act_kind, v1, act_param);
					if (vec_ok) *reinterpret_cast<double2*>(act_out + o) = make_double2(v0, v1);
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

template<int TN>
__global__ void __launch_bounds__(DF_THREADS, TN == 4 ? 2 : 1) dfma_wgrad_kernel(GatherGeom gg, const double* __restrict__ src,
		const double* __restrict__ plain, double* __restrict__ partial, long long m_per_split, long long dw_elems) {
	constexpr int BN = 16 * TN;
	constexpr int PITCH_A = DF_BM + 2, PITCH_B = BN + 2;
	constexpr int B_ITERS = BN / 32;
	__shared__ __align__(16) double As[2][DW_BKM][PITCH_A];
	__shared__ __align__(16) double Bs[2][DW_BKM][PITCH_B];

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

	double acc[8][TN];
	#pragma unroll
	for (int i = 0; i < 8; ++i)
		#pragma unroll
		for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;

	double pa[4], pb[B_ITERS];
	auto fetch = [&](long long mc) {
		const long long m = mc + l_mm;
		const bool ok = m < me;
		// M < 2^31 (dfma_wgrad_supported): 32-bit divisions
		const unsigned mu = (unsigned) m, pixu = mu / (unsigned) gg.N;
		const int n = (int) (mu - pixu * (unsigned) gg.N);
		const int ow = (int) (pixu / (unsigned) gg.OH), oh = (int) (pixu - (unsigned) ow * (unsigned) gg.OH);
		const int bh = oh * gg.ah, bw = ow * gg.aw;
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			double v = 0.0;
			const int th = bh + row_dh[i], tw = bw + row_dw[i];
			// weight gradients only ever gather forward-style (denh = denw = 1, dfma_wgrad_supported)
			if (ok && row_r[i] >= 0 && th >= 0 && tw >= 0 && th < gg.SH && tw < gg.SW)
				v = __ldg(src + n + (long long) gg.N * (th + (long long) gg.SH * tw) + row_r[i] * plane);
			pa[i] = v;
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int j = j0 + l_r0 + 32 * i;
			pb[i] = (ok && j < J) ? __ldg(plain + m + M * j) : 0.0;
		}
	};
	auto stash = [&](int buf) {
		#pragma unroll
		for (int i = 0; i < 4; ++i) As[buf][l_mm][l_r0 + 32 * i] = pa[i];
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) Bs[buf][l_mm][l_r0 + 32 * i] = pb[i];
	};

	const long long steps = me > ms ? (me - ms + DW_BKM - 1) / DW_BKM : 0;
	if (steps > 0) {
		fetch(ms);
		stash(0);
	}
	__syncthreads();
	for (long long st = 0; st < steps; ++st) {
		const int buf = (int) (st & 1);
		if (st + 1 < steps) fetch(ms + (st + 1) * DW_BKM);
		#pragma unroll
		for (int mk = 0; mk < DW_BKM; ++mk) {
			double a[8], b[TN];
			#pragma unroll
			for (int i = 0; i < 4; ++i)
				*reinterpret_cast<double2*>(&a[2 * i]) = *reinterpret_cast<const double2*>(&As[buf][mk][rm + 16 * i]);
			#pragma unroll
			for (int jq = 0; jq < TN / 2; ++jq)
				*reinterpret_cast<double2*>(&b[2 * jq]) = *reinterpret_cast<const double2*>(&Bs[buf][mk][cn + 8 * jq]);
			#pragma unroll
			for (int i = 0; i < 8; ++i)
				#pragma unroll
				for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
		}
		if (st + 1 < steps) stash(buf ^ 1);
		__syncthreads();
	}

	double* dst = partial + (long long) blockIdx.z * dw_elems;
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

__global__ void __launch_bounds__(256) dfma_wgrad_reduce_kernel(const double* __restrict__ partial, int splits, long long elems,
		double* __restrict__ dw) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < elems; i += (long long) gridDim.x * 256) {
		double s = 0;
		for (int z = 0; z < splits; ++z) s += partial[(long long) z * elems + i];
		dw[i] += s;
	}
}

} // namespace

// The big-tile kernels pay off once a tile is mostly real work: enough filters for a 64-wide tile and a reduction
// longer than a couple of k-blocks.
bool dfma_gather_gemm_supported(const GatherGeom& gg) {
	// reduce channels are walked in blocks of 8 per tap: tiny channel counts would mostly multiply padding
	return gg.J > 32 && gg.SC >= 6 && (long long) gg.RH * gg.RW * gg.SC >= 32;
}

int dfma_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* w, const double* bias,
		int bias_mode, double* out, const EpilogueArgs* ep) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	double* act_out = act ? (double*) ep->act_out : nullptr;
	const int vec_ok = M % 2 == 0 && (!out || aligned16(out)) && (!act_out || aligned16(act_out));
	const int act_kind = act ? ep->act_kind : CATTL3_ACT_NONE;
	const double act_param = act ? ep->act_param : 0.0;
	// measured at config 2 (profiles/README.md, r1e): 64 filters -> 128 x 64 tiles, two CTAs per SM (20 TFLOP/s against
	// 12 with half-empty 128-wide tiles); 256 filters -> 128 x 128 tiles (18.8 against 15.7)
	if (gg.J > 64) {
		dim3 grid((unsigned) ceil_div(M, DF_BM), (unsigned) ceil_div(gg.J, 128));
		dfma_gather_gemm_kernel<8><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, w, bias, bias_mode, out, act_kind, act_param,
				act_out, vec_ok);
	} else {
		dim3 grid((unsigned) ceil_div(M, DF_BM), (unsigned) ceil_div(gg.J, 64));
		dfma_gather_gemm_kernel<4><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, w, bias, bias_mode, out, act_kind, act_param,
				act_out, vec_ok);
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

bool dfma_wgrad_supported(const GatherGeom& gg) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	return gg.J > 32 && (long long) gg.RH * gg.RW * gg.SC >= 64 && M >= 1024 && M < (1ll << 31) && gg.denh == 1 && gg.denw == 1;
}

int dfma_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* plain, double* dw) {
	const int Ktot = gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) Ktot * gg.J;
	const int BN = gg.J > 64 ? 128 : 64;
	const long long gx = ceil_div(Ktot, DF_BM), gy = ceil_div(gg.J, BN);
	// one (BN = 128) or two (BN = 64) CTAs per SM: as many m-splits as fill the machine once
	long long splits = (BN == 128 ? 1 : 2) * ctx->sm_count / (gx * gy);
	if (splits < 1) splits = 1;
	const long long max_splits = ceil_div(M, 1024);
	if (splits > max_splits) splits = max_splits;
	long long m_per_split = ceil_div(ceil_div(M, splits), DW_BKM) * DW_BKM;
	splits = ceil_div(M, m_per_split);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (splits * elems) * sizeof(double)));
	// rows of a tile beyond Ktot and columns beyond J are never written: the reduce reads only real elements
	dim3 grid((unsigned) gx, (unsigned) gy, (unsigned) splits);
	if (BN == 128)
		dfma_wgrad_kernel<8><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	else
		dfma_wgrad_kernel<4><<<grid, DF_THREADS, 0, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	CATTL3_LAUNCHED(ctx);
	dfma_wgrad_reduce_kernel<<<ew_grid(ctx, elems, 256), 256, 0, ctx->stream>>>((const double*) ctx->ws, (int) splits, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

} // namespace cattl3
