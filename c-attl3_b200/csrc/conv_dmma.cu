// conv_dmma.cu -- the double instantiation of the kernel layers on the FP64 TENSOR pipe: warp-level
// mma.sync.m8n8k4.f64 (DMMA).  Measured on B200 (scripts/dmma_peak.cu): 37.0 TFLOP/s, against 33.7 for plain DFMA
// (scripts/dfma_peak.cu) -- but what matters more is what each instruction carries: one DMMA is 256 FMAs for two
// 8-byte fragment loads, where the FMA kernels (conv_dfma.cu) issue 64 DFMAs per eight 16-byte shared-memory loads
// and stall on exactly those loads (profiles/README.md, r1e: FP64 pipe 51 % active, short-scoreboard stalls).
// tcgen05 has no fp64 kind, so this is the tensor-core path there is for double.
//
// Same implicit GEMM, same tiling as conv_dfma.cu: 128 x 64 (weight gradient: also 128 x 128) block tiles, 256 threads =
// 8 warps as 2 (rows) x 4 (columns), a warp owns 64 x 16 (64 x 32) as 8 x 2 (8 x 4) m8n8k4 tiles, accumulators in registers;
// tiles staged global -> shared with cp.async, two buffers, one barrier per k-block of 8; no im2col buffer.
// Fragments (PTX ISA, mma.m8n8k4 .f64): A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4],
// C[row = lane / 4][col = 2 * (lane % 4) + {0, 1}].  Shared-memory rows are k (the reduction index) with a pitch of
// tile + 4 doubles, which makes both the fragment reads and the weight gradient's transposing writes conflict free.
#include "activations.cuh"

namespace cattl3 {

namespace {

constexpr int DM_BM = 128, DM_BK = 8, DM_THREADS = 256, DM_PAD = 4;

__device__ __forceinline__ void cp_async_f64(double* dst, const double* src, bool pred) {
	const uint32_t d = (uint32_t) __cvta_generic_to_shared(dst);
	const int bytes = pred ? 8 : 0;
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_f64() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_f64() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
			: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One k-block (8 reduction steps) of the warp's 64 x (8 * NI) tile: As / Bs point at [k][row] / [k][col] of the
// current buffer, row0 / col0 are the warp's offsets in the block tile.
template<int NI, int PA, int PB>
__device__ __forceinline__ void dmma_block(const double* __restrict__ As, const double* __restrict__ Bs, int row0, int col0,
		int lane, double (&acc)[8][NI][2]) {
	const int kq = lane & 3, g = lane >> 2;
	#pragma unroll
	for (int k4 = 0; k4 < DM_BK; k4 += 4) {
		double a[8], b[NI];
		#pragma unroll
		for (int mi = 0; mi < 8; ++mi) a[mi] = As[(k4 + kq) * PA + row0 + 8 * mi + g];
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) b[ni] = Bs[(k4 + kq) * PB + col0 + 8 * ni + g];
		#pragma unroll
		for (int mi = 0; mi < 8; ++mi)
			#pragma unroll
			for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
	}
}

// Gather GEMM: out[m + M*j] = bias + sum_{tap, r} src(m, tap, r) * w(tap, r, j) (+ fused activation).
template<int NI>
__global__ void __launch_bounds__(DM_THREADS, NI == 2 ? 2 : 1) dmma_gather_gemm_kernel(GatherGeom gg,
		const double* __restrict__ src, const double* __restrict__ w, const double* __restrict__ bias, int bias_mode,
		double* __restrict__ out, int act_kind, double act_param, double* __restrict__ act_out) {
	constexpr int BN = 32 * NI;
	constexpr int PA = DM_BM + DM_PAD, PB = BN + DM_PAD;
	constexpr int B_ITERS = DM_BK * BN / DM_THREADS;   // 4 (BN = 128) or 2 (BN = 64)
	constexpr int B_KSTEP = DM_THREADS / BN;           // 2 or 4
	__shared__ __align__(16) double As[2][DM_BK * PA];
	__shared__ __align__(16) double Bs[2][DM_BK * PB];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int row0 = (warp & 1) * 64, col0 = (warp >> 1) * (8 * NI);
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long m0 = (long long) blockIdx.x * DM_BM;
	const int j0 = blockIdx.y * BN;
	const int R = gg.SC, J = gg.J, T = gg.RH * gg.RW;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;
	const bool unit_den = gg.denh == 1 && gg.denw == 1;

	// loader: this thread always fetches row a_ml of the A tile and column b_j of the B tile
	const int a_ml = tid & (DM_BM - 1), a_k0 = tid >> 7;
	const long long am = m0 + a_ml;
	const bool m_ok = am < M;
	const int an = (int) (am % gg.N);
	const long long apix = am / gg.N;
	const int aoh = (int) (apix % gg.OH), aow = (int) (apix / gg.OH);
	const int b_j = tid & (BN - 1), b_k0 = tid / BN;
	const bool j_ok = j0 + b_j < J;
	const long long wj = (long long) (j0 + b_j) * gg.w_sj;

	double acc[8][NI][2];
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

	// taps outer, channel blocks of 8 inner (a partial last block is zero filled); coordinates once per tap
	auto tap_src = [&](int tap) -> const double* {
		const int rw = tap / gg.RH, rh = tap - rw * gg.RH;
		const int th = aoh * gg.ah + rh * gg.bh + gg.ch;
		const int tw = aow * gg.aw + rw * gg.bw + gg.cw;
		if (!m_ok || th < 0 || tw < 0) return nullptr;
		int ih = th, iw = tw;
		if (!unit_den) {
			if (th % gg.denh != 0 || tw % gg.denw != 0) return nullptr;
			ih = th / gg.denh; iw = tw / gg.denw;
		}
		if (ih >= gg.SH || iw >= gg.SW) return nullptr;
		return src + an + (long long) gg.N * (ih + (long long) gg.SH * iw);
	};
	const int rblocks = (R + DM_BK - 1) / DM_BK;
	const int ksteps = T * rblocks;
	int f_tap = 0, f_r0 = 0;
	const double* f_src = tap_src(0);
	const double* f_w = w + wj;
	auto fetch = [&](int buf) {
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int r = f_r0 + a_k0 + 2 * i;
			const bool ok = f_src && r < R;
			cp_async_f64(&As[buf][(a_k0 + 2 * i) * PA + a_ml], ok ? f_src + r * plane : src, ok);
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int r = f_r0 + b_k0 + B_KSTEP * i;
			const bool ok = j_ok && r < R;
			cp_async_f64(&Bs[buf][(b_k0 + B_KSTEP * i) * PB + b_j], ok ? f_w + r * gg.w_sr : w, ok);
		}
		cp_async_commit_f64();
		f_r0 += DM_BK;
		if (f_r0 >= R) {
			f_r0 = 0;
			if (++f_tap < T) {
				f_src = tap_src(f_tap);
				f_w = w + wj + f_tap * gg.w_stap;
			}
		}
	};
	fetch(0);
	cp_async_wait_f64();
	__syncthreads();
	for (int ks = 0; ks < ksteps; ++ks) {
		const int buf = ks & 1;
		if (ks + 1 < ksteps) fetch(buf ^ 1);
		dmma_block<NI, PA, PB>(As[buf], Bs[buf], row0, col0, lane, acc);
		cp_async_wait_f64();
		__syncthreads();
	}

	// epilogue: lane holds rows 8 mi + lane / 4, columns 8 ni + 2 (lane % 4) + {0, 1} of the warp tile
	const long long P = (long long) gg.OH * gg.OW;
	const int g = lane >> 2, kq = lane & 3;
	#pragma unroll
	for (int ni = 0; ni < NI; ++ni) {
		#pragma unroll
		for (int e = 0; e < 2; ++e) {
			const int j = j0 + col0 + 8 * ni + 2 * kq + e;
			if (j >= J) continue;
			const double bj = bias_mode == 1 ? __ldg(bias + j) : 0.0;
			#pragma unroll
			for (int mi = 0; mi < 8; ++mi) {
				const long long m = m0 + row0 + 8 * mi + g;
				if (m >= M) continue;
				double v = acc[mi][ni][e] + bj;
				if (bias_mode == 2) v += __ldg(bias + m / gg.N + P * j);
				const long long o = m + M * j;
				if (out) out[o] = v;
				if (act_out) act_out[o] = act_fwd_rt<double>(act_kind, v, act_param);
			}
		}
	}
}

// Weight gradient: dw(tap, r, j) += sum_m src(m, tap, r) * plain[m + M*j]; rows of the output tile are the flattened
// (tap, r), the reduction runs over m in blocks of 8 (both operands are m-contiguous in HBM: 64-byte runs, transposed
// on the way into shared memory); split over m across the grid, per-split partials, deterministic reduce.
template<int NI>
__global__ void __launch_bounds__(DM_THREADS, NI == 2 ? 2 : 1) dmma_wgrad_kernel(GatherGeom gg, const double* __restrict__ src,
		const double* __restrict__ plain, double* __restrict__ partial, long long m_per_split, long long dw_elems) {
	constexpr int BN = 32 * NI;
	constexpr int PA = DM_BM + DM_PAD, PB = BN + DM_PAD;
	constexpr int B_ITERS = BN / 32;
	__shared__ __align__(16) double As[2][DM_BK * PA];
	__shared__ __align__(16) double Bs[2][DM_BK * PB];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int row0 = (warp & 1) * 64, col0 = (warp >> 1) * (8 * NI);
	const int R = gg.SC, J = gg.J;
	const int Ktot = gg.RH * gg.RW * R;
	const int k0 = blockIdx.x * DM_BM, j0 = blockIdx.y * BN;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long ms = (long long) blockIdx.z * m_per_split;
	const long long me = ms + m_per_split < M ? ms + m_per_split : M;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;

	// loader: 8 consecutive m per (tap, r) row / filter column; this thread owns m offset l_mm and rows l_r0 + 32 i
	const int l_mm = tid & 7, l_r0 = tid >> 3;
	int row_dh[4], row_dw[4], row_r[4];
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

	double acc[8][NI][2];
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

	auto fetch = [&](long long mc, int buf) {
		const long long m = mc + l_mm;
		const bool ok = m < me;
		const unsigned mu = (unsigned) m, pixu = mu / (unsigned) gg.N;   // M < 2^31
		const int n = (int) (mu - pixu * (unsigned) gg.N);
		const int ow = (int) (pixu / (unsigned) gg.OH), oh = (int) (pixu - (unsigned) ow * (unsigned) gg.OH);
		const int bh = oh * gg.ah, bw = ow * gg.aw;
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int th = bh + row_dh[i], tw = bw + row_dw[i];
			const bool aok = ok && row_r[i] >= 0 && th >= 0 && tw >= 0 && th < gg.SH && tw < gg.SW;   // denh = denw = 1
			cp_async_f64(&As[buf][l_mm * PA + l_r0 + 32 * i],
					aok ? src + n + (long long) gg.N * (th + (long long) gg.SH * tw) + row_r[i] * plane : src, aok);
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int j = j0 + l_r0 + 32 * i;
			const bool bok = ok && j < J;
			cp_async_f64(&Bs[buf][l_mm * PB + l_r0 + 32 * i], bok ? plain + m + M * j : plain, bok);
		}
		cp_async_commit_f64();
	};

	const long long steps = me > ms ? (me - ms + DM_BK - 1) / DM_BK : 0;
	if (steps > 0) {
		fetch(ms, 0);
		cp_async_wait_f64();
	}
	__syncthreads();
	for (long long st = 0; st < steps; ++st) {
		const int buf = (int) (st & 1);
		if (st + 1 < steps) fetch(ms + (st + 1) * DM_BK, buf ^ 1);
		dmma_block<NI, PA, PB>(As[buf], Bs[buf], row0, col0, lane, acc);
		cp_async_wait_f64();
		__syncthreads();
	}

	double* dst = partial + (long long) blockIdx.z * dw_elems;
	const int g = lane >> 2, kq = lane & 3;
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi) {
		const int k = k0 + row0 + 8 * mi + g;
		if (k >= Ktot) continue;
		const int tap = k / R, r = k - tap * R;
		const long long base = tap * gg.w_stap + r * gg.w_sr;
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) {
			#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int j = j0 + col0 + 8 * ni + 2 * kq + e;
				if (j < J) dst[base + j * gg.w_sj] = acc[mi][ni][e];
			}
		}
	}
}

__global__ void __launch_bounds__(256) dmma_wgrad_reduce_kernel(const double* __restrict__ partial, int splits, long long elems,
		double* __restrict__ dw) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < elems; i += (long long) gridDim.x * 256) {
		double s = 0;
		for (int z = 0; z < splits; ++z) s += partial[(long long) z * elems + i];
		dw[i] += s;
	}
}

} // namespace

bool dmma_gather_gemm_supported(const GatherGeom& gg) {
	return gg.J > 32 && gg.SC >= 6 && (long long) gg.RH * gg.RW * gg.SC >= 32;
}

int dmma_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* w, const double* bias,
		int bias_mode, double* out, const EpilogueArgs* ep) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	double* act_out = act ? (double*) ep->act_out : nullptr;
	const int act_kind = act ? ep->act_kind : CATTL3_ACT_NONE;
	const double act_param = act ? ep->act_param : 0.0;
	// 128 x 64 tiles whatever the filter count: 122 registers -> two CTAs (16 warps) per SM hide the fragment-load
	// latency; measured at config 2 (256 filters): 23.3 TFLOP/s against 18.5 with 128 x 128 tiles and one CTA per SM
	dim3 grid((unsigned) ceil_div(M, DM_BM), (unsigned) ceil_div(gg.J, 64));
	dmma_gather_gemm_kernel<2><<<grid, DM_THREADS, 0, ctx->stream>>>(gg, src, w, bias, bias_mode, out, act_kind, act_param,
			act_out);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

bool dmma_wgrad_supported(const GatherGeom& gg) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	return gg.J > 32 && (long long) gg.RH * gg.RW * gg.SC >= 64 && M >= 1024 && M < (1ll << 31) && gg.denh == 1 && gg.denw == 1;
}

int dmma_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* plain, double* dw) {
	const int Ktot = gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) Ktot * gg.J;
	// here the wide tile wins (19.3 against 17.0 TFLOP/s at 256 filters; a 512-thread 128 x 128 variant with 64 x 16 warp
	// tiles measured 17.2): half as many passes over the gathered rows
	const int BN = gg.J > 64 ? 128 : 64;
	const long long gx = ceil_div(Ktot, DM_BM), gy = ceil_div(gg.J, BN);
	long long splits = (long long) (BN == 128 ? 1 : 2) * ctx->sm_count / (gx * gy);
	if (splits < 1) splits = 1;
	const long long max_splits = ceil_div(M, 1024);
	if (splits > max_splits) splits = max_splits;
	long long m_per_split = ceil_div(ceil_div(M, splits), DM_BK) * DM_BK;
	splits = ceil_div(M, m_per_split);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (splits * elems) * sizeof(double)));
	dim3 grid((unsigned) gx, (unsigned) gy, (unsigned) splits);
	if (BN == 128)
		dmma_wgrad_kernel<4><<<grid, DM_THREADS, 0, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	else
		dmma_wgrad_kernel<2><<<grid, DM_THREADS, 0, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	CATTL3_LAUNCHED(ctx);
	dmma_wgrad_reduce_kernel<<<ew_grid(ctx, elems, 256), 256, 0, ctx->stream>>>((const double*) ctx->ws, (int) splits, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

} // namespace cattl3
