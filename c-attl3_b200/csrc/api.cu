// api.cu -- context / memory management and the kernel-layer entry points of the C ABI
// (include/cattl3_b200.h).  Each kernel-layer call is lowered to one or more "gather GEMM" passes
// (common.cuh: GatherGeom) that either the tcgen05 path (conv_tc.cu) or the SIMT path
// (conv_simt.cu) executes; there is no CPU fallback.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace cattl3 {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_error, sizeof(g_error), fmt, ap);
	va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
	set_error("CUDA error %d (%s) at %s:%d: %s", (int) e, cudaGetErrorString(e), file, line, what);
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
		return CATTL3_ERR_NO_DEVICE;
	return CATTL3_ERR_CUDA;
}

int ensure_buffer(cattl3_ctx* ctx, void** buf, size_t* cur, size_t need) {
	if (*cur >= need)
		return CATTL3_OK;
	if (ctx->capturing) {
		set_error("scratch buffer would have to grow during graph capture (run the step eagerly at this shape first)");
		return CATTL3_ERR_UNSUPPORTED;
	}
	if (*buf) {
		// the old buffer may still be in use by work queued on the stream
		CATTL3_CUDA(cudaStreamSynchronize(ctx->stream));
		CATTL3_CUDA(cudaFree(*buf));
		*buf = nullptr;
		*cur = 0;
		ctx->scratch_generation++;   // step graphs captured so far hold the old address: cattl3_graph_launch refuses them
	}
	size_t bytes = (need + ((size_t) 1 << 20) - 1) & ~(((size_t) 1 << 20) - 1);
	CATTL3_CUDA(cudaMalloc(buf, bytes));
	*cur = bytes;
	return CATTL3_OK;
}

static int conv_out(int in, int r, int p, int d, int s) { return (in - r - (r - 1) * d + 2 * p) / s + 1; }
static int tconv_out(int in, int r, int p, int d, int s) { return (in - 1) * s + r + (r - 1) * d - 2 * p; }

static int check_geom(const cattl3_conv_geom* g, int transposed, int* oh, int* ow) {
	CATTL3_REQUIRE(g, "null geometry");
	CATTL3_REQUIRE(g->n > 0 && g->h > 0 && g->w > 0 && g->c > 0 && g->f > 0, "conv geometry: non-positive size");
	CATTL3_REQUIRE(g->rh > 0 && g->rw > 0 && g->sh > 0 && g->sw > 0, "conv geometry: receptor/stride must be > 0");
	CATTL3_REQUIRE(g->ph >= 0 && g->pw >= 0 && g->dh >= 0 && g->dw >= 0, "conv geometry: negative padding/dilation");
	const int erh = g->rh + (g->rh - 1) * g->dh, erw = g->rw + (g->rw - 1) * g->dw;
	if (!transposed) {
		CATTL3_REQUIRE(g->h + 2 * g->ph >= erh && g->w + 2 * g->pw >= erw, "conv geometry: receptor larger than padded input");
		*oh = conv_out(g->h, g->rh, g->ph, g->dh, g->sh);
		*ow = conv_out(g->w, g->rw, g->pw, g->dw, g->sw);
	} else {
		*oh = tconv_out(g->h, g->rh, g->ph, g->dh, g->sh);
		*ow = tconv_out(g->w, g->rw, g->pw, g->dw, g->sw);
		CATTL3_REQUIRE(*oh > 0 && *ow > 0, "transposed conv geometry: empty output");
	}
	return CATTL3_OK;
}

// ---- lowering of the kernel layers to gather-GEMM passes ------------------------------------------
// Forward-style gather: source coordinate = o * stride + r * (dilation + 1) - pad.
static GatherGeom fwd_gather(int N, int SH, int SW, int SC, int OH, int OW, int J, const cattl3_conv_geom* g) {
	GatherGeom gg;
	gg.N = N; gg.SH = SH; gg.SW = SW; gg.SC = SC; gg.OH = OH; gg.OW = OW; gg.J = J;
	gg.RH = g->rh; gg.RW = g->rw;
	gg.ah = g->sh; gg.bh = g->dh + 1; gg.ch = -g->ph; gg.denh = 1;
	gg.aw = g->sw; gg.bw = g->dw + 1; gg.cw = -g->pw; gg.denw = 1;
	gg.w_stap = gg.w_sr = gg.w_sj = 0;
	return gg;
}
// Transposed-style gather: source coordinate = (o + pad - r * (dilation + 1)) / stride when divisible.
static GatherGeom bwd_gather(int N, int SH, int SW, int SC, int OH, int OW, int J, const cattl3_conv_geom* g) {
	GatherGeom gg;
	gg.N = N; gg.SH = SH; gg.SW = SW; gg.SC = SC; gg.OH = OH; gg.OW = OW; gg.J = J;
	gg.RH = g->rh; gg.RW = g->rw;
	gg.ah = 1; gg.bh = -(g->dh + 1); gg.ch = g->ph; gg.denh = g->sh;
	gg.aw = 1; gg.bw = -(g->dw + 1); gg.cw = g->pw; gg.denw = g->sw;
	gg.w_stap = gg.w_sr = gg.w_sj = 0;
	return gg;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template<typename S> struct IsFloat { static constexpr bool value = false; };
template<> struct IsFloat<float> { static constexpr bool value = true; };

// One dimension of a strided transposed gather (source = (o + c - r*d) / s where divisible), split by the residue
// class a = (o + c) mod s: outputs o = o0 + s*i take only the taps r = r0 + q*k, and for those
// source = i - (d/g)*k + c_sub -- an ordinary stride-1 gather.  (g = gcd(d, s), q = s/g.)
struct LatticeDim { int o0, n_o, r0, q, n_r, b_sub, c_sub; };
static LatticeDim lattice_dim(int a, int O, int R, int c, int d, int s) {
	LatticeDim L;
	L.o0 = ((a - c) % s + s) % s;
	L.n_o = L.o0 < O ? (O - L.o0 + s - 1) / s : 0;
	L.r0 = -1; L.q = 1; L.n_r = 0;
	for (int r = 0; r < R; ++r)
		if ((r * d) % s == a) {
			if (L.n_r == 0) L.r0 = r;
			else if (L.n_r == 1) L.q = r - L.r0;
			++L.n_r;
		}
	int g = d, t = s;
	while (t) { int u = g % t; g = t; t = u; }   // gcd(d, s)
	L.b_sub = -(d / g);
	L.c_sub = L.n_r ? (L.o0 + c - L.r0 * d) / s : 0;  // exact: the numerator is a multiple of s
	return L;
}

template<typename S>
static int run_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias,
		int bias_mode, S* out, const EpilogueArgs* ep = nullptr) {
	// ep: fused activation and / or column statistics.  The tcgen05 kernel does both in its epilogue; the SIMT kernel
	// fuses the activation, and the statistics are then one reduction pass over the finished output.
	// (a tensor at an address that is not 16-byte aligned -- a slice of a larger buffer -- takes the FMA / SIMT kernels
	// instead of failing in the TMA descriptor)
	const bool tensor_ok = ctx->conv_path != CATTL3_PATH_SIMT && ctx->conv_path != CATTL3_PATH_FMA && aligned16(src) &&
			(!out || aligned16(out));
	if (IsFloat<S>::value && tensor_ok && tc_gather_gemm_supported(ctx, gg)) {
		ctx->last_path = "tcgen05";
		if (ep && ep->col_stats && gg.J >= 256 && out && bias_mode == 1) {
			// Wide layers: the statistics epilogue needs 128-wide tiles (two accumulators to hide behind), which converts the
			// big operand twice -- measured slower than the plain 256-wide kernel followed by one reduction pass over the
			// finished output (profiles/README.md, r2: 1.73 against 1.67 ms layer by layer).  The activation stays fused.
			EpilogueArgs plain = *ep;
			plain.col_stats = nullptr;
			CATTL3_CHECK(tc_gather_gemm_f32(ctx, gg, (const float*) src, (const float*) w, (const float*) bias, bias_mode,
					(float*) out, plain.act_kind != CATTL3_ACT_NONE ? &plain : nullptr));
			return colstats_shifted<S>(ctx, (int64_t) gg.N * gg.OH * gg.OW, gg.J, out, bias, ep->col_stats);
		}
		return tc_gather_gemm_f32(ctx, gg, (const float*) src, (const float*) w, (const float*) bias, bias_mode,
				(float*) out, ep);
	}
	CATTL3_REQUIRE(!(ep && ep->col_stats) || (bias_mode == 1 && out), "column statistics need a per-column bias and y");
	if (IsFloat<S>::value && tensor_ok && (gg.denh > 1 || gg.denw > 1) && gg.ah == 1 && gg.aw == 1) {
		// strided transposed gather (input gradient of a strided convolution, forward of a strided transposed
		// convolution): denh * denw stride-1 sub-problems, one per residue class of the output pixel, each writing
		// its own sub-lattice of the output through the tensor-core kernel
		GatherGeom probe = gg;
		probe.denh = probe.denw = 1;
		if (tc_gather_gemm_supported(ctx, probe)) {
			bool need_zero = false;
			for (int a = 0; a < gg.denh && !need_zero; ++a)
				for (int b = 0; b < gg.denw && !need_zero; ++b) {
					const LatticeDim Lh = lattice_dim(a, gg.OH, gg.RH, gg.ch, -gg.bh, gg.denh);
					const LatticeDim Lw = lattice_dim(b, gg.OW, gg.RW, gg.cw, -gg.bw, gg.denw);
					if (Lh.n_o > 0 && Lw.n_o > 0 && (Lh.n_r == 0 || Lw.n_r == 0)) need_zero = bias_mode == 0;
					if (Lh.n_o > 0 && Lw.n_o > 0 && (Lh.n_r == 0 || Lw.n_r == 0) && bias_mode != 0) goto simt;  // bias-only pixels
				}
			if (need_zero)
				CATTL3_CUDA(cudaMemsetAsync(out, 0, sizeof(S) * (size_t) gg.N * gg.OH * gg.OW * gg.J, ctx->stream));
			const long long srw = gg.w_srw ? gg.w_srw : gg.w_stap * gg.RH;
			for (int a = 0; a < gg.denh; ++a)
				for (int b = 0; b < gg.denw; ++b) {
					const LatticeDim Lh = lattice_dim(a, gg.OH, gg.RH, gg.ch, -gg.bh, gg.denh);
					const LatticeDim Lw = lattice_dim(b, gg.OW, gg.RW, gg.cw, -gg.bw, gg.denw);
					if (Lh.n_o == 0 || Lw.n_o == 0 || Lh.n_r == 0 || Lw.n_r == 0) continue;
					GatherGeom sub = gg;
					sub.OH = Lh.n_o; sub.OW = Lw.n_o; sub.RH = Lh.n_r; sub.RW = Lw.n_r;
					sub.ah = 1; sub.bh = Lh.b_sub; sub.ch = Lh.c_sub; sub.denh = 1;
					sub.aw = 1; sub.bw = Lw.b_sub; sub.cw = Lw.c_sub; sub.denw = 1;
					sub.w_off = gg.w_off + Lh.r0 * gg.w_stap + Lw.r0 * srw;
					sub.w_stap = gg.w_stap * Lh.q; sub.w_srw = srw * Lw.q;
					sub.out_h0 = Lh.o0; sub.out_hs = gg.denh; sub.out_H = gg.OH;
					sub.out_w0 = Lw.o0; sub.out_ws = gg.denw; sub.out_W = gg.OW;
					CATTL3_CHECK(tc_gather_gemm_f32(ctx, sub, (const float*) src, (const float*) w, (const float*) bias,
							bias_mode, (float*) out, ep));  // every output pixel belongs to exactly one sub-lattice
				}
			ctx->last_path = "tcgen05";
			return CATTL3_OK;
		}
	}
simt:
	if (ctx->conv_path == CATTL3_PATH_TCGEN05) {
		set_error("tcgen05 path requested but the shape/type does not qualify (needs float, batch %% 32 == 0)");
		return CATTL3_ERR_UNSUPPORTED;
	}
	if (ctx->conv_path != CATTL3_PATH_SIMT && skinny_gather_gemm_supported(ctx, gg)) {
		// a classifier head: small batch, long reduction, a few outputs -- split over the reduction
		ctx->last_path = "skinny";
		CATTL3_CHECK(skinny_gather_gemm<S>(ctx, gg, src, w, bias, bias_mode, out, ep));
	} else if (ctx->conv_path != CATTL3_PATH_SIMT && tiny_gather_gemm_supported(gg, sizeof(S))) {
		// a handful of output columns: the streaming kernel (conv_simt.cu)
		ctx->last_path = "tiny";
		CATTL3_CHECK(tiny_gather_gemm<S>(ctx, gg, src, w, bias, bias_mode, out, ep));
	} else if (!IsFloat<S>::value && tensor_ok && dmma_gather_gemm_supported(gg)) {
		// double at GEMM-sized shapes: FP64 tensor cores (conv_dmma.cu)
		ctx->last_path = "dmma";
		CATTL3_CHECK(dmma_gather_gemm(ctx, gg, (const double*) src, (const double*) w, (const double*) bias, bias_mode,
				(double*) out, ep));
	} else if (ctx->conv_path != CATTL3_PATH_SIMT && fma_gather_gemm_supported<S>(gg)) {
		// GEMM-sized shapes off the tensor-core path (double; float with few channels or a ragged batch): the
		// big-tile FMA kernels (conv_dfma.cu)
		ctx->last_path = IsFloat<S>::value ? "ffma" : "dfma";
		CATTL3_CHECK(fma_gather_gemm<S>(ctx, gg, src, w, bias, bias_mode, out, ep));
	} else {
		ctx->last_path = "simt";
		CATTL3_CHECK(simt_gather_gemm<S>(ctx, gg, src, w, bias, bias_mode, out, ep));
	}
	if (ep && ep->col_stats)
		CATTL3_CHECK(colstats_shifted<S>(ctx, (int64_t) gg.N * gg.OH * gg.OW, gg.J, out, bias, ep->col_stats));
	return CATTL3_OK;
}

// db_colsum != null: the caller also wants db += column sums of `plain`; the tcgen05 kernel produces them from
// the same read of `plain` (*db_done = true), otherwise the caller runs colsum_accumulate.
template<typename S>
static int run_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw, S* db_colsum = nullptr,
		bool* db_done = nullptr) {
	if (db_done) *db_done = false;
	const bool tensor_ok = ctx->conv_path != CATTL3_PATH_SIMT && ctx->conv_path != CATTL3_PATH_FMA && aligned16(src) && aligned16(plain);
	if (IsFloat<S>::value && tensor_ok && tc_wgrad_supported(ctx, gg)) {
		ctx->last_path = "tcgen05";
		if (db_done) *db_done = db_colsum != nullptr;
		return tc_wgrad_f32(ctx, gg, (const float*) src, (const float*) plain, (float*) dw, (float*) db_colsum);
	}
	if (ctx->conv_path == CATTL3_PATH_TCGEN05) {
		set_error("tcgen05 weight-gradient path requested but the shape/type does not qualify");
		return CATTL3_ERR_UNSUPPORTED;
	}
	if (ctx->conv_path != CATTL3_PATH_SIMT && tiny_wgrad_supported(gg, sizeof(S))) {
		ctx->last_path = "tiny";
		return tiny_wgrad<S>(ctx, gg, src, plain, dw);
	}
	if (!IsFloat<S>::value && tensor_ok && dmma_wgrad_supported(gg)) {
		ctx->last_path = "dmma";
		return dmma_wgrad(ctx, gg, (const double*) src, (const double*) plain, (double*) dw);
	}
	if (ctx->conv_path != CATTL3_PATH_SIMT && fma_wgrad_supported<S>(gg)) {
		ctx->last_path = IsFloat<S>::value ? "ffma" : "dfma";
		return fma_wgrad<S>(ctx, gg, src, plain, dw);
	}
	ctx->last_path = "simt";
	return simt_wgrad<S>(ctx, gg, src, plain, dw);
}

static int check_epilogue(const cattl3_epilogue* ep, const void* y, EpilogueArgs* ea) {
	if (!ep) {
		CATTL3_REQUIRE(y, "forward: null output tensor");
		return CATTL3_OK;
	}
	CATTL3_REQUIRE(ep->act_kind >= CATTL3_ACT_NONE && ep->act_kind < CATTL3_ACT_SOFTMAX,
			"epilogue: activation kind %d cannot be fused (element-wise kinds 0-6 only)", ep->act_kind);
	CATTL3_REQUIRE(ep->act_kind == CATTL3_ACT_NONE || ep->act_out, "epilogue: activation without act_out");
	CATTL3_REQUIRE(y || ep->act_kind != CATTL3_ACT_NONE, "forward: neither y nor act_out given");
	CATTL3_REQUIRE(!ep->col_stats || y, "epilogue: column statistics need y (the batch-norm input)");
	ea->act_kind = ep->act_kind; ea->act_param = ep->act_param; ea->act_out = ep->act_out; ea->col_stats = ep->col_stats;
	return CATTL3_OK;
}

template<typename S>
static int conv_forward(cattl3_ctx* ctx, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y,
		const cattl3_epilogue* ep = nullptr) {
	CATTL3_CHECK(check_ctx(ctx));
	int oh, ow;
	CATTL3_CHECK(check_geom(g, 0, &oh, &ow));
	CATTL3_REQUIRE(x && w && b, "conv_forward: null tensor");
	EpilogueArgs ea;
	CATTL3_CHECK(check_epilogue(ep, y, &ea));
	const long long T = (long long) g->rh * g->rw;
	GatherGeom gg = fwd_gather(g->n, g->h, g->w, g->c, oh, ow, g->f, g);
	gg.w_stap = 1; gg.w_sr = T; gg.w_sj = T * g->c;
	return run_gather_gemm<S>(ctx, gg, x, w, b, 1, y, ep ? &ea : nullptr);
}

template<typename S>
static int conv_backward(cattl3_ctx* ctx, const cattl3_conv_geom* g, const S* x, const S* w, const S* dy, S* dw,
		S* db, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	int oh, ow;
	CATTL3_CHECK(check_geom(g, 0, &oh, &ow));
	// dw == db == NULL: the input gradient alone (an unrolled recurrent network takes its shared kernels' weight gradients
	// once over all time steps instead of once per step)
	CATTL3_REQUIRE(w && dy && ((x && dw && db) || (!dw && !db && dx)), "conv_backward: null tensor");
	const long long T = (long long) g->rh * g->rw;
	if (dw) {
		// dW += cols^T dY (ConvKernelLayer.hpp:154)
		GatherGeom gw = fwd_gather(g->n, g->h, g->w, g->c, oh, ow, g->f, g);
		gw.w_stap = 1; gw.w_sr = T; gw.w_sj = T * g->c;
		bool db_done = false;
		CATTL3_CHECK(run_wgrad<S>(ctx, gw, x, dy, dw, db, &db_done));
		// db += colsum(dY) (:155), unless the weight-gradient kernel already produced it from its own read of dY
		if (!db_done)
			CATTL3_CHECK(colsum_accumulate<S>(ctx, (int64_t) g->n * oh * ow, g->f, dy, db));
	}
	if (!dx)
		return CATTL3_OK;  // input layer (:156-157)
	// dX = crop(col2im(dY W^T)) (:159-188) as a gather over (tap, f)
	GatherGeom gd = bwd_gather(g->n, oh, ow, g->f, g->h, g->w, g->c, g);
	gd.w_stap = 1; gd.w_sr = T * g->c; gd.w_sj = T;
	return run_gather_gemm<S>(ctx, gd, dy, w, (const S*) nullptr, 0, dx);
}

template<typename S>
static int transconv_forward(cattl3_ctx* ctx, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y,
		const cattl3_epilogue* ep = nullptr) {
	CATTL3_CHECK(check_ctx(ctx));
	int oh, ow;
	CATTL3_CHECK(check_geom(g, 1, &oh, &ow));
	CATTL3_REQUIRE(x && w && b, "transconv_forward: null tensor");
	CATTL3_REQUIRE(!ep || !ep->col_stats, "transconv_forward: column statistics need a per-filter bias (convolution / dense only)");
	EpilogueArgs ea;
	CATTL3_CHECK(check_epilogue(ep, y, &ea));
	const long long T = (long long) g->rh * g->rw;
	GatherGeom gg = bwd_gather(g->n, g->h, g->w, g->c, oh, ow, g->f, g);
	gg.w_stap = g->c; gg.w_sr = 1; gg.w_sj = (long long) g->c * T;
	return run_gather_gemm<S>(ctx, gg, x, w, b, 2, y, ep ? &ea : nullptr);
}

template<typename S>
static int transconv_backward(cattl3_ctx* ctx, const cattl3_conv_geom* g, const S* x, const S* w, const S* dy,
		S* dw, S* db, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	int oh, ow;
	CATTL3_CHECK(check_geom(g, 1, &oh, &ow));
	CATTL3_REQUIRE(x && w && dy && dw && db, "transconv_backward: null tensor");
	const long long T = (long long) g->rh * g->rw;
	// db += sum_n dY, one bias per output element (TransConvKernelLayer.hpp:160-161)
	CATTL3_CHECK(colsum_accumulate<S>(ctx, g->n, (int64_t) oh * ow * g->f, dy, db));
	// G = im2col(pad(dY)); dW += x^T G (:165-182); dX = G W^T (:185-187): an ordinary conv of dY
	GatherGeom gg = fwd_gather(g->n, oh, ow, g->f, g->h, g->w, g->c, g);
	gg.w_stap = g->c; gg.w_sr = (long long) g->c * T; gg.w_sj = 1;
	CATTL3_CHECK(run_wgrad<S>(ctx, gg, dy, x, dw));
	if (!dx)
		return CATTL3_OK;
	return run_gather_gemm<S>(ctx, gg, dy, w, (const S*) nullptr, 0, dx);
}

static cattl3_conv_geom dense_geom(int n, int in, int out) {
	cattl3_conv_geom g;
	g.n = n; g.h = 1; g.w = 1; g.c = in; g.f = out;
	g.rh = g.rw = 1; g.ph = g.pw = 0; g.sh = g.sw = 1; g.dh = g.dw = 0;
	return g;
}

} // namespace cattl3

using namespace cattl3;

extern "C" {

int cattl3_abi_version(void) { return CATTL3_ABI_VERSION; }
const char* cattl3_last_error(void) { return g_error; }

int cattl3_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int cattl3_ctx_create(cattl3_ctx** out, int device, void* cuda_stream) {
	CATTL3_REQUIRE(out, "ctx_create: null out pointer");
	*out = nullptr;
	int n = cattl3_device_count();
	if (n <= 0) {
		set_error("no CUDA device available (libcattl3_b200 has no CPU fallback)");
		return CATTL3_ERR_NO_DEVICE;
	}
	CATTL3_REQUIRE(device >= 0 && device < n, "ctx_create: device %d out of range (%d devices)", device, n);
	CATTL3_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	CATTL3_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) {
		set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
		return CATTL3_ERR_UNSUPPORTED;
	}
	{
		// keep freed blocks cached in the stream-ordered pool instead of returning them to the driver
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			uint64_t keep = UINT64_MAX;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		}
	}
	cattl3_ctx* ctx = new cattl3_ctx();
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	if (cuda_stream) {
		ctx->stream = (cudaStream_t) cuda_stream;
	} else {
		cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
		if (e != cudaSuccess) {
			delete ctx;
			return cuda_fail(e, "cudaStreamCreateWithFlags", __FILE__, __LINE__);
		}
		ctx->own_stream = true;
	}
	*out = ctx;
	return CATTL3_OK;
}

int cattl3_ctx_destroy(cattl3_ctx* ctx) {
	if (!ctx)
		return CATTL3_OK;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->ws) cudaFree(ctx->ws);
	if (ctx->tc_w) cudaFree(ctx->tc_w);
	for (int i = 0; i < cattl3_ctx::MAX_PACK_SLOTS; ++i)
		if (ctx->pack_slots[i].buf) cudaFree(ctx->pack_slots[i].buf);
	if (ctx->stat_ws) cudaFree(ctx->stat_ws);
	if (ctx->reg_ws) cudaFree(ctx->reg_ws);
	if (ctx->con_ws) cudaFree(ctx->con_ws);
	for (int a = 0; a < ctx->arena_count; ++a)
		cudaFree(ctx->arenas[a].base);
	for (int i = 0; i < 8; ++i)
		if (ctx->throttle_ev[i]) cudaEventDestroy(ctx->throttle_ev[i]);
	for (int i = 0; i < 5; ++i)
		if (ctx->stage_dev[i]) cudaFree(ctx->stage_dev[i]);
	if (ctx->up_stream) {
		cudaStreamSynchronize(ctx->up_stream);
		cudaStreamSynchronize(ctx->down_stream);
		for (int i = 0; i < cattl3_ctx::HOST_EVENTS; ++i)
			if (ctx->host_ev[i]) cudaEventDestroy(ctx->host_ev[i]);
		cudaStreamDestroy(ctx->up_stream);
		cudaStreamDestroy(ctx->down_stream);
	}
	if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
	return CATTL3_OK;
}

} // extern "C"

// ---- step graphs ------------------------------------------------------------------------------------------------
struct cattl3_graph {
	cattl3_ctx* ctx = nullptr;
	cudaGraph_t graph = nullptr;
	cudaGraphExec_t exec = nullptr;
	int arena = -1;
	int64_t scratch_generation = 0;   // the context's when the graph was captured
};

namespace {

inline size_t granule(size_t bytes) { return ((bytes ? bytes : 1) + 255) & ~(size_t) 255; }

// index of the arena that holds p, or -1
inline int arena_of(const cattl3_ctx* ctx, const void* p) {
	for (int a = 0; a < ctx->arena_count; ++a) {
		if ((const char*) p >= ctx->arenas[a].base && (const char*) p < ctx->arenas[a].base + ctx->arenas[a].size)
			return a;
	}
	return -1;
}

}

extern "C" {

int64_t cattl3_ctx_allocated_bytes(cattl3_ctx* ctx) {
	return ctx ? ctx->allocated_bytes : 0;
}

int cattl3_graph_begin(cattl3_ctx* ctx, size_t arena_bytes) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(!ctx->capturing, "graph_begin: already capturing");
	CATTL3_REQUIRE(arena_bytes > 0, "graph_begin: the arena needs a size (cattl3_ctx_allocated_bytes over one eager step)");
	arena_bytes = granule(arena_bytes);
	// smallest retired arena that is large enough, else a new one
	int pick = -1;
	for (int a = 0; a < ctx->arena_count; ++a) {
		if (!ctx->arenas[a].in_use && ctx->arenas[a].size >= arena_bytes &&
				(pick < 0 || ctx->arenas[a].size < ctx->arenas[pick].size))
			pick = a;
	}
	if (pick < 0) {
		if (ctx->arena_count == cattl3_ctx::MAX_ARENAS) {
			set_error("graph_begin: all %d step-graph arenas are taken", cattl3_ctx::MAX_ARENAS);
			return CATTL3_ERR_UNSUPPORTED;
		}
		void* base = nullptr;
		CATTL3_CUDA(cudaMalloc(&base, arena_bytes));
		pick = ctx->arena_count++;
		ctx->arenas[pick].base = (char*) base;
		ctx->arenas[pick].size = arena_bytes;
	}
	CATTL3_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
	ctx->arenas[pick].in_use = true;
	ctx->cap_arena = pick;
	ctx->cap_used = 0;
	ctx->cap_block_count = 0;
	ctx->capturing = true;
	return CATTL3_OK;
}

int cattl3_graph_end(cattl3_ctx* ctx, cattl3_graph** out) {
	CATTL3_REQUIRE(ctx && out, "graph_end: bad arguments");
	*out = nullptr;
	CATTL3_REQUIRE(ctx->capturing, "graph_end: not capturing");
	const int arena = ctx->cap_arena;
	ctx->capturing = false;
	ctx->cap_arena = -1;
	ctx->cap_block_count = 0;
	cudaGraph_t graph = nullptr;
	cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
	// what the host released during the capture goes back to the pool now, in stream order behind everything queued
	for (int i = 0; i < ctx->deferred_free_count; ++i) {
		if (cudaFreeAsync(ctx->deferred_free[i], ctx->stream) != cudaSuccess)
			cudaGetLastError();
	}
	ctx->deferred_free_count = 0;
	if (e != cudaSuccess || !graph) {
		cudaGetLastError();
		if (graph) cudaGraphDestroy(graph);
		ctx->arenas[arena].in_use = false;
		return cuda_fail(e != cudaSuccess ? e : cudaErrorUnknown, "cudaStreamEndCapture (the capture was invalidated)", __FILE__, __LINE__);
	}
	cudaGraphExec_t exec = nullptr;
	e = cudaGraphInstantiate(&exec, graph, 0);
	if (e != cudaSuccess) {
		cudaGetLastError();
		cudaGraphDestroy(graph);
		ctx->arenas[arena].in_use = false;
		return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
	}
	cattl3_graph* g = new cattl3_graph();
	g->ctx = ctx; g->graph = graph; g->exec = exec; g->arena = arena;
	g->scratch_generation = ctx->scratch_generation;
	*out = g;
	return CATTL3_OK;
}

int cattl3_graph_launch(cattl3_ctx* ctx, cattl3_graph* g) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(g && g->exec && g->ctx == ctx && !ctx->capturing, "graph_launch: bad arguments");
	if (g->scratch_generation != ctx->scratch_generation) {
		set_error("graph_launch: library scratch moved since the capture (a larger problem ran in between); capture again");
		return CATTL3_ERR_UNSUPPORTED;
	}
	CATTL3_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
	ctx->launches++;
	return CATTL3_OK;
}

int cattl3_graph_destroy(cattl3_graph* g) {
	if (!g) return CATTL3_OK;
	cudaSetDevice(g->ctx->device);
	cudaStreamSynchronize(g->ctx->stream);
	if (g->exec) cudaGraphExecDestroy(g->exec);
	if (g->graph) cudaGraphDestroy(g->graph);
	if (g->arena >= 0)
		g->ctx->arenas[g->arena].in_use = false;   // retired: the next graph_begin may take it
	delete g;
	return CATTL3_OK;
}

int cattl3_ctx_throttle(cattl3_ctx* ctx, int max_in_flight) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(max_in_flight >= 1 && max_in_flight <= 7, "ctx_throttle: 1..7 steps in flight");
	const int ring = 8;
	const long long p = ctx->throttle_calls++;
	cudaEvent_t& ev = ctx->throttle_ev[p % ring];
	if (!ev) CATTL3_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventBlockingSync));
	CATTL3_CUDA(cudaEventRecord(ev, ctx->stream));
	if (p >= max_in_flight) {
		cudaEvent_t old = ctx->throttle_ev[(p - max_in_flight) % ring];
		if (old) CATTL3_CUDA(cudaEventSynchronize(old));
	}
	return CATTL3_OK;
}

int cattl3_ctx_synchronize(cattl3_ctx* ctx) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_CUDA(cudaStreamSynchronize(ctx->stream));
	return CATTL3_OK;
}

int cattl3_weights_stable_begin(cattl3_ctx* ctx) {
	CATTL3_CHECK(check_ctx(ctx));
	ctx->pack_stable = true;
	ctx->pack_used = 0;
	return CATTL3_OK;
}
int cattl3_weights_stable_end(cattl3_ctx* ctx) {
	CATTL3_CHECK(check_ctx(ctx));
	ctx->pack_stable = false;
	ctx->pack_used = 0;
	return CATTL3_OK;
}

int cattl3_ctx_set_conv_path(cattl3_ctx* ctx, int path) {
	CATTL3_REQUIRE(ctx && path >= CATTL3_PATH_AUTO && path <= CATTL3_PATH_FMA, "set_conv_path: bad arguments");
	ctx->conv_path = path;
	return CATTL3_OK;
}

int64_t cattl3_ctx_launch_count(const cattl3_ctx* ctx) { return ctx ? ctx->launches : 0; }
const char* cattl3_ctx_last_path(const cattl3_ctx* ctx) { return ctx ? ctx->last_path : "none"; }
void* cattl3_ctx_stream(const cattl3_ctx* ctx) { return ctx ? (void*) ctx->stream : nullptr; }

int cattl3_malloc(cattl3_ctx* ctx, void** p, size_t bytes) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(p, "malloc: null out pointer");
	const size_t size = granule(bytes);
	ctx->allocated_bytes += (int64_t) size;
	if (ctx->capturing) {
		// a step graph's activations come out of its arena (fixed addresses, no allocation nodes in the graph)
		cattl3_ctx::Arena& arena = ctx->arenas[ctx->cap_arena];
		for (int i = 0; i < ctx->cap_block_count; ++i) {
			cattl3_ctx::CapBlock& blk = ctx->cap_blocks[i];
			if (blk.free && blk.size == size) {   // released earlier in the step: stream order makes the reuse safe
				blk.free = false;
				*p = blk.ptr;
				return CATTL3_OK;
			}
		}
		if (ctx->cap_used + size > arena.size || ctx->cap_block_count == cattl3_ctx::MAX_CAP_BLOCKS) {
			set_error("step-graph arena exhausted (%zu of %zu bytes, %d blocks)", ctx->cap_used, arena.size, ctx->cap_block_count);
			return CATTL3_ERR_UNSUPPORTED;
		}
		cattl3_ctx::CapBlock& blk = ctx->cap_blocks[ctx->cap_block_count++];
		blk.ptr = arena.base + ctx->cap_used;
		blk.size = size;
		blk.free = false;
		ctx->cap_used += size;
		*p = blk.ptr;
		return CATTL3_OK;
	}
	// stream-ordered allocation from the device's default pool (kept warm: see ctx_create), so a
	// layer's activation buffers are recycled without a cudaMalloc / cudaFree synchronisation per call
	CATTL3_CUDA(cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream));
	return CATTL3_OK;
}
int cattl3_free(cattl3_ctx* ctx, void* p) {
	CATTL3_CHECK(check_ctx(ctx));
	if (!p)
		return CATTL3_OK;
	const int a = arena_of(ctx, p);
	if (a >= 0) {
		// arena memory belongs to a step graph: recycled within the capture that handed it out, otherwise nothing to do
		if (ctx->capturing && a == ctx->cap_arena) {
			for (int i = 0; i < ctx->cap_block_count; ++i) {
				if (ctx->cap_blocks[i].ptr == (char*) p) {
					ctx->cap_blocks[i].free = true;
					break;
				}
			}
		}
		return CATTL3_OK;
	}
	if (ctx->capturing) {
		// memory from before the capture: released once the capture has ended (see cattl3_graph_end)
		if (ctx->deferred_free_count == cattl3_ctx::MAX_DEFERRED_FREES) {
			set_error("free: too many releases of pool memory during one graph capture");
			return CATTL3_ERR_UNSUPPORTED;   // the block stays allocated; the caller's capture is abandoned by the host side
		}
		ctx->deferred_free[ctx->deferred_free_count++] = p;
		return CATTL3_OK;
	}
	cudaError_t e = cudaFreeAsync(p, ctx->stream);
	if (e != cudaSuccess) {
		cudaGetLastError();   // callers are destructors: do not leave the error for the next launch check
		return cuda_fail(e, "cudaFreeAsync", __FILE__, __LINE__);
	}
	return CATTL3_OK;
}
int cattl3_memset(cattl3_ctx* ctx, void* p, int value, size_t bytes) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_CUDA(cudaMemsetAsync(p, value, bytes, ctx->stream));
	return CATTL3_OK;
}
int cattl3_memcpy_h2d(cattl3_ctx* ctx, void* dst, const void* src, size_t bytes) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	return CATTL3_OK;
}
} // extern "C"

// ---- input feed ---------------------------------------------------------------------------------------------
struct cattl3_feed {
	cattl3_ctx* ctx = nullptr;
	int slots = 0;
	long long pushes = 0;
	void** dev = nullptr;          // [slots] device buffers, grown on demand
	size_t* dev_bytes = nullptr;
	cudaEvent_t* marks = nullptr;  // [slots] marks[p % slots] = "everything on the compute stream at push p"
	cudaStream_t copy = nullptr;
	void* pinned[2] = { nullptr, nullptr };
	cudaEvent_t pinned_done[2] = { nullptr, nullptr };
	cudaEvent_t ready = nullptr;
};
static constexpr size_t FEED_CHUNK = (size_t) 8 << 20;

extern "C" {

int cattl3_feed_create(cattl3_feed** out, cattl3_ctx* ctx, int slots) {
	CATTL3_REQUIRE(out, "feed_create: null out pointer");
	*out = nullptr;
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(slots >= 2 && slots <= 16, "feed_create: 2..16 slots");
	cattl3_feed* f = new cattl3_feed();
	f->ctx = ctx; f->slots = slots;
	f->dev = new void*[slots](); f->dev_bytes = new size_t[slots](); f->marks = new cudaEvent_t[slots]();
	cudaError_t e = cudaStreamCreateWithFlags(&f->copy, cudaStreamNonBlocking);
	for (int i = 0; i < slots && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&f->marks[i], cudaEventDisableTiming);
	for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
		e = cudaMallocHost(&f->pinned[i], FEED_CHUNK);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->pinned_done[i], cudaEventDisableTiming);
	}
	if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ready, cudaEventDisableTiming);
	if (e != cudaSuccess) {
		cattl3_feed_destroy(f);
		return cuda_fail(e, "feed_create", __FILE__, __LINE__);
	}
	*out = f;
	return CATTL3_OK;
}

int cattl3_feed_destroy(cattl3_feed* f) {
	if (!f) return CATTL3_OK;
	cudaSetDevice(f->ctx->device);
	if (f->copy) cudaStreamSynchronize(f->copy);
	cudaStreamSynchronize(f->ctx->stream);
	for (int i = 0; i < f->slots; ++i) {
		if (f->dev && f->dev[i]) cudaFree(f->dev[i]);
		if (f->marks && f->marks[i]) cudaEventDestroy(f->marks[i]);
	}
	for (int i = 0; i < 2; ++i) {
		if (f->pinned[i]) cudaFreeHost(f->pinned[i]);
		if (f->pinned_done[i]) cudaEventDestroy(f->pinned_done[i]);
	}
	if (f->ready) cudaEventDestroy(f->ready);
	if (f->copy) cudaStreamDestroy(f->copy);
	delete[] f->dev; delete[] f->dev_bytes; delete[] f->marks;
	delete f;
	return CATTL3_OK;
}

int cattl3_feed_push(cattl3_feed* f, const void* src, size_t bytes, void** dev_ptr) {
	CATTL3_REQUIRE(f && src && bytes > 0 && dev_ptr, "feed_push: bad arguments");
	cattl3_ctx* ctx = f->ctx;
	CATTL3_CHECK(check_ctx(ctx));
	const int s = (int) (f->pushes % f->slots);
	if (f->dev_bytes[s] < bytes) {
		// growing a slot (first batches, or a larger batch): its old contents may still be in use
		CATTL3_CUDA(cudaStreamSynchronize(ctx->stream));
		CATTL3_CUDA(cudaStreamSynchronize(f->copy));
		if (f->dev[s]) CATTL3_CUDA(cudaFree(f->dev[s]));
		f->dev[s] = nullptr; f->dev_bytes[s] = 0;
		const size_t rounded = (bytes + ((size_t) 1 << 20) - 1) & ~(((size_t) 1 << 20) - 1);
		CATTL3_CUDA(cudaMalloc(&f->dev[s], rounded));
		f->dev_bytes[s] = rounded;
	}
	// the previous contents of this slot were consumed by work enqueued before the push `slots - 1` pushes ago
	if (f->pushes >= f->slots - 1)
		CATTL3_CUDA(cudaStreamWaitEvent(f->copy, f->marks[(f->pushes + 1) % f->slots], 0));
	size_t off = 0;
	for (int chunk = 0; off < bytes; ++chunk) {
		const int b = chunk & 1;
		const size_t n = bytes - off < FEED_CHUNK ? bytes - off : FEED_CHUNK;
		CATTL3_CUDA(cudaEventSynchronize(f->pinned_done[b]));   // the copy engine has drained this staging buffer
		memcpy(f->pinned[b], (const char*) src + off, n);
		CATTL3_CUDA(cudaMemcpyAsync((char*) f->dev[s] + off, f->pinned[b], n, cudaMemcpyHostToDevice, f->copy));
		CATTL3_CUDA(cudaEventRecord(f->pinned_done[b], f->copy));
		off += n;
	}
	CATTL3_CUDA(cudaEventRecord(f->ready, f->copy));
	CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, f->ready, 0));
	CATTL3_CUDA(cudaEventRecord(f->marks[s], ctx->stream));
	*dev_ptr = f->dev[s];
	++f->pushes;
	return CATTL3_OK;
}

int cattl3_memcpy_d2h(cattl3_ctx* ctx, void* dst, const void* src, size_t bytes) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(!ctx->capturing, "memcpy_d2h synchronises: not possible during graph capture");
	CATTL3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	CATTL3_CUDA(cudaStreamSynchronize(ctx->stream));
	return CATTL3_OK;
}
int cattl3_memcpy_d2d(cattl3_ctx* ctx, void* dst, const void* src, size_t bytes) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
	return CATTL3_OK;
}
int cattl3_memcpy_2d(cattl3_ctx* ctx, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width, size_t height) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(dst && src && width > 0 && height > 0 && dst_pitch >= width && src_pitch >= width, "memcpy_2d: bad arguments");
	if (height == 1 || (dst_pitch == width && src_pitch == width)) {
		CATTL3_CUDA(cudaMemcpyAsync(dst, src, width * height, cudaMemcpyDeviceToDevice, ctx->stream));
	} else {
		CATTL3_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, height, cudaMemcpyDeviceToDevice, ctx->stream));
	}
	return CATTL3_OK;
}
int cattl3_host_alloc(void** p, size_t bytes) {
	CATTL3_REQUIRE(p, "host_alloc: null out pointer");
	CATTL3_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
	return CATTL3_OK;
}
int cattl3_host_free(void* p) {
	if (p) CATTL3_CUDA(cudaFreeHost(p));
	return CATTL3_OK;
}

int cattl3_conv_output_dims(const cattl3_conv_geom* g, int transposed, int32_t* oh, int32_t* ow) {
	int a, b;
	CATTL3_CHECK(check_geom(g, transposed, &a, &b));
	if (oh) *oh = a;
	if (ow) *ow = b;
	return CATTL3_OK;
}

int cattl3_pool_output_dims(const cattl3_pool_geom* g, int32_t* oh, int32_t* ow) {
	CATTL3_REQUIRE(g, "null pool geometry");
	CATTL3_REQUIRE(g->n > 0 && g->h > 0 && g->w > 0 && g->c > 0 && g->rh > 0 && g->rw > 0 && g->sh > 0 && g->sw > 0,
			"pool geometry: non-positive size");
	CATTL3_REQUIRE(g->h >= g->rh && g->w >= g->rw, "pool geometry: receptor larger than input");
	if (oh) *oh = (g->h - g->rh) / g->sh + 1;
	if (ow) *ow = (g->w - g->rw) / g->sw + 1;
	return CATTL3_OK;
}

#define KERNEL_LAYER_API(S, SUF) \
int cattl3_conv_forward_##SUF(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y) { \
	return conv_forward<S>(c, g, x, w, b, y); } \
int cattl3_conv_backward_##SUF(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* dy, S* dw, S* db, S* dx) { \
	return conv_backward<S>(c, g, x, w, dy, dw, db, dx); } \
int cattl3_conv_forward_fused_##SUF(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y, \
		const cattl3_epilogue* ep) { \
	CATTL3_REQUIRE(ep, "conv_forward_fused: null epilogue"); \
	return conv_forward<S>(c, g, x, w, b, y, ep); } \
int cattl3_dense_forward_fused_##SUF(cattl3_ctx* c, int32_t n, int32_t in, int32_t out, const S* x, const S* w, const S* b, \
		S* y, const cattl3_epilogue* ep) { \
	CATTL3_REQUIRE(n > 0 && in > 0 && out > 0, "dense_forward: non-positive size"); \
	CATTL3_REQUIRE(ep, "dense_forward_fused: null epilogue"); \
	cattl3_conv_geom g = dense_geom(n, in, out); \
	return conv_forward<S>(c, &g, x, w, b, y, ep); } \
int cattl3_transconv_forward_fused_##SUF(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y, \
		const cattl3_epilogue* ep) { \
	CATTL3_REQUIRE(ep, "transconv_forward_fused: null epilogue"); \
	return transconv_forward<S>(c, g, x, w, b, y, ep); } \
int cattl3_transconv_forward_##SUF(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y) { \
	return transconv_forward<S>(c, g, x, w, b, y); } \
int cattl3_transconv_backward_##SUF(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* dy, S* dw, S* db, S* dx) { \
	return transconv_backward<S>(c, g, x, w, dy, dw, db, dx); } \
int cattl3_dense_forward_##SUF(cattl3_ctx* c, int32_t n, int32_t in, int32_t out, const S* x, const S* w, const S* b, S* y) { \
	CATTL3_REQUIRE(n > 0 && in > 0 && out > 0, "dense_forward: non-positive size"); \
	cattl3_conv_geom g = dense_geom(n, in, out); \
	return conv_forward<S>(c, &g, x, w, b, y); } \
int cattl3_dense_backward_##SUF(cattl3_ctx* c, int32_t n, int32_t in, int32_t out, const S* x, const S* w, const S* dy, S* dw, S* db, S* dx) { \
	CATTL3_REQUIRE(n > 0 && in > 0 && out > 0, "dense_backward: non-positive size"); \
	cattl3_conv_geom g = dense_geom(n, in, out); \
	return conv_backward<S>(c, &g, x, w, dy, dw, db, dx); }

KERNEL_LAYER_API(float, f32)
KERNEL_LAYER_API(double, f64)

// Host-buffer convolution (what the reference's Layer API hands over: host tensors in, host tensors out), as a
// three-stream pipeline: uploads on `up`, kernels on the context's stream, downloads on `down`, chained by events.
// The work is cut into chunks of FILTERS -- the slowest dimension of y and dY, so a chunk of either is one contiguous
// block, and a chunk of W / b / dW / db as well -- so that PCIe runs in both directions while the kernels work:
//   forward : x up | chunk k: y[:, k] = conv(x, W[:, k]) -> y[:, k] down while chunk k + 1 computes
//   backward: chunk k: dY[:, k] up while chunk k - 1 computes | dW[:, k], db[k] += ...; dX += dY[:, k] * W[:, k]^T | dX down
// The *_async entry points return once everything is enqueued (a forward's download of y then overlaps the
// following backward's upload of dY: full-duplex PCIe); cattl3_host_wait() is the point after which the host
// buffers hold the results.  The plain entry points are *_async + cattl3_host_wait.
// x_dev_keep (n*h*w*c floats, optional) receives the device copy of x so that the backward call needs no second upload.
namespace {

enum { ST_X = 0, ST_Y = 1, ST_DX = 2, ST_DY = 3, ST_DXT = 4 };

int host_pipe_init(cattl3_ctx* ctx) {
	if (ctx->up_stream) return CATTL3_OK;
	CATTL3_CUDA(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
	CATTL3_CUDA(cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
	for (int i = 0; i < cattl3_ctx::HOST_EVENTS; ++i)
		CATTL3_CUDA(cudaEventCreateWithFlags(&ctx->host_ev[i], cudaEventDisableTiming));
	return CATTL3_OK;
}

// filters per chunk: whole tensor-core tiles, at most CATTL3_HOST_CHUNKS (default 4) chunks
int host_chunk_filters(const cattl3_conv_geom* g) {
	const char* env = getenv("CATTL3_HOST_CHUNKS");
	const int want = env ? atoi(env) : 4;
	int chunks = want < 1 ? 1 : (want > cattl3_ctx::HOST_MAX_CHUNKS ? cattl3_ctx::HOST_MAX_CHUNKS : want);
	while (chunks > 1 && (g->f % chunks != 0 || (g->f / chunks) % 64 != 0)) --chunks;
	return g->f / chunks;
}

// ---- strips of image columns --------------------------------------------------------------------------------------
// Chunks of filters leave a serial head (all of x must be up before anything computes) and tail (dX comes down after the
// last chunk).  With a stride-1 convolution on the tensor-core path the work can be cut along the image WIDTH instead: a
// strip of columns of x / y / dY / dX is, per channel, one contiguous block (N * H * columns elements), i.e. a 2-D copy of
// `channels` long rows; a strip of y needs the strip of x plus a halo column either side, a strip of dW is the reduction over
// the strip's rows of dY (accumulated strip by strip), a strip of dX needs the strip of dY plus its halo.  Uploads, kernels
// and downloads then overlap from the first strip to the last, and only one strip's kernels and download are exposed at the end.
struct Strips { int count; int w0[cattl3_ctx::HOST_MAX_CHUNKS + 1]; };

static Strips cut_strips(int width) {
	const char* env = getenv("CATTL3_HOST_STRIPS");
	int want = env ? atoi(env) : 4;
	if (want > cattl3_ctx::HOST_MAX_CHUNKS / 2) want = cattl3_ctx::HOST_MAX_CHUNKS / 2;
	if (want > width / 2) want = width / 2;
	Strips st;
	st.count = want < 1 ? 1 : want;
	for (int k = 0; k <= st.count; ++k) st.w0[k] = (int) ((long long) width * k / st.count);
	return st;
}

// stride 1, undilated or dilated, every pass on the tcgen05 kernels (they are the ones that take sub-lattices / row ranges)
static bool host_strips_ok(cattl3_ctx* ctx, const cattl3_conv_geom* g, int oh, int ow) {
	if (getenv("CATTL3_NO_HOST_STRIPS") || g->sh != 1 || g->sw != 1 || ctx->conv_path == CATTL3_PATH_SIMT || ctx->conv_path == CATTL3_PATH_FMA)
		return false;
	if (g->w < 4 || ow < 4) return false;
	const long long T = (long long) g->rh * g->rw;
	GatherGeom gf = fwd_gather(g->n, g->h, g->w, g->c, oh, ow, g->f, g);
	gf.w_stap = 1; gf.w_sr = T; gf.w_sj = T * g->c;
	GatherGeom gd = bwd_gather(g->n, oh, ow, g->f, g->h, g->w, g->c, g);
	gd.w_stap = 1; gd.w_sr = T * g->c; gd.w_sj = T;
	return tc_gather_gemm_supported(ctx, gf) && tc_gather_gemm_supported(ctx, gd) && tc_wgrad_supported(ctx, gf);
}

static int host_forward_strips(cattl3_ctx* ctx, const cattl3_conv_geom* g, int oh, int ow, const float* x_host, const float* w_dev,
		const float* b_dev, float* y_host, float* xd, float* yd) {
	cudaEvent_t* ev = ctx->host_ev;
	const Strips xs = cut_strips(g->w), ys = cut_strips(ow);
	const size_t x_pitch = sizeof(float) * (size_t) g->n * g->h * g->w, y_pitch = sizeof(float) * (size_t) g->n * oh * ow;
	const long long T = (long long) g->rh * g->rw;
	for (int k = 0; k < xs.count; ++k) {
		const size_t off = (size_t) g->n * g->h * xs.w0[k];
		CATTL3_CUDA(cudaMemcpy2DAsync(xd + off, x_pitch, x_host + off, x_pitch, sizeof(float) * (size_t) g->n * g->h * (xs.w0[k + 1] - xs.w0[k]),
				(size_t) g->c, cudaMemcpyHostToDevice, ctx->up_stream));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_CHUNK + k], ctx->up_stream));
	}
	int waited = 0;   // strips of x the compute stream has waited for
	for (int k = 0; k < ys.count; ++k) {
		// the last source column this strip of outputs reads
		int need = (ys.w0[k + 1] - 1) + (g->rw - 1) * (g->dw + 1) - g->pw;
		if (need > g->w - 1) need = g->w - 1;
		while (waited < xs.count && xs.w0[waited] <= need) {
			CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_CHUNK + waited], 0));
			++waited;
		}
		GatherGeom gg = fwd_gather(g->n, g->h, g->w, g->c, oh, ys.w0[k + 1] - ys.w0[k], g->f, g);
		gg.w_stap = 1; gg.w_sr = T; gg.w_sj = T * g->c;
		gg.cw += ys.w0[k] * gg.aw;                         // the strip's first output column
		gg.out_H = oh; gg.out_W = ow; gg.out_w0 = ys.w0[k];
		CATTL3_CHECK(tc_gather_gemm_f32(ctx, gg, xd, w_dev, b_dev, 1, yd, nullptr));
		ctx->last_path = "tcgen05";
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_CHUNK + xs.count + k], ctx->stream));
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->down_stream, ev[cattl3_ctx::EV_CHUNK + xs.count + k], 0));
		const size_t off = (size_t) g->n * oh * ys.w0[k];
		CATTL3_CUDA(cudaMemcpy2DAsync(y_host + off, y_pitch, yd + off, y_pitch, sizeof(float) * (size_t) g->n * oh * (ys.w0[k + 1] - ys.w0[k]),
				(size_t) g->f, cudaMemcpyDeviceToHost, ctx->down_stream));
	}
	while (waited < xs.count) {   // (strips of x no output needed: still part of this call's upload)
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_CHUNK + waited], 0));
		++waited;
	}
	return CATTL3_OK;
}

static int host_backward_strips(cattl3_ctx* ctx, const cattl3_conv_geom* g, int oh, int ow, const float* x_dev, const float* w_dev,
		const float* dy_host, float* dw_dev, float* db_dev, float* dx_host, float* dyd, float* dxd) {
	cudaEvent_t* ev = ctx->host_ev;
	const Strips xs = cut_strips(g->w), ys = cut_strips(ow);
	const size_t x_pitch = sizeof(float) * (size_t) g->n * g->h * g->w, y_pitch = sizeof(float) * (size_t) g->n * oh * ow;
	const long long T = (long long) g->rh * g->rw;
	for (int k = 0; k < ys.count; ++k) {
		const size_t off = (size_t) g->n * oh * ys.w0[k];
		CATTL3_CUDA(cudaMemcpy2DAsync(dyd + off, y_pitch, dy_host + off, y_pitch, sizeof(float) * (size_t) g->n * oh * (ys.w0[k + 1] - ys.w0[k]),
				(size_t) g->f, cudaMemcpyHostToDevice, ctx->up_stream));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_CHUNK + k], ctx->up_stream));
	}
	int done_dx = 0;   // strips of dX computed so far
	for (int k = 0; k < ys.count; ++k) {
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_CHUNK + k], 0));
		if (dw_dev) {
			// dW += (columns of the gathered x)^T dY over the strip's rows m = n + N * (oh + OH * ow); db from the same read
			GatherGeom gw = fwd_gather(g->n, g->h, g->w, g->c, oh, ow, g->f, g);
			gw.w_stap = 1; gw.w_sr = T; gw.w_sj = T * g->c;
			gw.m_first = (long long) g->n * oh * ys.w0[k];
			gw.m_count = (long long) g->n * oh * (ys.w0[k + 1] - ys.w0[k]);
			CATTL3_CHECK(tc_wgrad_f32(ctx, gw, x_dev, dyd, dw_dev, db_dev));
			ctx->last_path = "tcgen05";
		}
		// every strip of dX whose last needed column of dY has now arrived
		while (dxd && done_dx < xs.count) {
			int need = (xs.w0[done_dx + 1] - 1) + g->pw;   // source column = iw + pw - rw * (dw + 1), largest at rw = 0
			if (need > ow - 1) need = ow - 1;
			if (need >= ys.w0[k + 1]) break;
			const int j = done_dx++;
			GatherGeom gd = bwd_gather(g->n, oh, ow, g->f, g->h, xs.w0[j + 1] - xs.w0[j], g->c, g);
			gd.w_stap = 1; gd.w_sr = T * g->c; gd.w_sj = T;
			gd.cw += xs.w0[j] * gd.aw;
			gd.out_H = g->h; gd.out_W = g->w; gd.out_w0 = xs.w0[j];
			CATTL3_CHECK(tc_gather_gemm_f32(ctx, gd, dyd, w_dev, nullptr, 0, dxd, nullptr));
			CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_CHUNK + ys.count + j], ctx->stream));
			CATTL3_CUDA(cudaStreamWaitEvent(ctx->down_stream, ev[cattl3_ctx::EV_CHUNK + ys.count + j], 0));
			const size_t off = (size_t) g->n * g->h * xs.w0[j];
			CATTL3_CUDA(cudaMemcpy2DAsync(dx_host + off, x_pitch, dxd + off, x_pitch, sizeof(float) * (size_t) g->n * g->h * (xs.w0[j + 1] - xs.w0[j]),
					(size_t) g->c, cudaMemcpyDeviceToHost, ctx->down_stream));
		}
	}
	return CATTL3_OK;
}

}  // namespace

int cattl3_host_wait(cattl3_ctx* ctx) {
	CATTL3_CHECK(check_ctx(ctx));
	if (ctx->up_stream) {
		CATTL3_CUDA(cudaStreamSynchronize(ctx->up_stream));
		CATTL3_CUDA(cudaStreamSynchronize(ctx->stream));
		CATTL3_CUDA(cudaStreamSynchronize(ctx->down_stream));
	} else {
		CATTL3_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	return CATTL3_OK;
}

int cattl3_conv_forward_host_async_f32(cattl3_ctx* ctx, const cattl3_conv_geom* g, const float* x_host, const float* w_dev,
		const float* b_dev, float* y_host, float* x_dev_keep) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(!ctx->capturing, "conv_forward_host: not inside a step graph");
	int oh, ow;
	CATTL3_CHECK(check_geom(g, 0, &oh, &ow));
	CATTL3_REQUIRE(x_host && y_host, "conv_forward_host: null host tensor");
	CATTL3_CHECK(host_pipe_init(ctx));
	const size_t xb = sizeof(float) * (size_t) g->n * g->h * g->w * g->c;
	const size_t M = (size_t) g->n * oh * ow;
	const size_t K = (size_t) g->rh * g->rw * g->c;
	float* xd = x_dev_keep;
	if (!xd) {
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->stage_dev[ST_X], &ctx->stage_dev_bytes[ST_X], xb));
		xd = (float*) ctx->stage_dev[ST_X];
	}
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->stage_dev[ST_Y], &ctx->stage_dev_bytes[ST_Y], sizeof(float) * M * g->f));
	float* yd = (float*) ctx->stage_dev[ST_Y];
	cudaEvent_t* ev = ctx->host_ev;
	// the upload of x follows everything enqueued so far that may still read the buffer it lands in
	CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_ENTRY], ctx->stream));
	CATTL3_CUDA(cudaStreamWaitEvent(ctx->up_stream, ev[cattl3_ctx::EV_ENTRY], 0));
	if (ctx->host_ev_used[cattl3_ctx::EV_Y_FREE])   // an earlier forward's download of the y stage
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_Y_FREE], 0));
	if (b_dev && host_strips_ok(ctx, g, oh, ow)) {
		// strips of image columns: uploads, kernels and downloads overlap from the first strip on
		CATTL3_CHECK(host_forward_strips(ctx, g, oh, ow, x_host, w_dev, b_dev, y_host, xd, yd));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_Y_FREE], ctx->down_stream));
		ctx->host_ev_used[cattl3_ctx::EV_Y_FREE] = true;
		return CATTL3_OK;
	}
	CATTL3_CUDA(cudaMemcpyAsync(xd, x_host, xb, cudaMemcpyHostToDevice, ctx->up_stream));
	CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_UP], ctx->up_stream));
	CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_UP], 0));
	const int fc = host_chunk_filters(g);
	cattl3_conv_geom gc = *g;
	gc.f = fc;
	for (int f0 = 0, k = 0; f0 < g->f; f0 += fc, ++k) {
		CATTL3_CHECK(conv_forward<float>(ctx, &gc, xd, w_dev + K * f0, b_dev ? b_dev + f0 : nullptr, yd + M * f0));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_CHUNK + k], ctx->stream));
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->down_stream, ev[cattl3_ctx::EV_CHUNK + k], 0));
		CATTL3_CUDA(cudaMemcpyAsync(y_host + M * f0, yd + M * f0, sizeof(float) * M * fc, cudaMemcpyDeviceToHost, ctx->down_stream));
	}
	CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_Y_FREE], ctx->down_stream));
	ctx->host_ev_used[cattl3_ctx::EV_Y_FREE] = true;
	return CATTL3_OK;
}

int cattl3_conv_backward_host_async_f32(cattl3_ctx* ctx, const cattl3_conv_geom* g, const float* x_dev, const float* w_dev,
		const float* dy_host, float* dw_dev, float* db_dev, float* dx_host) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(!ctx->capturing, "conv_backward_host: not inside a step graph");
	int oh, ow;
	CATTL3_CHECK(check_geom(g, 0, &oh, &ow));
	CATTL3_REQUIRE(x_dev && dy_host, "conv_backward_host: null tensor");
	CATTL3_CHECK(host_pipe_init(ctx));
	const size_t x_elems = (size_t) g->n * g->h * g->w * g->c;
	const size_t xb = sizeof(float) * x_elems;
	const size_t M = (size_t) g->n * oh * ow;
	const size_t K = (size_t) g->rh * g->rw * g->c;
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->stage_dev[ST_DY], &ctx->stage_dev_bytes[ST_DY], sizeof(float) * M * g->f));
	float* dyd = (float*) ctx->stage_dev[ST_DY];
	const int fc = host_chunk_filters(g);
	float* dxd = nullptr;
	float* dxt = nullptr;
	if (dx_host) {
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->stage_dev[ST_DX], &ctx->stage_dev_bytes[ST_DX], xb));
		dxd = (float*) ctx->stage_dev[ST_DX];
		if (fc < g->f) {
			CATTL3_CHECK(ensure_buffer(ctx, &ctx->stage_dev[ST_DXT], &ctx->stage_dev_bytes[ST_DXT], xb));
			dxt = (float*) ctx->stage_dev[ST_DXT];
		}
	}
	cudaEvent_t* ev = ctx->host_ev;
	if (ctx->host_ev_used[cattl3_ctx::EV_DY_FREE])   // an earlier backward's kernels still reading the dY stage
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->up_stream, ev[cattl3_ctx::EV_DY_FREE], 0));
	if (dxd && ctx->host_ev_used[cattl3_ctx::EV_DX_FREE])   // an earlier backward's download of the dX stage
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_DX_FREE], 0));
	if (host_strips_ok(ctx, g, oh, ow) && (!dw_dev) == (!db_dev)) {
		CATTL3_CHECK(host_backward_strips(ctx, g, oh, ow, x_dev, w_dev, dy_host, dw_dev, db_dev, dx_host, dyd, dxd));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_DY_FREE], ctx->stream));
		ctx->host_ev_used[cattl3_ctx::EV_DY_FREE] = true;
		if (dx_host) {
			CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_DX_FREE], ctx->down_stream));
			ctx->host_ev_used[cattl3_ctx::EV_DX_FREE] = true;
		}
		return CATTL3_OK;
	}
	cattl3_conv_geom gc = *g;
	gc.f = fc;
	for (int f0 = 0, k = 0; f0 < g->f; f0 += fc, ++k) {
		CATTL3_CUDA(cudaMemcpyAsync(dyd + M * f0, dy_host + M * f0, sizeof(float) * M * fc, cudaMemcpyHostToDevice, ctx->up_stream));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_CHUNK + k], ctx->up_stream));
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->stream, ev[cattl3_ctx::EV_CHUNK + k], 0));
		// this chunk's share of every gradient: dW[:, chunk], db[chunk] (accumulating) and dY[:, chunk] * W[:, chunk]^T of dX
		CATTL3_CHECK(conv_backward<float>(ctx, &gc, x_dev, w_dev + K * f0, dyd + M * f0, dw_dev ? dw_dev + K * f0 : nullptr,
				db_dev ? db_dev + f0 : nullptr, dxd ? (k == 0 ? dxd : dxt) : nullptr));
		if (dxd && k > 0)
			CATTL3_CHECK(cattl3_add_inplace_f32(ctx, (int64_t) x_elems, dxd, dxt));
	}
	CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_DY_FREE], ctx->stream));
	ctx->host_ev_used[cattl3_ctx::EV_DY_FREE] = true;
	if (dx_host) {
		CATTL3_CUDA(cudaStreamWaitEvent(ctx->down_stream, ev[cattl3_ctx::EV_DY_FREE], 0));
		CATTL3_CUDA(cudaMemcpyAsync(dx_host, dxd, xb, cudaMemcpyDeviceToHost, ctx->down_stream));
		CATTL3_CUDA(cudaEventRecord(ev[cattl3_ctx::EV_DX_FREE], ctx->down_stream));
		ctx->host_ev_used[cattl3_ctx::EV_DX_FREE] = true;
	}
	return CATTL3_OK;
}

int cattl3_conv_forward_host_f32(cattl3_ctx* ctx, const cattl3_conv_geom* g, const float* x_host, const float* w_dev,
		const float* b_dev, float* y_host, float* x_dev_keep) {
	CATTL3_CHECK(cattl3_conv_forward_host_async_f32(ctx, g, x_host, w_dev, b_dev, y_host, x_dev_keep));
	return cattl3_host_wait(ctx);
}

int cattl3_conv_backward_host_f32(cattl3_ctx* ctx, const cattl3_conv_geom* g, const float* x_dev, const float* w_dev,
		const float* dy_host, float* dw_dev, float* db_dev, float* dx_host) {
	CATTL3_CHECK(cattl3_conv_backward_host_async_f32(ctx, g, x_dev, w_dev, dy_host, dw_dev, db_dev, dx_host));
	return cattl3_host_wait(ctx);
}

}
