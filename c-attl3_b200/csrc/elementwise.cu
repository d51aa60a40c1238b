// elementwise.cu -- the HBM-bound layers around the kernel layers: activations, pooling, batch
// normalisation, the fused optimizer step and the small network-glue helpers.  All kernels are
// grid-stride with 16-byte vector accesses where alignment allows, sized in multiples of the SM
// count (common.cuh: ew_grid); none of them stages through shared memory except reductions.
#include "activations.cuh"

namespace cattl3 {

template<typename S, int KIND>
__global__ void __launch_bounds__(256) act_fwd_kernel(long long count, int vec_ok, S alpha,
		const S* __restrict__ x, S* __restrict__ y) {
	typedef typename V16<S>::type V;
	constexpr int G = V16<S>::G;
	const long long nvec = vec_ok ? count / G : 0;
	const long long stride = (long long) gridDim.x * 256;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += stride) {
		V v = reinterpret_cast<const V*>(x)[i];
		S* e = reinterpret_cast<S*>(&v);
		#pragma unroll
		for (int k = 0; k < G; ++k) e[k] = act_fwd<S, KIND>(e[k], alpha);
		reinterpret_cast<V*>(y)[i] = v;
	}
	for (long long i = nvec * G + blockIdx.x * 256ll + threadIdx.x; i < count; i += stride)
		y[i] = act_fwd<S, KIND>(x[i], alpha);
}

template<typename S, int KIND>
__global__ void __launch_bounds__(256) act_bwd_kernel(long long count, int vec_ok, S alpha,
		const S* __restrict__ x, const S* __restrict__ y, const S* __restrict__ dy, S* __restrict__ dx) {
	typedef typename V16<S>::type V;
	constexpr int G = V16<S>::G;
	constexpr bool NEED_X = KIND == CATTL3_ACT_RELU || KIND == CATTL3_ACT_LEAKY_RELU ||
			KIND == CATTL3_ACT_ELU || KIND == CATTL3_ACT_SWISH || KIND == CATTL3_ACT_SOFTPLUS;
	constexpr bool NEED_Y = KIND == CATTL3_ACT_ELU || KIND == CATTL3_ACT_SIGMOID || KIND == CATTL3_ACT_TANH;
	const long long nvec = vec_ok ? count / G : 0;
	const long long stride = (long long) gridDim.x * 256;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += stride) {
		V vx, vy, vg = reinterpret_cast<const V*>(dy)[i];
		if (NEED_X) vx = reinterpret_cast<const V*>(x)[i];
		if (NEED_Y) vy = reinterpret_cast<const V*>(y)[i];
		S* ex = reinterpret_cast<S*>(&vx);
		S* ey = reinterpret_cast<S*>(&vy);
		S* eg = reinterpret_cast<S*>(&vg);
		#pragma unroll
		for (int k = 0; k < G; ++k)
			eg[k] = act_bwd<S, KIND>(NEED_X ? ex[k] : (S) 0, NEED_Y ? ey[k] : (S) 0, eg[k], alpha);
		reinterpret_cast<V*>(dx)[i] = vg;
	}
	for (long long i = nvec * G + blockIdx.x * 256ll + threadIdx.x; i < count; i += stride)
		dx[i] = act_bwd<S, KIND>(NEED_X ? x[i] : (S) 0, NEED_Y ? y[i] : (S) 0, dy[i], alpha);
}

// Softmax over `vol` for each of `rows` rows (rows fastest): one thread per row keeps the accesses
// of a warp coalesced (consecutive rows are adjacent in memory).  SoftmaxActivationLayer.hpp:49-78.
template<typename S>
__global__ void __launch_bounds__(128) softmax_fwd_kernel(long long rows, long long vol, S eps,
		const S* __restrict__ x, S* __restrict__ y) {
	for (long long r = blockIdx.x * 128ll + threadIdx.x; r < rows; r += (long long) gridDim.x * 128) {
		S mx = x[r];
		for (long long j = 1; j < vol; ++j) { S v = x[r + rows * j]; mx = v > mx ? v : mx; }
		S sum = 0;
		for (long long j = 0; j < vol; ++j) sum += dev_exp<S>(x[r + rows * j] - mx);
		const S den = sum + eps;
		for (long long j = 0; j < vol; ++j) y[r + rows * j] = dev_exp<S>(x[r + rows * j] - mx) / den;
	}
}

template<typename S>
__global__ void __launch_bounds__(128) softmax_bwd_kernel(long long rows, long long vol,
		const S* __restrict__ y, const S* __restrict__ dy, S* __restrict__ dx) {
	for (long long r = blockIdx.x * 128ll + threadIdx.x; r < rows; r += (long long) gridDim.x * 128) {
		S dot = 0;
		for (long long j = 0; j < vol; ++j) dot += y[r + rows * j] * dy[r + rows * j];
		for (long long j = 0; j < vol; ++j) dx[r + rows * j] = y[r + rows * j] * (dy[r + rows * j] - dot);
	}
}

template<typename S>
int activation_forward(cattl3_ctx* ctx, int kind, S alpha, int64_t rows, int64_t vol, const S* x, S* y) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(rows > 0 && vol > 0 && x && y, "activation_forward: bad arguments");
	const long long count = rows * vol;
	const int grid = ew_grid(ctx, ceil_div(count, V16<S>::G), 256);
	const int vec_ok = aligned16(x) && aligned16(y);
	switch (kind) {
#define CASE(K) case K: act_fwd_kernel<S, K><<<grid, 256, 0, ctx->stream>>>(count, vec_ok, alpha, x, y); break;
		CASE(CATTL3_ACT_RELU) CASE(CATTL3_ACT_LEAKY_RELU) CASE(CATTL3_ACT_ELU) CASE(CATTL3_ACT_SWISH)
		CASE(CATTL3_ACT_SIGMOID) CASE(CATTL3_ACT_TANH) CASE(CATTL3_ACT_SOFTPLUS)
#undef CASE
		case CATTL3_ACT_SOFTMAX:
			softmax_fwd_kernel<S><<<ew_grid(ctx, rows, 128), 128, 0, ctx->stream>>>(rows, vol, alpha, x, y);
			break;
		default:
			set_error("activation_forward: unknown kind %d", kind);
			return CATTL3_ERR_INVALID;
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

template<typename S>
int activation_backward(cattl3_ctx* ctx, int kind, S alpha, int64_t rows, int64_t vol, const S* x,
		const S* y, const S* dy, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(rows > 0 && vol > 0 && dy && dx, "activation_backward: bad arguments");
	const long long count = rows * vol;
	const int grid = ew_grid(ctx, ceil_div(count, V16<S>::G), 256);
	const int vec_ok = aligned16(x) && aligned16(y) && aligned16(dy) && aligned16(dx);
	switch (kind) {
#define CASE(K) case K: act_bwd_kernel<S, K><<<grid, 256, 0, ctx->stream>>>(count, vec_ok, alpha, x, y, dy, dx); break;
		CASE(CATTL3_ACT_RELU) CASE(CATTL3_ACT_LEAKY_RELU) CASE(CATTL3_ACT_ELU) CASE(CATTL3_ACT_SWISH)
		CASE(CATTL3_ACT_SIGMOID) CASE(CATTL3_ACT_TANH) CASE(CATTL3_ACT_SOFTPLUS)
#undef CASE
		case CATTL3_ACT_SOFTMAX:
			CATTL3_REQUIRE(y, "softmax backward needs the cached output");
			softmax_bwd_kernel<S><<<ew_grid(ctx, rows, 128), 128, 0, ctx->stream>>>(rows, vol, y, dy, dx);
			break;
		default:
			set_error("activation_backward: unknown kind %d", kind);
			return CATTL3_ERR_INVALID;
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- pooling -------------------------------------------------------------------------------------
// One thread per output element, n fastest => coalesced reads of every window tap and coalesced
// writes.  Max: strict '>' from lowest(), width-outer / height-inner scan (MaxPoolLayer.hpp:45-58).
template<typename S> __device__ __forceinline__ S lowest();
template<> __device__ __forceinline__ float lowest<float>() { return -3.402823466e+38f; }
template<> __device__ __forceinline__ double lowest<double>() { return -1.7976931348623157e+308; }

// Thread mapping shared by the two pooling kernels: the batch index n is the fastest dimension, so a thread owns
// G consecutive batch entries (one 16-byte vector; G = 1 when n or the alignment forbids it) of one
// (pixel, channel) row.  tv threads cover a row, 256 / tv rows share a CTA; the (h, w, c) decomposition of the row
// index is 32-bit and paid once per row, not per element.
template<typename S, bool VEC> struct PoolVec { typedef typename V16<S>::type type; static constexpr int G = V16<S>::G; };
template<typename S> struct PoolVec<S, false> { typedef S type; static constexpr int G = 1; };
template<int G> struct ByteVec;
template<> struct ByteVec<1> { typedef uint8_t type; };
template<> struct ByteVec<2> { typedef uchar2 type; };
template<> struct ByteVec<4> { typedef uchar4 type; };

template<typename S, int KIND, bool VEC>
__global__ void __launch_bounds__(256) pool_fwd_kernel(cattl3_pool_geom g, int OH, int OW, int tv,
		const S* __restrict__ x, S* __restrict__ y, uint8_t* __restrict__ argmax) {
	typedef typename PoolVec<S, VEC>::type V;
	constexpr int G = PoolVec<S, VEC>::G;
	typedef typename ByteVec<G>::type BV;
	const int nv = g.n / G;
	const int rows_per_cta = 256 / tv;
	const int v0 = threadIdx.x % tv, rsub = threadIdx.x / tv;
	if (rsub >= rows_per_cta) return;
	const unsigned rows = (unsigned) OH * OW * g.c;
	for (unsigned r = blockIdx.x * rows_per_cta + rsub; r < rows; r += gridDim.x * rows_per_cta) {
		const unsigned oh = r % OH, t = r / OH;
		const unsigned ow = t % OW, c = t / OW;
		const S* base = x + (long long) g.n * (oh * g.sh + (long long) g.h * (ow * g.sw + (long long) g.w * c));
		for (int v = v0; v < nv; v += tv) {
			__align__(16) S best[G];
			__align__(4) uint8_t bi[G];
			#pragma unroll
			for (int e = 0; e < G; ++e) { best[e] = KIND == CATTL3_POOL_MAX ? lowest<S>() : (S) 0; bi[e] = 0; }
			for (int k = 0; k < g.rw; ++k)
				for (int l = 0; l < g.rh; ++l) {
					const V vv = *reinterpret_cast<const V*>(base + (long long) g.n * (l + (long long) g.h * k) + v * G);
					const S* ev = reinterpret_cast<const S*>(&vv);
					#pragma unroll
					for (int e = 0; e < G; ++e) {
						if (KIND == CATTL3_POOL_MAX) {
							if (ev[e] > best[e]) { best[e] = ev[e]; bi[e] = (uint8_t) (k * g.rh + l); }
						} else {
							best[e] += ev[e];
						}
					}
				}
			if (KIND == CATTL3_POOL_MEAN) {
				#pragma unroll
				for (int e = 0; e < G; ++e) best[e] = best[e] / (S) (g.rh * g.rw);
			}
			const long long o = (long long) g.n * r + v * G;
			*reinterpret_cast<V*>(y + o) = *reinterpret_cast<const V*>(best);
			if (KIND == CATTL3_POOL_MAX && argmax) *reinterpret_cast<BV*>(argmax + o) = *reinterpret_cast<const BV*>(bi);
		}
	}
}

// Gather form of PoolLayer::_pass_back (PoolLayer.hpp:98-116): a thread owns G batch entries of one INPUT
// (pixel, channel) row and sums the contributions of every window that covers it (overlapping windows
// accumulate), no atomics.
template<typename S, int KIND, bool VEC>
__global__ void __launch_bounds__(256) pool_bwd_kernel(cattl3_pool_geom g, int OH, int OW, int tv,
		const S* __restrict__ dy, const uint8_t* __restrict__ argmax, S* __restrict__ dx) {
	typedef typename PoolVec<S, VEC>::type V;
	constexpr int G = PoolVec<S, VEC>::G;
	typedef typename ByteVec<G>::type BV;
	const int nv = g.n / G;
	const int rows_per_cta = 256 / tv;
	const int v0 = threadIdx.x % tv, rsub = threadIdx.x / tv;
	if (rsub >= rows_per_cta) return;
	const unsigned rows = (unsigned) g.h * g.w * g.c;
	const S inv_area = (S) 1 / (S) (g.rh * g.rw);
	for (unsigned r = blockIdx.x * rows_per_cta + rsub; r < rows; r += gridDim.x * rows_per_cta) {
		const int h = (int) (r % g.h);
		const unsigned t = r / g.h;
		const int w = (int) (t % g.w), c = (int) (t / g.w);
		int oh_lo = h - g.rh + 1; oh_lo = oh_lo <= 0 ? 0 : (oh_lo + g.sh - 1) / g.sh;
		int ow_lo = w - g.rw + 1; ow_lo = ow_lo <= 0 ? 0 : (ow_lo + g.sw - 1) / g.sw;
		int oh_hi = h / g.sh; if (oh_hi > OH - 1) oh_hi = OH - 1;
		int ow_hi = w / g.sw; if (ow_hi > OW - 1) ow_hi = OW - 1;
		for (int v = v0; v < nv; v += tv) {
			__align__(16) S s[G];
			#pragma unroll
			for (int e = 0; e < G; ++e) s[e] = (S) 0;
			for (int ow = ow_lo; ow <= ow_hi; ++ow)
				for (int oh = oh_lo; oh <= oh_hi; ++oh) {
					const long long o = (long long) g.n * (oh + (long long) OH * (ow + (long long) OW * c)) + v * G;
					const V vg = *reinterpret_cast<const V*>(dy + o);
					const S* eg = reinterpret_cast<const S*>(&vg);
					if (KIND == CATTL3_POOL_MAX) {
						const int idx = (w - ow * g.sw) * g.rh + (h - oh * g.sh);
						const BV vb = *reinterpret_cast<const BV*>(argmax + o);
						const uint8_t* eb = reinterpret_cast<const uint8_t*>(&vb);
						#pragma unroll
						for (int e = 0; e < G; ++e) if ((int) eb[e] == idx) s[e] += eg[e];
					} else {
						#pragma unroll
						for (int e = 0; e < G; ++e) s[e] += eg[e] * inv_area;
					}
				}
			*reinterpret_cast<V*>(dx + (long long) g.n * r + v * G) = *reinterpret_cast<const V*>(s);
		}
	}
}

// Disjoint windows (stride >= window in both directions, the usual 2x2 / 2 case): every input element belongs to
// at most one window, so the backward pass streams: a thread owns G batch entries of one OUTPUT row, reads dy and
// the argmax once and writes its whole stride cell of dx (zeros outside the window; the last cell of a row / column
// extends to the input's edge), so dx is written exactly once and nothing is read twice.
template<typename S, int KIND, bool VEC>
__global__ void __launch_bounds__(256) pool_bwd_disjoint_kernel(cattl3_pool_geom g, int OH, int OW, int tv,
		const S* __restrict__ dy, const uint8_t* __restrict__ argmax, S* __restrict__ dx) {
	typedef typename PoolVec<S, VEC>::type V;
	constexpr int G = PoolVec<S, VEC>::G;
	typedef typename ByteVec<G>::type BV;
	const int nv = g.n / G;
	const int rows_per_cta = 256 / tv;
	const int v0 = threadIdx.x % tv, rsub = threadIdx.x / tv;
	if (rsub >= rows_per_cta) return;
	const unsigned rows = (unsigned) OH * OW * g.c;
	const S inv_area = (S) 1 / (S) (g.rh * g.rw);
	for (unsigned r = blockIdx.x * rows_per_cta + rsub; r < rows; r += gridDim.x * rows_per_cta) {
		const int oh = (int) (r % OH);
		const unsigned t = r / OH;
		const int ow = (int) (t % OW), c = (int) (t / OW);
		const int h0 = oh * g.sh, w0 = ow * g.sw;
		const int h1 = oh == OH - 1 ? g.h : h0 + g.sh, w1 = ow == OW - 1 ? g.w : w0 + g.sw;
		for (int v = v0; v < nv; v += tv) {
			const long long o = (long long) g.n * r + v * G;
			const V vg = *reinterpret_cast<const V*>(dy + o);
			const S* eg = reinterpret_cast<const S*>(&vg);
			BV vb;
			if (KIND == CATTL3_POOL_MAX) vb = *reinterpret_cast<const BV*>(argmax + o);
			const uint8_t* eb = reinterpret_cast<const uint8_t*>(&vb);
			for (int w = w0; w < w1; ++w)
				for (int h = h0; h < h1; ++h) {
					const int k = w - w0, l = h - h0;
					const bool inside = k < g.rw && l < g.rh;
					const int idx = k * g.rh + l;
					__align__(16) S out[G];
					#pragma unroll
					for (int e = 0; e < G; ++e) {
						if (KIND == CATTL3_POOL_MAX) out[e] = (inside && (int) eb[e] == idx) ? eg[e] : (S) 0;
						else out[e] = inside ? eg[e] * inv_area : (S) 0;
					}
					*reinterpret_cast<V*>(dx + (long long) g.n * (h + (long long) g.h * (w + (long long) g.w * c)) + v * G) =
							*reinterpret_cast<const V*>(out);
				}
		}
	}
	// columns / rows of the input in front of which no window starts cannot exist (windows start at 0); cells cover
	// [0, H) x [0, W) completely because the last cell extends to the edge.
}

// tv = threads along the batch vectors of a row; grid = enough CTAs for all rows, capped at 8 per SM
static inline void pool_launch_shape(const cattl3_ctx* ctx, int nv, long long rows, int* tv, int* grid) {
	*tv = nv < 256 ? nv : 256;
	const int rows_per_cta = 256 / *tv;
	long long blocks = ceil_div(rows, rows_per_cta);
	const long long cap = (long long) ctx->sm_count * 8;
	*grid = (int) (blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

template<typename S>
int pool_forward(cattl3_ctx* ctx, int kind, const cattl3_pool_geom* g, const S* x, S* y, uint8_t* argmax) {
	CATTL3_CHECK(check_ctx(ctx));
	int32_t oh, ow;
	CATTL3_CHECK(cattl3_pool_output_dims(g, &oh, &ow));
	CATTL3_REQUIRE(x && y, "pool_forward: null tensor");
	CATTL3_REQUIRE(kind == CATTL3_POOL_MAX || kind == CATTL3_POOL_MEAN, "pool_forward: unknown kind %d", kind);
	if (kind == CATTL3_POOL_MAX && g->rh * g->rw > 256) {
		set_error("pool_forward: max-pool windows above 256 elements are unsupported");
		return CATTL3_ERR_UNSUPPORTED;
	}
	const long long rows = (long long) oh * ow * g->c;
	CATTL3_REQUIRE(rows < (1ll << 31) && (long long) g->h * g->w * g->c < (1ll << 31), "pool: more than 2^31 rows");
	constexpr int G = V16<S>::G;
	const bool vec = g->n % G == 0 && aligned16(x) && aligned16(y) && (!argmax || (reinterpret_cast<uintptr_t>(argmax) % G) == 0);
	int tv, grid;
	pool_launch_shape(ctx, vec ? g->n / G : g->n, rows, &tv, &grid);
	if (kind == CATTL3_POOL_MAX) {
		if (vec) pool_fwd_kernel<S, CATTL3_POOL_MAX, true><<<grid, 256, 0, ctx->stream>>>(*g, oh, ow, tv, x, y, argmax);
		else pool_fwd_kernel<S, CATTL3_POOL_MAX, false><<<grid, 256, 0, ctx->stream>>>(*g, oh, ow, tv, x, y, argmax);
	} else {
		if (vec) pool_fwd_kernel<S, CATTL3_POOL_MEAN, true><<<grid, 256, 0, ctx->stream>>>(*g, oh, ow, tv, x, y, nullptr);
		else pool_fwd_kernel<S, CATTL3_POOL_MEAN, false><<<grid, 256, 0, ctx->stream>>>(*g, oh, ow, tv, x, y, nullptr);
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

template<typename S>
int pool_backward(cattl3_ctx* ctx, int kind, const cattl3_pool_geom* g, const S* dy, const uint8_t* argmax, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	int32_t oh, ow;
	CATTL3_CHECK(cattl3_pool_output_dims(g, &oh, &ow));
	CATTL3_REQUIRE(dy && dx, "pool_backward: null tensor");
	CATTL3_REQUIRE(kind == CATTL3_POOL_MAX || kind == CATTL3_POOL_MEAN, "pool_backward: unknown kind %d", kind);
	CATTL3_REQUIRE(kind != CATTL3_POOL_MAX || argmax, "pool_backward: max pooling needs the argmax cache");
	const bool disjoint = g->sh >= g->rh && g->sw >= g->rw;
	const long long rows = disjoint ? (long long) oh * ow * g->c : (long long) g->h * g->w * g->c;
	CATTL3_REQUIRE((long long) g->h * g->w * g->c < (1ll << 31), "pool: more than 2^31 rows");
	constexpr int G = V16<S>::G;
	const bool vec = g->n % G == 0 && aligned16(dy) && aligned16(dx) && (!argmax || (reinterpret_cast<uintptr_t>(argmax) % G) == 0);
	int tv, grid;
	pool_launch_shape(ctx, vec ? g->n / G : g->n, rows, &tv, &grid);
#define POOL_BWD(KERNEL, K, AM) do { \
		if (vec) KERNEL<S, K, true><<<grid, 256, 0, ctx->stream>>>(*g, oh, ow, tv, dy, AM, dx); \
		else KERNEL<S, K, false><<<grid, 256, 0, ctx->stream>>>(*g, oh, ow, tv, dy, AM, dx); } while (0)
	if (kind == CATTL3_POOL_MAX) {
		if (disjoint) POOL_BWD(pool_bwd_disjoint_kernel, CATTL3_POOL_MAX, argmax);
		else POOL_BWD(pool_bwd_kernel, CATTL3_POOL_MAX, argmax);
	} else {
		if (disjoint) POOL_BWD(pool_bwd_disjoint_kernel, CATTL3_POOL_MEAN, nullptr);
		else POOL_BWD(pool_bwd_kernel, CATTL3_POOL_MEAN, nullptr);
	}
#undef POOL_BWD
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- batch normalisation -------------------------------------------------------------------------
// Every group (a channel, or one activation) is L contiguous elements.  Statistics use a two-stage
// deterministic reduction in double: stage 1 (grid = chunks x groups) produces shifted partial sums
// sum(x - K), sum((x - K)^2) with K = first element of the group (guards the E[x^2] - mu^2
// cancellation); stage 2 (one thread per group) combines them in chunk order.
struct BnPartial { double s1, s2; };

__device__ __forceinline__ void block_reduce2(double& a, double& b) {
	__shared__ double ra[8], rb[8];
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		a += __shfl_xor_sync(0xffffffffu, a, o);
		b += __shfl_xor_sync(0xffffffffu, b, o);
	}
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0) { ra[warp] = a; rb[warp] = b; }
	__syncthreads();
	if (warp == 0) {
		a = lane < (int) (blockDim.x >> 5) ? ra[lane] : 0.0;
		b = lane < (int) (blockDim.x >> 5) ? rb[lane] : 0.0;
		#pragma unroll
		for (int o = 4; o > 0; o >>= 1) {
			a += __shfl_xor_sync(0xffffffffu, a, o);
			b += __shfl_xor_sync(0xffffffffu, b, o);
		}
	}
}

// Applies fn(element index within the group range, G consecutive elements) over [lo, hi) of one group: 16-byte
// vectors when `vec` (group base, lo and hi are then multiples of G elements and the tensors 16-byte aligned).
#define BN_FOREACH(S, vec, lo, hi, BODY_VEC, BODY_SCALAR) \
	if (vec) { \
		constexpr int G_ = V16<S>::G; \
		_Pragma("unroll 4") \
		for (long long i = (lo) + (long long) threadIdx.x * G_; i < (hi); i += (long long) blockDim.x * G_) { BODY_VEC } \
	} else { \
		for (long long i = (lo) + threadIdx.x; i < (hi); i += blockDim.x) { BODY_SCALAR } \
	}

template<typename S>
__global__ void __launch_bounds__(256) bn_stats_partial_kernel(long long L, long long chunk, int vec,
		const S* __restrict__ x, BnPartial* __restrict__ part) {
	typedef typename V16<S>::type V;
	const long long g = blockIdx.x;
	const S* xg = x + g * L;
	const double K = (double) xg[0];
	const long long lo = (long long) blockIdx.y * chunk;
	const long long hi = lo + chunk < L ? lo + chunk : L;
	double s1 = 0, s2 = 0;
	BN_FOREACH(S, vec, lo, hi,
		const V v = *reinterpret_cast<const V*>(xg + i);
		const S* e = reinterpret_cast<const S*>(&v);
		_Pragma("unroll")
		for (int k = 0; k < G_; ++k) { const double d = (double) e[k] - K; s1 += d; s2 += d * d; },
		const double d = (double) xg[i] - K; s1 += d; s2 += d * d;)
	block_reduce2(s1, s2);
	if (threadIdx.x == 0) part[g * gridDim.y + blockIdx.y] = BnPartial{ s1, s2 };
}

template<typename S>
__global__ void __launch_bounds__(128) bn_stats_final_kernel(long long groups, long long L, int chunks,
		const S* __restrict__ x, const BnPartial* __restrict__ part, S eps, S decay, int running_init,
		S* __restrict__ running_mean, S* __restrict__ running_inv_sd, S* __restrict__ saved_mean,
		S* __restrict__ saved_inv_sd) {
	const long long g = blockIdx.x * 128ll + threadIdx.x;
	if (g >= groups) return;
	double s1 = 0, s2 = 0;
	for (int c = 0; c < chunks; ++c) { s1 += part[g * chunks + c].s1; s2 += part[g * chunks + c].s2; }
	const double K = (double) x[g * L];
	const double m1 = s1 / (double) L;
	double var = s2 / (double) L - m1 * m1;
	var = var < 0 ? 0 : var;
	const S mean = (S) (K + m1);
	const S inv_sd = (S) (1.0 / sqrt(var + (double) eps));
	saved_mean[g] = mean;
	saved_inv_sd[g] = inv_sd;
	if (running_init) {  // BatchNormLayer.hpp:234-238
		running_mean[g] = ((S) 1 - decay) * running_mean[g] + decay * mean;
		running_inv_sd[g] = ((S) 1 - decay) * running_inv_sd[g] + decay * inv_sd;
	} else {             // :239-242
		running_mean[g] = mean;
		running_inv_sd[g] = inv_sd;
	}
}

template<typename S>
__global__ void __launch_bounds__(256) bn_apply_kernel(long long L, long long chunk, int vec, const S* __restrict__ x,
		const S* __restrict__ mean, const S* __restrict__ inv_sd, const S* __restrict__ gamma,
		const S* __restrict__ beta, S* __restrict__ y) {
	typedef typename V16<S>::type V;
	const long long g = blockIdx.x;
	const S mu = mean[g], sc = inv_sd[g], gm = gamma[g], bt = beta[g];
	const long long lo = (long long) blockIdx.y * chunk;
	const long long hi = lo + chunk < L ? lo + chunk : L;
	const S* xg = x + g * L;
	S* yg = y + g * L;
	BN_FOREACH(S, vec, lo, hi,
		V v = *reinterpret_cast<const V*>(xg + i);
		S* e = reinterpret_cast<S*>(&v);
		_Pragma("unroll")
		for (int k = 0; k < G_; ++k) e[k] = ((e[k] - mu) * sc) * gm + bt;
		*reinterpret_cast<V*>(yg + i) = v;,
		yg[i] = ((xg[i] - mu) * sc) * gm + bt;)
}

// Statistics from the column sums a kernel layer's epilogue produced (cattl3_epilogue::col_stats): S1 = sum (x - shift),
// S2 = sum (x - shift)^2 per group.  Same outputs as bn_stats_final_kernel.
template<typename S>
__global__ void __launch_bounds__(128) bn_stats_from_sums_kernel(long long groups, long long L_local,
		const double* __restrict__ global_count, const double* __restrict__ col_stats, const S* __restrict__ shift,
		S eps, S decay, int running_init,
		S* __restrict__ running_mean, S* __restrict__ running_inv_sd, S* __restrict__ saved_mean,
		S* __restrict__ saved_inv_sd) {
	const long long g = blockIdx.x * 128ll + threadIdx.x;
	if (g >= groups) return;
	const double L = global_count ? *global_count : (double) L_local;   // elements behind the sums (all ranks)
	const double m1 = col_stats[g] / L;
	double var = col_stats[groups + g] / L - m1 * m1;
	var = var < 0 ? 0 : var;
	const S mean = (S) ((double) shift[g] + m1);
	const S inv_sd = (S) (1.0 / sqrt(var + (double) eps));
	saved_mean[g] = mean;
	saved_inv_sd[g] = inv_sd;
	if (running_init) {
		running_mean[g] = ((S) 1 - decay) * running_mean[g] + decay * mean;
		running_inv_sd[g] = ((S) 1 - decay) * running_inv_sd[g] + decay * inv_sd;
	} else {
		running_mean[g] = mean;
		running_inv_sd[g] = inv_sd;
	}
}

// bn_apply_kernel with the following element-wise activation layer applied in the same pass: y (the activation's
// cached input; may be null) and act_out = f(y).
template<typename S>
__global__ void __launch_bounds__(256) bn_apply_act_kernel(long long L, long long chunk, int vec, const S* __restrict__ x,
		const S* __restrict__ mean, const S* __restrict__ inv_sd, const S* __restrict__ gamma,
		const S* __restrict__ beta, S* __restrict__ y, int act_kind, S act_param, S* __restrict__ act_out) {
	typedef typename V16<S>::type V;
	const long long g = blockIdx.x;
	const S mu = mean[g], sc = inv_sd[g], gm = gamma[g], bt = beta[g];
	const long long lo = (long long) blockIdx.y * chunk;
	const long long hi = lo + chunk < L ? lo + chunk : L;
	const S* xg = x + g * L;
	S* yg = y ? y + g * L : nullptr;
	S* ag = act_out + g * L;
	BN_FOREACH(S, vec, lo, hi,
		V v = *reinterpret_cast<const V*>(xg + i);
		S* e = reinterpret_cast<S*>(&v);
		_Pragma("unroll")
		for (int k = 0; k < G_; ++k) e[k] = ((e[k] - mu) * sc) * gm + bt;
		if (yg) *reinterpret_cast<V*>(yg + i) = v;
		_Pragma("unroll")
		for (int k = 0; k < G_; ++k) e[k] = act_fwd_rt<S>(act_kind, e[k], act_param);
		*reinterpret_cast<V*>(ag + i) = v;,
		const S t = ((xg[i] - mu) * sc) * gm + bt;
		if (yg) yg[i] = t;
		ag[i] = act_fwd_rt<S>(act_kind, t, act_param);)
}

template<typename S>
__global__ void __launch_bounds__(256) bn_bwd_partial_kernel(long long L, long long chunk, int vec,
		const S* __restrict__ x, const S* __restrict__ mean, const S* __restrict__ inv_sd,
		const S* __restrict__ dy, BnPartial* __restrict__ part) {
	typedef typename V16<S>::type V;
	const long long g = blockIdx.x;
	const S mu = mean[g], sc = inv_sd[g];
	const long long lo = (long long) blockIdx.y * chunk;
	const long long hi = lo + chunk < L ? lo + chunk : L;
	const S* xg = x + g * L;
	const S* dyg = dy + g * L;
	double s1 = 0, s2 = 0;  // sum dy, sum dy * xhat
	BN_FOREACH(S, vec, lo, hi,
		const V vx = *reinterpret_cast<const V*>(xg + i);
		const V vg = *reinterpret_cast<const V*>(dyg + i);
		const S* ex = reinterpret_cast<const S*>(&vx);
		const S* eg = reinterpret_cast<const S*>(&vg);
		_Pragma("unroll")
		for (int k = 0; k < G_; ++k) { s1 += (double) eg[k]; s2 += (double) (eg[k] * ((ex[k] - mu) * sc)); },
		const S gy = dyg[i]; s1 += (double) gy; s2 += (double) (gy * ((xg[i] - mu) * sc));)
	block_reduce2(s1, s2);
	if (threadIdx.x == 0) part[g * gridDim.y + blockIdx.y] = BnPartial{ s1, s2 };
}

// sums[g] = (sum dy, sum dy*xhat); dgamma/dbeta accumulate (BatchNormLayer.hpp:252-254).
template<typename S>
__global__ void __launch_bounds__(128) bn_bwd_final_kernel(long long groups, int chunks,
		const BnPartial* __restrict__ part, S* __restrict__ dgamma, S* __restrict__ dbeta, double* __restrict__ sums) {
	const long long g = blockIdx.x * 128ll + threadIdx.x;
	if (g >= groups) return;
	double s1 = 0, s2 = 0;
	for (int c = 0; c < chunks; ++c) { s1 += part[g * chunks + c].s1; s2 += part[g * chunks + c].s2; }
	sums[g] = s1;
	sums[groups + g] = s2;
	dbeta[g] += (S) s1;
	dgamma[g] += (S) s2;
}

// dx = (L*g - sum g - xhat * sum(xhat g)) * inv_sd / L with g = gamma * dy (BatchNormLayer.hpp:257-261).
// `sums` = (sum dy, sum dy*xhat) per group over ALL `Ltot` elements of the statistic (the local L, or the whole
// data-parallel batch when the sums were all-reduced).
template<typename S>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(long long L, const double* __restrict__ global_count, long long chunk, int vec,
		const S* __restrict__ x, const S* __restrict__ mean, const S* __restrict__ inv_sd,
		const S* __restrict__ gamma, const double* __restrict__ sums, const S* __restrict__ dy,
		S* __restrict__ dx) {
	typedef typename V16<S>::type V;
	const long long g = blockIdx.x;
	const S mu = mean[g], sc = inv_sd[g], gm = gamma[g];
	const S sum_g = gm * (S) sums[g], sum_xg = gm * (S) sums[gridDim.x + g];
	const S Ltot = global_count ? (S) *global_count : (S) L;
	const S scale = ((S) 1 / Ltot) * sc;
	const long long lo = (long long) blockIdx.y * chunk;
	const long long hi = lo + chunk < L ? lo + chunk : L;
	const S* xg = x + g * L;
	const S* dyg = dy + g * L;
	S* dxg = dx + g * L;
	BN_FOREACH(S, vec, lo, hi,
		const V vx = *reinterpret_cast<const V*>(xg + i);
		V vg = *reinterpret_cast<const V*>(dyg + i);
		const S* ex = reinterpret_cast<const S*>(&vx);
		S* eg = reinterpret_cast<S*>(&vg);
		_Pragma("unroll")
		for (int k = 0; k < G_; ++k) {
			const S xh = (ex[k] - mu) * sc;
			eg[k] = ((Ltot * (gm * eg[k]) - sum_g) - xh * sum_xg) * scale;
		}
		*reinterpret_cast<V*>(dxg + i) = vg;,
		const S xh = (xg[i] - mu) * sc;
		dxg[i] = ((Ltot * (gm * dyg[i]) - sum_g) - xh * sum_xg) * scale;)
}

static void bn_partition(const cattl3_ctx* ctx, long long groups, long long L, int* chunks, long long* chunk,
		int* threads) {
	// aim for ~8 CTAs per SM in total; chunks of at least 2048 elements
	long long want = ceil_div(8ll * ctx->sm_count, groups);
	long long maxc = ceil_div(L, 2048);
	long long c = want < maxc ? want : maxc;
	if (c < 1) c = 1;
	*chunk = ceil_div(ceil_div(L, c), 4) * 4;   // chunk starts stay 16-byte aligned for the vector loops
	*chunks = (int) ceil_div(L, *chunk);
	*threads = L >= 1024 ? 256 : (L >= 128 ? 128 : 32);
}

template<typename S>
int batchnorm_forward(cattl3_ctx* ctx, int per_channel, int n, int h, int w, int c, int training,
		int running_init, S decay, S eps, const S* x, const S* gamma, const S* beta, S* running_mean,
		S* running_inv_sd, S* saved_mean, S* saved_inv_sd, S* y) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && x && y && gamma && beta && running_mean &&
			running_inv_sd, "batchnorm_forward: bad arguments");
	const long long groups = per_channel ? c : (long long) h * w * c;
	const long long L = per_channel ? (long long) n * h * w : n;
	int chunks, threads;
	long long chunk;
	bn_partition(ctx, groups, L, &chunks, &chunk, &threads);
	CATTL3_REQUIRE(groups <= 2147483647ll, "batchnorm: too many groups");
	dim3 grid((unsigned) groups, (unsigned) chunks);
	const int vec = L % V16<S>::G == 0 && aligned16(x) && aligned16(y);
	if (training) {
		CATTL3_REQUIRE(saved_mean && saved_inv_sd, "batchnorm_forward: training needs saved statistics");
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, sizeof(BnPartial) * groups * chunks));
		bn_stats_partial_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, chunk, vec, x, (BnPartial*) ctx->ws);
		CATTL3_LAUNCHED(ctx);
		bn_stats_final_kernel<S><<<(unsigned) ceil_div(groups, 128), 128, 0, ctx->stream>>>(groups, L, chunks, x,
				(const BnPartial*) ctx->ws, eps, decay, running_init, running_mean, running_inv_sd, saved_mean,
				saved_inv_sd);
		CATTL3_LAUNCHED(ctx);
		bn_apply_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, chunk, vec, x, saved_mean, saved_inv_sd, gamma, beta, y);
	} else {
		bn_apply_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, chunk, vec, x, running_mean, running_inv_sd, gamma, beta, y);
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

template<typename S>
int batchnorm_forward_stats(cattl3_ctx* ctx, int per_channel, int n, int h, int w, int c, int running_init, S decay,
		S eps, const S* x, const double* col_stats, const double* global_count, const S* shift, const S* gamma,
		const S* beta, S* running_mean, S* running_inv_sd, S* saved_mean, S* saved_inv_sd, S* y, int act_kind,
		S act_param, S* act_out) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && x && col_stats && shift && gamma && beta && running_mean &&
			running_inv_sd && saved_mean && saved_inv_sd, "batchnorm_forward_stats: bad arguments");
	CATTL3_REQUIRE(act_kind >= CATTL3_ACT_NONE && act_kind < CATTL3_ACT_SOFTMAX, "batchnorm_forward_stats: activation kind %d "
			"cannot be fused", act_kind);
	CATTL3_REQUIRE(act_kind == CATTL3_ACT_NONE ? y != nullptr : act_out != nullptr, "batchnorm_forward_stats: no output tensor");
	const long long groups = per_channel ? c : (long long) h * w * c;
	const long long L = per_channel ? (long long) n * h * w : n;
	CATTL3_REQUIRE(groups <= 2147483647ll, "batchnorm: too many groups");
	int chunks, threads;
	long long chunk;
	bn_partition(ctx, groups, L, &chunks, &chunk, &threads);
	dim3 grid((unsigned) groups, (unsigned) chunks);
	bn_stats_from_sums_kernel<S><<<(unsigned) ceil_div(groups, 128), 128, 0, ctx->stream>>>(groups, L, global_count, col_stats,
			shift, eps, decay, running_init, running_mean, running_inv_sd, saved_mean, saved_inv_sd);
	CATTL3_LAUNCHED(ctx);
	const int vec = L % V16<S>::G == 0 && aligned16(x) && (!y || aligned16(y)) && (!act_out || aligned16(act_out));
	if (act_kind == CATTL3_ACT_NONE)
		bn_apply_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, chunk, vec, x, saved_mean, saved_inv_sd, gamma, beta, y);
	else
		bn_apply_act_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, chunk, vec, x, saved_mean, saved_inv_sd, gamma, beta, y,
				act_kind, act_param, act_out);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// The backward pass in two halves (they are one call for a single process): `sums` = (sum dy, sum dy*xhat) per group,
// with the LOCAL sums accumulated into dgamma / dbeta; then dx from sums that may have been all-reduced over the
// data-parallel ranks (synchronised batch statistics), `total_count` = elements per group over all ranks.
template<typename S>
int batchnorm_backward_sums(cattl3_ctx* ctx, int per_channel, int n, int h, int w, int c, const S* x,
		const S* saved_mean, const S* saved_inv_sd, const S* dy, S* dgamma, S* dbeta, double* sums) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && x && saved_mean && saved_inv_sd && dy && dgamma && dbeta && sums,
			"batchnorm_backward_sums: bad arguments");
	const long long groups = per_channel ? c : (long long) h * w * c;
	const long long L = per_channel ? (long long) n * h * w : n;
	CATTL3_REQUIRE(groups <= 2147483647ll, "batchnorm: too many groups");
	int chunks, threads;
	long long chunk;
	bn_partition(ctx, groups, L, &chunks, &chunk, &threads);
	dim3 grid((unsigned) groups, (unsigned) chunks);
	const int vec = L % V16<S>::G == 0 && aligned16(x) && aligned16(dy);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, sizeof(BnPartial) * groups * chunks));
	bn_bwd_partial_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, chunk, vec, x, saved_mean, saved_inv_sd, dy,
			(BnPartial*) ctx->ws);
	CATTL3_LAUNCHED(ctx);
	bn_bwd_final_kernel<S><<<(unsigned) ceil_div(groups, 128), 128, 0, ctx->stream>>>(groups, chunks,
			(const BnPartial*) ctx->ws, dgamma, dbeta, sums);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

template<typename S>
int batchnorm_backward_apply(cattl3_ctx* ctx, int per_channel, int n, int h, int w, int c, const double* global_count,
		const S* x, const S* gamma, const S* saved_mean, const S* saved_inv_sd, const S* dy, const double* sums, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && x && gamma && saved_mean && saved_inv_sd && dy && sums && dx,
			"batchnorm_backward_apply: bad arguments");
	const long long groups = per_channel ? c : (long long) h * w * c;
	const long long L = per_channel ? (long long) n * h * w : n;
	CATTL3_REQUIRE(groups <= 2147483647ll, "batchnorm: too many groups");
	int chunks, threads;
	long long chunk;
	bn_partition(ctx, groups, L, &chunks, &chunk, &threads);
	dim3 grid((unsigned) groups, (unsigned) chunks);
	const int vec = L % V16<S>::G == 0 && aligned16(x) && aligned16(dy) && aligned16(dx);
	bn_bwd_apply_kernel<S><<<grid, threads, 0, ctx->stream>>>(L, global_count, chunk, vec, x, saved_mean, saved_inv_sd,
			gamma, sums, dy, dx);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

template<typename S>
int batchnorm_backward(cattl3_ctx* ctx, int per_channel, int n, int h, int w, int c, const S* x,
		const S* gamma, const S* saved_mean, const S* saved_inv_sd, const S* dy, S* dgamma, S* dbeta, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && gamma, "batchnorm_backward: bad arguments");
	const long long groups = per_channel ? c : (long long) h * w * c;
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->stat_ws, &ctx->stat_ws_bytes, sizeof(double) * 2 * groups));
	CATTL3_CHECK(batchnorm_backward_sums<S>(ctx, per_channel, n, h, w, c, x, saved_mean, saved_inv_sd, dy, dgamma, dbeta,
			(double*) ctx->stat_ws));
	if (!dx)
		return CATTL3_OK;
	return batchnorm_backward_apply<S>(ctx, per_channel, n, h, w, c, (const double*) nullptr, x, gamma, saved_mean,
			saved_inv_sd, dy, (const double*) ctx->stat_ws, dx);
}

// Per-group shifted sums of a batch-norm input (the first reduction of the training forward pass) on their own: what a
// kernel layer's fused epilogue would have produced (cattl3_epilogue::col_stats), for inputs that come from
// elsewhere and for synchronised statistics (all-reduce them, then cattl3_batchnorm_forward_stats).
template<typename S>
int batchnorm_stats(cattl3_ctx* ctx, int per_channel, int n, int h, int w, int c, const S* x, const S* shift,
		double* col_stats) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && x && shift && col_stats, "batchnorm_stats: bad arguments");
	const long long groups = per_channel ? c : (long long) h * w * c;
	const long long L = per_channel ? (long long) n * h * w : n;
	return colstats_shifted<S>(ctx, L, groups, x, shift, col_stats);
}

// ---- fused optimizer step ------------------------------------------------------------------------
// One launch over the whole parameter arena: L2 regularisation (g += lambda p), the update rule and
// the gradient reset, i.e. SGDOptimizer::_train lines 57-70 collapsed.  Expression order follows the
// reference's Eigen expressions so that float results agree to rounding.
template<typename S>
struct OptScalars { S lr, a, b, eps, lr_epoch, c1, c1n, c2, l2; int reset; };

// One element of the update: reads (p, g, state), writes them back in place.
template<typename S, int KIND>
__device__ __forceinline__ void opt_update(const OptScalars<S>& h, S& pv, S& gv, S& a1, S& a2, S& a3) {
	if (h.l2 > (S) 0) gv += pv * h.l2;
	if (KIND == CATTL3_OPT_VANILLA_SGD) {
		pv = pv - gv * h.lr;
	} else if (KIND == CATTL3_OPT_MOMENTUM) {
		const S v = a1 * h.b + gv * h.lr_epoch;
		a1 = v;
		pv = pv - v;
	} else if (KIND == CATTL3_OPT_NESTEROV) {
		const S old = a1;
		const S v = old * h.b - gv * h.lr_epoch;
		a1 = v;
		pv = pv + old * -h.b + v * ((S) 1 + h.b);
	} else if (KIND == CATTL3_OPT_ADAGRAD) {
		const S s = a1 + gv * gv;
		a1 = s;
		pv = pv - gv * h.lr / (dev_sqrt<S>(s) + h.eps);
	} else if (KIND == CATTL3_OPT_RMSPROP) {
		const S s = a1 * ((S) 1 - h.b) + gv * gv * h.b;
		a1 = s;
		pv = pv - gv * h.lr / (dev_sqrt<S>(s) + h.eps);
	} else if (KIND == CATTL3_OPT_ADADELTA) {
		const S s = a1 * ((S) 1 - h.a) + gv * gv * h.a;
		a1 = s;
		const S u = -gv * dev_sqrt<S>(a2 + h.eps) / dev_sqrt<S>(s + h.eps);
		pv = pv + u;
		a2 = a2 * ((S) 1 - h.a) + u * u * h.a;
	} else if (KIND == CATTL3_OPT_ADAM) {
		const S m = a1 * ((S) 1 - h.a) + gv * h.a;
		const S v = a2 * ((S) 1 - h.b) + gv * gv * h.b;
		a1 = m; a2 = v;
		pv = pv - (m * (h.lr * h.c1)) / dev_sqrt<S>(v * h.c2 + h.eps);
	} else if (KIND == CATTL3_OPT_ADAMAX) {
		const S m = a1 * ((S) 1 - h.a) + gv * h.a;
		const S d = a2 * ((S) 1 - h.b);
		const S ag = gv < (S) 0 ? -gv : gv;
		const S v = d > ag ? d : ag;
		a1 = m; a2 = v;
		pv = pv - (m * (h.lr * h.c1)) / (v + h.eps);
	} else if (KIND == CATTL3_OPT_NADAM) {
		const S m = a1 * ((S) 1 - h.a) + gv * h.a;
		const S v = a2 * ((S) 1 - h.b) + gv * gv * h.b;
		a1 = m; a2 = v;
		pv = pv - (gv * (h.a * h.c1) + m * (((S) 1 - h.a) * h.c1n)) * h.lr / dev_sqrt<S>(v * h.c2 + h.eps);
	} else if (KIND == CATTL3_OPT_AMSGRAD) {
		const S m = a1 * ((S) 1 - h.a) + gv * h.a;
		const S v = a2 * ((S) 1 - h.b) + gv * gv * h.b;
		const S mx = v > a3 ? v : a3;
		a1 = m; a2 = v; a3 = mx;
		pv = pv - m * h.lr / dev_sqrt<S>(mx + h.eps);
	}
}

// 16-byte vector body (every stream of the arena is 16-byte aligned at the same element offsets) + scalar tail.
template<typename S, int KIND>
__global__ void __launch_bounds__(256) opt_step_kernel(long long count, int vec_ok, OptScalars<S> h_value,
		const cattl3_opt_step* __restrict__ dev_step, S* __restrict__ p, S* __restrict__ g, S* __restrict__ s1,
		S* __restrict__ s2, S* __restrict__ s3) {
	typedef typename V16<S>::type V;
	// dev_step != null: the step scalars live in device memory (a captured step graph replays this launch with new
	// scalars every step: cattl3_optimizer_step_indirect)
	OptScalars<S> h = h_value;
	if (dev_step) {
		h.lr = (S) dev_step->lr; h.a = (S) dev_step->a; h.b = (S) dev_step->b; h.eps = (S) dev_step->eps;
		h.lr_epoch = (S) dev_step->lr_epoch; h.c1 = (S) dev_step->c1; h.c1n = (S) dev_step->c1n; h.c2 = (S) dev_step->c2;
		h.l2 = (S) dev_step->l2_lambda; h.reset = dev_step->reset_grad;
	}
	constexpr int G = V16<S>::G;
	constexpr int NS = KIND == CATTL3_OPT_VANILLA_SGD ? 0 : (KIND <= CATTL3_OPT_RMSPROP ? 1 : (KIND == CATTL3_OPT_AMSGRAD ? 3 : 2));
	const long long nvec = vec_ok ? count / G : 0;
	const long long stride = (long long) gridDim.x * 256;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += stride) {
		V vp = reinterpret_cast<V*>(p)[i], vg = reinterpret_cast<V*>(g)[i], v1, v2, v3;
		if (NS >= 1) v1 = reinterpret_cast<V*>(s1)[i];
		if (NS >= 2) v2 = reinterpret_cast<V*>(s2)[i];
		if (NS >= 3) v3 = reinterpret_cast<V*>(s3)[i];
		S* ep = reinterpret_cast<S*>(&vp); S* eg = reinterpret_cast<S*>(&vg);
		S* e1 = reinterpret_cast<S*>(&v1); S* e2 = reinterpret_cast<S*>(&v2); S* e3 = reinterpret_cast<S*>(&v3);
		#pragma unroll
		for (int k = 0; k < G; ++k) {
			S d1 = NS >= 1 ? e1[k] : (S) 0, d2 = NS >= 2 ? e2[k] : (S) 0, d3 = NS >= 3 ? e3[k] : (S) 0;
			opt_update<S, KIND>(h, ep[k], eg[k], d1, d2, d3);
			if (NS >= 1) e1[k] = d1;
			if (NS >= 2) e2[k] = d2;
			if (NS >= 3) e3[k] = d3;
			eg[k] = (S) 0;
		}
		reinterpret_cast<V*>(p)[i] = vp;
		if (NS >= 1) reinterpret_cast<V*>(s1)[i] = v1;
		if (NS >= 2) reinterpret_cast<V*>(s2)[i] = v2;
		if (NS >= 3) reinterpret_cast<V*>(s3)[i] = v3;
		if (h.reset) reinterpret_cast<V*>(g)[i] = vg;
	}
	for (long long i = nvec * G + blockIdx.x * 256ll + threadIdx.x; i < count; i += stride) {
		S pv = p[i], gv = g[i];
		S d1 = NS >= 1 ? s1[i] : (S) 0, d2 = NS >= 2 ? s2[i] : (S) 0, d3 = NS >= 3 ? s3[i] : (S) 0;
		opt_update<S, KIND>(h, pv, gv, d1, d2, d3);
		p[i] = pv;
		if (NS >= 1) s1[i] = d1;
		if (NS >= 2) s2[i] = d2;
		if (NS >= 3) s3[i] = d3;
		if (h.reset) g[i] = (S) 0;
	}
}

// st: host scalars (dev_step == null), or only `kind` is taken from the host and the scalars from dev_step.
template<typename S>
int optimizer_step(cattl3_ctx* ctx, const cattl3_opt_step* st, int64_t count, S* p, S* g, S* s1, S* s2, S* s3,
		const cattl3_opt_step* dev_step = nullptr) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(st && count > 0 && p && g, "optimizer_step: bad arguments");
	OptScalars<S> h{ (S) st->lr, (S) st->a, (S) st->b, (S) st->eps, (S) st->lr_epoch, (S) st->c1, (S) st->c1n,
			(S) st->c2, (S) st->l2_lambda, st->reset_grad };
	const int grid = ew_grid(ctx, ceil_div(count, V16<S>::G), 256);
	const int k = st->kind;
	const int need = (k == 0) ? 0 : (k <= 4 ? 1 : (k == 9 ? 3 : 2));
	CATTL3_REQUIRE((need < 1 || s1) && (need < 2 || s2) && (need < 3 || s3), "optimizer_step: missing state vector");
	const int vec_ok = aligned16(p) && aligned16(g) && (need < 1 || aligned16(s1)) && (need < 2 || aligned16(s2)) &&
			(need < 3 || aligned16(s3));
	switch (k) {
#define CASE(K) case K: opt_step_kernel<S, K><<<grid, 256, 0, ctx->stream>>>(count, vec_ok, h, dev_step, p, g, s1, s2, s3); break;
		CASE(CATTL3_OPT_VANILLA_SGD) CASE(CATTL3_OPT_MOMENTUM) CASE(CATTL3_OPT_NESTEROV) CASE(CATTL3_OPT_ADAGRAD)
		CASE(CATTL3_OPT_RMSPROP) CASE(CATTL3_OPT_ADADELTA) CASE(CATTL3_OPT_ADAM) CASE(CATTL3_OPT_ADAMAX)
		CASE(CATTL3_OPT_NADAM) CASE(CATTL3_OPT_AMSGRAD)
#undef CASE
		default:
			set_error("optimizer_step: unknown kind %d", k);
			return CATTL3_ERR_INVALID;
	}
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- glue ----------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// y = OP(x, y) over 16-byte vectors with a scalar tail; OP 0: y + x, 1: x * alpha, 2: fma(alpha, x, y), 3: y * x
template<typename S, int OP>
__global__ void __launch_bounds__(256) glue_kernel(long long count, int vec_ok, S alpha, const S* __restrict__ x, S* __restrict__ y) {
	typedef typename V16<S>::type V;
	constexpr int G = V16<S>::G;
	const long long nvec = vec_ok ? count / G : 0;
	const long long stride = (long long) gridDim.x * 256;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += stride) {
		V vx = reinterpret_cast<const V*>(x)[i], vy;
		if (OP != 1) vy = reinterpret_cast<V*>(y)[i];
		S* ex = reinterpret_cast<S*>(&vx); S* ey = reinterpret_cast<S*>(&vy);
		#pragma unroll
		for (int k = 0; k < G; ++k) ey[k] = OP == 0 ? ey[k] + ex[k] : (OP == 1 ? ex[k] * alpha : (OP == 2 ? fma(alpha, ex[k], ey[k]) : ey[k] * ex[k]));
		reinterpret_cast<V*>(y)[i] = vy;
	}
	for (long long i = nvec * G + blockIdx.x * 256ll + threadIdx.x; i < count; i += stride)
		y[i] = OP == 0 ? y[i] + x[i] : (OP == 1 ? x[i] * alpha : (OP == 2 ? fma(alpha, x[i], y[i]) : y[i] * x[i]));
}
template<typename S, int OP>
static int glue(cattl3_ctx* ctx, const char* what, int64_t count, S alpha, const S* x, S* y) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && y && x, "%s: bad arguments", what);
	glue_kernel<S, OP><<<ew_grid(ctx, ceil_div(count, V16<S>::G), 256), 256, 0, ctx->stream>>>(count,
			aligned16(x) && aligned16(y), alpha, x, y);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template<typename S> int axpy(cattl3_ctx* ctx, int64_t count, S alpha, const S* x, S* y) { return glue<S, 2>(ctx, "axpy", count, alpha, x, y); }
template<typename S> int add_inplace(cattl3_ctx* ctx, int64_t count, S* y, const S* x) { return glue<S, 0>(ctx, "add_inplace", count, (S) 0, x, y); }
template<typename S> int scale(cattl3_ctx* ctx, int64_t count, S alpha, const S* x, S* y) { return glue<S, 1>(ctx, "scale", count, alpha, x, y); }
template<typename S> int mul_inplace(cattl3_ctx* ctx, int64_t count, S* y, const S* x) { return glue<S, 3>(ctx, "mul_inplace", count, (S) 0, x, y); }

// out = (accumulate ? out : 0) + a * b + (c ? c * d : 0): the gate arithmetic of an LSTM cell (state = forget * previous +
// write * candidate, hidden = read * activated state, and the products of its backward pass) in one pass over 16-byte
// vectors.  Products and sums are rounded separately (no contraction), as Eigen's expression `a * b + c * d` is.
template<typename S>
__global__ void __launch_bounds__(256) muladd_kernel(long long count, int vec_ok, int accumulate, const S* __restrict__ a,
		const S* __restrict__ b, const S* __restrict__ c, const S* __restrict__ d, S* out) {
	typedef typename V16<S>::type V;
	constexpr int G = V16<S>::G;
	const long long nvec = vec_ok ? count / G : 0;
	const long long stride = (long long) gridDim.x * 256;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < nvec; i += stride) {
		V va = reinterpret_cast<const V*>(a)[i], vb = reinterpret_cast<const V*>(b)[i], vc, vd, vo;
		if (c) { vc = reinterpret_cast<const V*>(c)[i]; vd = reinterpret_cast<const V*>(d)[i]; }
		if (accumulate) vo = reinterpret_cast<const V*>(out)[i];
		S* ea = reinterpret_cast<S*>(&va); S* eb = reinterpret_cast<S*>(&vb); S* ec = reinterpret_cast<S*>(&vc);
		S* ed = reinterpret_cast<S*>(&vd); S* eo = reinterpret_cast<S*>(&vo);
		#pragma unroll
		for (int k = 0; k < G; ++k) {
			S r = mul_rn(ea[k], eb[k]);
			if (c) r = add_rn(r, mul_rn(ec[k], ed[k]));
			eo[k] = accumulate ? add_rn(eo[k], r) : r;
		}
		reinterpret_cast<V*>(out)[i] = vo;
	}
	for (long long i = nvec * G + blockIdx.x * 256ll + threadIdx.x; i < count; i += stride) {
		S r = mul_rn(a[i], b[i]);
		if (c) r = add_rn(r, mul_rn(c[i], d[i]));
		out[i] = accumulate ? add_rn(out[i], r) : r;
	}
}
template<typename S>
static int muladd(cattl3_ctx* ctx, int64_t count, int accumulate, const S* a, const S* b, const S* c, const S* d, S* out) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && a && b && out && ((c == nullptr) == (d == nullptr)), "muladd: bad arguments");
	const int vec_ok = aligned16(a) && aligned16(b) && aligned16(out) && (!c || (aligned16(c) && aligned16(d)));
	muladd_kernel<S><<<ew_grid(ctx, ceil_div(count, V16<S>::G), 256), 256, 0, ctx->stream>>>(count, vec_ok, accumulate, a, b, c,
			d, out);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// Parameter regularisation on the device (L1 / L2 / ElasticNet: C-ATTL3/parameter_regularization/*.hpp):
//   grad[i] += (v >= 0 ? l1 : -l1) + l2 * v      (Parameters::regularize, StandardParameters.hpp:133-136)
//   penalty += l1 * sum |v| + l2 / 2 * sum v^2   (get_regularization_penalty; accumulated in double on the device, so that
//                                                 the batch loop needs no host read of the parameters per step)
// Per-block partial penalties go to scratch; the last block to finish (a counter) adds them in block order: deterministic.
template<typename S>
__global__ void __launch_bounds__(256) regularize_kernel(long long count, S l1, S l2, const S* __restrict__ v, S* __restrict__ g,
		double* __restrict__ partial, unsigned int* __restrict__ done, double* __restrict__ penalty) {
	__shared__ double red[256];
	__shared__ bool last;
	double abs_sum = 0, sq_sum = 0;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) {
		const S x = v[i];
		if (g) g[i] = add_rn(g[i], add_rn(x >= (S) 0 ? l1 : -l1, mul_rn(x, l2)));
		abs_sum += fabs((double) x);
		sq_sum += (double) x * (double) x;
	}
	if (!penalty) return;
	red[threadIdx.x] = (double) l1 * abs_sum + 0.5 * (double) l2 * sq_sum;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int) threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		partial[blockIdx.x] = red[0];
		__threadfence();
		last = atomicAdd(done, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (last && threadIdx.x == 0) {
		__threadfence();
		double s = 0;
		for (unsigned int b = 0; b < gridDim.x; ++b) s += ((volatile double*) partial)[b];
		*penalty += s;
		*done = 0;   // ready for the next launch
	}
}
template<typename S>
static int regularize(cattl3_ctx* ctx, int64_t count, S l1, S l2, const S* values, S* grad, double* penalty) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && values && (grad || penalty), "regularize: bad arguments");
	int grid = ew_grid(ctx, count, 1024);
	if (grid > 256) grid = 256;
	// scratch: 256 partials + the counter, private to this kernel (zeroed once: the kernel leaves the counter at 0)
	if (!ctx->reg_ws) {
		CATTL3_REQUIRE(!ctx->capturing, "regularize: first use during graph capture (run the step eagerly first)");
		CATTL3_CUDA(cudaMalloc(&ctx->reg_ws, 257 * sizeof(double)));
		CATTL3_CUDA(cudaMemsetAsync(ctx->reg_ws, 0, 257 * sizeof(double), ctx->stream));
	}
	regularize_kernel<S><<<grid, 256, 0, ctx->stream>>>(count, l1, l2, values, grad, (double*) ctx->reg_ws,
			(unsigned int*) ((double*) ctx->reg_ws + 256), penalty);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- value / gradient constraints (C-ATTL3/parameters/StandardParameters.hpp:150-182) ------------------------------
// In the reference's order and with its definitions: clip every element to [-clip, clip]; the "L1" limit is compared with
// the FROBENIUS norm and rescales by max / norm; the "L2" limit is compared with the SQUARED norm of the result and
// rescales by max / squared norm.  The second rescaling's squared norm is f1^2 times the first's, so one reduction
// (of the clipped elements, in double, per-block partials added in block order by the last block: deterministic)
// decides both factors; a second pass applies their product.  0 switches a limit off.
template<typename S>
__global__ void __launch_bounds__(256) constrain_reduce_kernel(long long count, S clip, S max_l1, S max_l2, S* __restrict__ x,
		double* __restrict__ partial, unsigned int* __restrict__ done, S* __restrict__ factor) {
	__shared__ double red[256];
	__shared__ bool last;
	double sq_sum = 0;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) {
		S v = x[i];
		if (clip > (S) 0) {
			v = v > clip ? clip : (v < -clip ? -clip : v);
			x[i] = v;
		}
		sq_sum += (double) v * (double) v;
	}
	red[threadIdx.x] = sq_sum;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int) threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		partial[blockIdx.x] = red[0];
		__threadfence();
		last = atomicAdd(done, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (last && threadIdx.x == 0) {
		__threadfence();
		double sq = 0;
		for (unsigned int b = 0; b < gridDim.x; ++b) sq += ((volatile double*) partial)[b];
		S f = (S) 1;
		if (max_l1 > (S) 0) {
			const S norm = (S) sqrt(sq);
			if (norm > max_l1) {
				const S f1 = max_l1 / norm;
				f = f1;
				sq *= (double) f1 * (double) f1;
			}
		}
		if (max_l2 > (S) 0) {
			const S sq_norm = (S) sq;
			if (sq_norm > max_l2) f *= max_l2 / sq_norm;
		}
		*factor = f;
		*done = 0;   // ready for the next launch
	}
}
template<typename S>
__global__ void __launch_bounds__(256) constrain_scale_kernel(long long count, const S* __restrict__ factor, S* __restrict__ x) {
	const S f = *factor;
	if (f == (S) 1) return;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) x[i] = mul_rn(x[i], f);
}
template<typename S>
static int constrain(cattl3_ctx* ctx, int64_t count, S clip, S max_l1, S max_l2, S* x) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && x && clip >= (S) 0 && max_l1 >= (S) 0 && max_l2 >= (S) 0, "constrain: bad arguments");
	if (!(clip > (S) 0) && !(max_l1 > (S) 0) && !(max_l2 > (S) 0))
		return CATTL3_OK;
	int grid = ew_grid(ctx, count, 1024);
	if (grid > 256) grid = 256;
	// scratch shared with cattl3_regularize's layout: 256 partials, a counter, then the factor (one stream: calls are ordered)
	if (!ctx->con_ws) {
		CATTL3_REQUIRE(!ctx->capturing, "constrain: first use during graph capture (run the step eagerly first)");
		CATTL3_CUDA(cudaMalloc(&ctx->con_ws, 258 * sizeof(double)));
		CATTL3_CUDA(cudaMemsetAsync(ctx->con_ws, 0, 258 * sizeof(double), ctx->stream));
	}
	double* ws = (double*) ctx->con_ws;
	if (max_l1 > (S) 0 || max_l2 > (S) 0) {
		constrain_reduce_kernel<S><<<grid, 256, 0, ctx->stream>>>(count, clip, max_l1, max_l2, x, ws, (unsigned int*) (ws + 256),
				(S*) (ws + 257));
		CATTL3_LAUNCHED(ctx);
		constrain_scale_kernel<S><<<ew_grid(ctx, count, 256), 256, 0, ctx->stream>>>(count, (const S*) (ws + 257), x);
		CATTL3_LAUNCHED(ctx);
	} else {
		// clip only: no reduction needed; the reduce kernel's first loop does the clipping, its tail is cheap
		constrain_reduce_kernel<S><<<grid, 256, 0, ctx->stream>>>(count, clip, (S) 0, (S) 0, x, ws, (unsigned int*) (ws + 256),
				(S*) (ws + 257));
		CATTL3_LAUNCHED(ctx);
	}
	return CATTL3_OK;
}

// y[i] = value: device-side constants (e.g. the element count that travels with synchronised batch-norm sums) without a
// host -> device copy, which from pageable memory would synchronise the host with the stream.
template<typename S>
__global__ void __launch_bounds__(256) fill_kernel(long long count, S value, S* __restrict__ y) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) y[i] = value;
}
template<typename S>
int fill(cattl3_ctx* ctx, int64_t count, S value, S* y) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && y, "fill: bad arguments");
	fill_kernel<S><<<ew_grid(ctx, count, 256), 256, 0, ctx->stream>>>(count, value, y);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- mini-batch rows out of a device-resident data set (MemoryDataProvider::get_data, :72-83) -------------------
// dst[n + rows * j] = src[first + n + total * j]: a contiguous run per column, 16-byte vectors where everything lines up.
template<typename S>
__global__ void __launch_bounds__(256) slice_rows_kernel(long long total, long long vol, long long first, long long rows,
		int vec, const S* __restrict__ src, S* __restrict__ dst) {
	typedef typename V16<S>::type V;
	constexpr int G = V16<S>::G;
	if (vec) {
		const long long rv = rows / G, count = rv * vol;
		for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) {
			const long long j = i / rv, n = (i - j * rv) * G;
			*reinterpret_cast<V*>(dst + n + rows * j) = *reinterpret_cast<const V*>(src + first + n + total * j);
		}
	} else {
		const long long count = rows * vol;
		for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) {
			const long long j = i / rows, n = i - j * rows;
			dst[n + rows * j] = src[first + n + total * j];
		}
	}
}

template<typename S>
int slice_rows(cattl3_ctx* ctx, int64_t total, int64_t vol, int64_t first, int64_t rows, const S* src, S* dst) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(total > 0 && vol > 0 && rows > 0 && first >= 0 && first + rows <= total && src && dst,
			"slice_rows: bad arguments");
	constexpr int G = V16<S>::G;
	const int vec = rows % G == 0 && total % G == 0 && first % G == 0 && aligned16(src) && aligned16(dst);
	slice_rows_kernel<S><<<ew_grid(ctx, rows * vol / (vec ? G : 1), 256), 256, 0, ctx->stream>>>(total, vol, first, rows, vec,
			src, dst);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- dropout (C-ATTL3/layer/DropoutLayer.hpp:74-94) ------------------------------------------------------
// Inverted dropout: mask = u <= p ? 0 : 1 / (1 - p + eps), y = x * mask, with u uniform in [0, 1) from a
// counter-based generator (a 64-bit mix of seed and element index: the same (seed, index) always gives the
// same draw, whatever the launch geometry).  The mask is kept as one byte per element for the backward pass.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}
__device__ __forceinline__ float uniform01(uint64_t seed, long long index) {
	const uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t) (index + 1));
	return (float) (uint32_t) (h >> 40) * (1.0f / 16777216.0f);
}

template<typename S>
__global__ void __launch_bounds__(256) dropout_fwd_kernel(long long count, float prob, S scale, uint64_t seed,
		const S* __restrict__ x, S* __restrict__ y, uint8_t* __restrict__ mask) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256) {
		const bool keep = uniform01(seed, i) > prob;
		mask[i] = keep ? 1 : 0;
		y[i] = keep ? x[i] * scale : (S) 0;
	}
}
template<typename S>
__global__ void __launch_bounds__(256) dropout_bwd_kernel(long long count, S scale, const S* __restrict__ dy,
		const uint8_t* __restrict__ mask, S* __restrict__ dx) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < count; i += (long long) gridDim.x * 256)
		dx[i] = mask[i] ? dy[i] * scale : (S) 0;
}

template<typename S>
int dropout_forward(cattl3_ctx* ctx, int64_t count, S prob, S eps, uint64_t seed, const S* x, S* y, uint8_t* mask) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && x && y && mask, "dropout_forward: bad arguments");
	CATTL3_REQUIRE(prob > (S) 0 && prob <= (S) 1 && eps > (S) 0, "dropout_forward: probability must be in (0, 1], epsilon > 0");
	const S scale = (S) 1 / ((S) 1 - prob + eps);
	dropout_fwd_kernel<S><<<ew_grid(ctx, count, 256), 256, 0, ctx->stream>>>(count, (float) prob, scale, seed, x, y, mask);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}
template<typename S>
int dropout_backward(cattl3_ctx* ctx, int64_t count, S prob, S eps, const S* dy, const uint8_t* mask, S* dx) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(count > 0 && dy && dx && mask, "dropout_backward: bad arguments");
	const S scale = (S) 1 / ((S) 1 - prob + eps);
	dropout_bwd_kernel<S><<<ew_grid(ctx, count, 256), 256, 0, ctx->stream>>>(count, scale, dy, mask, dx);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

// ---- losses (C-ATTL3/loss/SquaredLoss.hpp:25-34, CrossEntropyLoss.hpp:33-42, UniversalLoss.hpp:24-58) ----
// out, obj: rows x vol (rows = batch, fastest).  loss[r] = sum_j (out - obj)^2  |  -sum_j log(out + eps) * obj;
// grad = 2 (out - obj)  |  -obj / (out + eps), divided by grad_div (the batch loop's nominal batch size,
// SGDOptimizer.hpp:55-56).
// A CTA = 32 rows x 8 column lanes over one chunk of columns: loads and stores are coalesced along the rows, the per-row
// loss of the chunk is reduced over the column lanes in shared memory and written to partial[chunk][row]; the chunks
// are then added in order (deterministic).  One thread per row would leave a 512-row batch with 512 threads.
template<typename S, int KIND>
__global__ void __launch_bounds__(256) loss_kernel(long long rows, long long vol, long long chunk_cols, S eps, S grad_div,
		const S* __restrict__ out, const S* __restrict__ obj, S* __restrict__ partial, S* __restrict__ grad) {
	__shared__ S red[8][33];
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
	const long long r = blockIdx.x * 32ll + tx;
	const long long j0 = (long long) blockIdx.y * chunk_cols;
	const long long j1 = j0 + chunk_cols < vol ? j0 + chunk_cols : vol;
	S acc = 0;
	if (r < rows) {
		for (long long j = j0 + ty; j < j1; j += 8) {
			const S o = out[r + rows * j], t = obj[r + rows * j];
			if (KIND == CATTL3_LOSS_SQUARED) {
				const S d = o - t;
				acc += d * d;
				if (grad) grad[r + rows * j] = ((S) 2 * d) / grad_div;
			} else {
				acc += dev_log<S>(o + eps) * t;
				if (grad) grad[r + rows * j] = (-t / (o + eps)) / grad_div;
			}
		}
	}
	red[ty][tx] = acc;
	__syncthreads();
	if (ty == 0 && r < rows && partial) {
		S s = red[0][tx];
		#pragma unroll
		for (int k = 1; k < 8; ++k) s += red[k][tx];
		partial[(long long) blockIdx.y * rows + r] = KIND == CATTL3_LOSS_SQUARED ? s : -s;
	}
}

template<typename S>
__global__ void __launch_bounds__(256) loss_sum_kernel(long long rows, int chunks, const S* __restrict__ partial,
		S* __restrict__ loss) {
	const long long r = blockIdx.x * 256ll + threadIdx.x;
	if (r >= rows) return;
	S s = 0;
	for (int c = 0; c < chunks; ++c) s += partial[(long long) c * rows + r];
	loss[r] = s;
}

template<typename S>
int loss_forward_backward(cattl3_ctx* ctx, int kind, int64_t rows, int64_t vol, S eps, S grad_div, const S* out,
		const S* obj, S* loss, S* grad) {
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(rows > 0 && vol > 0 && out && obj && (loss || grad) && grad_div != (S) 0, "loss: bad arguments");
	CATTL3_REQUIRE(kind == CATTL3_LOSS_SQUARED || kind == CATTL3_LOSS_CROSS_ENTROPY, "loss: unknown kind %d", kind);
	const long long row_blocks = ceil_div(rows, 32);
	long long chunks = ceil_div(4ll * ctx->sm_count, row_blocks);
	const long long max_chunks = ceil_div(vol, 16);
	if (chunks > max_chunks) chunks = max_chunks;
	if (chunks < 1) chunks = 1;
	const long long chunk_cols = ceil_div(vol, chunks);
	chunks = ceil_div(vol, chunk_cols);
	S* partial = loss;
	if (loss && chunks > 1) {
		CATTL3_CHECK(ensure_buffer(ctx, &ctx->stat_ws, &ctx->stat_ws_bytes, sizeof(S) * (size_t) (chunks * rows)));
		partial = (S*) ctx->stat_ws;
	}
	dim3 grid((unsigned) row_blocks, (unsigned) chunks);
	if (kind == CATTL3_LOSS_SQUARED)
		loss_kernel<S, CATTL3_LOSS_SQUARED><<<grid, 256, 0, ctx->stream>>>(rows, vol, chunk_cols, eps, grad_div, out, obj, partial, grad);
	else
		loss_kernel<S, CATTL3_LOSS_CROSS_ENTROPY><<<grid, 256, 0, ctx->stream>>>(rows, vol, chunk_cols, eps, grad_div, out, obj, partial, grad);
	CATTL3_LAUNCHED(ctx);
	if (loss && chunks > 1) {
		loss_sum_kernel<S><<<(unsigned) ceil_div(rows, 256), 256, 0, ctx->stream>>>(rows, (int) chunks, partial, loss);
		CATTL3_LAUNCHED(ctx);
	}
	return CATTL3_OK;
}

} // namespace cattl3

using namespace cattl3;

extern "C" {

int cattl3_fill_f32(cattl3_ctx* c, int64_t count, float value, float* y) { return fill<float>(c, count, value, y); }
int cattl3_fill_f64(cattl3_ctx* c, int64_t count, double value, double* y) { return fill<double>(c, count, value, y); }
int cattl3_slice_rows_f32(cattl3_ctx* c, int64_t total, int64_t vol, int64_t first, int64_t rows, const float* src, float* dst) {
	return slice_rows<float>(c, total, vol, first, rows, src, dst); }
int cattl3_slice_rows_f64(cattl3_ctx* c, int64_t total, int64_t vol, int64_t first, int64_t rows, const double* src, double* dst) {
	return slice_rows<double>(c, total, vol, first, rows, src, dst); }
int cattl3_dropout_forward_f32(cattl3_ctx* c, int64_t count, float prob, float eps, uint64_t seed, const float* x, float* y, uint8_t* mask) {
	return dropout_forward<float>(c, count, prob, eps, seed, x, y, mask); }
int cattl3_dropout_forward_f64(cattl3_ctx* c, int64_t count, double prob, double eps, uint64_t seed, const double* x, double* y, uint8_t* mask) {
	return dropout_forward<double>(c, count, prob, eps, seed, x, y, mask); }
int cattl3_dropout_backward_f32(cattl3_ctx* c, int64_t count, float prob, float eps, const float* dy, const uint8_t* mask, float* dx) {
	return dropout_backward<float>(c, count, prob, eps, dy, mask, dx); }
int cattl3_dropout_backward_f64(cattl3_ctx* c, int64_t count, double prob, double eps, const double* dy, const uint8_t* mask, double* dx) {
	return dropout_backward<double>(c, count, prob, eps, dy, mask, dx); }
int cattl3_loss_f32(cattl3_ctx* c, int kind, int64_t rows, int64_t vol, float eps, float grad_div, const float* out, const float* obj, float* loss, float* grad) {
	return loss_forward_backward<float>(c, kind, rows, vol, eps, grad_div, out, obj, loss, grad); }
int cattl3_loss_f64(cattl3_ctx* c, int kind, int64_t rows, int64_t vol, double eps, double grad_div, const double* out, const double* obj, double* loss, double* grad) {
	return loss_forward_backward<double>(c, kind, rows, vol, eps, grad_div, out, obj, loss, grad); }

int cattl3_activation_forward_f32(cattl3_ctx* c, int kind, float alpha, int64_t rows, int64_t vol, const float* x, float* y) {
	return activation_forward<float>(c, kind, alpha, rows, vol, x, y); }
int cattl3_activation_forward_f64(cattl3_ctx* c, int kind, double alpha, int64_t rows, int64_t vol, const double* x, double* y) {
	return activation_forward<double>(c, kind, alpha, rows, vol, x, y); }
int cattl3_activation_backward_f32(cattl3_ctx* c, int kind, float alpha, int64_t rows, int64_t vol, const float* x, const float* y, const float* dy, float* dx) {
	return activation_backward<float>(c, kind, alpha, rows, vol, x, y, dy, dx); }
int cattl3_activation_backward_f64(cattl3_ctx* c, int kind, double alpha, int64_t rows, int64_t vol, const double* x, const double* y, const double* dy, double* dx) {
	return activation_backward<double>(c, kind, alpha, rows, vol, x, y, dy, dx); }

int cattl3_pool_forward_f32(cattl3_ctx* c, int kind, const cattl3_pool_geom* g, const float* x, float* y, uint8_t* am) {
	return pool_forward<float>(c, kind, g, x, y, am); }
int cattl3_pool_forward_f64(cattl3_ctx* c, int kind, const cattl3_pool_geom* g, const double* x, double* y, uint8_t* am) {
	return pool_forward<double>(c, kind, g, x, y, am); }
int cattl3_pool_backward_f32(cattl3_ctx* c, int kind, const cattl3_pool_geom* g, const float* dy, const uint8_t* am, float* dx) {
	return pool_backward<float>(c, kind, g, dy, am, dx); }
int cattl3_pool_backward_f64(cattl3_ctx* c, int kind, const cattl3_pool_geom* g, const double* dy, const uint8_t* am, double* dx) {
	return pool_backward<double>(c, kind, g, dy, am, dx); }

int cattl3_batchnorm_forward_f32(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, int training, int rinit, float decay, float eps, const float* x, const float* gamma, const float* beta, float* rm, float* rs, float* sm, float* ss, float* y) {
	return batchnorm_forward<float>(c, pc, n, h, w, ch, training, rinit, decay, eps, x, gamma, beta, rm, rs, sm, ss, y); }
int cattl3_batchnorm_forward_f64(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, int training, int rinit, double decay, double eps, const double* x, const double* gamma, const double* beta, double* rm, double* rs, double* sm, double* ss, double* y) {
	return batchnorm_forward<double>(c, pc, n, h, w, ch, training, rinit, decay, eps, x, gamma, beta, rm, rs, sm, ss, y); }
int cattl3_batchnorm_forward_stats_f32(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, int rinit, float decay, float eps, const float* x, const double* cs, const double* gc, const float* shift, const float* gamma, const float* beta, float* rm, float* rs, float* sm, float* ss, float* y, int ak, float ap, float* ao) {
	return batchnorm_forward_stats<float>(c, pc, n, h, w, ch, rinit, decay, eps, x, cs, gc, shift, gamma, beta, rm, rs, sm, ss, y, ak, ap, ao); }
int cattl3_batchnorm_forward_stats_f64(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, int rinit, double decay, double eps, const double* x, const double* cs, const double* gc, const double* shift, const double* gamma, const double* beta, double* rm, double* rs, double* sm, double* ss, double* y, int ak, double ap, double* ao) {
	return batchnorm_forward_stats<double>(c, pc, n, h, w, ch, rinit, decay, eps, x, cs, gc, shift, gamma, beta, rm, rs, sm, ss, y, ak, ap, ao); }
int cattl3_batchnorm_stats_f32(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const float* x, const float* shift, double* cs) {
	return batchnorm_stats<float>(c, pc, n, h, w, ch, x, shift, cs); }
int cattl3_batchnorm_stats_f64(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const double* x, const double* shift, double* cs) {
	return batchnorm_stats<double>(c, pc, n, h, w, ch, x, shift, cs); }
int cattl3_batchnorm_backward_sums_f32(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const float* x, const float* sm, const float* ss, const float* dy, float* dgamma, float* dbeta, double* sums) {
	return batchnorm_backward_sums<float>(c, pc, n, h, w, ch, x, sm, ss, dy, dgamma, dbeta, sums); }
int cattl3_batchnorm_backward_sums_f64(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const double* x, const double* sm, const double* ss, const double* dy, double* dgamma, double* dbeta, double* sums) {
	return batchnorm_backward_sums<double>(c, pc, n, h, w, ch, x, sm, ss, dy, dgamma, dbeta, sums); }
int cattl3_batchnorm_backward_apply_f32(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const double* total, const float* x, const float* gamma, const float* sm, const float* ss, const float* dy, const double* sums, float* dx) {
	return batchnorm_backward_apply<float>(c, pc, n, h, w, ch, total, x, gamma, sm, ss, dy, sums, dx); }
int cattl3_batchnorm_backward_apply_f64(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const double* total, const double* x, const double* gamma, const double* sm, const double* ss, const double* dy, const double* sums, double* dx) {
	return batchnorm_backward_apply<double>(c, pc, n, h, w, ch, total, x, gamma, sm, ss, dy, sums, dx); }
int cattl3_batchnorm_backward_f32(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const float* x, const float* gamma, const float* sm, const float* ss, const float* dy, float* dgamma, float* dbeta, float* dx) {
	return batchnorm_backward<float>(c, pc, n, h, w, ch, x, gamma, sm, ss, dy, dgamma, dbeta, dx); }
int cattl3_batchnorm_backward_f64(cattl3_ctx* c, int pc, int32_t n, int32_t h, int32_t w, int32_t ch, const double* x, const double* gamma, const double* sm, const double* ss, const double* dy, double* dgamma, double* dbeta, double* dx) {
	return batchnorm_backward<double>(c, pc, n, h, w, ch, x, gamma, sm, ss, dy, dgamma, dbeta, dx); }

int cattl3_optimizer_step_f32(cattl3_ctx* c, const cattl3_opt_step* st, int64_t count, float* p, float* g, float* s1, float* s2, float* s3) {
	return optimizer_step<float>(c, st, count, p, g, s1, s2, s3); }
int cattl3_optimizer_step_f64(cattl3_ctx* c, const cattl3_opt_step* st, int64_t count, double* p, double* g, double* s1, double* s2, double* s3) {
	return optimizer_step<double>(c, st, count, p, g, s1, s2, s3); }

int cattl3_optimizer_step_indirect_f32(cattl3_ctx* c, int kind, const cattl3_opt_step* dev_step, int64_t count, float* p, float* g, float* s1, float* s2, float* s3) {
	CATTL3_REQUIRE(dev_step, "optimizer_step_indirect: null scalars");
	cattl3_opt_step st = {}; st.kind = kind;
	return optimizer_step<float>(c, &st, count, p, g, s1, s2, s3, dev_step); }
int cattl3_optimizer_step_indirect_f64(cattl3_ctx* c, int kind, const cattl3_opt_step* dev_step, int64_t count, double* p, double* g, double* s1, double* s2, double* s3) {
	CATTL3_REQUIRE(dev_step, "optimizer_step_indirect: null scalars");
	cattl3_opt_step st = {}; st.kind = kind;
	return optimizer_step<double>(c, &st, count, p, g, s1, s2, s3, dev_step); }
int cattl3_constrain_f32(cattl3_ctx* c, int64_t count, float clip, float max_l1, float max_l2, float* x) {
	return constrain<float>(c, count, clip, max_l1, max_l2, x); }
int cattl3_constrain_f64(cattl3_ctx* c, int64_t count, double clip, double max_l1, double max_l2, double* x) {
	return constrain<double>(c, count, clip, max_l1, max_l2, x); }
int cattl3_add_inplace_f32(cattl3_ctx* c, int64_t count, float* y, const float* x) { return add_inplace<float>(c, count, y, x); }
int cattl3_add_inplace_f64(cattl3_ctx* c, int64_t count, double* y, const double* x) { return add_inplace<double>(c, count, y, x); }
int cattl3_muladd_f32(cattl3_ctx* c, int64_t count, int accumulate, const float* a, const float* b, const float* cc, const float* d, float* out) {
	return muladd<float>(c, count, accumulate, a, b, cc, d, out); }
int cattl3_muladd_f64(cattl3_ctx* c, int64_t count, int accumulate, const double* a, const double* b, const double* cc, const double* d, double* out) {
	return muladd<double>(c, count, accumulate, a, b, cc, d, out); }
int cattl3_regularize_f32(cattl3_ctx* c, int64_t count, float l1, float l2, const float* values, float* grad, double* penalty) {
	return regularize<float>(c, count, l1, l2, values, grad, penalty); }
int cattl3_regularize_f64(cattl3_ctx* c, int64_t count, double l1, double l2, const double* values, double* grad, double* penalty) {
	return regularize<double>(c, count, l1, l2, values, grad, penalty); }
int cattl3_mul_inplace_f32(cattl3_ctx* c, int64_t count, float* y, const float* x) { return mul_inplace<float>(c, count, y, x); }
int cattl3_mul_inplace_f64(cattl3_ctx* c, int64_t count, double* y, const double* x) { return mul_inplace<double>(c, count, y, x); }
int cattl3_scale_f32(cattl3_ctx* c, int64_t count, float alpha, const float* x, float* y) { return scale<float>(c, count, alpha, x, y); }
int cattl3_scale_f64(cattl3_ctx* c, int64_t count, double alpha, const double* x, double* y) { return scale<double>(c, count, alpha, x, y); }
int cattl3_axpy_f32(cattl3_ctx* c, int64_t count, float alpha, const float* x, float* y) { return axpy<float>(c, count, alpha, x, y); }
int cattl3_axpy_f64(cattl3_ctx* c, int64_t count, double alpha, const double* x, double* y) { return axpy<double>(c, count, alpha, x, y); }

}
