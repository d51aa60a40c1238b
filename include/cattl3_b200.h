/*
 * cattl3_b200.h -- C ABI of the B200-native C-ATTL3 hot path (libcattl3_b200.so).
 *
 * The reference (ViktorC/C-ATTL3) is a header-only C++ template library whose plug-in boundary
 * is the virtual cattle::Layer / cattle::Parameters / cattle::SGDOptimizer API; it has no FFI.
 * This header is the thin extern "C" layer underneath the header-only replacement classes in
 * c-attl3_b200/cattle/ (namespace cattle, same class names and constructor signatures as the
 * reference).  Each entry point names the reference function it replaces (paths relative to the
 * reference checkout, C-ATTL3/...).
 *
 * Conventions
 *  - All tensors use the reference's memory layout: Eigen column-major rank-4 tensors
 *    (N, H, W, C) with N FASTEST: offset(n,h,w,c) = n + N*(h + H*(w + W*c))
 *    (C-ATTL3/core/EigenProxy.hpp:56-57).  Parameter matrices are column-major.
 *  - Plain pointers and sizes only.  Unless a function name ends in `_host`, data pointers are
 *    DEVICE pointers valid on the context's device; work is enqueued on the context's stream and
 *    the call returns without synchronising.
 *  - `_f32` / `_f64` mirror the reference's float / double Scalar template instantiations.
 *  - Every function returns CATTL3_OK (0) or an error code; cattl3_last_error() gives the message
 *    of the calling thread's last failure.  Nothing throws across this boundary; the C++ headers
 *    turn non-zero codes into std::runtime_error (the convention of the reference's own
 *    C-ATTL3/core/gpu/cuda/CUDAError.hpp:17-52).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *    CATTL3_ERR_NO_DEVICE.
 *  - Entry points are re-entrant across distinct contexts (one context per host thread / stream;
 *    the reference runs network lanes on pthreads, C-ATTL3/neural_network/ParallelNeuralNetwork.hpp:142-197).
 */
#ifndef CATTL3_B200_H_
#define CATTL3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATTL3_ABI_VERSION 1

enum {
	CATTL3_OK = 0,
	CATTL3_ERR_INVALID = 1,      /* bad argument / inconsistent geometry */
	CATTL3_ERR_CUDA = 2,         /* a CUDA runtime / driver call failed */
	CATTL3_ERR_UNSUPPORTED = 3,  /* valid request outside what the kernels cover */
	CATTL3_ERR_NO_DEVICE = 4     /* no usable CUDA device (there is no CPU fallback) */
};

/* Activation kinds: ReLU / LeakyReLU / ELU / Swish are the hot-path set;
 * Sigmoid / Tanh / Softplus / Softmax are the "next" set (section 8f rank 1). */
enum {
	CATTL3_ACT_RELU = 0,      /* C-ATTL3/layer/activation/ReLUActivationLayer.hpp:45-57 */
	CATTL3_ACT_LEAKY_RELU = 1,/* C-ATTL3/layer/activation/LeakyReLUActivationLayer.hpp:50-62 (alpha) */
	CATTL3_ACT_ELU = 2,       /* C-ATTL3/layer/activation/ELUActivationLayer.hpp:55-78 (alpha) */
	CATTL3_ACT_SWISH = 3,     /* C-ATTL3/layer/activation/SwishActivationLayer.hpp:45-61 (alpha = beta) */
	CATTL3_ACT_SIGMOID = 4,   /* C-ATTL3/layer/activation/SigmoidActivationLayer.hpp */
	CATTL3_ACT_TANH = 5,      /* C-ATTL3/layer/activation/TanhActivationLayer.hpp */
	CATTL3_ACT_SOFTPLUS = 6,  /* C-ATTL3/layer/activation/SoftplusActivationLayer.hpp:40-52 */
	CATTL3_ACT_SOFTMAX = 7,   /* C-ATTL3/layer/activation/SoftmaxActivationLayer.hpp:49-78 (alpha = epsilon) */
	CATTL3_ACT_NONE = -1      /* cattl3_epilogue: no activation fused */
};

enum {
	CATTL3_POOL_MAX = 0,      /* C-ATTL3/layer/pool/MaxPoolLayer.hpp:38-80 */
	CATTL3_POOL_MEAN = 1      /* C-ATTL3/layer/pool/MeanPoolLayer.hpp:35-41 */
};

enum {
	CATTL3_OPT_VANILLA_SGD = 0, /* C-ATTL3/optimizer/VanillaSGDOptimizer.hpp:38-43 */
	CATTL3_OPT_MOMENTUM = 1,    /* C-ATTL3/optimizer/MomentumSGDOptimizer.hpp:54-63 */
	CATTL3_OPT_NESTEROV = 2,    /* C-ATTL3/optimizer/NesterovMomentumSGDOptimizer.hpp:43-55 */
	CATTL3_OPT_ADAGRAD = 3,     /* C-ATTL3/optimizer/AdaGradOptimizer.hpp:49-71 */
	CATTL3_OPT_RMSPROP = 4,     /* C-ATTL3/optimizer/RMSPropOptimizer.hpp:42-46 */
	CATTL3_OPT_ADADELTA = 5,    /* C-ATTL3/optimizer/AdaDeltaOptimizer.hpp:53-67 */
	CATTL3_OPT_ADAM = 6,        /* C-ATTL3/optimizer/AdamOptimizer.hpp:66-82 */
	CATTL3_OPT_ADAMAX = 7,      /* C-ATTL3/optimizer/AdaMaxOptimizer.hpp:44-60 */
	CATTL3_OPT_NADAM = 8,       /* C-ATTL3/optimizer/NadamOptimizer.hpp:44-63 */
	CATTL3_OPT_AMSGRAD = 9      /* C-ATTL3/optimizer/AMSGradOptimizer.hpp:50-67 */
};

/* Device losses. */
enum {
	CATTL3_LOSS_SQUARED = 0,        /* C-ATTL3/loss/SquaredLoss.hpp:25-34 */
	CATTL3_LOSS_CROSS_ENTROPY = 1   /* C-ATTL3/loss/CrossEntropyLoss.hpp:33-42 (eps inside the log and the quotient) */
};

/* Which kernel family a kernel-layer call may use (cattl3_ctx_set_conv_path). */
enum {
	CATTL3_PATH_AUTO = 0,       /* float: tcgen05 where the shape allows it; double: FP64 tensor cores (DMMA); then the
	                             * big-tile FMA, streaming ("tiny", "skinny") and small-tile SIMT kernels */
	CATTL3_PATH_SIMT = 1,       /* FFMA / DFMA implicit GEMM (any shape, float and double) */
	CATTL3_PATH_TCGEN05 = 2,    /* TMA + tcgen05 3xTF32 implicit GEMM; CATTL3_ERR_UNSUPPORTED if not applicable */
	CATTL3_PATH_FMA = 3         /* no tensor cores: the big-tile FFMA / DFMA kernels where they apply, else SIMT */
};

/*
 * Convolution geometry, in the reference's own terms (constructor arguments of
 * ConvKernelLayer / TransConvKernelLayer, C-ATTL3/layer/kernel/ConvKernelLayer.hpp:282-296):
 * input n x h x w x c, f filters, receptor rh x rw, padding (ph, pw) on both sides, stride
 * (sh, sw), dilation (dh, dw) where 0 means dense (effective tap step = d + 1).
 */
typedef struct cattl3_conv_geom {
	int32_t n, h, w, c;
	int32_t f;
	int32_t rh, rw;
	int32_t ph, pw;
	int32_t sh, sw;
	int32_t dh, dw;
} cattl3_conv_geom;

/* Pooling geometry (PoolLayer constructor, C-ATTL3/layer/PoolLayer.hpp:47-70): no padding. */
typedef struct cattl3_pool_geom {
	int32_t n, h, w, c;
	int32_t rh, rw;
	int32_t sh, sw;
} cattl3_pool_geom;

/*
 * One optimizer update.  lr/a/b/eps are the constructor hyper-parameters in the order the
 * reference declares them (kind 1,2: a = annealing_rate, b = momentum; 4: b = l2_decay;
 * 5: a = decay; 6-9: a = l1_decay, b = l2_decay).  The host evaluates the step-dependent scalars
 * exactly as the reference does and passes them in:
 *   lr_epoch = init_learning_rate / (1 + annealing_rate * epoch)      (MomentumSGDOptimizer.hpp:70-72)
 *   c1  = 1 / (1 - (1-a)^(t+1) + eps), c1n = 1 / (1 - (1-a)^(t+2) + eps),
 *   c2  = 1 / (1 - (1-b)^(t+1) + eps)                                 (AdamOptimizer.hpp:67-68, NadamOptimizer.hpp:45-47)
 * l2_lambda > 0 folds Parameters::regularize() of an L2ParameterRegularization into the step
 * (g += lambda * p, C-ATTL3/parameter_regularization/L2ParameterRegularization.hpp:31-33);
 * reset_grad != 0 zeroes the gradient afterwards (SGDOptimizer.hpp:69-70).
 */
typedef struct cattl3_opt_step {
	int32_t kind;
	int32_t reset_grad;
	double lr, a, b, eps;
	double lr_epoch, c1, c1n, c2;
	double l2_lambda;
} cattl3_opt_step;

/*
 * What a kernel layer's forward pass does in its epilogue besides adding the bias -- the work of the layer that
 * FOLLOWS it in the reference's layer loop (FeedforwardNeuralNetwork.hpp:112-118), done while the output tile is
 * still on chip:
 *  - act_kind != CATTL3_ACT_NONE: the element-wise ActivationLayer (kinds 0-6, formulas as cattl3_activation_forward):
 *    act_out = f(y).  The pre-activation y is still written when the y argument of the call is non-null (ReLU,
 *    LeakyReLU, ELU, Swish and Softplus layers cache their input for pass_back,
 *    e.g. ReLUActivationLayer.hpp:45-57); pass y = NULL to skip it (inference, Sigmoid / Tanh).
 *  - col_stats != NULL: the first reduction of a following BatchNormLayer (BatchNormLayer.hpp:225-233): per output
 *    column j (filter / dense output), over all rows m = N*OH*OW: col_stats[j] = sum_m (y(m,j) - b(j)),
 *    col_stats[J + j] = sum_m (y(m,j) - b(j))^2, in double whatever the scalar type (shifted by the bias so that the
 *    later E[d^2] - E[d]^2 does not cancel).  cattl3_batchnorm_forward_stats_* consumes them.
 */
typedef struct cattl3_epilogue {
	int32_t act_kind;
	int32_t reserved;
	double act_param;   /* alpha / beta of the activation */
	void* act_out;      /* device, same type and shape as y */
	double* col_stats;  /* device, 2 * J doubles */
} cattl3_epilogue;

typedef struct cattl3_ctx cattl3_ctx; /* opaque: device, stream, workspaces, cached TMA descriptors */

/* ---- library / context ------------------------------------------------------------------- */
int cattl3_abi_version(void);
const char* cattl3_last_error(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int cattl3_device_count(void);
/* cuda_stream: a cudaStream_t to enqueue on (e.g. the caller's framework stream; pass
 * cudaStreamLegacy = (cudaStream_t) 0x1 for the default stream) or NULL to let the context create
 * its own non-blocking stream. */
int cattl3_ctx_create(cattl3_ctx** out, int device, void* cuda_stream);
int cattl3_ctx_destroy(cattl3_ctx* ctx);
int cattl3_ctx_synchronize(cattl3_ctx* ctx);
/* Bounded run-ahead for an asynchronous step loop: marks "now" on the context's stream and blocks the host until the
 * mark made max_in_flight calls earlier has been reached (call once per training step). */
int cattl3_ctx_throttle(cattl3_ctx* ctx, int max_in_flight);
/* Between _begin and _end the caller promises that no weight array changes (a training step up to its optimizer update): the
 * kernel layers then keep the repacked (K-major, hi | lo split) weights of each (array, geometry) they meet and reuse them, e.g.
 * across the time steps of an unrolled LSTM whose cells share their kernels (LSTMNeuralNetwork.hpp:249-252).  Scopes do not nest. */
int cattl3_weights_stable_begin(cattl3_ctx*);
int cattl3_weights_stable_end(cattl3_ctx*);
int cattl3_ctx_set_conv_path(cattl3_ctx* ctx, int path);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t cattl3_ctx_launch_count(const cattl3_ctx* ctx);
/* Name of the kernel family the last kernel-layer call on this context used ("tcgen05"/"simt"). */
const char* cattl3_ctx_last_path(const cattl3_ctx* ctx);
void* cattl3_ctx_stream(const cattl3_ctx* ctx);

/* ---- memory helpers (so g++-only hosts need no CUDA headers) ------------------------------ */
/* Stream-ordered (cudaMallocAsync / cudaFreeAsync on the context's stream, pool kept warm): neither
 * call synchronises, and a freed block may be reused by later work on the same stream. */
int cattl3_malloc(cattl3_ctx* ctx, void** dev_ptr, size_t bytes);
int cattl3_free(cattl3_ctx* ctx, void* dev_ptr);
int cattl3_memset(cattl3_ctx* ctx, void* dev_ptr, int value, size_t bytes);
int cattl3_memcpy_h2d(cattl3_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
int cattl3_memcpy_d2h(cattl3_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes); /* synchronises */
int cattl3_memcpy_d2d(cattl3_ctx* ctx, void* dev_dst, const void* dev_src, size_t bytes);
/* `height` blocks of `width` bytes, device to device, with row pitches: joining / splitting tensors along a rank
 * (DenseNeuralNetwork.hpp:131-160 concatenate / slice; a concatenation along rank r of these column-major tensors is
 * one contiguous block per index of the ranks above r). */
int cattl3_memcpy_2d(cattl3_ctx* ctx, void* dev_dst, size_t dst_pitch, const void* dev_src, size_t src_pitch, size_t width, size_t height);
int cattl3_host_alloc(void** host_ptr, size_t bytes); /* pinned */
int cattl3_host_free(void* host_ptr);

/*
 * Input feed: the mini-batch upload of the batch loop (SGDOptimizer.hpp:44-47 hands each batch to propagate) taken
 * off the compute stream.  A ring of `slots` device buffers is filled through pinned staging on a dedicated copy
 * stream, so the upload of batch i+1 overlaps the kernels of batch i; cattl3_feed_push returns as soon as the host
 * data has been staged (the source may be reused), work enqueued afterwards on the context's stream sees the data, and
 * the slot is overwritten `slots` pushes later -- ordered after everything the context's stream had been given
 * `slots - 1` pushes earlier, i.e. after the step that consumed it.  The returned device pointer belongs to the feed.
 */
typedef struct cattl3_feed cattl3_feed;
int cattl3_feed_create(cattl3_feed** out, cattl3_ctx* ctx, int slots);
int cattl3_feed_destroy(cattl3_feed* feed);
int cattl3_feed_push(cattl3_feed* feed, const void* host_src, size_t bytes, void** dev_ptr);

/* y[i] = value on the device (no host -> device copy, hence no host synchronisation). */
int cattl3_fill_f32(cattl3_ctx*, int64_t count, float value, float* y);
int cattl3_fill_f64(cattl3_ctx*, int64_t count, double value, double* y);
/* Rows [first, first + rows) of a device-resident (total x vol) data set, rows fastest (MemoryDataProvider::get_data,
 * C-ATTL3/data_provider/MemoryDataProvider.hpp:72-83, without the host slice and the per-step upload). */
int cattl3_slice_rows_f32(cattl3_ctx*, int64_t total, int64_t vol, int64_t first, int64_t rows, const float* src, float* dst);
int cattl3_slice_rows_f64(cattl3_ctx*, int64_t total, int64_t vol, int64_t first, int64_t rows, const double* src, double* dst);

/*
 * Step graphs: a launch-bound training step (tens of small kernels: configs 1 and 3) captured once into a CUDA graph and
 * replayed with one launch per step (the batch loop of C-ATTL3/optimizer/SGDOptimizer.hpp:34-81 pays one launch per
 * layer pass otherwise).  Between _begin and _end everything enqueued on the context's stream is recorded instead of run:
 * kernels, memsets and device-to-device copies.  cattl3_malloc hands out blocks of a private arena of `arena_bytes`
 * (fixed addresses, recycled by cattl3_free within the capture; the graph holds no allocation nodes) -- size it with
 * cattl3_ctx_allocated_bytes() taken before and after one eager run of the same step.  Nothing may synchronise or
 * grow library scratch (the step must have run eagerly at the same shape before), and host-side arguments are frozen
 * into the graph -- step-dependent scalars therefore go through device memory (cattl3_optimizer_step_indirect).
 * Blocks still held when the capture ends stay valid for the life of the graph; freeing them later is a no-op.  Memory
 * from before the capture that is released during it is returned to the pool when the capture ends (never by the graph).
 * Errors: CATTL3_ERR_UNSUPPORTED when the arena is too small or scratch would have to grow (the capture must still be
 * closed with _end, which then reports the invalidated capture).
 */
typedef struct cattl3_graph cattl3_graph;
int64_t cattl3_ctx_allocated_bytes(cattl3_ctx* ctx);   /* running total handed out by cattl3_malloc, in 256-byte granules */
int cattl3_graph_begin(cattl3_ctx* ctx, size_t arena_bytes);
int cattl3_graph_end(cattl3_ctx* ctx, cattl3_graph** out);   /* *out = NULL and an error if the capture was invalidated */
int cattl3_graph_launch(cattl3_ctx* ctx, cattl3_graph* graph);   /* CATTL3_ERR_UNSUPPORTED: library scratch moved since the capture -- destroy and capture again */
int cattl3_graph_destroy(cattl3_graph* graph);

/* ---- shape helpers ------------------------------------------------------------------------ */
/* ConvKernelLayer.hpp:194-197 (transposed = 0) / TransConvKernelLayer.hpp:200-203 (transposed = 1). */
int cattl3_conv_output_dims(const cattl3_conv_geom* g, int transposed, int32_t* oh, int32_t* ow);
/* PoolLayer.hpp:145-148. */
int cattl3_pool_output_dims(const cattl3_pool_geom* g, int32_t* oh, int32_t* ow);

/* ---- kernel layers ------------------------------------------------------------------------ */
/*
 * ConvKernelLayerBase::_pass_forward (ConvKernelLayer.hpp:115-148) without the im2col buffer:
 *   y(n,oh,ow,f) = b(f) + sum_{c,rw,rh} x(n, oh*sh + rh*(dh+1) - ph, ow*sw + rw*(dw+1) - pw, c) * W(rh + RH*(rw + RW*c), f)
 * x: n*h*w*c, w: (rh*rw*c) x f col-major, b: f, y: n*oh*ow*f.
 */
int cattl3_conv_forward_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x, const float* w, const float* b, float* y);
int cattl3_conv_forward_f64(cattl3_ctx*, const cattl3_conv_geom*, const double* x, const double* w, const double* b, double* y);
/*
 * ConvKernelLayerBase::_pass_back (ConvKernelLayer.hpp:149-189).  dw and db ACCUMULATE (beta = 1,
 * Parameters::accumulate_grad, StandardParameters.hpp:115-123); dx is overwritten and may be NULL
 * for an input layer (Layer.hpp:82-90).  The weight gradient is a deterministic split-K reduction.
 * dw = db = NULL (then x may be NULL too): only the input gradient -- the cells of an unrolled recurrent network
 * share their kernels' parameters (LSTMNeuralNetwork.hpp:537-573), so their weight gradients are taken in one call
 * over all time steps (n = samples * steps: the sequence layout makes that a plain batch).  The same holds for
 * cattl3_dense_backward.
 */
int cattl3_conv_backward_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x, const float* w, const float* dy, float* dw, float* db, float* dx);
int cattl3_conv_backward_f64(cattl3_ctx*, const cattl3_conv_geom*, const double* x, const double* w, const double* dy, double* dw, double* db, double* dx);
/*
 * TransConvKernelLayerBase::_pass_forward / _pass_back (TransConvKernelLayer.hpp:115-151, 152-187).
 * geom.h/w/c describe the INPUT (ih x iw x c); w: c x (rh*rw*f) col-major with column
 * rh + RH*(rw + RW*f); b: one bias PER OUTPUT ELEMENT, oh*ow*f; y: n*oh*ow*f.
 */
int cattl3_transconv_forward_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x, const float* w, const float* b, float* y);
int cattl3_transconv_forward_f64(cattl3_ctx*, const cattl3_conv_geom*, const double* x, const double* w, const double* b, double* y);
int cattl3_transconv_backward_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x, const float* w, const float* dy, float* dw, float* db, float* dx);
int cattl3_transconv_backward_f64(cattl3_ctx*, const cattl3_conv_geom*, const double* x, const double* w, const double* dy, double* dw, double* db, double* dx);
/*
 * DenseKernelLayer::pass_forward / pass_back (DenseKernelLayer.hpp:92-102, 103-115):
 * x: n x in col-major (the free view of an (N,H,W,C) tensor), w: in x out, b: out, y: n x out.
 */
int cattl3_dense_forward_f32(cattl3_ctx*, int32_t n, int32_t in, int32_t out, const float* x, const float* w, const float* b, float* y);
int cattl3_dense_forward_f64(cattl3_ctx*, int32_t n, int32_t in, int32_t out, const double* x, const double* w, const double* b, double* y);
int cattl3_dense_backward_f32(cattl3_ctx*, int32_t n, int32_t in, int32_t out, const float* x, const float* w, const float* dy, float* dw, float* db, float* dx);
int cattl3_dense_backward_f64(cattl3_ctx*, int32_t n, int32_t in, int32_t out, const double* x, const double* w, const double* dy, double* dw, double* db, double* dx);

/*
 * Forward passes with a fused epilogue (cattl3_epilogue above): ConvKernelLayer / DenseKernelLayer followed by an
 * ActivationLayer and / or the statistics pass of a BatchNormLayer.  y may be NULL when ep->act_out is given.
 * On the tcgen05 path everything happens in the GEMM's epilogue; on the SIMT path the activation is fused and the
 * column statistics are a separate reduction over y (same results, one more read).
 */
int cattl3_conv_forward_fused_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x, const float* w, const float* b, float* y, const cattl3_epilogue* ep);
int cattl3_conv_forward_fused_f64(cattl3_ctx*, const cattl3_conv_geom*, const double* x, const double* w, const double* b, double* y, const cattl3_epilogue* ep);
int cattl3_transconv_forward_fused_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x, const float* w, const float* b, float* y, const cattl3_epilogue* ep);   /* activation only */
int cattl3_transconv_forward_fused_f64(cattl3_ctx*, const cattl3_conv_geom*, const double* x, const double* w, const double* b, double* y, const cattl3_epilogue* ep);
int cattl3_dense_forward_fused_f32(cattl3_ctx*, int32_t n, int32_t in, int32_t out, const float* x, const float* w, const float* b, float* y, const cattl3_epilogue* ep);
int cattl3_dense_forward_fused_f64(cattl3_ctx*, int32_t n, int32_t in, int32_t out, const double* x, const double* w, const double* b, double* y, const cattl3_epilogue* ep);

/* Host-buffer forms of the convolution layer: what the reference's Layer API hands over
 * (pass_forward(Data in, bool) / pass_back(Data out_grad), Layer.hpp:126,137) -- host tensors in,
 * host tensors out, host<->device copies inside the call.  Parameters stay device resident.
 * The copies are pipelined with the kernels over chunks of filters on an upload and a download stream (PCIe in both
 * directions at once).  The plain forms return with the host tensors complete (the reference's semantics); the
 * _async forms return once the work is enqueued and cattl3_host_wait() completes every outstanding transfer, so a
 * forward's download of y overlaps the following backward's upload of dY.  Host tensors should be pinned
 * (cattl3_host_alloc).  x_dev_keep (optional) receives the device copy of x for the backward call. */
int cattl3_conv_forward_host_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x_host, const float* w_dev, const float* b_dev, float* y_host, float* x_dev_keep);
int cattl3_conv_backward_host_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x_dev, const float* w_dev, const float* dy_host, float* dw_dev, float* db_dev, float* dx_host);
int cattl3_conv_forward_host_async_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x_host, const float* w_dev, const float* b_dev, float* y_host, float* x_dev_keep);
int cattl3_conv_backward_host_async_f32(cattl3_ctx*, const cattl3_conv_geom*, const float* x_dev, const float* w_dev, const float* dy_host, float* dw_dev, float* db_dev, float* dx_host);
int cattl3_host_wait(cattl3_ctx*);

/* ---- activation layers -------------------------------------------------------------------- */
/* x, y: rows x vol elements, rows (= batch) fastest.  Softmax normalises each row over `vol`. */
int cattl3_activation_forward_f32(cattl3_ctx*, int kind, float alpha, int64_t rows, int64_t vol, const float* x, float* y);
int cattl3_activation_forward_f64(cattl3_ctx*, int kind, double alpha, int64_t rows, int64_t vol, const double* x, double* y);
/* dx = f'(x) * dy, using the cached input x and output y exactly where the reference does. */
int cattl3_activation_backward_f32(cattl3_ctx*, int kind, float alpha, int64_t rows, int64_t vol, const float* x, const float* y, const float* dy, float* dx);
int cattl3_activation_backward_f64(cattl3_ctx*, int kind, double alpha, int64_t rows, int64_t vol, const double* x, const double* y, const double* dy, double* dx);

/* ---- dropout (DropoutLayer.hpp:74-94) ---------------------------------- */
/* Inverted dropout in training mode: mask = u <= prob ? 0 : 1 / (1 - prob + eps), y = x * mask; u is a uniform [0,1)
 * draw from a counter-based generator keyed by (seed, element index) -- deterministic per seed, NOT the reference's
 * host RNG stream (its masks are not reproducible either way).  mask: one byte per element,
 * kept for backward (dx = dy * mask). */
int cattl3_dropout_forward_f32(cattl3_ctx*, int64_t count, float prob, float eps, uint64_t seed, const float* x, float* y, uint8_t* mask);
int cattl3_dropout_forward_f64(cattl3_ctx*, int64_t count, double prob, double eps, uint64_t seed, const double* x, double* y, uint8_t* mask);
int cattl3_dropout_backward_f32(cattl3_ctx*, int64_t count, float prob, float eps, const float* dy, const uint8_t* mask, float* dx);
int cattl3_dropout_backward_f64(cattl3_ctx*, int64_t count, double prob, double eps, const double* dy, const uint8_t* mask, double* dx);

/* ---- losses (UniversalLoss.hpp:24-58 -> SquaredLoss / CrossEntropyLoss) --------------------------------- */
/* out, obj: rows x vol, rows (= batch) fastest.  loss (rows, may be NULL): per-sample loss; grad (rows x vol, may be
 * NULL): d loss / d out divided by grad_div -- the batch loop's nominal batch size (SGDOptimizer.hpp:55-56). */
int cattl3_loss_f32(cattl3_ctx*, int kind, int64_t rows, int64_t vol, float eps, float grad_div, const float* out, const float* obj, float* loss, float* grad);
int cattl3_loss_f64(cattl3_ctx*, int kind, int64_t rows, int64_t vol, double eps, double grad_div, const double* out, const double* obj, double* loss, double* grad);

/* ---- pooling layers (PoolLayer.hpp:77-116) -------------------------------------------------- */
/* argmax: one byte per output element, index rw*RH + rh of the first maximum in the reference's
 * scan order (MaxPoolLayer.hpp:45-58); required for CATTL3_POOL_MAX, ignored for MEAN. */
int cattl3_pool_forward_f32(cattl3_ctx*, int kind, const cattl3_pool_geom*, const float* x, float* y, uint8_t* argmax);
int cattl3_pool_forward_f64(cattl3_ctx*, int kind, const cattl3_pool_geom*, const double* x, double* y, uint8_t* argmax);
int cattl3_pool_backward_f32(cattl3_ctx*, int kind, const cattl3_pool_geom*, const float* dy, const uint8_t* argmax, float* dx);
int cattl3_pool_backward_f64(cattl3_ctx*, int kind, const cattl3_pool_geom*, const double* dy, const uint8_t* argmax, double* dx);

/* ---- batch normalisation (BatchNormLayer.hpp:170-262 per channel, 337-391 per activation) ---- */
/*
 * groups = c (per_channel) or h*w*c; every group is L = n*h*w (or n) contiguous elements.
 * training != 0: batch statistics; saved_mean / saved_inv_sd (groups) are written for backward and
 * the running averages are updated (assigned when *first* batch, i.e. running_initialised == 0,
 * else (1-decay)*avg + decay*new).  training == 0: running averages are used.
 */
int cattl3_batchnorm_forward_f32(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, int training, int running_initialised, float decay, float eps, const float* x, const float* gamma, const float* beta, float* running_mean, float* running_inv_sd, float* saved_mean, float* saved_inv_sd, float* y);
int cattl3_batchnorm_forward_f64(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, int training, int running_initialised, double decay, double eps, const double* x, const double* gamma, const double* beta, double* running_mean, double* running_inv_sd, double* saved_mean, double* saved_inv_sd, double* y);
/*
 * Training forward pass whose statistics reduction already happened in the producing kernel layer's epilogue:
 * col_stats (2 * groups doubles, cattl3_epilogue) are the sums of (x - shift) and (x - shift)^2 per group, shift = the
 * producer's bias (groups elements).  mean = shift + S1/L, var = S2/L - (S1/L)^2 in double, where L = n*h*w (or n)
 * unless global_count (a DEVICE scalar: the elements behind all-reduced sums, see cattl3_batchnorm_stats) is
 * given; the rest (saved
 * statistics, running averages, y) is cattl3_batchnorm_forward with training = 1.  The normalise pass can apply a
 * following element-wise activation as well: act_kind != CATTL3_ACT_NONE writes act_out = f(y) (y, the activation's
 * cached input, is still written unless NULL).
 */
int cattl3_batchnorm_forward_stats_f32(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, int running_initialised, float decay, float eps, const float* x, const double* col_stats, const double* global_count, const float* shift, const float* gamma, const float* beta, float* running_mean, float* running_inv_sd, float* saved_mean, float* saved_inv_sd, float* y, int act_kind, float act_param, float* act_out);
int cattl3_batchnorm_forward_stats_f64(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, int running_initialised, double decay, double eps, const double* x, const double* col_stats, const double* global_count, const double* shift, const double* gamma, const double* beta, double* running_mean, double* running_inv_sd, double* saved_mean, double* saved_inv_sd, double* y, int act_kind, double act_param, double* act_out);
/*
 * The per-group shifted sums on their own (what cattl3_epilogue::col_stats holds): col_stats[g] = sum (x - shift[g]),
 * col_stats[groups + g] = sum (x - shift[g])^2, in double.  For batch-norm inputs that do not come out of a kernel
 * layer's epilogue and for SYNCHRONISED statistics in data-parallel training: all-reduce the sums
 * over the ranks together with the element count and call cattl3_batchnorm_forward_stats with global_count.
 */
int cattl3_batchnorm_stats_f32(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const float* x, const float* shift, double* col_stats);
int cattl3_batchnorm_stats_f64(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const double* x, const double* shift, double* col_stats);
/*
 * cattl3_batchnorm_backward in two halves, for synchronised statistics: _sums writes sums[g] = sum dy,
 * sums[groups + g] = sum dy * xhat (double) and ACCUMULATES the local sums into dbeta / dgamma (the gradient
 * all-reduce adds the ranks' shares later); _apply computes dx from sums that were all-reduced in between,
 * global_count = a DEVICE scalar holding the elements per group over all ranks, all-reduced with the sums so that
 * no host synchronisation is needed (NULL = the local count) (BatchNormLayer.hpp:257-261 with L = the global batch).
 */
int cattl3_batchnorm_backward_sums_f32(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const float* x, const float* saved_mean, const float* saved_inv_sd, const float* dy, float* dgamma, float* dbeta, double* sums);
int cattl3_batchnorm_backward_sums_f64(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const double* x, const double* saved_mean, const double* saved_inv_sd, const double* dy, double* dgamma, double* dbeta, double* sums);
int cattl3_batchnorm_backward_apply_f32(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const double* global_count, const float* x, const float* gamma, const float* saved_mean, const float* saved_inv_sd, const float* dy, const double* sums, float* dx);
int cattl3_batchnorm_backward_apply_f64(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const double* global_count, const double* x, const double* gamma, const double* saved_mean, const double* saved_inv_sd, const double* dy, const double* sums, double* dx);
/* dgamma / dbeta ACCUMULATE; dx may be NULL (input layer). */
int cattl3_batchnorm_backward_f32(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const float* x, const float* gamma, const float* saved_mean, const float* saved_inv_sd, const float* dy, float* dgamma, float* dbeta, float* dx);
int cattl3_batchnorm_backward_f64(cattl3_ctx*, int per_channel, int32_t n, int32_t h, int32_t w, int32_t c, const double* x, const double* gamma, const double* saved_mean, const double* saved_inv_sd, const double* dy, double* dgamma, double* dbeta, double* dx);

/* ---- optimizer step (SGDOptimizer::_train's regularize -> _update_params -> reset_grad) ----- */
/* One fused launch over `count` contiguous parameters (a whole network's parameter arena).
 * s1/s2/s3: optimizer state vectors (unused ones may be NULL). */
int cattl3_optimizer_step_f32(cattl3_ctx*, const cattl3_opt_step*, int64_t count, float* p, float* g, float* s1, float* s2, float* s3);
int cattl3_optimizer_step_f64(cattl3_ctx*, const cattl3_opt_step*, int64_t count, double* p, double* g, double* s1, double* s2, double* s3);
/* The same update with the scalars of `cattl3_opt_step` read from DEVICE memory at run time (only `kind` is a host
 * argument): what a captured step graph replays. */
int cattl3_optimizer_step_indirect_f32(cattl3_ctx*, int kind, const cattl3_opt_step* dev_step, int64_t count, float* p, float* g, float* s1, float* s2, float* s3);
int cattl3_optimizer_step_indirect_f64(cattl3_ctx*, int kind, const cattl3_opt_step* dev_step, int64_t count, double* p, double* g, double* s1, double* s2, double* s3);

/* Parameter regularisation (C-ATTL3/parameter_regularization/{L1,L2,ElasticNet}ParameterRegularization.hpp behind
 * Parameters::regularize / get_regularization_penalty, C-ATTL3/parameters/StandardParameters.hpp:126-136):
 *   grad[i]  += (values[i] >= 0 ? l1 : -l1) + l2 * values[i]            (grad may be NULL: penalty only)
 *   *penalty += l1 * sum |values| + l2 / 2 * sum values^2               (device double; may be NULL; deterministic)
 * L1 = (lambda, 0), L2 = (0, lambda), ElasticNet = (l1_lambda, l2_lambda). */
int cattl3_regularize_f32(cattl3_ctx*, int64_t count, float l1, float l2, const float* values, float* grad, double* penalty);
int cattl3_regularize_f64(cattl3_ctx*, int64_t count, double l1, double l2, const double* values, double* grad, double* penalty);

/* ---- small element-wise helpers for the network glue ---------------------------------------- */
/* The constraints of StandardParameters on a device array, values (set_values, StandardParameters.hpp:105-111) or
 * gradient (accumulate_grad, :115-123), in the reference's order and with its definitions (:150-182): clip every element to
 * [-clip, clip]; the "L1" limit is compared with the FROBENIUS norm and rescales by max / norm; the "L2" limit is compared
 * with the SQUARED norm and rescales by max / squared norm.  0 switches a limit off. */
int cattl3_constrain_f32(cattl3_ctx*, int64_t count, float clip, float max_l1_norm, float max_l2_norm, float* x);
int cattl3_constrain_f64(cattl3_ctx*, int64_t count, double clip, double max_l1_norm, double max_l2_norm, double* x);
/* y += x (ResidualNeuralNetwork::propagate, C-ATTL3/neural_network/ResidualNeuralNetwork.hpp:112-117). */
int cattl3_add_inplace_f32(cattl3_ctx*, int64_t count, float* y, const float* x);
int cattl3_add_inplace_f64(cattl3_ctx*, int64_t count, double* y, const double* x);
/* y *= x: the PARALLEL_MUL merge of ParallelNeuralNetwork (ParallelNeuralNetwork.hpp:170-173, 286-291). */
int cattl3_mul_inplace_f32(cattl3_ctx*, int64_t count, float* y, const float* x);
int cattl3_mul_inplace_f64(cattl3_ctx*, int64_t count, double* y, const double* x);
/* out = (accumulate ? out : 0) + a * b + (c ? c * d : 0), c and d both NULL or both set; `out` may alias an operand.  The gate
 * arithmetic of an LSTM cell and of its backward pass (C-ATTL3/neural_network/LSTMNeuralNetwork.hpp:290-296, 369-371,
 * 441-447): products and sums are rounded separately, as in the reference's Eigen expressions. */
int cattl3_muladd_f32(cattl3_ctx*, int64_t count, int accumulate, const float* a, const float* b, const float* c, const float* d, float* out);
int cattl3_muladd_f64(cattl3_ctx*, int64_t count, int accumulate, const double* a, const double* b, const double* c, const double* d, double* out);
/* y = alpha * x (the 1/batch_size scaling of the loss gradient, SGDOptimizer.hpp:55-56). */
int cattl3_scale_f32(cattl3_ctx*, int64_t count, float alpha, const float* x, float* y);
int cattl3_scale_f64(cattl3_ctx*, int64_t count, double alpha, const double* x, double* y);
/* y += alpha * x (Parameters::regularize() with an L2 penalty: grad += lambda * values,
 * C-ATTL3/parameter_regularization/L2ParameterRegularization.hpp:31-33). */
int cattl3_axpy_f32(cattl3_ctx*, int64_t count, float alpha, const float* x, float* y);
int cattl3_axpy_f64(cattl3_ctx*, int64_t count, double alpha, const double* x, double* y);

/* ---- data-parallel exchange (no counterpart in the reference, which is single-process: SURVEY.md F6) --- */
/*
 * One communicator per process (one process per GPU).  The only exchange of the hot path is the sum
 * all-reduce of the parameter gradients between ConvKernelLayer::pass_back and the optimizer step
 * (between C-ATTL3/optimizer/SGDOptimizer.hpp:56 and :59), done in place with NCCL on the context's
 * stream.  world_size 1 makes every call a no-op, so single-GPU code needs no NCCL library.
 */
typedef struct cattl3_comm cattl3_comm;
/* 128-byte ncclUniqueId for rank 0 to hand to its peers. */
int cattl3_comm_unique_id(void* id128);
int cattl3_comm_create(cattl3_comm** out, cattl3_ctx* ctx, int world_size, int rank, const void* id128);
/* WORLD_SIZE / RANK from the environment (torchrun convention); the id travels through the file
 * CATTL3_COMM_ID_FILE (default /tmp/cattl3_nccl_id.<MASTER_PORT>.<parent pid>[.<TORCHELASTIC_RUN_ID>]: tied to the launch). */
int cattl3_comm_create_from_env(cattl3_comm** out, cattl3_ctx* ctx);
int cattl3_comm_destroy(cattl3_comm* comm);
int cattl3_comm_world_size(const cattl3_comm* comm);
int cattl3_comm_rank(const cattl3_comm* comm);
int cattl3_comm_group_start(cattl3_comm* comm);
int cattl3_comm_group_end(cattl3_comm* comm);
int cattl3_comm_allreduce_sum_f32(cattl3_comm* comm, float* dev_buf, int64_t count);
int cattl3_comm_allreduce_sum_f64(cattl3_comm* comm, double* dev_buf, int64_t count);
/* The same exchange on the communicator's own high-priority stream: it starts once everything enqueued on the
 * context's stream so far has finished and runs beside whatever is enqueued next (a layer's weight gradients travel
 * while its input gradient and the layers behind it compute); cattl3_comm_wait() makes the context's stream wait for
 * every exchange started this way (call it before the optimizer step). */
int cattl3_comm_allreduce_sum_async_f32(cattl3_comm* comm, float* buf, int64_t count);
int cattl3_comm_allreduce_sum_async_f64(cattl3_comm* comm, double* buf, int64_t count);
int cattl3_comm_wait(cattl3_comm* comm);

#ifdef __cplusplus
}
#endif

#endif /* CATTL3_B200_H_ */
