// common.cuh -- context, error handling and launch helpers shared by all translation units of
// libcattl3_b200.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cattl3_b200.h"

struct cattl3_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	int conv_path = CATTL3_PATH_AUTO;
	int64_t launches = 0;
	const char* last_path = "none";
	int sm_count = 148;
	// general scratch (split-K partials, batch-norm partial sums); grown on demand, never shrunk
	void* ws = nullptr;
	size_t ws_bytes = 0;
	// tcgen05 path scratch: the weights repacked K-major and split hi | lo for TMA
	void* tc_w = nullptr;
	size_t tc_w_bytes = 0;
	// Between cattl3_weights_stable_begin and _end the caller promises not to change any weights: a kernel layer's packed
	// weights are then kept per (weight array, geometry) and reused by later calls (the cells of an unrolled LSTM share their
	// kernels' weights: one repack per pass instead of one per time step).  Slots are handed out in call order and keep their
	// buffers from scope to scope, so a step that repeats the previous one allocates nothing (a captured step never may).
	struct PackKey { const void* w; int RH, RW, SC, J, r_pad, j_pad, tapbox, pad_; long long w_off, w_stap, w_srw, w_sr, w_sj; };
	struct PackSlot { void* buf; size_t bytes; PackKey key; };
	static constexpr int MAX_PACK_SLOTS = 96;
	PackSlot pack_slots[MAX_PACK_SLOTS] = {};
	int pack_used = 0;
	bool pack_stable = false;
	// per-CTA column sums of a fused batch-norm statistics epilogue (conv_tc.cu)
	void* stat_ws = nullptr;
	size_t stat_ws_bytes = 0;
	// cattl3_regularize: 256 per-block partial penalties + the "blocks done" counter
	void* reg_ws = nullptr;
	// cattl3_constrain: 256 per-block partial squared norms, the counter, the scale factor
	void* con_ws = nullptr;
	// between cattl3_graph_begin and cattl3_graph_end: the stream is capturing (nothing may synchronise or grow scratch)
	bool capturing = false;
	// Step graphs keep their activations in a private arena (graphs with allocation nodes launch slowly): while capturing,
	// cattl3_malloc carves blocks out of arena `cap_arena` and cattl3_free recycles them (one stream: program order makes
	// the reuse safe); outside a capture cattl3_free of an arena address is a no-op.  An arena is owned by its graph;
	// cattl3_graph_destroy retires it (in_use = false) for the next cattl3_graph_begin, the memory goes at ctx_destroy.
	static constexpr int MAX_CAP_BLOCKS = 8192, MAX_ARENAS = 64;
	struct Arena { char* base; size_t size; bool in_use; };
	Arena arenas[MAX_ARENAS];
	int arena_count = 0;
	int cap_arena = -1;
	size_t cap_used = 0;
	struct CapBlock { char* ptr; size_t size; bool free; };
	CapBlock cap_blocks[MAX_CAP_BLOCKS];
	int cap_block_count = 0;
	// pool memory released while the stream is capturing: the release must not become a node of the graph (a replay would
	// free it again), so it waits here until the capture has ended
	static constexpr int MAX_DEFERRED_FREES = 4096;
	void* deferred_free[MAX_DEFERRED_FREES];
	int deferred_free_count = 0;
	int64_t scratch_generation = 0;   // bumped whenever ensure_buffer moves a scratch buffer (captured graphs hold the old address)
	int64_t allocated_bytes = 0;   // running total of cattl3_malloc in 256-byte granules (sizes an arena from an eager step)
	// cattl3_ctx_throttle: one event per recent call
	cudaEvent_t throttle_ev[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	long long throttle_calls = 0;
	// device staging for the *_host entry points (x, y, dX, dY, a second dX for chunked input gradients) and their
	// transfer pipeline: an upload and a download stream beside `stream`, chained by events (api.cu)
	void* stage_dev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
	size_t stage_dev_bytes[5] = { 0, 0, 0, 0, 0 };
	cudaStream_t up_stream = nullptr, down_stream = nullptr;
	static constexpr int HOST_MAX_CHUNKS = 16;
	enum { EV_ENTRY = 0, EV_UP, EV_Y_FREE, EV_DY_FREE, EV_DX_FREE, EV_CHUNK, HOST_EVENTS = EV_CHUNK + HOST_MAX_CHUNKS };
	cudaEvent_t host_ev[HOST_EVENTS] = {};
	bool host_ev_used[HOST_EVENTS] = {};
};

namespace cattl3 {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int ensure_buffer(cattl3_ctx* ctx, void** buf, size_t* cur, size_t need);

#define CATTL3_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
	return cattl3::cuda_fail(e__, #call, __FILE__, __LINE__); } while (0)

#define CATTL3_REQUIRE(cond, ...) do { if (!(cond)) { cattl3::set_error(__VA_ARGS__); \
	return CATTL3_ERR_INVALID; } } while (0)

#define CATTL3_CHECK(expr) do { int rc__ = (expr); if (rc__ != CATTL3_OK) return rc__; } while (0)

// Checks the launch and counts it (bench.py reports the count as gpu_launches).
#define CATTL3_LAUNCHED(ctx) do { (ctx)->launches++; cudaError_t e__ = cudaGetLastError(); \
	if (e__ != cudaSuccess) return cattl3::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); } while (0)

inline int check_ctx(cattl3_ctx* ctx) {
	if (!ctx) {
		set_error("null context");
		return CATTL3_ERR_INVALID;
	}
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e != cudaSuccess)
		return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__);
	return CATTL3_OK;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid for a grid-stride element-wise kernel: enough CTAs for `work` items at `per_block`
// items each, capped at a multiple of the SM count (8 resident 256-thread CTAs per SM).
inline int ew_grid(const cattl3_ctx* ctx, int64_t work, int per_block) {
	int64_t blocks = ceil_div(work, per_block);
	int64_t cap = (int64_t) ctx->sm_count * 8;
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	return (int) blocks;
}

// ---- kernel-layer entry points implemented per translation unit ------------------------------

// Description of one implicit-GEMM pass over a gathered tensor.  See conv_simt.cu.
struct GatherGeom {
	int N;              // batch (fastest dim of every tensor)
	int SH, SW, SC;     // gathered ("source") tensor spatial dims and channels (reduce channels R)
	int OH, OW;         // grid of the output / plain tensor; M = N*OH*OW
	int J;              // output channels
	int RH, RW;         // taps
	// source coordinate: t = o*a + r*b + c; valid iff t % den == 0 and 0 <= t/den < S
	int ah, bh, ch, denh;
	int aw, bw, cw, denw;
	// weight element (tap = rh + RH*rw, reduce channel r, output channel j)
	long long w_stap, w_sr, w_sj;
	// ---- used by the tcgen05 gather GEMM only (sub-problems of a strided transposed gather, api.cu) ----
	// weight element = w_off + rh*w_stap + rw*w_srw + r*w_sr + j*w_sj; w_srw == 0 means w_stap*RH
	long long w_off = 0, w_srw = 0;
	// output lattice: element (n, i, j) of the OH x OW grid is stored at pixel (out_h0 + out_hs*i, out_w0 + out_ws*j)
	// of an out_H x out_W tensor; out_H == 0 means the dense OH x OW tensor itself
	int out_h0 = 0, out_hs = 1, out_H = 0, out_w0 = 0, out_ws = 1, out_W = 0;
	// ---- used by the tcgen05 weight gradient only: the reduction covers rows [m_first, m_first + m_count) of the M = N*OH*OW rows
	// of the plain tensor (a strip of output columns while the rest is still being uploaded, api.cu); m_count == 0 means all
	long long m_first = 0, m_count = 0;
};

// A kernel layer's fused epilogue (cattl3_epilogue, already validated): activation and / or column statistics.
struct EpilogueArgs {
	int act_kind = CATTL3_ACT_NONE;
	double act_param = 0;
	void* act_out = nullptr;
	double* col_stats = nullptr;
};

// out may be null when ep->act_out is given.  The SIMT kernels fuse the activation only (ep->col_stats is ignored:
// the caller runs colstats_shifted over the finished output).
template<typename S>
int simt_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias,
		int bias_mode, S* out, const EpilogueArgs* ep = nullptr);
// col_stats[j] = sum_m (a[m + rows*j] - shift[j]), col_stats[cols + j] = the sum of squares; double accumulation.
template<typename S>
int colstats_shifted(cattl3_ctx* ctx, int64_t rows, int64_t cols, const S* a, const S* shift, double* col_stats);
template<typename S>
int simt_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw);
template<typename S>
int colsum_accumulate(cattl3_ctx* ctx, int64_t rows, int64_t cols, const S* a, S* out);

// Tiny-channel streaming kernels (conv_simt.cu): 1-8 output columns / K x J <= 2048 gradients (configs 1 and 3).
bool tiny_gather_gemm_supported(const GatherGeom& gg, size_t scalar_bytes);
template<typename S>
int tiny_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias, int bias_mode, S* out,
		const EpilogueArgs* ep = nullptr);
// small batch x long reduction x <= 16 outputs (classifier heads): split-K dense forward
bool skinny_gather_gemm_supported(const cattl3_ctx* ctx, const GatherGeom& gg);
template<typename S>
int skinny_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias, int bias_mode, S* out,
		const EpilogueArgs* ep = nullptr);
bool tiny_wgrad_supported(const GatherGeom& gg, size_t scalar_bytes);
template<typename S> int tiny_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw);

// Big-tile FMA path (conv_dfma.cu): double at GEMM-sized shapes; float where the tensor-core path does not apply.
template<typename S> bool fma_gather_gemm_supported(const GatherGeom& gg);
template<typename S>
int fma_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* w, const S* bias, int bias_mode, S* out,
		const EpilogueArgs* ep = nullptr);
template<typename S> bool fma_wgrad_supported(const GatherGeom& gg);
template<typename S> int fma_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const S* src, const S* plain, S* dw);

// FP64 tensor-core path (conv_dmma.cu): mma.sync.m8n8k4.f64.
bool dmma_gather_gemm_supported(const GatherGeom& gg);
int dmma_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* w, const double* bias,
		int bias_mode, double* out, const EpilogueArgs* ep = nullptr);
bool dmma_wgrad_supported(const GatherGeom& gg);
int dmma_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* plain, double* dw);

// tcgen05 path (conv_tc.cu): returns CATTL3_ERR_UNSUPPORTED when the shape does not qualify.
bool tc_gather_gemm_supported(const cattl3_ctx* ctx, const GatherGeom& gg);
int tc_gather_gemm_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* w,
		const float* bias, int bias_mode, float* out, const EpilogueArgs* ep = nullptr);
bool tc_wgrad_supported(const cattl3_ctx* ctx, const GatherGeom& gg);
// db != null: also accumulates the column sums of `plain` (the bias gradient of a convolution) into db.
int tc_wgrad_f32(cattl3_ctx* ctx, const GatherGeom& gg, const float* src, const float* plain, float* dw, float* db);

} // namespace cattl3
