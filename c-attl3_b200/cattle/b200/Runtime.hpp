/*
 * b200/Runtime.hpp -- C++ face of the C ABI in include/cattl3_b200.h: status codes become
 * exceptions, device memory becomes RAII, and the float / double entry points are selected by
 * the Scalar template parameter.  Header-only, no CUDA headers needed: g++ is enough.
 *
 * The shape follows the reference's own (unfinished) device plumbing: status -> exception as in
 * C-ATTL3/core/gpu/cuda/CUDAError.hpp:17-52, owning device array as in
 * C-ATTL3/core/gpu/cuda/CUDAArray.hpp:46-118 -- but every transfer here is enqueued on one
 * stream and only device-to-host reads synchronise.
 */
#ifndef C_ATTL3_B200_RUNTIME_H_
#define C_ATTL3_B200_RUNTIME_H_

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>

#include "cattl3_b200.h"

namespace cattle {
namespace b200 {

/** Thrown whenever a libcattl3_b200 call fails; there is no fallback path to fall back to. */
class Error : public std::runtime_error {
public:
	inline Error(int code, const std::string& what) :
			std::runtime_error(what),
			code(code) { }
	const int code;
};

inline void check(int rc, const char* expr, const char* file, int line) {
	if (rc != CATTL3_OK) {
		throw Error(rc, std::string("libcattl3_b200: ") + cattl3_last_error() + " [" + expr + " at " +
				file + ":" + std::to_string(line) + "]");
	}
}

#define CATTLE_B200_CHECK(expr) ::cattle::b200::check((expr), #expr, __FILE__, __LINE__)

/**
 * The process-wide device context: one device (CATTL3_DEVICE, else LOCAL_RANK, else 0 -- one
 * process per GPU), one stream.  All layers of all networks enqueue on that stream, so work
 * issued from the pthread lanes of a ParallelNeuralNetwork
 * (C-ATTL3/neural_network/ParallelNeuralNetwork.hpp:142-197) is ordered with respect to the
 * optimizer's parameter updates without any cross-stream events; the lock serialises the
 * host-side bookkeeping of the context's workspaces.
 */
class Context {
public:
	typedef std::unique_lock<std::recursive_mutex> Lock;
	inline static Context& get() {
		static Context instance;
		return instance;
	}
	inline cattl3_ctx* handle() const {
		return ctx;
	}
	inline Lock lock() {
		return Lock(mutex);
	}
	inline int device() const {
		return device_index;
	}
	inline void synchronize() {
		Lock l(mutex);
		CATTLE_B200_CHECK(cattl3_ctx_synchronize(ctx));
	}
	inline std::int64_t launch_count() const {
		return cattl3_ctx_launch_count(ctx);
	}
	inline const char* last_path() const {
		return cattl3_ctx_last_path(ctx);
	}
	/** CATTL3_PATH_AUTO / _SIMT / _TCGEN05. */
	inline void set_conv_path(int path) {
		Lock l(mutex);
		CATTLE_B200_CHECK(cattl3_ctx_set_conv_path(ctx, path));
	}
	Context(const Context&) = delete;
	Context& operator=(const Context&) = delete;
private:
	inline Context() :
			ctx(nullptr),
			device_index(0) {
		const char* dev = std::getenv("CATTL3_DEVICE");
		if (!dev)
			dev = std::getenv("LOCAL_RANK");
		if (dev)
			device_index = std::atoi(dev);
		CATTLE_B200_CHECK(cattl3_ctx_create(&ctx, device_index, nullptr));
	}
	inline ~Context() {
		cattl3_ctx_destroy(ctx);
	}
	cattl3_ctx* ctx;
	int device_index;
	std::recursive_mutex mutex;
};

/** Scalar -> `_f32` / `_f64` entry points. */
/**
 * While one of these lives no parameter changes (a training step up to its update): the kernel layers keep their repacked
 * weights per (array, geometry) and reuse them, e.g. across the time steps of an unrolled LSTM (cattl3_weights_stable_begin).
 */
struct WeightsStable {
	inline WeightsStable() {
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_weights_stable_begin(c.handle()));
	}
	inline ~WeightsStable() {
		Context& c = Context::get();
		Context::Lock l = c.lock();
		cattl3_weights_stable_end(c.handle());
	}
	WeightsStable(const WeightsStable&) = delete;
	WeightsStable& operator=(const WeightsStable&) = delete;
};

template<typename Scalar> struct Api;

#define CATTLE_B200_API(S, SUF) \
template<> struct Api<S> { \
	static int conv_forward(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y) { \
		return cattl3_conv_forward_##SUF(c, g, x, w, b, y); } \
	static int conv_forward_fused(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y, const cattl3_epilogue* ep) { \
		return cattl3_conv_forward_fused_##SUF(c, g, x, w, b, y, ep); } \
	static int transconv_forward_fused(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y, const cattl3_epilogue* ep) { \
		return cattl3_transconv_forward_fused_##SUF(c, g, x, w, b, y, ep); } \
	static int dense_forward_fused(cattl3_ctx* c, std::int32_t n, std::int32_t in, std::int32_t out, const S* x, const S* w, const S* b, S* y, const cattl3_epilogue* ep) { \
		return cattl3_dense_forward_fused_##SUF(c, n, in, out, x, w, b, y, ep); } \
	static int batchnorm_forward_stats(cattl3_ctx* c, int pc, std::int32_t n, std::int32_t h, std::int32_t w, std::int32_t ch, int init, S decay, S eps, const S* x, const double* cs, const double* gc, const S* shift, const S* gamma, const S* beta, S* rm, S* rs, S* sm, S* ss, S* y, int ak, S ap, S* ao) { \
		return cattl3_batchnorm_forward_stats_##SUF(c, pc, n, h, w, ch, init, decay, eps, x, cs, gc, shift, gamma, beta, rm, rs, sm, ss, y, ak, ap, ao); } \
	static int fill(cattl3_ctx* c, std::int64_t count, S value, S* y) { \
		return cattl3_fill_##SUF(c, count, value, y); } \
	static int constrain(cattl3_ctx* c, std::int64_t count, S clip, S max_l1, S max_l2, S* x) { \
		return cattl3_constrain_##SUF(c, count, clip, max_l1, max_l2, x); } \
	static int slice_rows(cattl3_ctx* c, std::int64_t total, std::int64_t vol, std::int64_t first, std::int64_t rows, const S* src, S* dst) { \
		return cattl3_slice_rows_##SUF(c, total, vol, first, rows, src, dst); } \
	static int dropout_forward(cattl3_ctx* c, std::int64_t count, S prob, S eps, std::uint64_t seed, const S* x, S* y, std::uint8_t* mask) { \
		return cattl3_dropout_forward_##SUF(c, count, prob, eps, seed, x, y, mask); } \
	static int dropout_backward(cattl3_ctx* c, std::int64_t count, S prob, S eps, const S* dy, const std::uint8_t* mask, S* dx) { \
		return cattl3_dropout_backward_##SUF(c, count, prob, eps, dy, mask, dx); } \
	static int loss(cattl3_ctx* c, int kind, std::int64_t rows, std::int64_t vol, S eps, S grad_div, const S* out, const S* obj, S* loss, S* grad) { \
		return cattl3_loss_##SUF(c, kind, rows, vol, eps, grad_div, out, obj, loss, grad); } \
	static int batchnorm_stats(cattl3_ctx* c, int pc, std::int32_t n, std::int32_t h, std::int32_t w, std::int32_t ch, const S* x, const S* shift, double* cs) { \
		return cattl3_batchnorm_stats_##SUF(c, pc, n, h, w, ch, x, shift, cs); } \
	static int batchnorm_backward_sums(cattl3_ctx* c, int pc, std::int32_t n, std::int32_t h, std::int32_t w, std::int32_t ch, const S* x, const S* sm, const S* ss, const S* dy, S* dg, S* db, double* sums) { \
		return cattl3_batchnorm_backward_sums_##SUF(c, pc, n, h, w, ch, x, sm, ss, dy, dg, db, sums); } \
	static int batchnorm_backward_apply(cattl3_ctx* c, int pc, std::int32_t n, std::int32_t h, std::int32_t w, std::int32_t ch, const double* total, const S* x, const S* gamma, const S* sm, const S* ss, const S* dy, const double* sums, S* dx) { \
		return cattl3_batchnorm_backward_apply_##SUF(c, pc, n, h, w, ch, total, x, gamma, sm, ss, dy, sums, dx); } \
	static int conv_backward(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* dy, S* dw, S* db, S* dx) { \
		return cattl3_conv_backward_##SUF(c, g, x, w, dy, dw, db, dx); } \
	static int transconv_forward(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* b, S* y) { \
		return cattl3_transconv_forward_##SUF(c, g, x, w, b, y); } \
	static int transconv_backward(cattl3_ctx* c, const cattl3_conv_geom* g, const S* x, const S* w, const S* dy, S* dw, S* db, S* dx) { \
		return cattl3_transconv_backward_##SUF(c, g, x, w, dy, dw, db, dx); } \
	static int dense_forward(cattl3_ctx* c, std::int32_t n, std::int32_t in, std::int32_t out, const S* x, const S* w, const S* b, S* y) { \
		return cattl3_dense_forward_##SUF(c, n, in, out, x, w, b, y); } \
	static int dense_backward(cattl3_ctx* c, std::int32_t n, std::int32_t in, std::int32_t out, const S* x, const S* w, const S* dy, S* dw, S* db, S* dx) { \
		return cattl3_dense_backward_##SUF(c, n, in, out, x, w, dy, dw, db, dx); } \
	static int activation_forward(cattl3_ctx* c, int kind, S alpha, std::int64_t rows, std::int64_t vol, const S* x, S* y) { \
		return cattl3_activation_forward_##SUF(c, kind, alpha, rows, vol, x, y); } \
	static int activation_backward(cattl3_ctx* c, int kind, S alpha, std::int64_t rows, std::int64_t vol, const S* x, const S* y, const S* dy, S* dx) { \
		return cattl3_activation_backward_##SUF(c, kind, alpha, rows, vol, x, y, dy, dx); } \
	static int pool_forward(cattl3_ctx* c, int kind, const cattl3_pool_geom* g, const S* x, S* y, std::uint8_t* arg) { \
		return cattl3_pool_forward_##SUF(c, kind, g, x, y, arg); } \
	static int pool_backward(cattl3_ctx* c, int kind, const cattl3_pool_geom* g, const S* dy, const std::uint8_t* arg, S* dx) { \
		return cattl3_pool_backward_##SUF(c, kind, g, dy, arg, dx); } \
	static int batchnorm_forward(cattl3_ctx* c, int pc, std::int32_t n, std::int32_t h, std::int32_t w, std::int32_t ch, int training, int init, S decay, S eps, const S* x, const S* gamma, const S* beta, S* rm, S* rs, S* sm, S* ss, S* y) { \
		return cattl3_batchnorm_forward_##SUF(c, pc, n, h, w, ch, training, init, decay, eps, x, gamma, beta, rm, rs, sm, ss, y); } \
	static int batchnorm_backward(cattl3_ctx* c, int pc, std::int32_t n, std::int32_t h, std::int32_t w, std::int32_t ch, const S* x, const S* gamma, const S* sm, const S* ss, const S* dy, S* dg, S* db, S* dx) { \
		return cattl3_batchnorm_backward_##SUF(c, pc, n, h, w, ch, x, gamma, sm, ss, dy, dg, db, dx); } \
	static int optimizer_step(cattl3_ctx* c, const cattl3_opt_step* st, std::int64_t count, S* p, S* g, S* s1, S* s2, S* s3) { \
		return cattl3_optimizer_step_##SUF(c, st, count, p, g, s1, s2, s3); } \
	static int optimizer_step_indirect(cattl3_ctx* c, int kind, const cattl3_opt_step* dev, std::int64_t count, S* p, S* g, S* s1, S* s2, S* s3) { \
		return cattl3_optimizer_step_indirect_##SUF(c, kind, dev, count, p, g, s1, s2, s3); } \
	static int add_inplace(cattl3_ctx* c, std::int64_t count, S* y, const S* x) { \
		return cattl3_add_inplace_##SUF(c, count, y, x); } \
	static int mul_inplace(cattl3_ctx* c, std::int64_t count, S* y, const S* x) { \
		return cattl3_mul_inplace_##SUF(c, count, y, x); } \
	static int regularize(cattl3_ctx* c, std::int64_t count, S l1, S l2, const S* values, S* grad, double* penalty) { \
		return cattl3_regularize_##SUF(c, count, l1, l2, values, grad, penalty); } \
	static int muladd(cattl3_ctx* c, std::int64_t count, int accumulate, const S* a, const S* b, const S* cc, const S* d, S* out) { \
		return cattl3_muladd_##SUF(c, count, accumulate, a, b, cc, d, out); } \
	static int scale(cattl3_ctx* c, std::int64_t count, S alpha, const S* x, S* y) { \
		return cattl3_scale_##SUF(c, count, alpha, x, y); } \
	static int axpy(cattl3_ctx* c, std::int64_t count, S alpha, const S* x, S* y) { \
		return cattl3_axpy_##SUF(c, count, alpha, x, y); } \
};

CATTLE_B200_API(float, f32)
CATTLE_B200_API(double, f64)
#undef CATTLE_B200_API

/**
 * An owning array of `count` Scalars in the HBM of the context's device.  Copying is a deep
 * device-to-device copy (layers are cloned with their caches, e.g.
 * C-ATTL3/layer/kernel/DenseKernelLayer.hpp:78-82); moving transfers ownership.
 */
template<typename Scalar>
class DeviceBuffer {
public:
	inline DeviceBuffer() :
			ptr(nullptr),
			count(0),
			owned(true) { }
	inline explicit DeviceBuffer(std::size_t count, bool zero = false) :
			ptr(nullptr),
			count(count),
			owned(true) {
		if (count == 0)
			return;
		Context& c = Context::get();
		Context::Lock l = c.lock();
		void* p = nullptr;
		CATTLE_B200_CHECK(cattl3_malloc(c.handle(), &p, count * sizeof(Scalar)));
		ptr = static_cast<Scalar*>(p);
		if (zero)
			CATTLE_B200_CHECK(cattl3_memset(c.handle(), ptr, 0, count * sizeof(Scalar)));
	}
	/** A non-owning view of `count` elements at `borrowed` (memory owned by an InputFeed); never freed here. */
	inline static DeviceBuffer<Scalar> view(Scalar* borrowed, std::size_t count) {
		DeviceBuffer<Scalar> buffer;
		buffer.ptr = borrowed;
		buffer.count = count;
		buffer.owned = false;
		return buffer;
	}
	inline DeviceBuffer(const DeviceBuffer<Scalar>& other) :
			DeviceBuffer(other.count) {
		if (count > 0) {
			Context& c = Context::get();
			Context::Lock l = c.lock();
			CATTLE_B200_CHECK(cattl3_memcpy_d2d(c.handle(), ptr, other.ptr, count * sizeof(Scalar)));
		}
	}
	inline DeviceBuffer(DeviceBuffer<Scalar>&& other) noexcept :
			ptr(other.ptr),
			count(other.count),
			owned(other.owned) {
		other.ptr = nullptr;
		other.count = 0;
		other.owned = true;
	}
	inline ~DeviceBuffer() {
		release();
	}
	inline DeviceBuffer<Scalar>& operator=(DeviceBuffer<Scalar> other) noexcept {
		std::swap(ptr, other.ptr);
		std::swap(count, other.count);
		std::swap(owned, other.owned);
		return *this;
	}
	inline Scalar* data() {
		return ptr;
	}
	inline const Scalar* data() const {
		return ptr;
	}
	inline std::size_t size() const {
		return count;
	}
	inline bool empty() const {
		return count == 0;
	}
	inline void zero() {
		if (count == 0)
			return;
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_memset(c.handle(), ptr, 0, count * sizeof(Scalar)));
	}
	/** Host -> device, `n` elements into [offset, offset + n). */
	inline void upload(const Scalar* host, std::size_t n, std::size_t offset = 0) {
		if (n == 0)
			return;
		if (offset + n > count)
			throw Error(CATTL3_ERR_INVALID, "DeviceBuffer::upload out of range");
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_memcpy_h2d(c.handle(), ptr + offset, host, n * sizeof(Scalar)));
	}
	/** Device -> host (synchronises the stream). */
	inline void download(Scalar* host, std::size_t n, std::size_t offset = 0) const {
		if (n == 0)
			return;
		if (offset + n > count)
			throw Error(CATTL3_ERR_INVALID, "DeviceBuffer::download out of range");
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_memcpy_d2h(c.handle(), host, ptr + offset, n * sizeof(Scalar)));
	}
private:
	inline void release() {
		if (ptr && !owned) {
			ptr = nullptr;
			count = 0;
			owned = true;
		}
		if (ptr) {
			Context& c = Context::get();
			Context::Lock l = c.lock();
			cattl3_free(c.handle(), ptr);  // destructors must not throw
			ptr = nullptr;
			count = 0;
		}
	}
	Scalar* ptr;
	std::size_t count;
	bool owned;
};

/**
 * The mini-batch upload of the batch loop, off the compute stream (cattl3_feed, include/cattl3_b200.h): a ring of
 * device buffers filled through pinned staging on a copy stream, so that the upload of the next batch overlaps the
 * kernels of the current one.  push() returns once the host data has been staged; the returned view stays valid for
 * the next `slots - 1` pushes, i.e. for the whole training step that consumes it.
 */
class InputFeed {
public:
	inline explicit InputFeed(int slots = 3) :
			feed(nullptr) {
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_feed_create(&feed, c.handle(), slots));
	}
	inline ~InputFeed() {
		Context::Lock l = Context::get().lock();
		cattl3_feed_destroy(feed);
	}
	InputFeed(const InputFeed&) = delete;
	InputFeed& operator=(const InputFeed&) = delete;
	/** The process-wide feeds of the batch loop (0: observations, 1: objectives). */
	inline static InputFeed& shared(int which) {
		static InputFeed feeds[2];
		return feeds[which & 1];
	}
	template<typename Scalar>
	inline DeviceBuffer<Scalar> push(const Scalar* host, std::size_t count) {
		void* dev = nullptr;
		Context::Lock l = Context::get().lock();
		CATTLE_B200_CHECK(cattl3_feed_push(feed, host, count * sizeof(Scalar), &dev));
		return DeviceBuffer<Scalar>::view(static_cast<Scalar*>(dev), count);
	}
private:
	cattl3_feed* feed;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_RUNTIME_H_ */
