/*
 * b200/DeviceLayer.hpp -- the device-resident face of a layer.
 *
 * The reference's plug-in boundary is the virtual cattle::Layer API
 * (C-ATTL3/core/Layer.hpp:31-155): host Eigen tensors in, host Eigen tensors out.  Every
 * B200 layer implements that API *and* this mix-in, which passes activations as device tensors so
 * that adjacent B200 layers (and the FeedforwardNeuralNetwork layer loop,
 * C-ATTL3/neural_network/FeedforwardNeuralNetwork.hpp:112-127) never bounce through host memory.
 * It is the same two-level shape the reference sketched for its own GPU layers
 * (C-ATTL3/core/gpu/GPULayer.hpp:19-59: host overloads implemented as H2D -> device call -> D2H).
 */
#ifndef C_ATTL3_B200_DEVICELAYER_H_
#define C_ATTL3_B200_DEVICELAYER_H_

#include <array>
#include <cstddef>
#include <memory>
#include <utility>

#include "core/EigenProxy.hpp"
#include "b200/Runtime.hpp"

namespace cattle {
namespace b200 {

/**
 * A batch of activations in HBM in the reference's own memory order: Eigen column-major,
 * sample index fastest (C-ATTL3/core/EigenProxy.hpp:56-57).  `rows` is the batch size; the
 * remaining extents are implied by the layer that produced / consumes it.
 */
template<typename Scalar>
struct DeviceTensor {
	inline DeviceTensor() :
			rows(0) { }
	inline DeviceTensor(std::size_t rows, std::size_t sample_volume, bool zero = false) :
			buf(std::make_shared<DeviceBuffer<Scalar>>(rows * sample_volume, zero)),
			rows(rows) { }
	inline bool empty() const {
		return !buf || buf->empty();
	}
	inline std::size_t size() const {
		return buf ? buf->size() : 0;
	}
	/** Null for an empty tensor.  Writers must hold the only reference (see make_exclusive()). */
	inline Scalar* data() {
		return buf ? buf->data() : nullptr;
	}
	inline const Scalar* data() const {
		return buf ? buf->data() : nullptr;
	}
	/**
	 * Layers treat their inputs as read-only, so a tensor handed to the next layer and the copy a
	 * layer keeps for its backward pass share one buffer.  Anything that writes in place (the
	 * residual add) first makes sure it is the only owner.
	 */
	inline void make_exclusive() {
		if (buf && buf.use_count() > 1)
			buf = std::make_shared<DeviceBuffer<Scalar>>(*buf);
	}
	std::shared_ptr<DeviceBuffer<Scalar>> buf;
	std::size_t rows;
};

/** Host tensor -> device (one H2D on the context's stream). */
template<typename Scalar, std::size_t DataRank>
inline DeviceTensor<Scalar> to_device(const Tensor<Scalar,DataRank>& t) {
	DeviceTensor<Scalar> d;
	if (t.size() == 0)
		return d;
	d.rows = t.dimension(0);
	d.buf = std::make_shared<DeviceBuffer<Scalar>>(t.size());
	d.buf->upload(t.data(), t.size());
	// the copy engine reads the pageable source through a staging buffer before returning, so `t`
	// may be released by the caller immediately
	return d;
}

/** Device tensor -> host tensor of the given extents (synchronises). */
template<typename Scalar, std::size_t DataRank>
inline Tensor<Scalar,DataRank> to_host(const DeviceTensor<Scalar>& d, const std::array<std::size_t,DataRank>& extents) {
	if (d.empty())
		return Tensor<Scalar,DataRank>();
	Tensor<Scalar,DataRank> t(extents);
	if ((std::size_t) t.size() != d.size())
		throw Error(CATTL3_ERR_INVALID, "to_host: extents do not match the device tensor");
	d.buf->download(t.data(), t.size());
	return t;
}

template<typename Scalar, std::size_t Rank>
class DeviceLayer {
public:
	virtual ~DeviceLayer() = default;
	/**
	 * Layer::pass_forward on device tensors.  The layer takes ownership of `in` (it usually
	 * becomes the cache needed by the backward pass).
	 */
	virtual DeviceTensor<Scalar> pass_forward_dev(DeviceTensor<Scalar> in, bool training) = 0;
	/**
	 * Layer::pass_back on device tensors.  Parameter gradients accumulate on the device; an input
	 * layer returns an empty tensor (C-ATTL3/core/Layer.hpp:82-90).
	 */
	virtual DeviceTensor<Scalar> pass_back_dev(DeviceTensor<Scalar> out_grad) = 0;
};

/**
 * A kernel layer whose backward pass can be taken apart: the input gradient now, the parameter gradients later and for
 * many batches at once.  The cells of an unrolled recurrent network share their kernels' parameters, so
 * dW = sum over steps of in_t^T dY_t is ONE weight-gradient GEMM over the rows of all steps instead of one small GEMM
 * (plus its split-K reduction) per step.
 */
template<typename Scalar, std::size_t Rank>
class SplitBackwardLayer {
public:
	virtual ~SplitBackwardLayer() = default;
	/** Whether the two halves below are available for this layer instance. */
	virtual bool can_split_backward() const = 0;
	/** pass_back_dev without the parameter gradients; empty for an input layer.  Needs no forward cache. */
	virtual DeviceTensor<Scalar> pass_back_input_dev(const DeviceTensor<Scalar>& out_grad) = 0;
	/** dW += in^T dY, db += column sums of dY for `in.rows` (= out_grad.rows, any number of) samples. */
	virtual void accumulate_param_grads_dev(const DeviceTensor<Scalar>& in, const DeviceTensor<Scalar>& out_grad) = 0;
};

/**
 * What the layer FOLLOWING a kernel layer asks that layer's epilogue to do while the output tile is still
 * on chip (cattl3_epilogue, include/cattl3_b200.h): apply its element-wise activation and / or produce the
 * per-column sums a BatchNormLayer starts with.  The consumer fills in the request, the producer the results.
 */
template<typename Scalar>
struct FusedEpilogue {
	// request
	int act_kind = CATTL3_ACT_NONE;
	Scalar act_param = 0;
	bool keep_pre = true;      // the consumer caches the producer's plain output for its backward pass
	bool want_stats = false;
	// results
	DeviceTensor<Scalar> act_out;                      // f(plain output)
	std::shared_ptr<DeviceBuffer<double>> col_stats;   // 2 * columns sums, shifted by `shift`
	const Scalar* shift = nullptr;                     // the producer's bias (device), valid during the step
};

/** A layer whose forward kernel can run a FusedEpilogue (the kernel layers). */
template<typename Scalar>
class EpilogueProducer {
public:
	virtual ~EpilogueProducer() = default;
	virtual bool can_fuse_epilogue() const = 0;
	/** Columns of the layer's GEMM (filters / outputs): the groups column statistics are produced for. */
	virtual std::size_t stat_columns() const = 0;
	/**
	 * pass_forward_dev with the epilogue `ep`.  Returns the plain output (empty when !ep.keep_pre and an
	 * activation was requested); ep.act_out / ep.col_stats / ep.shift are filled in.
	 */
	virtual DeviceTensor<Scalar> pass_forward_dev_fused(DeviceTensor<Scalar> in, bool training, FusedEpilogue<Scalar>& ep) = 0;
};

/** A layer whose forward pass can be (partly) done by its producer's epilogue (activations, batch norm). */
template<typename Scalar>
class EpilogueConsumer {
public:
	virtual ~EpilogueConsumer() = default;
	/** Fills in the request; false = run the layer on its own. */
	virtual bool request_epilogue(FusedEpilogue<Scalar>& ep, std::size_t producer_stat_columns, bool training) const = 0;
	/** Whether accept_epilogue() can itself apply a following activation (`next`). */
	virtual bool chains_epilogue() const {
		return false;
	}
	/**
	 * Completes the layer's forward pass from the producer's results and sets up its backward caches exactly
	 * as pass_forward_dev would.  `next` (only if chains_epilogue()): an activation request of the layer
	 * after this one, to be applied in the same pass.  Returns the layer's own output.
	 */
	virtual DeviceTensor<Scalar> accept_epilogue(DeviceTensor<Scalar> pre, FusedEpilogue<Scalar>& ep, bool training,
			FusedEpilogue<Scalar>* next) = 0;
};

/** extents = { rows, dims... } */
template<std::size_t Rank, typename Dims>
inline std::array<std::size_t,Rank + 1> batch_extents(std::size_t rows, const Dims& dims) {
	std::array<std::size_t,Rank + 1> e;
	e[0] = rows;
	for (std::size_t i = 0; i < Rank; ++i)
		e[i + 1] = dims(i);
	return e;
}

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_DEVICELAYER_H_ */
