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
