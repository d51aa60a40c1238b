/*
 * DeviceSequenceNetwork.hpp -- the device face of sequential networks (NeuralNetwork<Scalar,Rank,true>,
 * C-ATTL3/core/NeuralNetwork.hpp:89-99 on tensors of rank Rank + 2: samples x time steps x dims).
 *
 * In the library's layout (column major, samples fastest) a sequence batch of N samples and T time steps is an
 * (N * T) x volume matrix whose row n + N * t holds time step t of sample n: folding the time steps into the batch
 * (SequentialNeuralNetwork) is a free view, and one time step is a strided block of N-element runs (one 2-D copy).
 * A sequence therefore travels as a DeviceTensor with rows = N * T plus the sample count N.
 */
#ifndef C_ATTL3_B200_DEVICESEQUENCENETWORK_H_
#define C_ATTL3_B200_DEVICESEQUENCENETWORK_H_

#include <array>
#include <cstddef>

#include "core/Layer.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar, std::size_t Rank>
class DeviceSequenceNetwork {
public:
	virtual ~DeviceSequenceNetwork() = default;
	/**
	 * NeuralNetwork<Scalar,Rank,true>::propagate on a device sequence.
	 *
	 * @param input (samples * input time steps) x input volume.
	 * @return (samples * output time steps) x output volume.
	 */
	virtual DeviceTensor<Scalar> propagate_seq_dev(DeviceTensor<Scalar> input, std::size_t samples, bool training) = 0;
	/** backpropagate on a device sequence; an empty tensor for a foremost network. */
	virtual DeviceTensor<Scalar> backpropagate_seq_dev(DeviceTensor<Scalar> out_grad, std::size_t samples) = 0;
	/**
	 * Whether a training step through this network may be captured as a step graph (cattl3_graph): false if anything
	 * on the device is carried from one step to the next behind the graph's back (a stateful LSTM's hidden state).
	 */
	virtual bool graph_safe() const {
		return true;
	}
};

/** Host sequence tensor (samples x steps x dims) -> device sequence (rows = samples * steps). */
template<typename Scalar, std::size_t DataRank>
inline DeviceTensor<Scalar> sequence_to_device(const Tensor<Scalar,DataRank>& t) {
	DeviceTensor<Scalar> d = to_device<Scalar,DataRank>(t);
	if (!d.empty())
		d.rows = (std::size_t) t.dimension(0) * (std::size_t) t.dimension(1);
	return d;
}

/** Device sequence -> host tensor samples x (rows / samples) x dims (synchronises). */
template<typename Scalar, std::size_t Rank, typename Dims>
inline Tensor<Scalar,Rank + 2> sequence_to_host(const DeviceTensor<Scalar>& d, std::size_t samples, const Dims& dims) {
	if (d.empty())
		return Tensor<Scalar,Rank + 2>();
	std::array<std::size_t,Rank + 2> extents;
	extents[0] = samples;
	extents[1] = d.rows / samples;
	for (std::size_t i = 0; i < Rank; ++i)
		extents[i + 2] = dims(i);
	return to_host<Scalar,Rank + 2>(d, extents);
}

/**
 * Time step `step` of a device sequence of `steps` steps as a tensor of its own (samples x volume): one strided copy.
 * A sequence of a single step is returned as it is (shared buffer).
 */
template<typename Scalar>
inline DeviceTensor<Scalar> time_step_of(const DeviceTensor<Scalar>& seq, std::size_t samples, std::size_t steps,
		std::size_t step) {
	if (steps == 1)
		return seq;
	const std::size_t volume = seq.size() / seq.rows;
	DeviceTensor<Scalar> out(samples, volume);
	Context& c = Context::get();
	Context::Lock l = c.lock();
	CATTLE_B200_CHECK(cattl3_memcpy_2d(c.handle(), out.data(), samples * sizeof(Scalar), seq.data() + samples * step,
			samples * steps * sizeof(Scalar), samples * sizeof(Scalar), volume));
	return out;
}

/** Writes `slice` (samples x volume) into time step `step` of `seq` (`steps` steps; the caller owns `seq` exclusively). */
template<typename Scalar>
inline void set_time_step(DeviceTensor<Scalar>& seq, std::size_t samples, std::size_t steps, std::size_t step,
		const DeviceTensor<Scalar>& slice) {
	const std::size_t volume = seq.size() / seq.rows;
	if (slice.size() != samples * volume)
		throw Error(CATTL3_ERR_INVALID, "set_time_step: the slice does not match the sequence");
	Context& c = Context::get();
	Context::Lock l = c.lock();
	CATTLE_B200_CHECK(cattl3_memcpy_2d(c.handle(), seq.data() + samples * step, samples * steps * sizeof(Scalar),
			slice.data(), samples * sizeof(Scalar), samples * sizeof(Scalar), volume));
}

// ---- what the cells of an unrolled recurrent network are made of (LSTMNeuralNetwork, RecurrentNeuralNetwork) ----------

/** Layer::pass_forward on the device; a layer without a device face is bridged through the host. */
template<typename Scalar, std::size_t Rank>
inline DeviceTensor<Scalar> layer_forward_dev(Layer<Scalar,Rank>& layer, DeviceTensor<Scalar> in, bool training) {
	if (DeviceLayer<Scalar,Rank>* dev = dynamic_cast<DeviceLayer<Scalar,Rank>*>(&layer))
		return dev->pass_forward_dev(std::move(in), training);
	return to_device<Scalar,Rank + 1>(layer.pass_forward(to_host<Scalar,Rank + 1>(in,
			batch_extents<Rank>(in.rows, layer.get_input_dims())), training));
}

/** Layer::pass_back on the device (host bridge as above); an empty tensor where the layer returns none. */
template<typename Scalar, std::size_t Rank>
inline DeviceTensor<Scalar> layer_backward_dev(Layer<Scalar,Rank>& layer, DeviceTensor<Scalar> out_grad) {
	if (DeviceLayer<Scalar,Rank>* dev = dynamic_cast<DeviceLayer<Scalar,Rank>*>(&layer))
		return dev->pass_back_dev(std::move(out_grad));
	Tensor<Scalar,Rank + 1> prev_out_grad = layer.pass_back(to_host<Scalar,Rank + 1>(out_grad,
			batch_extents<Rank>(out_grad.rows, layer.get_output_dims())));
	if (prev_out_grad.size() == 0)
		return DeviceTensor<Scalar>();
	return to_device<Scalar,Rank + 1>(prev_out_grad);
}

/** out = (accumulate ? out : 0) + a * b (+ c * d): cattl3_muladd; `out` may be one of the operands. */
template<typename Scalar>
inline void tensor_muladd(bool accumulate, const DeviceTensor<Scalar>& a, const DeviceTensor<Scalar>& b,
		const DeviceTensor<Scalar>* c_factor, const DeviceTensor<Scalar>* d_factor, DeviceTensor<Scalar>& out) {
	if (a.size() != out.size() || b.size() != out.size() || (c_factor && (!d_factor || c_factor->size() != out.size() ||
			d_factor->size() != out.size())))
		throw Error(CATTL3_ERR_INVALID, "tensor_muladd: the tensors differ in size");
	Context& c = Context::get();
	Context::Lock l = c.lock();
	CATTLE_B200_CHECK(Api<Scalar>::muladd(c.handle(), (std::int64_t) out.size(), accumulate ? 1 : 0, a.data(), b.data(),
			c_factor ? c_factor->data() : nullptr, d_factor ? d_factor->data() : nullptr, out.data()));
}

/** a * b as a new tensor. */
template<typename Scalar>
inline DeviceTensor<Scalar> tensor_product(const DeviceTensor<Scalar>& a, const DeviceTensor<Scalar>& b) {
	DeviceTensor<Scalar> out(a.rows, a.size() / a.rows);
	tensor_muladd<Scalar>(false, a, b, nullptr, nullptr, out);
	return out;
}

/** y += x; y is exclusively owned afterwards (copy on write if it shared its buffer). */
template<typename Scalar>
inline void tensor_add(DeviceTensor<Scalar>& y, const DeviceTensor<Scalar>& x) {
	if (x.size() != y.size())
		throw Error(CATTL3_ERR_INVALID, "tensor_add: the tensors differ in size");
	y.make_exclusive();
	Context& c = Context::get();
	Context::Lock l = c.lock();
	CATTLE_B200_CHECK(Api<Scalar>::add_inplace(c.handle(), (std::int64_t) y.size(), y.data(), x.data()));
}

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_DEVICESEQUENCENETWORK_H_ */
