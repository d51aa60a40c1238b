/*
 * neural_network/SequentialNeuralNetwork.hpp -- B200 replacement of the reference's sequence wrapper
 * (C-ATTL3/neural_network/SequentialNeuralNetwork.hpp:24-148), same class template, constructor and interface;
 * defines the reference header's include guard.
 *
 * The wrapper applies a non-sequential network to every time step by joining the samples and time-steps ranks
 * (:95-124).  With samples fastest in memory that join is a view of the same array, so on the device the wrapped
 * network simply sees a batch of samples * steps rows (b200::DeviceSequenceNetwork): no copy in either direction.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_SEQUENTIALNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_SEQUENTIALNEURALNETWORK_H_

#include <array>
#include <cassert>
#include <utility>
#include <vector>

#include "neural_network/CompositeNeuralNetwork.hpp"
#include "b200/DeviceNetwork.hpp"
#include "b200/DeviceSequenceNetwork.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class SequentialNeuralNetwork :
		public CompositeNeuralNetwork<Scalar,Rank,true,NeuralNetwork<Scalar,Rank,false>>,
		public b200::DeviceSequenceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,true> Base;
	typedef NeuralNetwork<Scalar,Rank,false> Wrapped;
	typedef SequentialNeuralNetwork<Scalar,Rank> Self;
	typedef b200::DeviceNetwork<Scalar,Rank> DevNet;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * @param network The non-sequential network to apply to every time step.
	 * @param foremost Whether the network is the first module of a composite.
	 */
	inline SequentialNeuralNetwork(NeuralNetPtr<Scalar,Rank,false>&& network, bool foremost = true) :
			wrapped(std::move(network)),
			first(foremost) {
		assert(wrapped);
		wrapped->set_foremost(first);
	}
	inline SequentialNeuralNetwork(const Self& other) :
			wrapped(other.wrapped->clone()),
			first(other.first) { }
	inline SequentialNeuralNetwork(Self&& other) :
			wrapped(std::move(other.wrapped)),
			first(other.first) { }
	~SequentialNeuralNetwork() = default;
	inline Self& operator=(Self other) {
		swap(*this, other);
		return *this;
	}
	inline Base* clone() const {
		return new Self(*this);
	}
	// the wrapped network's dimensions and layers are the wrapper's (:75-92)
	inline const typename Base::Dims& get_input_dims() const {
		return wrapped->get_input_dims();
	}
	inline const typename Base::Dims& get_output_dims() const {
		return wrapped->get_output_dims();
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		return static_cast<const Wrapped&>(*wrapped).get_layers();
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		return wrapped->get_layers();
	}
	inline std::vector<Wrapped*> get_modules() {
		return std::vector<Wrapped*>(1, wrapped.get());
	}
	inline bool is_foremost() const {
		return first;
	}
	inline void set_foremost(bool foremost) {
		first = foremost;
		wrapped->set_foremost(foremost);
	}
	inline void empty_caches() {
		wrapped->empty_caches();
	}
	/** Host API: one upload, the device path, one download. */
	inline typename Base::Data propagate(typename Base::Data input, bool training) {
		assert(get_input_dims() == (Dimensions<std::size_t,Base::DATA_RANK>(input.dimensions()).template demote<2>()));
		const std::size_t samples = input.dimension(0);
		return b200::sequence_to_host<Scalar,Rank>(propagate_seq_dev(
				b200::sequence_to_device<Scalar,Base::DATA_RANK>(input), samples, training), samples, get_output_dims());
	}
	inline typename Base::Data backpropagate(typename Base::Data out_grad) {
		assert(get_output_dims() == (Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<2>()));
		const std::size_t samples = out_grad.dimension(0);
		return b200::sequence_to_host<Scalar,Rank>(backpropagate_seq_dev(
				b200::sequence_to_device<Scalar,Base::DATA_RANK>(out_grad), samples), samples, get_input_dims());
	}
	/** b200::DeviceSequenceNetwork: the wrapped network sees samples * steps rows (the join is a view). */
	inline DevTensor propagate_seq_dev(DevTensor input, std::size_t samples, bool training) {
		if (DevNet* dev = dynamic_cast<DevNet*>(wrapped.get()))
			return dev->propagate_dev(std::move(input), training);
		// a network that only speaks the host API: one round trip around it
		return b200::to_device<Scalar,Rank + 1>(wrapped->propagate(b200::to_host<Scalar,Rank + 1>(input,
				b200::batch_extents<Rank>(input.rows, get_input_dims())), training));
	}
	inline DevTensor backpropagate_seq_dev(DevTensor out_grad, std::size_t samples) {
		if (DevNet* dev = dynamic_cast<DevNet*>(wrapped.get()))
			return dev->backpropagate_dev(std::move(out_grad));
		Tensor<Scalar,Rank + 1> prev_out_grad = wrapped->backpropagate(
				b200::to_host<Scalar,Rank + 1>(out_grad, b200::batch_extents<Rank>(out_grad.rows, get_output_dims())));
		if (first || prev_out_grad.size() == 0)
			return DevTensor();
		return b200::to_device<Scalar,Rank + 1>(prev_out_grad);
	}
	inline bool graph_safe() const {
		return dynamic_cast<const DevNet*>(wrapped.get()) != nullptr;
	}
	inline friend void swap(Self& a, Self& b) {
		std::swap(a.wrapped, b.wrapped);
		std::swap(a.first, b.first);
	}
private:
	NeuralNetPtr<Scalar,Rank,false> wrapped;
	bool first;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_SEQUENTIALNEURALNETWORK_H_ */
