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
	typedef SequentialNeuralNetwork<Scalar,Rank> Self;
	typedef NeuralNetPtr<Scalar,Rank,false> Net;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * @param network The non-sequential network to apply to every time step.
	 * @param foremost Whether the network is the first module of a composite.
	 */
	inline SequentialNeuralNetwork(Net&& network, bool foremost = true) :
			net(std::move(network)),
			foremost(foremost) {
		assert(net);
		input_dims = net->get_input_dims();
		output_dims = net->get_output_dims();
		set_foremost(foremost);
	}
	inline SequentialNeuralNetwork(const Self& network) :
			net(Net(network.net->clone())),
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims) { }
	inline SequentialNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~SequentialNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Base* clone() const {
		return new SequentialNeuralNetwork(*this);
	}
	inline const typename Base::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return output_dims;
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		return ((const NeuralNetwork<Scalar,Rank,false>&) *net).get_layers();
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		return net->get_layers();
	}
	inline std::vector<NeuralNetwork<Scalar,Rank,false>*> get_modules() {
		std::vector<NeuralNetwork<Scalar,Rank,false>*> modules;
		modules.push_back(net.get());
		return modules;
	}
	inline bool is_foremost() const {
		return foremost;
	}
	inline void set_foremost(bool foremost) {
		net->set_foremost(foremost);
		this->foremost = foremost;
	}
	inline void empty_caches() {
		net->empty_caches();
	}
	inline typename Base::Data propagate(typename Base::Data input, bool training) {
		assert(input_dims == (Dimensions<std::size_t,Base::DATA_RANK>(input.dimensions()).template demote<2>()));
		const std::size_t samples = input.dimension(0);
		DevTensor out = propagate_seq_dev(b200::sequence_to_device<Scalar,Base::DATA_RANK>(input), samples, training);
		return b200::sequence_to_host<Scalar,Rank>(out, samples, output_dims);
	}
	inline typename Base::Data backpropagate(typename Base::Data out_grad) {
		assert(output_dims == (Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<2>()));
		const std::size_t samples = out_grad.dimension(0);
		DevTensor prev_out_grad = backpropagate_seq_dev(b200::sequence_to_device<Scalar,Base::DATA_RANK>(out_grad), samples);
		return b200::sequence_to_host<Scalar,Rank>(prev_out_grad, samples, input_dims);
	}
	/** b200::DeviceSequenceNetwork: the wrapped network sees samples * steps rows (the join is a view). */
	inline DevTensor propagate_seq_dev(DevTensor input, std::size_t samples, bool training) {
		if (b200::DeviceNetwork<Scalar,Rank>* dev_net = dynamic_cast<b200::DeviceNetwork<Scalar,Rank>*>(net.get()))
			return dev_net->propagate_dev(std::move(input), training);
		// a network that only speaks the host API: one round trip around it
		return b200::to_device<Scalar,Rank + 1>(net->propagate(b200::to_host<Scalar,Rank + 1>(input,
				b200::batch_extents<Rank>(input.rows, input_dims)), training));
	}
	inline DevTensor backpropagate_seq_dev(DevTensor out_grad, std::size_t samples) {
		if (b200::DeviceNetwork<Scalar,Rank>* dev_net = dynamic_cast<b200::DeviceNetwork<Scalar,Rank>*>(net.get()))
			return dev_net->backpropagate_dev(std::move(out_grad));
		Tensor<Scalar,Rank + 1> prev_out_grad = net->backpropagate(
				b200::to_host<Scalar,Rank + 1>(out_grad, b200::batch_extents<Rank>(out_grad.rows, output_dims)));
		if (foremost || prev_out_grad.size() == 0)
			return DevTensor();
		return b200::to_device<Scalar,Rank + 1>(prev_out_grad);
	}
	inline bool graph_safe() const {
		return dynamic_cast<const b200::DeviceNetwork<Scalar,Rank>*>(net.get()) != nullptr;
	}
	inline friend void swap(Self& network1, Self& network2) {
		using std::swap;
		swap(network1.net, network2.net);
		swap(network1.foremost, network2.foremost);
		swap(network1.input_dims, network2.input_dims);
		swap(network1.output_dims, network2.output_dims);
	}
private:
	Net net;
	bool foremost;
	typename Base::Dims input_dims, output_dims;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_SEQUENTIALNEURALNETWORK_H_ */
