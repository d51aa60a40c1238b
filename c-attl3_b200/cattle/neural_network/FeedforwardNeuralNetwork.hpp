/*
 * neural_network/FeedforwardNeuralNetwork.hpp -- B200 replacement of the reference's layer loop
 * (C-ATTL3/neural_network/FeedforwardNeuralNetwork.hpp:38-127), same class template, constructors and
 * NeuralNetwork interface; defines the reference header's include guard.
 *
 * The difference is where the activations live between layers.  The reference moves an Eigen tensor
 * from layer to layer (:112-127).  Here every maximal run of consecutive B200 layers (those
 * implementing b200::DeviceLayer) is executed on the device end to end: one upload in front of the
 * run, one download behind it, nothing in between; layers that only know the host API (any reference
 * layer: PReLU, Dropout, Reshape, ...) still work, at the price of a round trip around them.  Through
 * b200::DeviceNetwork the whole network can also be driven with device tensors, which is what the
 * batch loop (optimizer/SGDOptimizer.hpp) and the composite networks do.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_FEEDFORWARDNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_FEEDFORWARDNEURALNETWORK_H_

#include <cassert>
#include <cstdlib>
#include <memory>
#include <utility>
#include <vector>

#include "core/NeuralNetwork.hpp"
#include "b200/DeviceNetwork.hpp"

namespace cattle {

/**
 * An alias for a unique pointer to a layer of arbitrary rank and scalar type.
 */
template<typename Scalar, std::size_t Rank>
using LayerPtr = std::unique_ptr<Layer<Scalar,Rank>>;

template<typename Scalar, std::size_t Rank>
class FeedforwardNeuralNetwork : public NeuralNetwork<Scalar,Rank,false>, public b200::DeviceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,false> Base;
	typedef FeedforwardNeuralNetwork<Scalar,Rank> Self;
	typedef b200::DeviceLayer<Scalar,Rank> DevLayer;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * @param layers The layers, in order; consecutive dimensions must match.
	 * @param foremost Whether the network is the first module of a composite: its first layer then
	 * need not produce an input gradient.
	 */
	inline FeedforwardNeuralNetwork(std::vector<LayerPtr<Scalar,Rank>>&& layers, bool foremost = true) :
			layers(std::move(layers)),
			foremost(foremost) {
		assert(this->layers.size() > 0 && "layers must contain at least 1 element");
		for (std::size_t i = 0; i < this->layers.size(); ++i) {
			assert(this->layers[i] != nullptr && "layers contains null pointers");
			assert((i == 0 || this->layers[i - 1]->get_output_dims() == this->layers[i]->get_input_dims()) &&
					"incompatible layer dimensions");
		}
		input_dims = this->layers.front()->get_input_dims();
		output_dims = this->layers.back()->get_output_dims();
		this->layers.front()->set_input_layer(foremost);
		find_device_layers();
	}
	inline FeedforwardNeuralNetwork(LayerPtr<Scalar,Rank>&& layer, bool foremost = true) :
			FeedforwardNeuralNetwork(single(std::move(layer)), foremost) { }
	inline FeedforwardNeuralNetwork(const Self& network) :
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims) {
		for (const LayerPtr<Scalar,Rank>& layer : network.layers)
			layers.push_back(LayerPtr<Scalar,Rank>(layer->clone()));
		find_device_layers();
	}
	inline FeedforwardNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~FeedforwardNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Base* clone() const {
		return new FeedforwardNeuralNetwork(*this);
	}
	inline const typename Base::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return output_dims;
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		std::vector<const Layer<Scalar,Rank>*> layer_ptrs;
		for (const LayerPtr<Scalar,Rank>& layer : layers)
			layer_ptrs.push_back(layer.get());
		return layer_ptrs;
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		std::vector<Layer<Scalar,Rank>*> layer_ptrs;
		for (const LayerPtr<Scalar,Rank>& layer : layers)
			layer_ptrs.push_back(layer.get());
		return layer_ptrs;
	}
	inline bool is_foremost() const {
		return foremost;
	}
	inline void set_foremost(bool foremost) {
		layers.front()->set_input_layer(foremost);
		this->foremost = foremost;
	}
	inline virtual void empty_caches() {
		for (const LayerPtr<Scalar,Rank>& layer : layers)
			layer->empty_cache();
	}
	inline typename Base::Data propagate(typename Base::Data input, bool training) {
		assert(input_dims == (Dimensions<std::size_t,Base::DATA_RANK>(input.dimensions()).template demote<>()));
		std::size_t i = 0;
		while (i < layers.size()) {
			if (!device_layers[i]) {
				input = layers[i]->pass_forward(std::move(input), training);
				++i;
				continue;
			}
			// a run of device layers: upload once, chain in HBM, download once
			DevTensor act = b200::to_device<Scalar,Base::DATA_RANK>(input);
			input = typename Base::Data();
			while (i < layers.size() && device_layers[i])
				act = forward_from(i, std::move(act), training);
			input = b200::to_host<Scalar,Base::DATA_RANK>(act,
					b200::batch_extents<Rank>(act.rows, layers[i - 1]->get_output_dims()));
		}
		return input;
	}
	inline typename Base::Data backpropagate(typename Base::Data out_grad) {
		assert(output_dims == (Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<>()));
		std::size_t i = layers.size();
		while (i > 0) {
			if (!device_layers[i - 1]) {
				out_grad = layers[i - 1]->pass_back(std::move(out_grad));
				--i;
				continue;
			}
			DevTensor grad = b200::to_device<Scalar,Base::DATA_RANK>(out_grad);
			out_grad = typename Base::Data();
			for (; i > 0 && device_layers[i - 1]; --i)
				grad = device_layers[i - 1]->pass_back_dev(std::move(grad));
			if (grad.empty())  // an input layer ended the chain (Layer.hpp:82-90)
				return typename Base::Data();
			out_grad = b200::to_host<Scalar,Base::DATA_RANK>(grad,
					b200::batch_extents<Rank>(grad.rows, layers[i]->get_input_dims()));
		}
		return out_grad;
	}
	inline DevTensor propagate_dev(DevTensor input, bool training) {
		std::size_t i = 0;
		while (i < layers.size()) {
			if (device_layers[i]) {
				input = forward_from(i, std::move(input), training);
			} else {
				typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(input,
						b200::batch_extents<Rank>(input.rows, layers[i]->get_input_dims()));
				input = b200::to_device<Scalar,Base::DATA_RANK>(layers[i]->pass_forward(std::move(host), training));
				++i;
			}
		}
		return input;
	}
	inline DevTensor backpropagate_dev(DevTensor out_grad) {
		for (std::size_t i = layers.size(); i > 0; --i) {
			if (device_layers[i - 1]) {
				out_grad = device_layers[i - 1]->pass_back_dev(std::move(out_grad));
			} else {
				typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(out_grad,
						b200::batch_extents<Rank>(out_grad.rows, layers[i - 1]->get_output_dims()));
				out_grad = b200::to_device<Scalar,Base::DATA_RANK>(layers[i - 1]->pass_back(std::move(host)));
			}
			if (out_grad.empty())
				break;
		}
		return out_grad;
	}
	/** Whether every layer runs on the device (no host round trips inside the network). */
	inline bool is_device_resident() const {
		for (DevLayer* layer : device_layers) {
			if (!layer)
				return false;
		}
		return true;
	}
	inline friend void swap(Self& network1, Self& network2) {
		using std::swap;
		swap(network1.layers, network2.layers);
		swap(network1.device_layers, network2.device_layers);
		swap(network1.producers, network2.producers);
		swap(network1.consumers, network2.consumers);
		swap(network1.foremost, network2.foremost);
		swap(network1.input_dims, network2.input_dims);
		swap(network1.output_dims, network2.output_dims);
	}
private:
	inline static std::vector<LayerPtr<Scalar,Rank>> single(LayerPtr<Scalar,Rank>&& layer) {
		std::vector<LayerPtr<Scalar,Rank>> vec;
		vec.push_back(std::move(layer));
		return vec;
	}
	inline void find_device_layers() {
		device_layers.clear();
		producers.clear();
		consumers.clear();
		for (const LayerPtr<Scalar,Rank>& layer : layers) {
			device_layers.push_back(dynamic_cast<DevLayer*>(layer.get()));
			producers.push_back(dynamic_cast<b200::EpilogueProducer<Scalar>*>(layer.get()));
			consumers.push_back(dynamic_cast<b200::EpilogueConsumer<Scalar>*>(layer.get()));
		}
	}
	/**
	 * Forward pass of device layer i -- together with the one or two layers behind it when they can ride in
	 * its epilogue: kernel layer -> activation, kernel layer -> batch norm (statistics) [-> activation].
	 * Every layer still ends up with the caches its own pass_back needs.  Advances i past the layers run.
	 */
	inline DevTensor forward_from(std::size_t& i, DevTensor act, bool training) {
		const std::size_t n = layers.size();
		b200::EpilogueProducer<Scalar>* producer = producers[i];
		b200::EpilogueConsumer<Scalar>* consumer = i + 1 < n ? consumers[i + 1] : nullptr;
		b200::FusedEpilogue<Scalar> ep;
		if (fuse_epilogues() && producer && consumer && producer->can_fuse_epilogue() &&
				consumer->request_epilogue(ep, producer->stat_columns(), training)) {
			b200::FusedEpilogue<Scalar> next;
			b200::EpilogueConsumer<Scalar>* chained = nullptr;
			if (consumer->chains_epilogue() && i + 2 < n && consumers[i + 2] &&
					consumers[i + 2]->request_epilogue(next, 0, training))
				chained = consumers[i + 2];
			act = producer->pass_forward_dev_fused(std::move(act), training, ep);
			act = consumer->accept_epilogue(std::move(act), ep, training, chained ? &next : nullptr);
			i += 2;
			if (chained) {
				act = chained->accept_epilogue(std::move(act), next, training, nullptr);
				++i;
			}
			return act;
		}
		act = device_layers[i]->pass_forward_dev(std::move(act), training);
		++i;
		return act;
	}
	/** CATTL3_NO_FUSION=1 in the environment runs every layer on its own (A/B measurements, debugging). */
	inline static bool fuse_epilogues() {
		static const bool on = [] {
			const char* v = std::getenv("CATTL3_NO_FUSION");
			return !(v && v[0] && v[0] != '0');
		}();
		return on;
	}
	std::vector<LayerPtr<Scalar,Rank>> layers;
	// device_layers[i] is layers[i] seen through its device interface, or null for a host-only layer
	std::vector<DevLayer*> device_layers;
	// layers[i] seen as an epilogue producer / consumer (b200/DeviceLayer.hpp), or null
	std::vector<b200::EpilogueProducer<Scalar>*> producers;
	std::vector<b200::EpilogueConsumer<Scalar>*> consumers;
	bool foremost;
	typename Base::Dims input_dims, output_dims;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_FEEDFORWARDNEURALNETWORK_H_ */
