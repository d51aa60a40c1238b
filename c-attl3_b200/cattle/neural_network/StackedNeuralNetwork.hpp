/*
 * neural_network/StackedNeuralNetwork.hpp -- B200 replacement of the reference's stacked network
 * (C-ATTL3/neural_network/StackedNeuralNetwork.hpp:21-167), same class template, constructors and
 * interface; defines the reference header's include guard.
 *
 * A stack runs its blocks one after the other (:111-121).  The reference moves a host tensor from block to
 * block; here, for non-sequential stacks, blocks that are device networks (b200::DeviceNetwork: the B200
 * Feedforward / Residual / Stacked networks) pass their activations on in HBM, and the stack is itself a
 * device network, so e.g. the encoder / decoder pair of examples/mnist_autoencoder.cpp:26-46 trains without
 * a host round trip between the two halves.  Blocks that only speak the host API are bridged with a round
 * trip around them.  Sequential stacks (rank + 2 tensors) do the same over b200::DeviceSequenceNetwork blocks (the
 * B200 SequentialNeuralNetwork / LSTMNeuralNetwork / sequential stacks); a sequential stack without any such block
 * keeps the reference's host protocol.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_STACKEDNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_STACKEDNEURALNETWORK_H_

#include <cassert>
#include <type_traits>
#include <utility>
#include <vector>

#include "neural_network/CompositeNeuralNetwork.hpp"
#include "b200/DeviceNetwork.hpp"
#include "b200/DeviceSequenceNetwork.hpp"

namespace cattle {

namespace b200 {
/** The device face of a stack: a device network, or (sequential stacks) a device sequence network. */
template<typename Scalar, std::size_t Rank, bool Sequential> struct StackDeviceFace { };
template<typename Scalar, std::size_t Rank> struct StackDeviceFace<Scalar,Rank,false> : public DeviceNetwork<Scalar,Rank> { };
template<typename Scalar, std::size_t Rank> struct StackDeviceFace<Scalar,Rank,true> : public DeviceSequenceNetwork<Scalar,Rank> { };
}

template<typename Scalar, std::size_t Rank, bool Sequential>
class StackedNeuralNetwork :
		public CompositeNeuralNetwork<Scalar,Rank,Sequential,NeuralNetwork<Scalar,Rank,Sequential>>,
		public b200::StackDeviceFace<Scalar,Rank,Sequential> {
	typedef NeuralNetwork<Scalar,Rank,Sequential> Base;
	typedef StackedNeuralNetwork<Scalar,Rank,Sequential> Self;
	typedef NeuralNetPtr<Scalar,Rank,Sequential> Block;
	typedef b200::DeviceNetwork<Scalar,Rank> DevNet;
	typedef b200::DeviceSequenceNetwork<Scalar,Rank> DevSeqNet;
	typedef b200::DeviceTensor<Scalar> DevTensor;
	typedef std::integral_constant<bool,Sequential> IsSequential;
public:
	/**
	 * @param blocks The sub-networks, in order; consecutive dimensions must match.
	 * @param foremost Whether the stack is the first module of a composite.
	 */
	inline StackedNeuralNetwork(std::vector<Block>&& blocks, bool foremost = true) :
			blocks(std::move(blocks)),
			foremost(foremost) {
		assert(this->blocks.size() > 0 && "blocks must contain at least 1 element");
		for (std::size_t i = 0; i < this->blocks.size(); ++i) {
			assert(this->blocks[i] != nullptr && "blocks contains null pointers");
			assert((i == 0 || this->blocks[i - 1]->get_output_dims() == this->blocks[i]->get_input_dims()) &&
					"incompatible network dimensions");
			this->blocks[i]->set_foremost(i == 0 && foremost);
		}
		input_dims = this->blocks.front()->get_input_dims();
		output_dims = this->blocks.back()->get_output_dims();
	}
	inline StackedNeuralNetwork(Block&& block, bool foremost = true) :
			StackedNeuralNetwork(single(std::move(block)), foremost) { }
	inline StackedNeuralNetwork(const Self& network) :
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims) {
		for (const Block& block : network.blocks)
			blocks.push_back(Block(block->clone()));
	}
	inline StackedNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~StackedNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Base* clone() const {
		return new StackedNeuralNetwork(*this);
	}
	inline const typename Base::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return output_dims;
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		std::vector<const Layer<Scalar,Rank>*> layer_ptrs;
		for (const Block& block : blocks) {
			for (Layer<Scalar,Rank>* layer : block->get_layers())
				layer_ptrs.push_back(layer);
		}
		return layer_ptrs;
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		std::vector<Layer<Scalar,Rank>*> layer_ptrs;
		for (const Block& block : blocks) {
			for (Layer<Scalar,Rank>* layer : block->get_layers())
				layer_ptrs.push_back(layer);
		}
		return layer_ptrs;
	}
	inline std::vector<Base*> get_modules() {
		std::vector<Base*> module_ptrs;
		for (const Block& block : blocks)
			module_ptrs.push_back(block.get());
		return module_ptrs;
	}
	inline bool is_foremost() const {
		return foremost;
	}
	inline void set_foremost(bool foremost) {
		blocks.front()->set_foremost(foremost);
		this->foremost = foremost;
	}
	inline void empty_caches() {
		for (const Block& block : blocks)
			block->empty_caches();
	}
	inline typename Base::Data propagate(typename Base::Data input, bool training) {
		assert(input_dims == (Dimensions<std::size_t,Base::DATA_RANK>(input.dimensions()).template demote<Sequential + 1>()));
		return propagate_host(std::move(input), training, IsSequential());
	}
	inline typename Base::Data backpropagate(typename Base::Data out_grad) {
		assert(output_dims == (Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<Sequential + 1>()));
		return backpropagate_host(std::move(out_grad), IsSequential());
	}
	/** b200::DeviceNetwork (non-sequential stacks): block to block in HBM. */
	inline DevTensor propagate_dev(DevTensor input, bool training) {
		for (const Block& block : blocks) {
			if (DevNet* dev_block = dynamic_cast<DevNet*>(block.get())) {
				input = dev_block->propagate_dev(std::move(input), training);
			} else {
				input = upload(block->propagate(download(input, block->get_input_dims(), IsSequential()), training),
						IsSequential());
			}
		}
		return input;
	}
	inline DevTensor backpropagate_dev(DevTensor out_grad) {
		for (std::size_t i = blocks.size(); i > 0 && !out_grad.empty(); --i) {
			Base& block = *blocks[i - 1];
			if (DevNet* dev_block = dynamic_cast<DevNet*>(&block)) {
				out_grad = dev_block->backpropagate_dev(std::move(out_grad));
			} else {
				out_grad = upload(block.backpropagate(download(out_grad, block.get_output_dims(), IsSequential())),
						IsSequential());
			}
		}
		return out_grad;
	}
	/** b200::DeviceSequenceNetwork (sequential stacks): block to block in HBM, rows = samples * time steps. */
	inline DevTensor propagate_seq_dev(DevTensor input, std::size_t samples, bool training) {
		for (const Block& block : blocks) {
			if (DevSeqNet* dev_block = dynamic_cast<DevSeqNet*>(block.get())) {
				input = dev_block->propagate_seq_dev(std::move(input), samples, training);
			} else {
				input = b200::sequence_to_device<Scalar,Rank + 2>(block->propagate(
						b200::sequence_to_host<Scalar,Rank>(input, samples, block->get_input_dims()), training));
			}
		}
		return input;
	}
	inline DevTensor backpropagate_seq_dev(DevTensor out_grad, std::size_t samples) {
		for (std::size_t i = blocks.size(); i > 0 && !out_grad.empty(); --i) {
			Base& block = *blocks[i - 1];
			if (DevSeqNet* dev_block = dynamic_cast<DevSeqNet*>(&block)) {
				out_grad = dev_block->backpropagate_seq_dev(std::move(out_grad), samples);
			} else {
				out_grad = b200::sequence_to_device<Scalar,Rank + 2>(block.backpropagate(
						b200::sequence_to_host<Scalar,Rank>(out_grad, samples, block.get_output_dims())));
			}
		}
		return out_grad;
	}
	/** b200::DeviceSequenceNetwork: a step graph needs every block on the device and none carrying state across steps. */
	inline bool graph_safe() const {
		for (const Block& block : blocks) {
			const DevSeqNet* dev_block = dynamic_cast<const DevSeqNet*>(block.get());
			if (!dev_block || !dev_block->graph_safe())
				return false;
		}
		return true;
	}
	inline friend void swap(Self& network1, Self& network2) {
		using std::swap;
		swap(network1.blocks, network2.blocks);
		swap(network1.foremost, network2.foremost);
		swap(network1.input_dims, network2.input_dims);
		swap(network1.output_dims, network2.output_dims);
	}
private:
	inline static std::vector<Block> single(Block&& block) {
		std::vector<Block> vec;
		vec.push_back(std::move(block));
		return vec;
	}
	// sequential stacks: one upload, the device chain, one download -- or, without any device sequence block, the
	// reference's host protocol (StackedNeuralNetwork.hpp:111-121)
	inline bool has_device_sequence_block() const {
		for (const Block& block : blocks) {
			if (dynamic_cast<DevSeqNet*>(block.get()))
				return true;
		}
		return false;
	}
	inline typename Base::Data propagate_host(typename Base::Data input, bool training, std::true_type) {
		if (has_device_sequence_block()) {
			const std::size_t samples = input.dimension(0);
			DevTensor out = propagate_seq_dev(b200::sequence_to_device<Scalar,Rank + 2>(input), samples, training);
			return b200::sequence_to_host<Scalar,Rank>(out, samples, output_dims);
		}
		for (const Block& block : blocks)
			input = block->propagate(std::move(input), training);
		return input;
	}
	inline typename Base::Data backpropagate_host(typename Base::Data out_grad, std::true_type) {
		if (has_device_sequence_block()) {
			const std::size_t samples = out_grad.dimension(0);
			DevTensor prev_out_grad = backpropagate_seq_dev(b200::sequence_to_device<Scalar,Rank + 2>(out_grad), samples);
			return b200::sequence_to_host<Scalar,Rank>(prev_out_grad, samples, input_dims);
		}
		for (std::size_t i = blocks.size(); i > 0; --i)
			out_grad = blocks[i - 1]->backpropagate(std::move(out_grad));
		return out_grad;
	}
	// non-sequential stacks: one upload, the device chain, one download
	inline typename Base::Data propagate_host(typename Base::Data input, bool training, std::false_type) {
		DevTensor out = propagate_dev(upload(std::move(input), std::false_type()), training);
		return download(out, output_dims, std::false_type());
	}
	inline typename Base::Data backpropagate_host(typename Base::Data out_grad, std::false_type) {
		DevTensor prev_out_grad = backpropagate_dev(upload(std::move(out_grad), std::false_type()));
		if (prev_out_grad.empty())
			return typename Base::Data();
		return download(prev_out_grad, input_dims, std::false_type());
	}
	inline static DevTensor upload(typename Base::Data data, std::false_type) {
		return b200::to_device<Scalar,Base::DATA_RANK>(data);
	}
	inline static typename Base::Data download(const DevTensor& data, const typename Base::Dims& dims, std::false_type) {
		return b200::to_host<Scalar,Base::DATA_RANK>(data, b200::batch_extents<Rank>(data.rows, dims));
	}
	// never called (the device face does not exist for sequential stacks); present so that the class compiles
	inline static DevTensor upload(typename Base::Data, std::true_type) {
		return DevTensor();
	}
	inline static typename Base::Data download(const DevTensor&, const typename Base::Dims&, std::true_type) {
		return typename Base::Data();
	}
	std::vector<Block> blocks;
	bool foremost;
	typename Base::Dims input_dims, output_dims;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_STACKEDNEURALNETWORK_H_ */
