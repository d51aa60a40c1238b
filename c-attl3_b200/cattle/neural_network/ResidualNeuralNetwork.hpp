/*
 * neural_network/ResidualNeuralNetwork.hpp -- B200 replacement of the reference's residual network
 * (C-ATTL3/neural_network/ResidualNeuralNetwork.hpp:27-160), same class template, constructors and
 * interface; defines the reference header's include guard.
 *
 * Forward: x <- x + module_i(x) for every module (:112-117); backward: g <- g + module_i'(g), except
 * that the first module of a foremost network returns its result directly (:118-127).  When the
 * modules are device networks (b200::DeviceNetwork, e.g. the B200 FeedforwardNeuralNetwork) the
 * activations stay in HBM across the whole network and the skip additions are one element-wise kernel
 * each (cattl3_add_inplace); modules that only speak the host API are bridged with a round trip.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_RESIDUALNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_RESIDUALNEURALNETWORK_H_

#include <cassert>
#include <utility>
#include <vector>

#include "neural_network/CompositeNeuralNetwork.hpp"
#include "b200/DeviceNetwork.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class ResidualNeuralNetwork :
		public CompositeNeuralNetwork<Scalar,Rank,false,NeuralNetwork<Scalar,Rank,false>>,
		public b200::DeviceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,false> Base;
	typedef NeuralNetPtr<Scalar,Rank,false> Module;
	typedef ResidualNeuralNetwork<Scalar,Rank> Self;
	typedef b200::DeviceNetwork<Scalar,Rank> DevNet;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * @param modules The residual modules; the input and output dimensions of each must be equal.
	 * @param foremost Whether the network is the first module of a composite.
	 */
	inline ResidualNeuralNetwork(std::vector<Module>&& modules, bool foremost = true) :
			modules(std::move(modules)),
			foremost(foremost) {
		assert(this->modules.size() > 0 && "modules must contain at least 1 element");
		input_dims = this->modules.front()->get_input_dims();
		output_dims = this->modules.back()->get_output_dims();
		for (std::size_t i = 0; i < this->modules.size(); ++i) {
			Base& module = *this->modules[i];
			assert(module.get_input_dims() == module.get_output_dims() &&
					"residual module input-output dimension discrepancy");
			assert(input_dims == module.get_input_dims() && "incompatible module dimensions");
			module.set_foremost(i == 0 && foremost);
		}
	}
	inline ResidualNeuralNetwork(Module&& module, bool foremost = true) :
			ResidualNeuralNetwork(single(std::move(module)), foremost) { }
	inline ResidualNeuralNetwork(const Self& network) :
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims) {
		for (const Module& module : network.modules)
			modules.push_back(Module(module->clone()));
	}
	inline ResidualNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~ResidualNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Base* clone() const {
		return new ResidualNeuralNetwork(*this);
	}
	inline const typename Base::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return output_dims;
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		std::vector<const Layer<Scalar,Rank>*> layer_ptrs;
		for (const Module& module : modules) {
			for (Layer<Scalar,Rank>* layer : module->get_layers())
				layer_ptrs.push_back(layer);
		}
		return layer_ptrs;
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		std::vector<Layer<Scalar,Rank>*> layer_ptrs;
		for (const Module& module : modules) {
			for (Layer<Scalar,Rank>* layer : module->get_layers())
				layer_ptrs.push_back(layer);
		}
		return layer_ptrs;
	}
	inline std::vector<Base*> get_modules() {
		std::vector<Base*> module_ptrs;
		for (const Module& module : modules)
			module_ptrs.push_back(module.get());
		return module_ptrs;
	}
	inline bool is_foremost() const {
		return foremost;
	}
	inline void set_foremost(bool foremost) {
		modules.front()->set_foremost(foremost);
		this->foremost = foremost;
	}
	inline void empty_caches() {
		for (const Module& module : modules)
			module->empty_caches();
	}
	inline typename Base::Data propagate(typename Base::Data input, bool training) {
		assert(input_dims == (Dimensions<std::size_t,Base::DATA_RANK>(input.dimensions()).template demote<>()));
		DevTensor out = propagate_dev(b200::to_device<Scalar,Base::DATA_RANK>(input), training);
		return b200::to_host<Scalar,Base::DATA_RANK>(out, b200::batch_extents<Rank>(out.rows, output_dims));
	}
	inline typename Base::Data backpropagate(typename Base::Data out_grad) {
		assert(output_dims == (Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<>()));
		DevTensor prev_out_grad = backpropagate_dev(b200::to_device<Scalar,Base::DATA_RANK>(out_grad));
		if (prev_out_grad.empty())
			return typename Base::Data();
		return b200::to_host<Scalar,Base::DATA_RANK>(prev_out_grad,
				b200::batch_extents<Rank>(prev_out_grad.rows, input_dims));
	}
	inline DevTensor propagate_dev(DevTensor input, bool training) {
		for (const Module& module : modules) {
			// the module reads (and may cache) `input`; the sum goes into the module's own output buffer
			DevTensor branch = run_forward(*module, input, training);
			add_into(branch, input);
			input = std::move(branch);
		}
		return input;
	}
	inline DevTensor backpropagate_dev(DevTensor out_grad) {
		for (std::size_t i = modules.size(); i > 0; --i) {
			DevTensor branch = run_backward(*modules[i - 1], out_grad);
			if (foremost && i == 1)
				return branch;  // nothing upstream needs the skip path's share
			add_into(branch, out_grad);
			out_grad = std::move(branch);
		}
		return out_grad;
	}
	inline friend void swap(Self& network1, Self& network2) {
		using std::swap;
		swap(network1.modules, network2.modules);
		swap(network1.foremost, network2.foremost);
		swap(network1.input_dims, network2.input_dims);
		swap(network1.output_dims, network2.output_dims);
	}
private:
	inline static std::vector<Module> single(Module&& module) {
		std::vector<Module> vec;
		vec.push_back(std::move(module));
		return vec;
	}
	/** target += addend, in place on the device. */
	inline static void add_into(DevTensor& target, const DevTensor& addend) {
		if (target.size() != addend.size())
			throw b200::Error(CATTL3_ERR_INVALID, "ResidualNeuralNetwork: module changed the tensor size");
		target.make_exclusive();
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(b200::Api<Scalar>::add_inplace(c.handle(), (std::int64_t) target.size(), target.data(),
				addend.data()));
	}
	inline DevTensor run_forward(Base& module, const DevTensor& input, bool training) const {
		if (DevNet* dev_module = dynamic_cast<DevNet*>(&module))
			return dev_module->propagate_dev(input, training);
		typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(input,
				b200::batch_extents<Rank>(input.rows, input_dims));
		return b200::to_device<Scalar,Base::DATA_RANK>(module.propagate(std::move(host), training));
	}
	inline DevTensor run_backward(Base& module, const DevTensor& out_grad) const {
		if (DevNet* dev_module = dynamic_cast<DevNet*>(&module))
			return dev_module->backpropagate_dev(out_grad);
		typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(out_grad,
				b200::batch_extents<Rank>(out_grad.rows, output_dims));
		return b200::to_device<Scalar,Base::DATA_RANK>(module.backpropagate(std::move(host)));
	}
	std::vector<Module> modules;
	bool foremost;
	typename Base::Dims input_dims, output_dims;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_RESIDUALNEURALNETWORK_H_ */
