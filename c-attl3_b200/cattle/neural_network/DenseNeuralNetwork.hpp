/*
 * neural_network/DenseNeuralNetwork.hpp -- B200 replacement of the reference's DenseNet container
 * (C-ATTL3/neural_network/DenseNeuralNetwork.hpp:25-195), same class template, constructors and interface;
 * defines the reference header's include guard.
 *
 * Every module's input is concatenated to its output along the lowest or the highest rank and handed to the
 * next module (:131-137); the backward pass slices the gradient apart again and adds the module's input gradient
 * to the skip part (:138-160).  The reference does this with Eigen concatenate / slice on host tensors.  Here
 * the tensors stay in HBM: in the batch-fastest column-major layout a concatenation along rank r is, for every
 * index of the ranks above r, one contiguous block from each operand -- for the highest rank a plain append --
 * so joining and splitting are strided device-to-device copies (cattl3_memcpy_2d) and the skip addition is one
 * element-wise kernel.  Modules that are device networks chain in HBM; others are bridged through the host.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_DENSENEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_DENSENEURALNETWORK_H_

#include <array>
#include <cassert>
#include <utility>
#include <vector>

#include "neural_network/CompositeNeuralNetwork.hpp"
#include "b200/DeviceNetwork.hpp"

namespace cattle {

/**
 * The ways the input of a module may be concatenated to its output.
 */
enum DenseConcatType { DENSE_LOWEST_RANK, DENSE_HIGHEST_RANK };

template<typename Scalar, std::size_t Rank, DenseConcatType ConcatType = DENSE_HIGHEST_RANK>
class DenseNeuralNetwork :
		public CompositeNeuralNetwork<Scalar,Rank,false,NeuralNetwork<Scalar,Rank,false>>,
		public b200::DeviceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,false> Base;
	typedef NeuralNetPtr<Scalar,Rank,false> Module;
	typedef DenseNeuralNetwork<Scalar,Rank,ConcatType> Self;
	typedef b200::DeviceNetwork<Scalar,Rank> DevNet;
	typedef b200::DeviceTensor<Scalar> DevTensor;
	static_assert(ConcatType >= DENSE_LOWEST_RANK && ConcatType <= DENSE_HIGHEST_RANK, "illegal merge type value");
	static constexpr std::size_t CONCAT_RANK = ConcatType == DENSE_HIGHEST_RANK ? Rank - 1 : 0;
public:
	/**
	 * @param modules The dense modules: module i takes the concatenation of the network input and the outputs of
	 * modules 0 .. i-1.
	 * @param foremost Whether the network is the first module of a composite.
	 */
	inline DenseNeuralNetwork(std::vector<Module>&& modules, bool foremost = true) :
			modules(std::move(modules)),
			foremost(foremost) {
		assert(this->modules.size() > 0 && "modules must contain at least 1 element");
		input_dims = this->modules.front()->get_input_dims();
		typename Base::Dims dims = input_dims;
		for (std::size_t i = 0; i < this->modules.size(); ++i) {
			Base& module = *this->modules[i];
			assert(module.get_input_dims() == dims && "incompatible module dimensions");
			for (std::size_t r = 0; r < Rank; ++r)
				assert((r == +CONCAT_RANK || module.get_output_dims()(r) == dims(r)) && "modules may only change the concatenation rank");
			dims(+CONCAT_RANK) += module.get_output_dims()(+CONCAT_RANK);
			module.set_foremost(i == 0 && foremost);
		}
		output_dims = dims;
	}
	inline DenseNeuralNetwork(Module&& module, bool foremost = true) :
			DenseNeuralNetwork(single(std::move(module)), foremost) { }
	inline DenseNeuralNetwork(const Self& network) :
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims) {
		for (const Module& module : network.modules)
			modules.push_back(Module(module->clone()));
	}
	inline DenseNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~DenseNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Base* clone() const {
		return new DenseNeuralNetwork(*this);
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
			const typename Base::Dims& in_dims = module->get_input_dims();
			const typename Base::Dims& out_dims = module->get_output_dims();
			DevTensor branch = run_forward(*module, input, training);
			// [input | module(input)] along the concatenation rank
			DevTensor joined(input.rows, in_dims.get_volume() + out_dims.get_volume());
			const Blocks b = blocks(input.rows, in_dims, out_dims);
			copy_blocks(joined.data(), b.pitch, input.data(), b.a, b.a, b.outer);
			copy_blocks(joined.data() + b.a, b.pitch, branch.data(), b.b, b.b, b.outer);
			input = std::move(joined);
		}
		return input;
	}
	inline DevTensor backpropagate_dev(DevTensor out_grad) {
		for (std::size_t i = modules.size(); i > 0; --i) {
			Base& module = *modules[i - 1];
			const typename Base::Dims& in_dims = module.get_input_dims();
			const typename Base::Dims& out_dims = module.get_output_dims();
			const Blocks b = blocks(out_grad.rows, in_dims, out_dims);
			DevTensor branch_grad(out_grad.rows, out_dims.get_volume());
			copy_blocks(branch_grad.data(), b.b, out_grad.data() + b.a, b.pitch, b.b, b.outer);
			DevTensor prev = run_backward(module, std::move(branch_grad));
			if (foremost && i == 1)
				return prev;  // empty: nothing upstream needs the skip part either (:150-151)
			DevTensor skip(out_grad.rows, in_dims.get_volume());
			copy_blocks(skip.data(), b.a, out_grad.data(), b.pitch, b.a, b.outer);
			if (prev.size() != skip.size())
				throw b200::Error(CATTL3_ERR_INVALID, "DenseNeuralNetwork: module returned a gradient of the wrong size");
			b200::Context& c = b200::Context::get();
			{
				b200::Context::Lock l = c.lock();
				CATTLE_B200_CHECK(b200::Api<Scalar>::add_inplace(c.handle(), (std::int64_t) skip.size(), skip.data(), prev.data()));
			}
			out_grad = std::move(skip);
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
	/** A joined tensor seen as `outer` rows of [a | b] elements (pitch = a + b). */
	struct Blocks {
		std::size_t outer, a, b, pitch;
	};
	inline static Blocks blocks(std::size_t rows, const typename Base::Dims& in_dims, const typename Base::Dims& out_dims) {
		std::size_t inner = rows, outer = 1;
		for (std::size_t r = 0; r < +CONCAT_RANK; ++r)
			inner *= in_dims(r);
		for (std::size_t r = +CONCAT_RANK + 1; r < Rank; ++r)
			outer *= in_dims(r);
		Blocks b;
		b.outer = outer;
		b.a = inner * in_dims(+CONCAT_RANK);
		b.b = inner * out_dims(+CONCAT_RANK);
		b.pitch = b.a + b.b;
		return b;
	}
	/** `outer` blocks of `width` elements from src (pitch src_pitch) to dst (pitch dst_pitch), on the device. */
	inline static void copy_blocks(Scalar* dst, std::size_t dst_pitch, const Scalar* src, std::size_t src_pitch,
			std::size_t width, std::size_t outer) {
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_memcpy_2d(c.handle(), dst, dst_pitch * sizeof(Scalar), src, src_pitch * sizeof(Scalar),
				width * sizeof(Scalar), outer));
	}
	inline static std::vector<Module> single(Module&& module) {
		std::vector<Module> vec;
		vec.push_back(std::move(module));
		return vec;
	}
	inline DevTensor run_forward(Base& module, const DevTensor& input, bool training) const {
		if (DevNet* dev_module = dynamic_cast<DevNet*>(&module))
			return dev_module->propagate_dev(input, training);
		typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(input,
				b200::batch_extents<Rank>(input.rows, module.get_input_dims()));
		return b200::to_device<Scalar,Base::DATA_RANK>(module.propagate(std::move(host), training));
	}
	inline DevTensor run_backward(Base& module, DevTensor out_grad) const {
		if (DevNet* dev_module = dynamic_cast<DevNet*>(&module))
			return dev_module->backpropagate_dev(std::move(out_grad));
		typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(out_grad,
				b200::batch_extents<Rank>(out_grad.rows, module.get_output_dims()));
		return b200::to_device<Scalar,Base::DATA_RANK>(module.backpropagate(std::move(host)));
	}
	std::vector<Module> modules;
	bool foremost;
	typename Base::Dims input_dims, output_dims;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_DENSENEURALNETWORK_H_ */
