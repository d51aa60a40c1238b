/*
 * neural_network/RecurrentNeuralNetwork.hpp -- B200 replacement of the reference's simple recurrent network
 * (C-ATTL3/neural_network/RecurrentNeuralNetwork.hpp:40-440), same class template (multiplicative integration and
 * statefulness included), constructor and interface; defines the reference header's include guard.
 *
 * Per time step, as in the reference (:207-246 forward, :274-327 backward):
 *
 *     state_t  = act(W_state state_{t-1} (+ or *) W_in x_t)        (the input term only while there is input)
 *     out_t    = act_out(W_out state_t)                            (only inside the output window)
 *
 * unrolled over the time steps with clones that share the main cell's parameters (:356-384).  The sequence stays in
 * HBM: a time step is one strided copy out of the (samples * steps) x volume sequence, the kernels and activations
 * run through their device faces (layers that only speak the host API -- the reference's Identity / Softsign
 * activations -- are bridged with a round trip), the multiplicative integration is the cattl3_muladd kernel.  The
 * host API is one upload, the device path and one download; b200::DeviceSequenceNetwork lets sequential stacks and
 * the batch loop skip that.  As with the LSTM, a copy of an unrolled network is unrolled again on its first
 * training pass instead of receiving cells with parameters of their own.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_RECURRENTNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_RECURRENTNEURALNETWORK_H_

#include <algorithm>
#include <array>
#include <cassert>
#include <functional>
#include <memory>
#include <utility>
#include <vector>

#include "layer/ActivationLayer.hpp"
#include "layer/KernelLayer.hpp"
#include "neural_network/UnidirectionalNeuralNetwork.hpp"
#include "b200/DeviceLayer.hpp"
#include "b200/DeviceSequenceNetwork.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
using KernelPtr = std::unique_ptr<KernelLayer<Scalar,Rank>>;

template<typename Scalar, std::size_t Rank>
using ActivationPtr = std::unique_ptr<ActivationLayer<Scalar,Rank>>;

template<typename Scalar, std::size_t Rank, bool MulInt = false, bool Stateful = false>
class RecurrentNeuralNetwork : public UnidirectionalNeuralNetwork<Scalar,Rank>,
		public b200::DeviceSequenceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,true> Root;
	typedef RecurrentNeuralNetwork<Scalar,Rank,MulInt,Stateful> Self;
	typedef std::function<std::pair<std::size_t,std::size_t>(std::size_t)> OutputSeqSizeFunc;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * The arguments of the reference's constructor (:50-83): the kernels applied to the input, to the previous state
	 * and to the state for the output, the state and output activations, the function from the input sequence
	 * length to (output sequence length, output delay), and the two flags.
	 */
	inline RecurrentNeuralNetwork(KernelPtr<Scalar,Rank>&& input_kernel, KernelPtr<Scalar,Rank>&& state_kernel,
			KernelPtr<Scalar,Rank>&& output_kernel, ActivationPtr<Scalar,Rank>&& state_act,
			ActivationPtr<Scalar,Rank>&& output_act, OutputSeqSizeFunc output_seq_size_func, bool reversed = false,
			bool foremost = true) :
				output_seq_size_func(output_seq_size_func),
				reversed(reversed),
				foremost(foremost),
				batch_size(-1),
				input_seq_length(-1),
				output_seq_length(-1),
				output_seq_delay(-1) {
		assert(input_kernel && state_kernel && output_kernel && state_act && output_act);
		input_dims = input_kernel->get_input_dims();
		state_dims = input_kernel->get_output_dims();
		output_dims = output_kernel->get_output_dims();
		assert(state_dims == state_kernel->get_output_dims() && state_dims == output_kernel->get_input_dims() &&
				state_dims == state_act->get_input_dims() && output_dims == output_act->get_input_dims() &&
				state_kernel->get_input_dims() == state_kernel->get_output_dims());
		main_cell.input_kernel = std::move(input_kernel);
		main_cell.state_kernel = std::move(state_kernel);
		main_cell.output_kernel = std::move(output_kernel);
		main_cell.state_act = std::move(state_act);
		main_cell.output_act = std::move(output_act);
		set_foremost(foremost);
	}
	inline RecurrentNeuralNetwork(const Self& network) :
			main_cell(network.main_cell, 0),
			output_seq_size_func(network.output_seq_size_func),
			reversed(network.reversed),
			foremost(network.foremost),
			input_dims(network.input_dims),
			state_dims(network.state_dims),
			output_dims(network.output_dims),
			state(network.state),
			batch_size(network.batch_size),
			input_seq_length(-1),
			output_seq_length(-1),
			output_seq_delay(-1) {
		state.make_exclusive();
	}
	inline RecurrentNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~RecurrentNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Root* clone() const {
		return new RecurrentNeuralNetwork(*this);
	}
	inline bool is_reversed() const {
		return reversed;
	}
	inline void reverse() {
		reversed = !reversed;
	}
	inline const typename Root::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Root::Dims& get_output_dims() const {
		return output_dims;
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		return std::vector<const Layer<Scalar,Rank>*>({ main_cell.input_kernel.get(), main_cell.state_kernel.get(),
				main_cell.output_kernel.get(), main_cell.state_act.get(), main_cell.output_act.get() });
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		return std::vector<Layer<Scalar,Rank>*>({ main_cell.input_kernel.get(), main_cell.state_kernel.get(),
				main_cell.output_kernel.get(), main_cell.state_act.get(), main_cell.output_act.get() });
	}
	inline bool is_foremost() const {
		return foremost;
	}
	inline void set_foremost(bool foremost) {
		main_cell.input_kernel->set_input_layer(foremost);
		this->foremost = foremost;
	}
	inline void empty_caches() {
		main_cell.empty_caches();
		// the hidden state and the unrolled cells go as well (:160-166)
		batch_size = -1;
		state = DevTensor();
		input_seq_length = -1;
		output_seq_length = -1;
		output_seq_delay = -1;
		cells.clear();
	}
	inline typename Root::Data propagate(typename Root::Data input, bool training) {
		assert(input_dims == (Dimensions<std::size_t,Root::DATA_RANK>(input.dimensions()).template demote<2>()));
		const std::size_t samples = input.dimension(0);
		DevTensor out = propagate_seq_dev(b200::sequence_to_device<Scalar,Root::DATA_RANK>(input), samples, training);
		return b200::sequence_to_host<Scalar,Rank>(out, samples, output_dims);
	}
	inline typename Root::Data backpropagate(typename Root::Data out_grad) {
		assert(output_dims == (Dimensions<std::size_t,Root::DATA_RANK>(out_grad.dimensions()).template demote<2>()));
		const std::size_t samples = out_grad.dimension(0);
		DevTensor prev_out_grad = backpropagate_seq_dev(b200::sequence_to_device<Scalar,Root::DATA_RANK>(out_grad), samples);
		return b200::sequence_to_host<Scalar,Rank>(prev_out_grad, samples, input_dims);
	}
	/** b200::DeviceSequenceNetwork: the unrolled forward pass on a sequence in HBM (:168-253). */
	inline DevTensor propagate_seq_dev(DevTensor input, std::size_t samples, bool training) {
		if (input.empty() || samples == 0 || input.rows % samples != 0)
			throw b200::Error(CATTL3_ERR_INVALID, "RecurrentNeuralNetwork: the input is not a sequence batch");
		const int in_len = (int) (input.rows / samples);
		const std::pair<std::size_t,std::size_t> out_info = output_seq_size_func((std::size_t) in_len);
		const int out_len = (int) out_info.first, out_delay = (int) out_info.second;
		assert(out_len > 0);
		const int out_end = out_len + out_delay;
		const int time_steps = std::max(in_len, out_end);
		// unrolled only for training and only when the sequence alignment has changed (:184-186)
		if (training && (in_len != input_seq_length || out_len != output_seq_length || out_delay != output_seq_delay))
			unroll_network(time_steps, in_len, out_delay, out_end);
		setup_hidden_state(samples);
		DevTensor out;
		if (out_len > 1)
			out = DevTensor(samples * out_len, output_dims.get_volume());
		int out_step = 0;
		for (int i = 0; i < time_steps; ++i) {
			Cell& cell = !training || i == 0 ? main_cell : cells[i - 1];
			// the state kernel is always applied (:210-211)
			state = b200::layer_forward_dev<Scalar,Rank>(*cell.state_kernel, std::move(state), training);
			if (i < in_len) {
				DevTensor x = b200::time_step_of(input, samples, (std::size_t) in_len,
						(std::size_t) (reversed ? in_len - 1 - i : i));
				DevTensor from_input = b200::layer_forward_dev<Scalar,Rank>(*cell.input_kernel, std::move(x), training);
				if (MulInt) {
					DevTensor integrated(samples, state_dims.get_volume());
					b200::tensor_muladd<Scalar>(false, state, from_input, nullptr, nullptr, integrated);
					if (training) {
						// the factors of the product, for the backward pass (:222-227)
						cell.state_kernel_cache = std::move(state);
						cell.input_kernel_cache = std::move(from_input);
					}
					state = std::move(integrated);
				} else {
					b200::tensor_add<Scalar>(state, from_input);
				}
			}
			state = b200::layer_forward_dev<Scalar,Rank>(*cell.state_act, std::move(state), training);
			if (i >= out_delay && i < out_end) {
				DevTensor out_i = b200::layer_forward_dev<Scalar,Rank>(*cell.output_act,
						b200::layer_forward_dev<Scalar,Rank>(*cell.output_kernel, state, training), training);
				if (out_len > 1)
					b200::set_time_step(out, samples, (std::size_t) out_len, (std::size_t) out_step++, out_i);
				else
					out = std::move(out_i);
			}
		}
		batch_size = (int) samples;
		input_seq_length = in_len;
		output_seq_length = out_len;
		output_seq_delay = out_delay;
		return out;
	}
	/** b200::DeviceSequenceNetwork: back-propagation through time on the device (:254-330). */
	inline DevTensor backpropagate_seq_dev(DevTensor out_grad, std::size_t samples) {
		if (out_grad.empty() || (int) samples != batch_size || out_grad.rows != samples * (std::size_t) output_seq_length)
			throw b200::Error(CATTL3_ERR_INVALID, "RecurrentNeuralNetwork: the gradient does not match the last training pass");
		const int in_len = input_seq_length, out_len = output_seq_length, out_delay = output_seq_delay;
		const int out_end = out_len + out_delay;
		const int time_steps = std::max(in_len, out_end);
		DevTensor prev_out_grad;
		if (!foremost && in_len > 1)
			prev_out_grad = DevTensor(samples * in_len, input_dims.get_volume());
		DevTensor state_grad(samples, state_dims.get_volume(), true);
		int out_step = out_len - 1, in_step = in_len - 1;
		for (int i = time_steps - 1; i >= 0; --i) {
			Cell& cell = i == 0 ? main_cell : cells[i - 1];
			// an output at this step feeds its gradient into the state's (:283-295)
			if (i >= out_delay && i < out_end) {
				DevTensor out_grad_i = b200::time_step_of(out_grad, samples, (std::size_t) out_len, (std::size_t) out_step--);
				b200::tensor_add<Scalar>(state_grad, b200::layer_backward_dev<Scalar,Rank>(*cell.output_kernel,
						b200::layer_backward_dev<Scalar,Rank>(*cell.output_act, std::move(out_grad_i))));
			}
			state_grad = b200::layer_backward_dev<Scalar,Rank>(*cell.state_act, std::move(state_grad));
			if (i < in_len) {
				// through the input kernel (its parameter gradients always; its input gradient unless foremost), :299-320
				DevTensor input_grad = b200::layer_backward_dev<Scalar,Rank>(*cell.input_kernel,
						MulInt ? b200::tensor_product<Scalar>(cell.state_kernel_cache, state_grad) : state_grad);
				if (!foremost) {
					if (in_len > 1)
						b200::set_time_step(prev_out_grad, samples, (std::size_t) in_len, (std::size_t) in_step--, input_grad);
					else
						prev_out_grad = std::move(input_grad);
				}
				// through the state kernel (:322-326)
				state_grad = b200::layer_backward_dev<Scalar,Rank>(*cell.state_kernel,
						MulInt ? b200::tensor_product<Scalar>(cell.input_kernel_cache, state_grad) : std::move(state_grad));
			} else {
				state_grad = b200::layer_backward_dev<Scalar,Rank>(*cell.state_kernel, std::move(state_grad));
			}
		}
		return prev_out_grad;
	}
	/** A stateful network carries its hidden state from step to step: not for a captured step graph. */
	inline bool graph_safe() const {
		return !Stateful;
	}
	inline friend void swap(Self& network1, Self& network2) {
		using std::swap;
		swap(network1.main_cell, network2.main_cell);
		swap(network1.output_seq_size_func, network2.output_seq_size_func);
		swap(network1.reversed, network2.reversed);
		swap(network1.foremost, network2.foremost);
		swap(network1.input_dims, network2.input_dims);
		swap(network1.state_dims, network2.state_dims);
		swap(network1.output_dims, network2.output_dims);
		swap(network1.cells, network2.cells);
		swap(network1.state, network2.state);
		swap(network1.batch_size, network2.batch_size);
		swap(network1.input_seq_length, network2.input_seq_length);
		swap(network1.output_seq_length, network2.output_seq_length);
		swap(network1.output_seq_delay, network2.output_seq_delay);
	}
private:
	/** One cell of the unrolled network: its layers and, for multiplicative integration, the factors of its product. */
	struct Cell {
		inline Cell() { }
		/** Deep copy of the layers (Layer::clone()), without the caches. */
		inline Cell(const Cell& cell, int) :
				input_kernel(clone_of(cell.input_kernel)),
				state_kernel(clone_of(cell.state_kernel)),
				output_kernel(clone_of(cell.output_kernel)),
				state_act(clone_of(cell.state_act)),
				output_act(clone_of(cell.output_act)) { }
		Cell(Cell&&) = default;
		Cell& operator=(Cell&&) = default;
		inline void empty_caches() {
			if (input_kernel) input_kernel->empty_cache();
			if (state_kernel) state_kernel->empty_cache();
			if (output_kernel) output_kernel->empty_cache();
			if (state_act) state_act->empty_cache();
			if (output_act) output_act->empty_cache();
			state_kernel_cache = input_kernel_cache = DevTensor();
		}
		template<typename L>
		inline static std::unique_ptr<L> clone_of(const std::unique_ptr<L>& layer) {
			return std::unique_ptr<L>(layer ? static_cast<L*>(layer->clone()) : nullptr);
		}
		template<typename L>
		inline static std::unique_ptr<L> shared_clone_of(const std::unique_ptr<L>& layer) {
			return std::unique_ptr<L>(static_cast<L*>(layer->clone_with_shared_params()));
		}
		KernelPtr<Scalar,Rank> input_kernel, state_kernel, output_kernel;
		ActivationPtr<Scalar,Rank> state_act, output_act;
		DevTensor state_kernel_cache, input_kernel_cache;
	};
	/** Cells for the steps after the first: clones with shared parameters, only of what the step uses (:356-384). */
	inline void unroll_network(std::size_t time_steps, std::size_t in_len, std::size_t out_delay, std::size_t out_end) {
		if (time_steps <= 1) {
			cells.clear();
			return;
		}
		empty_caches();
		cells.resize(time_steps - 1);
		for (std::size_t j = 1; j < time_steps; ++j) {
			Cell& cell = cells[j - 1];
			cell.state_kernel = Cell::shared_clone_of(main_cell.state_kernel);
			cell.state_act = Cell::shared_clone_of(main_cell.state_act);
			if (j < in_len)
				cell.input_kernel = Cell::shared_clone_of(main_cell.input_kernel);
			if (j >= out_delay && j < out_end) {
				cell.output_kernel = Cell::shared_clone_of(main_cell.output_kernel);
				cell.output_act = Cell::shared_clone_of(main_cell.output_act);
			}
		}
	}
	/** Zero state for a new sequence; a stateful network keeps its state, grown or cut to the batch (:385-408). */
	inline void setup_hidden_state(std::size_t samples) {
		const std::size_t volume = state_dims.get_volume();
		if (!Stateful || batch_size == -1 || state.empty()) {
			state = DevTensor(samples, volume, true);
		} else if (samples != (std::size_t) batch_size) {
			DevTensor new_state(samples, volume, true);
			const std::size_t kept = std::min(samples, (std::size_t) batch_size);
			b200::Context& c = b200::Context::get();
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(cattl3_memcpy_2d(c.handle(), new_state.data(), samples * sizeof(Scalar), state.data(),
					(std::size_t) batch_size * sizeof(Scalar), kept * sizeof(Scalar), volume));
			state = std::move(new_state);
		}
	}
	Cell main_cell;
	OutputSeqSizeFunc output_seq_size_func;
	bool reversed, foremost;
	typename Root::Dims input_dims, state_dims, output_dims;
	std::vector<Cell> cells;
	DevTensor state;
	int batch_size, input_seq_length, output_seq_length, output_seq_delay;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_RECURRENTNEURALNETWORK_H_ */
