/*
 * neural_network/LSTMNeuralNetwork.hpp -- B200 replacement of the reference's LSTM network
 * (C-ATTL3/neural_network/LSTMNeuralNetwork.hpp:44-700), same class template (multiplicative integration and
 * statefulness included), constructor and interface; defines the reference header's include guard.
 *
 * The network is unrolled over the time steps as in the reference (:537-573: every step after the first gets clones
 * of the kernels and activations that SHARE the main cell's parameters and keep their own caches), and the same
 * equations run per step (:283-383 forward, :441-499 backward):
 *
 *     forget / write / read filters  = act(W_in x_t (+ or *) W_out h_{t-1})
 *     candidates                     = act(W_in x_t (+ or *) W_out h_{t-1})
 *     state_t  = forget * state_{t-1} + write * candidates
 *     h_t      = read * act(state_t)
 *
 * What changes is where they run: the whole sequence stays in HBM.  A time step is one strided copy out of the
 * (samples * steps) x volume sequence (b200::time_step_of), the kernels and activations are driven through their
 * device faces (b200::DeviceLayer; layers that only speak the host API are bridged with a round trip), and the gate
 * arithmetic is the cattl3_muladd kernel.  The cells share their kernels' parameters, so the eight weight gradients are
 * taken once over the whole sequence (one GEMM over samples * steps rows each, b200::SplitBackwardLayer) instead of
 * once per time step; CATTL3_LSTM_STEPWISE_WGRAD=1 keeps the per-step form (A/B, tests).  The host API (propagate / backpropagate on rank + 2 tensors) is one upload,
 * the device path and one download; b200::DeviceSequenceNetwork lets a sequential stack and the batch loop skip
 * even that.
 *
 * Deviations from the reference, both outside of what its own callers exercise: a copy of an unrolled network is
 * unrolled again on its first training pass (the reference's copy gives the copied cells parameters of their own,
 * :226-271 with Layer::clone(), which detaches them from the main cell), and a Stateful network whose state is kept
 * across sequences keeps it in HBM.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_LSTMNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_LSTMNEURALNETWORK_H_

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdlib>
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
class LSTMNeuralNetwork : public UnidirectionalNeuralNetwork<Scalar,Rank>, public b200::DeviceSequenceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,true> Root;
	typedef UnidirectionalNeuralNetwork<Scalar,Rank> Base;
	typedef LSTMNeuralNetwork<Scalar,Rank,MulInt,Stateful> Self;
	typedef std::function<std::pair<std::size_t,std::size_t>(std::size_t)> OutputSeqSizeFunc;
	typedef b200::DeviceTensor<Scalar> DevTensor;
	enum Gate { FORGET, WRITE, CANDIDATE, READ, GATES };
public:
	/**
	 * The arguments of the reference's constructor (:53-98): per gate (forget, write, candidate, read) the kernel
	 * applied to the input and the kernel applied to the previous hidden output, the five activations, the function
	 * from the input sequence length to (output sequence length, output delay), and the two flags.
	 */
	inline LSTMNeuralNetwork(KernelPtr<Scalar,Rank>&& input_forget_kernel,
			KernelPtr<Scalar,Rank>&& output_forget_kernel, KernelPtr<Scalar,Rank>&& input_write_kernel,
			KernelPtr<Scalar,Rank>&& output_write_kernel, KernelPtr<Scalar,Rank>&& input_candidate_kernel,
			KernelPtr<Scalar,Rank>&& output_candidate_kernel, KernelPtr<Scalar,Rank>&& input_read_kernel,
			KernelPtr<Scalar,Rank>&& output_read_kernel, ActivationPtr<Scalar,Rank>&& forget_act,
			ActivationPtr<Scalar,Rank>&& write_act, ActivationPtr<Scalar,Rank>&& candidate_act,
			ActivationPtr<Scalar,Rank>&& state_act, ActivationPtr<Scalar,Rank>&& read_act,
			OutputSeqSizeFunc output_seq_size_func, bool reversed = false, bool foremost = true) :
				output_seq_size_func(output_seq_size_func),
				reversed(reversed),
				foremost(foremost),
				batch_size(-1),
				input_seq_length(-1),
				output_seq_length(-1),
				output_seq_delay(-1) {
		assert(output_forget_kernel && input_forget_kernel && output_write_kernel && input_write_kernel &&
				output_candidate_kernel && input_candidate_kernel && output_read_kernel && input_read_kernel &&
				forget_act && write_act && candidate_act && state_act && read_act);
		input_dims = input_forget_kernel->get_input_dims();
		output_dims = input_forget_kernel->get_output_dims();
		main_cell.in_kernel[FORGET] = std::move(input_forget_kernel);
		main_cell.out_kernel[FORGET] = std::move(output_forget_kernel);
		main_cell.in_kernel[WRITE] = std::move(input_write_kernel);
		main_cell.out_kernel[WRITE] = std::move(output_write_kernel);
		main_cell.in_kernel[CANDIDATE] = std::move(input_candidate_kernel);
		main_cell.out_kernel[CANDIDATE] = std::move(output_candidate_kernel);
		main_cell.in_kernel[READ] = std::move(input_read_kernel);
		main_cell.out_kernel[READ] = std::move(output_read_kernel);
		main_cell.act[FORGET] = std::move(forget_act);
		main_cell.act[WRITE] = std::move(write_act);
		main_cell.act[CANDIDATE] = std::move(candidate_act);
		main_cell.act[READ] = std::move(read_act);
		main_cell.state_act = std::move(state_act);
		for (int g = 0; g < GATES; ++g) {
			assert(main_cell.in_kernel[g]->get_input_dims() == input_dims &&
					main_cell.in_kernel[g]->get_output_dims() == output_dims &&
					main_cell.out_kernel[g]->get_input_dims() == output_dims &&
					main_cell.out_kernel[g]->get_output_dims() == output_dims &&
					main_cell.act[g]->get_input_dims() == output_dims);
		}
		assert(main_cell.state_act->get_input_dims() == output_dims);
		set_foremost(foremost);
	}
	inline LSTMNeuralNetwork(const Self& network) :
			main_cell(network.main_cell, false),
			output_seq_size_func(network.output_seq_size_func),
			reversed(network.reversed),
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims),
			state(network.state),
			batch_size(network.batch_size),
			input_seq_length(-1),
			output_seq_length(-1),
			output_seq_delay(-1) {
		state.make_exclusive();
	}
	inline LSTMNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~LSTMNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Root* clone() const {
		return new LSTMNeuralNetwork(*this);
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
		std::vector<const Layer<Scalar,Rank>*> layer_ptrs(13);
		populate_layer_vector<const Layer<Scalar,Rank>*>(layer_ptrs);
		return layer_ptrs;
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		std::vector<Layer<Scalar,Rank>*> layer_ptrs(13);
		populate_layer_vector<Layer<Scalar,Rank>*>(layer_ptrs);
		return layer_ptrs;
	}
	inline bool is_foremost() const {
		return foremost;
	}
	inline void set_foremost(bool foremost) {
		for (int g = 0; g < GATES; ++g)
			main_cell.in_kernel[g]->set_input_layer(foremost);
		this->foremost = foremost;
	}
	inline void empty_caches() {
		main_cell.empty_caches();
		// the hidden state and the unrolled cells go as well (:243-249)
		batch_size = -1;
		state = DevTensor();
		input_seq = hidden_seq = DevTensor();
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
	/** b200::DeviceSequenceNetwork: the unrolled forward pass on a sequence in HBM (:251-389). */
	inline DevTensor propagate_seq_dev(DevTensor input, std::size_t samples, bool training) {
		if (input.empty() || samples == 0 || input.rows % samples != 0)
			throw b200::Error(CATTL3_ERR_INVALID, "LSTMNeuralNetwork: the input is not a sequence batch");
		const int in_len = (int) (input.rows / samples);
		const std::pair<std::size_t,std::size_t> out_info = output_seq_size_func((std::size_t) in_len);
		const int out_len = (int) out_info.first, out_delay = (int) out_info.second;
		assert(out_len > 0);
		const int out_end = out_len + out_delay;
		const int time_steps = std::max(in_len, out_end);
		// unrolled only for training and only when the sequence alignment has changed (:263-265)
		if (training && (in_len != input_seq_length || out_len != output_seq_length || out_delay != output_seq_delay))
			unroll_network(time_steps, in_len);
		setup_hidden_state(samples);
		const std::size_t out_volume = output_dims.get_volume();
		DevTensor out;
		if (out_len > 1)
			out = DevTensor(samples * out_len, out_volume);
		// batched weight gradients (see backpropagate_seq_dev): the input sequence and the hidden outputs as sequences
		const bool batched = training && batched_weight_gradients();
		input_seq = batched ? input : DevTensor();
		hidden_seq = batched && time_steps > 1 ? DevTensor(samples * (time_steps - 1), out_volume) : DevTensor();
		// The input kernels do not depend on the recurrence: in the batched mode each runs ONCE over the whole input
		// sequence (samples * steps rows), and a step takes its slice of that output.
		DevTensor input_parts[GATES];
		if (batched) {
			for (int g = 0; g < GATES; ++g)
				input_parts[g] = forward(*main_cell.in_kernel[g], input, training);
		}
		DevTensor hidden_out;
		int out_step = 0;
		for (int i = 0; i < time_steps; ++i) {
			Cell& cell = !training || i == 0 ? main_cell : cells[i - 1];
			const bool has_input = i < in_len, has_hidden = i > 0;
			const std::size_t in_slot = (std::size_t) (reversed ? in_len - 1 - i : i);
			DevTensor x, part[GATES];
			if (has_input && batched) {
				for (int g = 0; g < GATES; ++g)
					part[g] = b200::time_step_of(input_parts[g], samples, (std::size_t) in_len, in_slot);
			} else if (has_input) {
				x = b200::time_step_of(input, samples, (std::size_t) in_len, in_slot);
			}
			const bool sliced = has_input && batched;
			// state update: selective remembrance, then the filtered candidates (:290-296, :318-347, :352-361)
			cell.forget_filter = gate(cell, FORGET, x, hidden_out, has_input, has_hidden, training, sliced ? &part[FORGET] : nullptr);
			cell.prev_state = std::move(state);
			cell.write_filter = gate(cell, WRITE, x, hidden_out, has_input, has_hidden, training, sliced ? &part[WRITE] : nullptr);
			cell.candidates = gate(cell, CANDIDATE, x, hidden_out, has_input, has_hidden, training,
					sliced ? &part[CANDIDATE] : nullptr);
			state = DevTensor(samples, out_volume);
			muladd(false, cell.forget_filter, cell.prev_state, &cell.write_filter, &cell.candidates, state);
			// output computation (:364-385)
			cell.read_filter = gate(cell, READ, x, hidden_out, has_input, has_hidden, training, sliced ? &part[READ] : nullptr);
			cell.activated_state = forward(*cell.state_act, state, training);
			hidden_out = DevTensor(samples, out_volume);
			muladd(false, cell.read_filter, cell.activated_state, nullptr, nullptr, hidden_out);
			if (batched && i + 1 < time_steps)
				b200::set_time_step(hidden_seq, samples, (std::size_t) (time_steps - 1), (std::size_t) i, hidden_out);
			if (i >= out_delay && i < out_end) {
				if (out_len > 1)
					b200::set_time_step(out, samples, (std::size_t) out_len, (std::size_t) out_step++, hidden_out);
				else
					out = hidden_out;
			}
		}
		batch_size = (int) samples;
		input_seq_length = in_len;
		output_seq_length = out_len;
		output_seq_delay = out_delay;
		return out;
	}
	/** b200::DeviceSequenceNetwork: back-propagation through time on the device (:390-503). */
	inline DevTensor backpropagate_seq_dev(DevTensor out_grad, std::size_t samples) {
		if (out_grad.empty() || (int) samples != batch_size || out_grad.rows != samples * (std::size_t) output_seq_length)
			throw b200::Error(CATTL3_ERR_INVALID, "LSTMNeuralNetwork: the gradient does not match the last training pass");
		const int in_len = input_seq_length, out_len = output_seq_length, out_delay = output_seq_delay;
		const int out_end = out_len + out_delay;
		const int time_steps = std::max(in_len, out_end);
		const std::size_t out_volume = output_dims.get_volume();
		DevTensor prev_out_grad;
		if (!foremost && in_len > 1 && (input_seq.empty() || reversed))
			prev_out_grad = DevTensor(samples * in_len, input_dims.get_volume());
		DevTensor state_grad(samples, out_volume, true), hidden_out_grad(samples, out_volume, true);
		int out_step = out_len - 1, in_step = in_len - 1;
		// The cells share their kernels' parameters, so a kernel's weight gradient over the whole sequence is one GEMM
		// over samples * steps rows (b200::SplitBackwardLayer) instead of one small GEMM and split-K reduction per step:
		// per step only the input gradients are taken, the gate gradients are collected as sequences, and the eight
		// weight (and bias) gradients follow the loop.  Not with multiplicative integration (its kernels see products).
		const bool batched = !input_seq.empty();
		DevTensor gate_grads_in[GATES], gate_grads_out[GATES];
		if (batched) {
			for (int g = 0; g < GATES; ++g) {
				gate_grads_in[g] = DevTensor(samples * in_len, out_volume);
				if (time_steps > 1)
					gate_grads_out[g] = DevTensor(samples * (time_steps - 1), out_volume);
			}
		}
		for (int i = time_steps - 1; i >= 0; --i) {
			Cell& cell = i == 0 ? main_cell : cells[i - 1];
			// the gradient of a non-hidden output at this step joins the hidden output's (:420-428)
			if (i >= out_delay && i < out_end)
				add(hidden_out_grad, b200::time_step_of(out_grad, samples, (std::size_t) out_len, (std::size_t) out_step--));
			add(state_grad, backward(*cell.state_act, product(cell.read_filter, hidden_out_grad)));
			DevTensor grad[GATES];
			grad[READ] = backward(*cell.act[READ], product(cell.activated_state, hidden_out_grad));
			grad[CANDIDATE] = backward(*cell.act[CANDIDATE], product(cell.write_filter, state_grad));
			grad[WRITE] = backward(*cell.act[WRITE], product(cell.candidates, state_grad));
			grad[FORGET] = backward(*cell.act[FORGET], product(cell.prev_state, state_grad));
			state_grad.make_exclusive();
			muladd(false, state_grad, cell.forget_filter, nullptr, nullptr, state_grad);
			// through the kernels, in the reference's order of summation: read, candidate, write, forget (:437-494)
			static const Gate order[GATES] = { READ, CANDIDATE, WRITE, FORGET };
			const bool has_input = i < in_len, has_hidden = i > 0;
			const bool integrated = MulInt && has_input && has_hidden;
			if (batched) {
				for (int g = 0; g < GATES; ++g) {
					if (has_input)
						b200::set_time_step(gate_grads_in[g], samples, (std::size_t) in_len,
								(std::size_t) (reversed ? in_len - 1 - i : i), grad[g]);
					if (has_hidden)
						b200::set_time_step(gate_grads_out[g], samples, (std::size_t) (time_steps - 1), (std::size_t) (i - 1), grad[g]);
				}
			}
			if (has_hidden) {
				DevTensor sum;
				for (Gate g : order) {
					DevTensor part = batched ? split(*main_cell.out_kernel[g]).pass_back_input_dev(grad[g]) :
							backward(*cell.out_kernel[g], integrated ? product(cell.weighted_in[g], grad[g]) : grad[g]);
					if (sum.empty())
						sum = std::move(part);
					else
						add(sum, part);
				}
				hidden_out_grad = std::move(sum);
			}
			if (has_input && !(batched && !reversed)) {
				DevTensor sum;
				for (Gate g : order) {
					if (batched && foremost)
						continue;   // input layers: no input gradient, and the weight gradients follow the loop
					DevTensor part = batched ? split(*main_cell.in_kernel[g]).pass_back_input_dev(grad[g]) :
							backward(*cell.in_kernel[g], integrated ? product(cell.weighted_out[g], grad[g]) : grad[g]);
					if (foremost || part.empty())
						continue;   // input layers return nothing (C-ATTL3/core/Layer.hpp:82-90)
					if (sum.empty())
						sum = std::move(part);
					else
						add(sum, part);
				}
				if (!foremost) {
					if (in_len > 1)
						b200::set_time_step(prev_out_grad, samples, (std::size_t) in_len, (std::size_t) in_step--, sum);
					else
						prev_out_grad = std::move(sum);
				}
			}
		}
		if (batched && !reversed && !foremost) {
			// the input gradient of the input kernels, once over the whole sequence, in the per-step order of summation.
			// (A reversed network keeps the per-step form above: the reference returns its gradient in loop order, :483-489.)
			static const Gate order[GATES] = { READ, CANDIDATE, WRITE, FORGET };
			for (Gate g : order) {
				DevTensor part = split(*main_cell.in_kernel[g]).pass_back_input_dev(gate_grads_in[g]);
				if (prev_out_grad.empty() || g == READ)
					prev_out_grad = std::move(part);
				else
					add(prev_out_grad, part);
			}
		}
		if (batched) {
			for (int g = 0; g < GATES; ++g) {
				split(*main_cell.in_kernel[g]).accumulate_param_grads_dev(input_seq, gate_grads_in[g]);
				if (time_steps > 1)
					split(*main_cell.out_kernel[g]).accumulate_param_grads_dev(hidden_seq, gate_grads_out[g]);
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
		swap(network1.output_dims, network2.output_dims);
		swap(network1.cells, network2.cells);
		swap(network1.state, network2.state);
		swap(network1.input_seq, network2.input_seq);
		swap(network1.hidden_seq, network2.hidden_seq);
		swap(network1.batch_size, network2.batch_size);
		swap(network1.input_seq_length, network2.input_seq_length);
		swap(network1.output_seq_length, network2.output_seq_length);
		swap(network1.output_seq_delay, network2.output_seq_delay);
	}
private:
	/** One cell of the unrolled network: its layers and what the backward pass needs of its forward pass. */
	struct Cell {
		inline Cell() { }
		/** Deep copy of the layers (Layer::clone()); the caches only if asked for. */
		inline Cell(const Cell& cell, bool with_caches) :
				state_act(clone_of(cell.state_act)) {
			for (int g = 0; g < GATES; ++g) {
				in_kernel[g] = clone_of(cell.in_kernel[g]);
				out_kernel[g] = clone_of(cell.out_kernel[g]);
				act[g] = clone_of(cell.act[g]);
			}
			if (with_caches) {
				forget_filter = cell.forget_filter; prev_state = cell.prev_state; write_filter = cell.write_filter;
				candidates = cell.candidates; read_filter = cell.read_filter; activated_state = cell.activated_state;
				for (int g = 0; g < GATES; ++g) {
					weighted_in[g] = cell.weighted_in[g];
					weighted_out[g] = cell.weighted_out[g];
				}
			}
		}
		Cell(Cell&&) = default;
		Cell& operator=(Cell&&) = default;
		inline void empty_caches() {
			for (int g = 0; g < GATES; ++g) {
				if (in_kernel[g]) in_kernel[g]->empty_cache();
				if (out_kernel[g]) out_kernel[g]->empty_cache();
				if (act[g]) act[g]->empty_cache();
				weighted_in[g] = weighted_out[g] = DevTensor();
			}
			if (state_act) state_act->empty_cache();
			forget_filter = prev_state = write_filter = candidates = read_filter = activated_state = DevTensor();
		}
		template<typename L>
		inline static std::unique_ptr<L> clone_of(const std::unique_ptr<L>& layer) {
			return std::unique_ptr<L>(layer ? static_cast<L*>(layer->clone()) : nullptr);
		}
		template<typename L>
		inline static std::unique_ptr<L> shared_clone_of(const std::unique_ptr<L>& layer) {
			return std::unique_ptr<L>(static_cast<L*>(layer->clone_with_shared_params()));
		}
		KernelPtr<Scalar,Rank> in_kernel[GATES], out_kernel[GATES];
		ActivationPtr<Scalar,Rank> act[GATES], state_act;
		// the factors of the multiplicative filtering operations
		DevTensor forget_filter, prev_state, write_filter, candidates, read_filter, activated_state;
		// the factors of multiplicative integration
		DevTensor weighted_in[GATES], weighted_out[GATES];
	};
	template<typename _LayerPtr>
	inline void populate_layer_vector(std::vector<_LayerPtr>& layer_ptrs) const {
		// the reference's order (:520-534): kernel pairs per gate, then forget, write, candidate, read, state activations
		for (int g = 0; g < GATES; ++g) {
			layer_ptrs[2 * g] = main_cell.in_kernel[g].get();
			layer_ptrs[2 * g + 1] = main_cell.out_kernel[g].get();
		}
		layer_ptrs[8] = main_cell.act[FORGET].get();
		layer_ptrs[9] = main_cell.act[WRITE].get();
		layer_ptrs[10] = main_cell.act[CANDIDATE].get();
		layer_ptrs[11] = main_cell.act[READ].get();
		layer_ptrs[12] = main_cell.state_act.get();
	}
	/** Cells for the steps after the first: clones with shared parameters; input kernels only while there is input. */
	inline void unroll_network(std::size_t time_steps, std::size_t in_len) {
		empty_caches();
		if (time_steps <= 1)
			return;
		cells.resize(time_steps - 1);
		for (std::size_t j = 1; j < time_steps; ++j) {
			Cell& cell = cells[j - 1];
			for (int g = 0; g < GATES; ++g) {
				cell.out_kernel[g] = Cell::shared_clone_of(main_cell.out_kernel[g]);
				cell.act[g] = Cell::shared_clone_of(main_cell.act[g]);
				if (j < in_len)
					cell.in_kernel[g] = Cell::shared_clone_of(main_cell.in_kernel[g]);
			}
			cell.state_act = Cell::shared_clone_of(main_cell.state_act);
		}
	}
	/** Zero state for a new sequence; a stateful network keeps its state, grown or cut to the batch (:575-597). */
	inline void setup_hidden_state(std::size_t samples) {
		const std::size_t volume = output_dims.get_volume();
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
	/**
	 * One gate of one step: the activation of the input kernel's and / or the hidden kernel's output, summed or
	 * (multiplicative integration) multiplied where both exist.
	 */
	inline static DevTensor gate(Cell& cell, int g, const DevTensor& x, const DevTensor& hidden_out, bool has_input,
			bool has_hidden, bool training, const DevTensor* input_part = nullptr) {
		// input_part: this step's slice of the input kernel's output over the whole sequence (batched input kernels)
		DevTensor weighted;
		if (has_input && has_hidden) {
			DevTensor from_input = input_part ? *input_part : forward(*cell.in_kernel[g], x, training);
			DevTensor from_hidden = forward(*cell.out_kernel[g], hidden_out, training);
			if (MulInt) {
				weighted = DevTensor(from_input.rows, from_input.size() / from_input.rows);
				muladd(false, from_input, from_hidden, nullptr, nullptr, weighted);
				if (training) {
					cell.weighted_in[g] = std::move(from_input);
					cell.weighted_out[g] = std::move(from_hidden);
				}
			} else {
				from_input.make_exclusive();
				add(from_input, from_hidden);
				weighted = std::move(from_input);
			}
		} else if (has_input) {
			weighted = input_part ? *input_part : forward(*cell.in_kernel[g], x, training);
		} else {
			weighted = forward(*cell.out_kernel[g], hidden_out, training);
		}
		return forward(*cell.act[g], std::move(weighted), training);
	}
	// the cell's building blocks (b200/DeviceSequenceNetwork.hpp)
	inline static DevTensor forward(Layer<Scalar,Rank>& layer, DevTensor in, bool training) {
		return b200::layer_forward_dev<Scalar,Rank>(layer, std::move(in), training);
	}
	inline static DevTensor backward(Layer<Scalar,Rank>& layer, DevTensor out_grad) {
		return b200::layer_backward_dev<Scalar,Rank>(layer, std::move(out_grad));
	}
	inline static void muladd(bool accumulate, const DevTensor& a, const DevTensor& b, const DevTensor* c_factor,
			const DevTensor* d_factor, DevTensor& out) {
		b200::tensor_muladd<Scalar>(accumulate, a, b, c_factor, d_factor, out);
	}
	inline static DevTensor product(const DevTensor& a, const DevTensor& b) {
		return b200::tensor_product<Scalar>(a, b);
	}
	inline static void add(DevTensor& y, const DevTensor& x) {
		b200::tensor_add<Scalar>(y, x);
	}
	/** Whether every kernel's backward pass can be split (and the network does not integrate multiplicatively). */
	inline bool batched_weight_gradients() const {
		static const bool enabled = [] {
			const char* v = std::getenv("CATTL3_LSTM_STEPWISE_WGRAD");
			return !(v && v[0] && v[0] != '0');
		}();
		if (MulInt || !enabled)
			return false;
		for (int g = 0; g < GATES; ++g) {
			const b200::SplitBackwardLayer<Scalar,Rank>* in = dynamic_cast<const b200::SplitBackwardLayer<Scalar,Rank>*>(
					main_cell.in_kernel[g].get());
			const b200::SplitBackwardLayer<Scalar,Rank>* out = dynamic_cast<const b200::SplitBackwardLayer<Scalar,Rank>*>(
					main_cell.out_kernel[g].get());
			if (!in || !out || !in->can_split_backward() || !out->can_split_backward())
				return false;
		}
		return true;
	}
	inline static b200::SplitBackwardLayer<Scalar,Rank>& split(KernelLayer<Scalar,Rank>& kernel) {
		return dynamic_cast<b200::SplitBackwardLayer<Scalar,Rank>&>(kernel);
	}
	Cell main_cell;
	OutputSeqSizeFunc output_seq_size_func;
	bool reversed, foremost;
	typename Root::Dims input_dims, output_dims;
	std::vector<Cell> cells;
	DevTensor state;
	// the last training pass's input sequence and hidden outputs h_0 ... h_{steps - 2} as sequences (batched weight gradients)
	DevTensor input_seq, hidden_seq;
	int batch_size, input_seq_length, output_seq_length, output_seq_delay;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_LSTMNEURALNETWORK_H_ */
