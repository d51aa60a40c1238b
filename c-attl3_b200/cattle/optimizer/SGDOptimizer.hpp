/*
 * optimizer/SGDOptimizer.hpp -- B200 replacement of the reference's mini-batch loop
 * (C-ATTL3/optimizer/SGDOptimizer.hpp:21-128): same class template, same constructor, same protected
 * virtuals (_fit, _update_params), so user-defined optimizers derived from it keep working; defines the
 * reference header's include guard.
 *
 * What the loop does per mini-batch is what the reference does (:44-71) -- propagate, loss, back-propagate
 * the loss gradient divided by the NOMINAL batch size, regularise, update, reset the gradients, with a
 * time step that keeps counting across epochs and is only reset by fit() (:28-32,66-67).  What changes:
 *
 *  - the network runs on the device; parameter values, gradients and optimizer state never leave HBM
 *    (parameters/B200Parameters.hpp), and a derived optimizer's whole update rule is ONE fused kernel per
 *    parameter array (fused_step(): update + gradient reset, cattl3_optimizer_step in
 *    include/cattl3_b200.h) instead of a chain of whole-matrix Eigen expressions
 *    (e.g. C-ATTL3/optimizer/NadamOptimizer.hpp:44-63);
 *  - data parallelism (absent from the reference): with a b200::Communicator of world size G > 1, one
 *    process per GPU, every rank draws the same mini-batch from its (identical, deterministic) provider,
 *    keeps rows [r*B/G, (r+1)*B/G), and the parameter gradients are sum-all-reduced over NVLink before
 *    the update.  Because the loss gradient is divided by the nominal *global* batch size, the reduced
 *    gradient is exactly the single-process gradient, and every rank applies the identical update.
 */
#ifndef C_ATTL3_OPTIMIZER_SGDOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_SGDOPTIMIZER_H_

#include <array>
#include <cassert>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "core/Optimizer.hpp"
#include "parameters/B200Parameters.hpp"
#include "b200/Communicator.hpp"
#include "b200/DeviceDataSource.hpp"
#include "b200/DeviceLoss.hpp"
#include "b200/DeviceNetwork.hpp"
#include "b200/Runtime.hpp"
#include "layer/DropoutLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class SGDOptimizer : public Optimizer<Scalar,Rank,Sequential> {
	typedef Optimizer<Scalar,Rank,Sequential> Base;
public:
	/**
	 * @param loss The loss function to minimise.
	 * @param batch_size The nominal (global) mini-batch size.
	 */
	inline SGDOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size) :
			Base::Optimizer(loss),
			batch_size(batch_size),
			timestep(0),
			target_net_ptr(nullptr) {
		assert(batch_size > 0);
	}
	virtual ~SGDOptimizer() = default;
	inline void fit(typename Base::Net& net) {
		timestep = 0;
		target_net_ptr = &net;
		device_states.clear();
		host_states.clear();
		drop_graph();
		carried_graph_loss = 0;
		graphs_failed = false;
		fused_step_seen = false;
		eager_steps_at_shape = 0;
		exchange = GradientExchange();
		std::vector<Parameters<Scalar>*> params_vec = net.get_all_unique_params();
		pack_parameters(params_vec);
		_fit(params_vec);
	}
protected:
	inline Scalar _train(typename Base::Net& net, typename Base::Provider& training_prov, std::size_t epoch,
			bool verbose) {
		assert(target_net_ptr == &net);
		b200::Communicator& comm = b200::Communicator::get();
		double obj_loss = 0, reg_loss = 0;
		std::size_t instances = 0, updates = 0;
		std::vector<Parameters<Scalar>*> params_vec = net.get_all_unique_params();
		DeviceFace face(net, device_loop());
		DeviceFace* dev_net = face ? &face : nullptr;
		const b200::DeviceLoss<Scalar>* dev_loss = dynamic_cast<const b200::DeviceLoss<Scalar>*>(Base::loss.get());
		std::vector<b200::DeviceBuffer<Scalar>> step_losses;
		// A ragged last batch with fewer rows than ranks cannot be sharded without leaving ranks empty-handed, and a rank
		// that skips a step skips the collectives of a synchronised BatchNorm in it: every rank then takes the WHOLE batch,
		// with the loss gradient divided by world size as well (the summed gradients and statistics come out the same) and
		// the losses weighted 1 / world size.  step_loss_weights goes with step_losses.
		std::vector<double> step_loss_weights;
		const std::size_t world = comm.world_size();
		b200::DeviceDataSource<Scalar>* dev_data = dev_net && dev_loss ?
				dynamic_cast<b200::DeviceDataSource<Scalar>*>(&training_prov) : nullptr;
		if (dev_data && !dev_data->device_resident())
			dev_data = nullptr;
		const bool use_graphs = dev_data != nullptr && graph_eligible(net, params_vec);
		while (training_prov.has_more()) {
			if (dev_data) {
				// the data set lives in HBM: the (shard of the) mini-batch is cut out on the device
				b200::DeviceTensor<Scalar> obs, obj;
				if (graph && use_graphs) {
					// a nominal mini-batch is cut straight into the step graph's input buffers
					obs.buf = graph->obs; obs.rows = graph->rows;
					obj.buf = graph->obj; obj.rows = graph->rows;
				}
				const std::size_t batch_rows = dev_data->next_batch_dev(batch_size, comm.rank(), comm.world_size(), obs, obj);
				instances += batch_rows;
				const std::size_t replicas = world > 1 && batch_rows < world ? world : 1;
				if (!obs.empty() && use_graphs && !graphs_failed) {
					// launch-bound steps: after two eager steps at a shape (every scratch buffer has its size, BatchNorm
					// has seen a batch) the step is captured as a CUDA graph and replayed
					// (batch_rows counts the whole mini-batch, obs.rows this rank's shard of it)
					if (graph && graph->rows != obs.rows && batch_rows == batch_size)
						drop_graph();   // the nominal batch changed shape: capture again
					if (!graph && batch_rows == batch_size) {
						if (eager_shape_rows != obs.rows) {
							eager_shape_rows = obs.rows;
							eager_steps_at_shape = 0;
						}
						// (only optimizers whose _update_params goes through fused_step() can be replayed: a user-defined
						// subclass that updates through get_grad / set_values would apply a real host update per call)
						if (eager_steps_at_shape >= 2 && fused_step_seen)
							graph.reset(new StepGraph(obs.rows, obs.size(), obj.size()));
					}
					if (graph && graph->rows == obs.rows &&
							graph_step(*dev_net, *dev_loss, net, params_vec, obs, obj, epoch)) {
						++updates;
						++timestep;
						b200::Context& c = b200::Context::get();
						b200::Context::Lock l = c.lock();
						CATTLE_B200_CHECK(cattl3_ctx_throttle(c.handle(), 2));
						continue;
					}
					++eager_steps_at_shape;
				}
				const std::int64_t allocated_before = cattl3_ctx_allocated_bytes(b200::Context::get().handle());
				if (!obs.empty()) {
					b200::WeightsStable no_update_until_the_end_of_the_passes;
					b200::DeviceTensor<Scalar> out = dev_net->propagate_dev(std::move(obs), true);
					step_losses.emplace_back();
					step_loss_weights.push_back(1.0 / replicas);
					b200::DeviceTensor<Scalar> out_grad = dev_loss->loss_and_gradient_dev(out, obj, (Scalar) (batch_size * replicas),
							step_losses.back());
					exchange_begin(params_vec, comm);
					dev_net->backpropagate_dev(std::move(out_grad));
				}
				finish_step(params_vec, comm, epoch, reg_loss, updates);
				// what one step takes from cattl3_malloc: the size of a step graph's arena
				eager_step_bytes = (std::size_t) (cattl3_ctx_allocated_bytes(b200::Context::get().handle()) - allocated_before);
				continue;
			}
			DataPair<Scalar,Rank,Sequential> data_pair = training_prov.get_data(batch_size);
			const std::size_t batch_rows = data_pair.first.dimension(0);
			instances += batch_rows;
			const std::size_t replicas = world > 1 && batch_rows < world ? world : 1;
			if (world > 1 && replicas == 1)
				data_pair = shard(std::move(data_pair), comm);
			if (data_pair.first.dimension(0) > 0 && dev_net && dev_loss) {
				// the whole step in HBM: one upload of the mini-batch, then propagate -> loss -> back-propagate
				// without a host round trip; the per-sample losses are collected at the end of the epoch
				// (the uploads go through the input feeds: staged on a copy stream, they overlap the previous step)
				b200::WeightsStable no_update_until_the_end_of_the_passes;
				b200::DeviceTensor<Scalar> obj = fed(1, data_pair.second);
				b200::DeviceTensor<Scalar> out = dev_net->propagate_dev(fed(0, data_pair.first), true);
				step_losses.emplace_back();
				step_loss_weights.push_back(1.0 / replicas);
				b200::DeviceTensor<Scalar> out_grad = dev_loss->loss_and_gradient_dev(out, obj, (Scalar) (batch_size * replicas),
						step_losses.back());
				exchange_begin(params_vec, comm);
				dev_net->backpropagate_dev(std::move(out_grad));
			} else if (data_pair.first.dimension(0) > 0) {
				b200::WeightsStable no_update_until_the_end_of_the_passes;
				typename Base::Data out = net.propagate(std::move(data_pair.first), true);
				obj_loss += Base::loss->function(out, data_pair.second).sum() / (double) replicas;
				// dividing by the nominal batch size decouples the learning rate from the batch size and
				// makes the shard gradients add up to the full-batch gradient
				net.backpropagate(Base::loss->d_function(std::move(out), std::move(data_pair.second)) /
						(Scalar) (batch_size * replicas));
			}
			finish_step(params_vec, comm, epoch, reg_loss, updates);
		}
		obj_loss += collect_graph_loss() + carried_graph_loss;
		carried_graph_loss = 0;
		reg_loss += collect_device_penalty();
		for (std::size_t i = 0; i < step_losses.size(); ++i) {
			const b200::DeviceBuffer<Scalar>& losses = step_losses[i];
			std::vector<Scalar> host(losses.size());
			losses.download(host.data(), host.size());
			double sum = 0;
			for (Scalar l : host)
				sum += l;
			obj_loss += sum * step_loss_weights[i];
		}
		if (comm.world_size() > 1)
			obj_loss = comm.all_reduce_sum(obj_loss);
		const Scalar mean_obj_loss = (Scalar) (obj_loss / instances);
		const Scalar mean_reg_loss = (Scalar) (reg_loss / updates);
		if (verbose && comm.rank() == 0) {
			std::cout << std::left << std::setw(20) << "\ttraining obj loss: " << std::right <<
					std::to_string(mean_obj_loss) << std::endl;
			std::cout << std::left << std::setw(20) << "\ttraining reg loss: " << std::right <<
					std::to_string(mean_reg_loss) << std::endl;
		}
		return mean_obj_loss + mean_reg_loss;
	}
	inline Scalar _test(typename Base::Net& net, typename Base::Provider& test_prov, std::size_t epoch,
			bool verbose) {
		assert(target_net_ptr == &net);
		double obj_loss = 0;
		std::size_t instances = 0;
		while (test_prov.has_more()) {
			DataPair<Scalar,Rank,Sequential> data_pair = test_prov.get_data(batch_size);
			instances += data_pair.first.dimension(0);
			obj_loss += Base::loss->function(net.infer(std::move(data_pair.first)),
					std::move(data_pair.second)).sum();
		}
		const Scalar mean_obj_loss = (Scalar) (obj_loss / instances);
		Scalar reg_loss = 0;
		for (Parameters<Scalar>* params_ptr : net.get_all_unique_params()) {
			if (params_ptr->are_optimizable() && !params_ptr->are_frozen())
				reg_loss += params_ptr->get_regularization_penalty();
		}
		if (verbose && b200::Communicator::get().rank() == 0) {
			std::cout << std::left << std::setw(20) << "\ttest obj loss: " << std::right <<
					std::to_string(mean_obj_loss) << std::endl;
			std::cout << std::left << std::setw(20) << "\ttest reg loss: " << std::right <<
					std::to_string(reg_loss) << std::endl;
		}
		return mean_obj_loss + reg_loss;
	}
	/**
	 * It fits the optimizer to the parameters of the network (called by fit()).
	 */
	virtual void _fit(const std::vector<Parameters<Scalar>*>& params_vec) = 0;
	/**
	 * It updates the optimizable, non-frozen parameters using their accumulated gradients.
	 *
	 * @param epoch The index of the epoch, starting from 0.
	 * @param timestep The index of the update since fit(), starting from 0.
	 */
	virtual void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) = 0;
	/**
	 * One fused device kernel per parameter array: the update rule `step.kind` with the scalars in
	 * `step`, followed by the gradient reset.  Optimizer state (up to three arrays per parameter
	 * array, zero-initialised, discarded by fit()) lives in HBM.  Views into one shared array that are
	 * adjacent (BatchNormLayer's per-channel gamma / beta) are updated by a single launch.  Parameters
	 * that are not device resident (reference layers mixed into the network) are staged through the
	 * same kernel, so there is exactly one implementation of every update rule.
	 */
	inline void fused_step(const std::vector<Parameters<Scalar>*>& params_vec, cattl3_opt_step step) {
		step.reset_grad = 1;
		step.l2_lambda = 0;  // Parameters::regularize() has already added the penalty's derivative
		fused_step_seen = true;
		if (step_mode == STEP_SCALARS_ONLY) {
			// a step graph replays the launches; only this step's scalars travel (one small upload from pinned memory)
			graph->publish_scalars(step);
			return;
		}
		struct Run { B200Parameters<Scalar>* first; std::size_t count; std::vector<B200Parameters<Scalar>*> members; };
		std::vector<Run> runs;
		b200::Context& c = b200::Context::get();
		for (Parameters<Scalar>* params_ptr : params_vec) {
			if (!params_ptr->are_optimizable() || params_ptr->are_frozen())
				continue;
			B200Parameters<Scalar>* dev = dynamic_cast<B200Parameters<Scalar>*>(params_ptr);
			if (!dev) {
				staged_step(*params_ptr, step);
				continue;
			}
			bool merged = false;
			for (Run& run : runs) {
				if (run.first->device_values() + run.count == dev->device_values() &&
						run.first->device_grad() + run.count == dev->device_grad() &&
						!dev->has_value_constraints() && !run.first->has_value_constraints()) {
					run.count += dev->count();
					run.members.push_back(dev);
					merged = true;
					break;
				}
			}
			if (!merged)
				runs.push_back(Run{ dev, dev->count(), { dev } });
		}
		for (Run& run : runs) {
			StateArrays& state = device_states[run.first->device_values()];
			for (int s = 0; s < states_needed(step.kind); ++s) {
				if (state[s].size() < run.count) {
					if (!state[s].empty())
						throw b200::Error(CATTL3_ERR_INVALID, "optimizer state does not match the parameters; call fit()");
					state[s] = b200::DeviceBuffer<Scalar>(run.count, true);
				}
			}
			{
				b200::Context::Lock l = c.lock();
				if (step_mode == STEP_CAPTURE) {
					CATTLE_B200_CHECK(b200::Api<Scalar>::optimizer_step_indirect(c.handle(), step.kind, graph->device_scalars(),
							(std::int64_t) run.count, run.first->device_values(), run.first->device_grad(), state[0].data(),
							state[1].data(), state[2].data()));
				} else {
					CATTLE_B200_CHECK(b200::Api<Scalar>::optimizer_step(c.handle(), &step, (std::int64_t) run.count,
							run.first->device_values(), run.first->device_grad(), state[0].data(), state[1].data(),
							state[2].data()));
				}
			}
			for (B200Parameters<Scalar>* member : run.members) {
				member->values_written_on_device();
				member->grad_zeroed_on_device();
			}
		}
	}
	/** Fills the hyper-parameter part of a step descriptor (meaning of a / b per kind: cattl3_b200.h). */
	inline static cattl3_opt_step make_step(int kind, Scalar lr, Scalar a, Scalar b, Scalar eps) {
		cattl3_opt_step step;
		step.kind = kind;
		step.reset_grad = 1;
		step.lr = lr; step.a = a; step.b = b; step.eps = eps;
		step.lr_epoch = lr; step.c1 = 1; step.c1n = 1; step.c2 = 1;
		step.l2_lambda = 0;
		return step;
	}
	const std::size_t batch_size;
private:
	typedef std::array<b200::DeviceBuffer<Scalar>,3> StateArrays;
	/**
	 * The device face of the network being trained: a b200::DeviceNetwork, or for sequential data a
	 * b200::DeviceSequenceNetwork, whose tensors carry samples * time steps rows (b200/DeviceSequenceNetwork.hpp) while
	 * the data provider and the loss see one row per sample -- the same array, a different row count.
	 */
	class DeviceFace {
	public:
		inline DeviceFace(typename Base::Net& net, bool enabled) :
				plain(nullptr),
				seq(nullptr),
				in_volume(net.get_input_dims().get_volume()),
				out_volume(net.get_output_dims().get_volume()),
				samples(0) {
			if (!enabled)
				return;
			if (Sequential)
				seq = dynamic_cast<b200::DeviceSequenceNetwork<Scalar,Rank>*>(&net);
			else
				plain = dynamic_cast<b200::DeviceNetwork<Scalar,Rank>*>(&net);
		}
		inline explicit operator bool() const {
			return plain || seq;
		}
		inline b200::DeviceTensor<Scalar> propagate_dev(b200::DeviceTensor<Scalar> input, bool training) {
			if (plain)
				return plain->propagate_dev(std::move(input), training);
			samples = input.rows;
			input.rows = samples * (input.size() / samples / in_volume);
			b200::DeviceTensor<Scalar> out = seq->propagate_seq_dev(std::move(input), samples, training);
			out.rows = samples;
			return out;
		}
		inline b200::DeviceTensor<Scalar> backpropagate_dev(b200::DeviceTensor<Scalar> out_grad) {
			if (plain)
				return plain->backpropagate_dev(std::move(out_grad));
			out_grad.rows = samples * (out_grad.size() / samples / out_volume);
			return seq->backpropagate_seq_dev(std::move(out_grad), samples);
		}
	private:
		b200::DeviceNetwork<Scalar,Rank>* plain;
		b200::DeviceSequenceNetwork<Scalar,Rank>* seq;
		std::size_t in_volume, out_volume, samples;
	};
	enum StepMode { STEP_EAGER, STEP_SCALARS_ONLY, STEP_CAPTURE };
	/**
	 * A training step captured as a CUDA graph (cattl3_graph, include/cattl3_b200.h): forward, loss, backward and the
	 * fused update of one mini-batch shape, replayed with a single launch.  The mini-batch is copied into fixed input
	 * buffers in front of the graph, the optimizer's step scalars go through a device struct fed from a ring of pinned
	 * slots, the per-sample losses are accumulated on the device.
	 */
	struct StepGraph {
		inline StepGraph(std::size_t rows, std::size_t obs_count, std::size_t obj_count) :
				graph(nullptr),
				rows(rows),
				obs(std::make_shared<b200::DeviceBuffer<Scalar>>(obs_count)),
				obj(std::make_shared<b200::DeviceBuffer<Scalar>>(obj_count)),
				loss_rows(rows),
				loss_accum(rows, true),
				dev_scalars(1),
				pinned(nullptr),
				slot(0) {
			void* p = nullptr;
			CATTLE_B200_CHECK(cattl3_host_alloc(&p, SLOTS * sizeof(cattl3_opt_step)));
			pinned = static_cast<cattl3_opt_step*>(p);
		}
		inline ~StepGraph() {
			if (graph && trace())
				std::cerr << "cattl3: step graph of " << rows << " rows destroyed after " << replays << " replays" << std::endl;
			b200::Context::Lock l = b200::Context::get().lock();
			cattl3_graph_destroy(graph);
			cattl3_host_free(pinned);
		}
		StepGraph(const StepGraph&) = delete;
		StepGraph& operator=(const StepGraph&) = delete;
		inline void publish_scalars(const cattl3_opt_step& step) {
			// the slot is reused four steps later; the loop keeps the host at most two steps ahead of the device
			pinned[slot] = step;
			b200::Context& c = b200::Context::get();
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(cattl3_memcpy_h2d(c.handle(), dev_scalars.data(), &pinned[slot], sizeof(cattl3_opt_step)));
			slot = (slot + 1) % SLOTS;
		}
		inline const cattl3_opt_step* device_scalars() const {
			return dev_scalars.data();
		}
		/** CATTL3_GRAPH_TRACE=1: one line on stderr per capture and per destroyed graph. */
		inline static bool trace() {
			static const bool on = [] {
				const char* v = std::getenv("CATTL3_GRAPH_TRACE");
				return v && v[0] && v[0] != '0';
			}();
			return on;
		}
		static constexpr int SLOTS = 4;
		std::size_t replays = 0;
		cattl3_graph* graph;
		std::size_t rows;
		std::shared_ptr<b200::DeviceBuffer<Scalar>> obs, obj;   // the graph's fixed inputs
		b200::DeviceBuffer<Scalar> loss_rows, loss_accum;
		b200::DeviceBuffer<cattl3_opt_step> dev_scalars;
		cattl3_opt_step* pinned;
		int slot;
	};
	/**
	 * Whether the training step of `net` can be captured: one process, every layer and every parameter on the device,
	 * nothing whose host-side arguments change from step to step (dropout seeds), nothing that reads parameters on
	 * the host in the step (penalties other than L1 / L2 / ElasticNet, value or gradient constraints) and, for sequential networks,
	 * nothing carried across steps on the device (DeviceSequenceNetwork::graph_safe).  CATTL3_NO_GRAPH=1 disables it.
	 */
	inline static bool graph_eligible(typename Base::Net& net, const std::vector<Parameters<Scalar>*>& params_vec) {
		static const bool enabled = [] {
			const char* v = std::getenv("CATTL3_NO_GRAPH");
			return !(v && v[0] && v[0] != '0');
		}();
		static const bool dp_graphs = [] {
			const char* v = std::getenv("CATTL3_NO_DP_GRAPH");
			return !(v && v[0] && v[0] != '0');
		}();
		if (!enabled || (b200::Communicator::get().world_size() > 1 && !dp_graphs))
			return false;
		if (Sequential) {
			const b200::DeviceSequenceNetwork<Scalar,Rank>* seq = dynamic_cast<const b200::DeviceSequenceNetwork<Scalar,Rank>*>(&net);
			if (!seq || !seq->graph_safe())
				return false;
		}
		for (Layer<Scalar,Rank>* layer : net.get_layers()) {
			if (!dynamic_cast<b200::DeviceLayer<Scalar,Rank>*>(layer) || dynamic_cast<DropoutLayer<Scalar,Rank>*>(layer))
				return false;
		}
		for (Parameters<Scalar>* params_ptr : params_vec) {
			B200Parameters<Scalar>* dev = dynamic_cast<B200Parameters<Scalar>*>(params_ptr);
			if (!dev || dev->has_host_regularization())
				return false;
		}
		return true;
	}
	/** The per-sample losses the step graph has accumulated on the device since the last call (synchronises). */
	inline double collect_graph_loss() {
		double sum = 0;
		if (graph) {
			std::vector<Scalar> host(graph->loss_accum.size());
			graph->loss_accum.download(host.data(), host.size());
			for (Scalar l : host)
				sum += l;
			graph->loss_accum.zero();
		}
		return sum;
	}
	/** Destroys the step graph; the losses it holds are carried to the end of the epoch. */
	inline void drop_graph() {
		carried_graph_loss += collect_graph_loss();
		graph.reset();
	}
	/** Steps that allocate more than this are not captured (CATTL3_GRAPH_MAX_MB, default 2048): they are not launch bound. */
	inline static std::size_t max_arena_bytes() {
		static const std::size_t bytes = [] {
			const char* v = std::getenv("CATTL3_GRAPH_MAX_MB");
			const long mb = v && v[0] ? std::atol(v) : 2048;
			return (std::size_t) (mb > 0 ? mb : 0) << 20;
		}();
		return bytes;
	}
	/** One step at `graph`'s shape: captured on first use, replayed afterwards.  False = run this step eagerly. */
	inline bool graph_step(DeviceFace& dev_net, const b200::DeviceLoss<Scalar>& dev_loss,
			typename Base::Net& net, const std::vector<Parameters<Scalar>*>& params_vec, const b200::DeviceTensor<Scalar>& obs,
			const b200::DeviceTensor<Scalar>& obj, std::size_t epoch) {
		b200::Context& c = b200::Context::get();
		StepGraph& g = *graph;
		if (obs.data() != g.obs->data() || obj.data() != g.obj->data()) {
			if (obs.size() != g.obs->size() || obj.size() != g.obj->size())
				return false;
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(cattl3_memcpy_d2d(c.handle(), g.obs->data(), obs.data(), obs.size() * sizeof(Scalar)));
			CATTLE_B200_CHECK(cattl3_memcpy_d2d(c.handle(), g.obj->data(), obj.data(), obj.size() * sizeof(Scalar)));
		}
		step_mode = STEP_SCALARS_ONLY;
		_update_params(params_vec, epoch - 1, timestep);
		step_mode = STEP_EAGER;
		if (!g.graph) {
			// capture: allocations made before the capture must not be released inside it
			net.empty_caches();
			cattl3_graph* captured = nullptr;
			{
				// the arena holds everything an eager step allocated (blocks released inside the step are recycled, so
				// this is an upper bound) plus slack for the 256-byte granules of buffers the eager step did not make
				const std::size_t arena_bytes = eager_step_bytes + eager_step_bytes / 8 + (1u << 20);
				b200::Context::Lock l = c.lock();
				if (arena_bytes > max_arena_bytes() || cattl3_graph_begin(c.handle(), arena_bytes) != CATTL3_OK) {
					graphs_failed = true;   // too large to be launch bound, or no arena left: stay eager
					return false;
				}
			}
			bool ok = true;
			try {
				b200::DeviceTensor<Scalar> obs_view, obj_view;
				obs_view.rows = obj_view.rows = g.rows;
				obs_view.buf = std::make_shared<b200::DeviceBuffer<Scalar>>(
						b200::DeviceBuffer<Scalar>::view(g.obs->data(), g.obs->size()));
				obj_view.buf = std::make_shared<b200::DeviceBuffer<Scalar>>(
						b200::DeviceBuffer<Scalar>::view(g.obj->data(), g.obj->size()));
				std::unique_ptr<b200::WeightsStable> no_update(new b200::WeightsStable());
				b200::DeviceTensor<Scalar> out = dev_net.propagate_dev(std::move(obs_view), true);
				b200::DeviceTensor<Scalar> out_grad = dev_loss.loss_and_gradient_dev(out, obj_view, (Scalar) batch_size,
						g.loss_rows);
				out = b200::DeviceTensor<Scalar>();
				{
					b200::Context::Lock l = c.lock();
					CATTLE_B200_CHECK(b200::Api<Scalar>::add_inplace(c.handle(), (std::int64_t) g.rows, g.loss_accum.data(),
							g.loss_rows.data()));
				}
				dev_net.backpropagate_dev(std::move(out_grad));
				no_update.reset();
				// data parallel: the exchange (and the statistics all-reduces of synchronised BatchNorm layers inside the
				// passes above) are NCCL operations on the capturing stream -- they become nodes of the step graph
				if (b200::Communicator::get().world_size() > 1)
					all_reduce_gradients(params_vec, b200::Communicator::get());
				double no_host_penalty = 0;
				regularize_all(params_vec, no_host_penalty);   // device penalties only (graph_eligible)
				step_mode = STEP_CAPTURE;
				_update_params(params_vec, epoch - 1, timestep);
				step_mode = STEP_EAGER;
			} catch (...) {
				// whatever was thrown (a b200::Error, or anything a user-defined layer / optimizer throws): the stream must
				// leave capture mode below before the step can be retried eagerly
				step_mode = STEP_EAGER;
				ok = false;
			}
			{
				b200::Context::Lock l = c.lock();
				const int rc = cattl3_graph_end(c.handle(), &captured);
				ok = ok && rc == CATTL3_OK && captured != nullptr;
			}
			if (!ok) {
				// nothing was executed: forget every handle created during the capture and fall back for good
				if (captured) {
					b200::Context::Lock l = c.lock();
					cattl3_graph_destroy(captured);
				}
				net.empty_caches();
				graphs_failed = true;
				return false;
			}
			g.graph = captured;
			if (StepGraph::trace())
				std::cerr << "cattl3: step graph captured (" << g.rows << " rows, arena of " <<
						(eager_step_bytes + eager_step_bytes / 8 + (1u << 20)) << " bytes)" << std::endl;
		}
		{
			b200::Context::Lock l = c.lock();
			const int rc = cattl3_graph_launch(c.handle(), g.graph);
			if (rc == CATTL3_ERR_UNSUPPORTED) {
				// library scratch moved since the capture (a larger problem ran on this context in between): start over
				l.unlock();
				drop_graph();
				eager_steps_at_shape = 0;
				return false;
			}
			CATTLE_B200_CHECK(rc);
			++g.replays;
		}
		for (Parameters<Scalar>* params_ptr : params_vec) {
			B200Parameters<Scalar>* dev = static_cast<B200Parameters<Scalar>*>(params_ptr);
			dev->values_written_on_device();
			if (dev->are_optimizable() && !dev->are_frozen())
				dev->grad_zeroed_on_device();
		}
		return true;
	}
	/**
	 * Lays the value storages of all optimizable, non-frozen device parameters out next to each other in one array,
	 * and their gradient storages in another with the same offsets ("one contiguous gradient
	 * arena").  fused_step() and all_reduce_gradients() merge adjacent arrays, so a whole network is then updated
	 * by one or two kernel launches and all-reduced as one or two messages instead of one per Parameters object.
	 * CATTL3_NO_ARENA=1 leaves the storages where they are.
	 */
	inline static void pack_parameters(const std::vector<Parameters<Scalar>*>& params_vec) {
		static const bool enabled = [] {
			const char* v = std::getenv("CATTL3_NO_ARENA");
			return !(v && v[0] && v[0] != '0');
		}();
		if (!enabled)
			return;
		typedef std::shared_ptr<b200::ParameterStorage<Scalar>> StoragePtr;
		std::vector<std::pair<StoragePtr,StoragePtr>> stores;
		std::size_t total = 0;
		for (Parameters<Scalar>* params_ptr : params_vec) {
			B200Parameters<Scalar>* dev = dynamic_cast<B200Parameters<Scalar>*>(params_ptr);
			if (!dev || !dev->are_optimizable() || dev->are_frozen() || dev->has_value_constraints())
				continue;
			bool seen = false;
			for (const std::pair<StoragePtr,StoragePtr>& s : stores)
				seen = seen || s.first == dev->value_storage();
			if (seen)
				continue;
			stores.emplace_back(dev->value_storage(), dev->grad_storage());
			total += dev->value_storage()->size();
		}
		if (stores.size() < 2)
			return;
		auto values = std::make_shared<b200::DeviceBuffer<Scalar>>(total);
		auto grads = std::make_shared<b200::DeviceBuffer<Scalar>>(total);
		std::size_t offset = 0;
		for (const std::pair<StoragePtr,StoragePtr>& s : stores) {
			s.first->relocate(values, offset);
			s.second->relocate(grads, offset);
			offset += s.first->size();
		}
	}
	/** The tail of a training step: all-reduce, regularise, update, reset (SGDOptimizer.hpp:57-70). */
	inline void finish_step(const std::vector<Parameters<Scalar>*>& params_vec, b200::Communicator& comm, std::size_t epoch,
			double& reg_loss, std::size_t& updates) {
		if (comm.world_size() > 1 && !exchange_finish(params_vec, comm))
			all_reduce_gradients(params_vec, comm);
		regularize_all(params_vec, reg_loss);
		_update_params(params_vec, epoch - 1, timestep);
		++updates;
		++timestep;
		for (Parameters<Scalar>* params_ptr : params_vec)
			params_ptr->reset_grad();  // a no-op where the fused step already cleared the gradient
		// the loop is asynchronous: keep the host at most two steps ahead of the device (bounded queues and
		// scratch, ranks of a data-parallel job stay within a step of each other)
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_ctx_throttle(c.handle(), 2));
	}
	/**
	 * Penalty and penalty derivative of every optimizable, non-frozen parameter (SGDOptimizer.hpp:59-64).  L1 / L2 /
	 * ElasticNet penalties of device parameters are one kernel each (cattl3_regularize) that also adds the penalty to
	 * a device accumulator, fetched once per epoch; anything else takes the reference's host route.
	 */
	inline void regularize_all(const std::vector<Parameters<Scalar>*>& params_vec, double& reg_loss) {
		for (Parameters<Scalar>* params_ptr : params_vec) {
			if (!params_ptr->are_optimizable() || params_ptr->are_frozen())
				continue;
			B200Parameters<Scalar>* dev = dynamic_cast<B200Parameters<Scalar>*>(params_ptr);
			if (dev && dev->has_device_regularization()) {
				if (reg_accum.empty())
					reg_accum = b200::DeviceBuffer<double>(1, true);
				dev->regularize_dev(reg_accum.data());
			} else {
				reg_loss += params_ptr->get_regularization_penalty();
				params_ptr->regularize();
			}
		}
	}
	/** The penalties accumulated on the device since the last call (synchronises). */
	inline double collect_device_penalty() {
		if (reg_accum.empty())
			return 0;
		double host = 0;
		reg_accum.download(&host, 1);
		reg_accum.zero();
		return host;
	}
	/**
	 * Host tensor -> device through one of the process's two input feeds (0: observations, 1: objectives; created
	 * on first use and kept: their pinned staging and device ring are not worth re-creating per train() call).
	 */
	inline static b200::DeviceTensor<Scalar> fed(int which, const typename Base::Data& data) {
		b200::DeviceTensor<Scalar> tensor;
		tensor.rows = data.dimension(0);
		tensor.buf = std::make_shared<b200::DeviceBuffer<Scalar>>(
				b200::InputFeed::shared(which).push(data.data(), (std::size_t) data.size()));
		return tensor;
	}
	/** CATTL3_HOST_LOOP=1 keeps the reference's host protocol between network and loss (A/B, debugging). */
	inline static bool device_loop() {
		static const bool on = [] {
			const char* v = std::getenv("CATTL3_HOST_LOOP");
			return !(v && v[0] && v[0] != '0');
		}();
		return on;
	}
	inline static int states_needed(int kind) {
		if (kind == CATTL3_OPT_VANILLA_SGD)
			return 0;
		if (kind == CATTL3_OPT_AMSGRAD)
			return 3;
		return kind <= CATTL3_OPT_RMSPROP ? 1 : 2;
	}
	/** The same kernel for host-resident parameters: upload, step, download. */
	inline void staged_step(Parameters<Scalar>& params, cattl3_opt_step step) {
		const std::size_t count = params.get_rows() * params.get_cols();
		b200::DeviceBuffer<Scalar> values(count), grad(count);
		values.upload(params.get_values().data(), count);
		grad.upload(params.get_grad().data(), count);
		StateArrays& state = host_states[&params];
		for (int s = 0; s < states_needed(step.kind); ++s) {
			if (state[s].size() != count)
				state[s] = b200::DeviceBuffer<Scalar>(count, true);
		}
		step.reset_grad = 0;
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::optimizer_step(c.handle(), &step, (std::int64_t) count,
					values.data(), grad.data(), state[0].data(), state[1].data(), state[2].data()));
		}
		Matrix<Scalar> updated(params.get_rows(), params.get_cols());
		values.download(updated.data(), count);
		params.set_values(std::move(updated));
	}
	/** Rows [r*n/G, (r+1)*n/G) of the mini-batch (the last batch of an epoch may be short). */
	inline static DataPair<Scalar,Rank,Sequential> shard(DataPair<Scalar,Rank,Sequential> pair,
			const b200::Communicator& comm) {
		const std::size_t n = pair.first.dimension(0);
		const std::size_t lo = n * comm.rank() / comm.world_size();
		const std::size_t hi = n * (comm.rank() + 1) / comm.world_size();
		return std::make_pair(rows_of(pair.first, lo, hi), rows_of(pair.second, lo, hi));
	}
	inline static typename Base::Data rows_of(const typename Base::Data& data, std::size_t lo, std::size_t hi) {
		typename Base::Data::Dimensions offsets, extents = data.dimensions();
		for (std::size_t i = 0; i < (std::size_t) data.NumDimensions; ++i)
			offsets[i] = 0;
		offsets[0] = lo;
		extents[0] = hi - lo;
		if (hi == lo)
			return typename Base::Data();
		return data.slice(offsets, extents);
	}
	/**
	 * The device gradient arrays of all optimizable parameters as maximal contiguous stretches (sorted by address, adjacent
	 * and overlapping views merged): the packed arena is ONE stretch, a per-channel BatchNormLayer's 2 C one-element views
	 * over two vectors are two.
	 */
	inline static std::vector<std::pair<Scalar*,Scalar*>> gradient_stretches(const std::vector<Parameters<Scalar>*>& params_vec,
			std::vector<B200Parameters<Scalar>*>* devs = nullptr) {
		std::vector<std::pair<Scalar*,Scalar*>> ranges;
		for (Parameters<Scalar>* params_ptr : params_vec) {
			if (!params_ptr->are_optimizable())
				continue;
			B200Parameters<Scalar>* dev = dynamic_cast<B200Parameters<Scalar>*>(params_ptr);
			if (!dev)
				throw b200::Error(CATTL3_ERR_UNSUPPORTED, "data-parallel training needs device-resident parameters");
			ranges.emplace_back(dev->device_grad(), dev->device_grad() + dev->count());
			if (devs)
				devs->push_back(dev);
		}
		std::sort(ranges.begin(), ranges.end());
		std::vector<std::pair<Scalar*,Scalar*>> merged;
		for (const std::pair<Scalar*,Scalar*>& r : ranges) {
			if (!merged.empty() && r.first <= merged.back().second)
				merged.back().second = std::max(merged.back().second, r.second);
			else
				merged.push_back(r);
		}
		return merged;
	}
	/** Sum over ranks of every optimizable parameter gradient, in place, on the device: one message per stretch. */
	inline static void all_reduce_gradients(const std::vector<Parameters<Scalar>*>& params_vec,
			b200::Communicator& comm) {
		std::vector<B200Parameters<Scalar>*> reduced;
		const std::vector<std::pair<Scalar*,Scalar*>> stretches = gradient_stretches(params_vec, &reduced);
		comm.group_start();
		for (const std::pair<Scalar*,Scalar*>& r : stretches)
			comm.all_reduce_sum(r.first, (std::size_t) (r.second - r.first));
		comm.group_end();
		for (B200Parameters<Scalar>* dev : reduced)
			dev->grad_written_on_device();
	}
	/**
	 * The exchange overlapped with the backward pass.  Layers are back-propagated last to first
	 * (FeedforwardNeuralNetwork.hpp:120-127) and the arena holds their gradients first to last, so the finished part of
	 * the arena is a suffix that grows towards the front: whenever it has grown by a bucket, that stretch goes on its way
	 * on the communicator's side stream while the layers in front of it still compute.  A parameter written more than once
	 * per step (shared by the cells of an LSTM) would be sent before it is complete, so the first data-parallel step only
	 * COUNTS the writes and the overlap is used if every parameter is written exactly once; otherwise, and for unpacked or
	 * constrained parameters, the exchange stays behind the backward pass (all_reduce_gradients).
	 */
	struct GradientExchange : public B200Parameters<Scalar>::GradientListener {
		enum Mode { IDLE, COUNTING, SENDING };
		Mode mode = IDLE;
		bool decided = false, usable = false, violated = false;
		b200::Communicator* comm = nullptr;
		Scalar* lo = nullptr; Scalar* hi = nullptr;      // the arena
		Scalar* sent_from = nullptr;                      // [sent_from, hi) is on its way
		std::size_t bucket = 0;
		std::vector<std::pair<Scalar*,Scalar*>> done;     // written, not yet sent
		std::map<Scalar*,int> writes;
		inline void gradient_written(Scalar* p, std::size_t n) override {
			if (mode == COUNTING) {
				++writes[p];
			} else if (mode == SENDING) {
				if (p < lo || p + n > hi)
					return;
				if (p + n > sent_from) {
					violated = true;   // a second write into a stretch that has left: the warm-up count said this cannot happen
					return;
				}
				done.emplace_back(p, p + n);
				Scalar* frontier = sent_from;
				for (bool grew = true; grew;) {
					grew = false;
					for (std::size_t i = 0; i < done.size(); ++i) {
						if (done[i].second == frontier) {
							frontier = done[i].first;
							done[i] = done.back();
							done.pop_back();
							grew = true;
							break;
						}
					}
				}
				// a stretch smaller than a bucket waits at the frontier: keep it in `done` as one range
				if ((std::size_t) (sent_from - frontier) >= bucket) {
					comm->all_reduce_sum_async(frontier, (std::size_t) (sent_from - frontier));
					sent_from = frontier;
				} else if (frontier != sent_from) {
					done.emplace_back(frontier, sent_from);
				}
			}
		}
		inline void begin(Mode m) {
			mode = m;
			done.clear();
			sent_from = hi;
			violated = false;
			B200Parameters<Scalar>::gradient_listener() = this;
		}
		inline void end() {
			B200Parameters<Scalar>::gradient_listener() = nullptr;
		}
	};
	GradientExchange exchange;
	/** Before a data-parallel backward pass: count the writes (first step), or send finished stretches as they appear. */
	inline void exchange_begin(const std::vector<Parameters<Scalar>*>& params_vec, b200::Communicator& comm) {
		static const bool enabled = [] {
			const char* v = std::getenv("CATTL3_NO_OVERLAP");
			return !(v && v[0] && v[0] != '0');
		}();
		if (!enabled || comm.world_size() < 2)
			return;
		if (!exchange.decided) {
			exchange.writes.clear();
			exchange.begin(GradientExchange::COUNTING);
		} else if (exchange.usable) {
			exchange.begin(GradientExchange::SENDING);
		}
	}
	/** After the backward pass: the rest of the gradients; true if the exchange has been taken care of here. */
	inline bool exchange_finish(const std::vector<Parameters<Scalar>*>& params_vec, b200::Communicator& comm) {
		if (exchange.mode == GradientExchange::IDLE)
			return false;
		exchange.end();
		const typename GradientExchange::Mode mode = exchange.mode;
		exchange.mode = GradientExchange::IDLE;
		if (mode == GradientExchange::COUNTING) {
			// decide once: one packed stretch, every parameter written exactly once, no constraints evaluated on the host
			std::vector<B200Parameters<Scalar>*> devs;
			const std::vector<std::pair<Scalar*,Scalar*>> stretches = gradient_stretches(params_vec, &devs);
			bool ok = stretches.size() == 1;
			for (B200Parameters<Scalar>* dev : devs) {
				typename std::map<Scalar*,int>::const_iterator it = exchange.writes.find(dev->device_grad());
				ok = ok && it != exchange.writes.end() && it->second == 1 && !dev->has_grad_constraints();
			}
			exchange.decided = true;
			exchange.usable = ok;
			if (ok) {
				exchange.comm = &comm;
				exchange.lo = stretches[0].first;
				exchange.hi = stretches[0].second;
				const char* v = std::getenv("CATTL3_BUCKET_KB");
				const long kb = v && v[0] ? std::atol(v) : 1024;
				exchange.bucket = (std::size_t) (kb > 0 ? kb : 1) * 1024 / sizeof(Scalar);
			}
			return false;
		}
		if (exchange.violated)
			throw b200::Error(CATTL3_ERR_UNSUPPORTED, "data-parallel exchange: a gradient was written again after its bucket "
					"had been sent (set CATTL3_NO_OVERLAP=1)");
		if (exchange.sent_from > exchange.lo)
			comm.all_reduce_sum_async(exchange.lo, (std::size_t) (exchange.sent_from - exchange.lo));
		comm.wait();
		return true;
	}
	std::size_t timestep;
	const typename Base::Net* target_net_ptr;
	// the captured training step (device-resident loop only) and how the optimizer's update is being issued
	std::unique_ptr<StepGraph> graph;
	std::size_t eager_steps_at_shape = 0, eager_shape_rows = 0, eager_step_bytes = 0;
	bool graphs_failed = false;
	bool fused_step_seen = false;
	double carried_graph_loss = 0;
	b200::DeviceBuffer<double> reg_accum;   // penalties of the device-regularised parameters, summed over the epoch's steps
	StepMode step_mode = STEP_EAGER;
	// optimizer state per device parameter array (keyed by the device address of its first value) and
	// per host-resident Parameters object
	std::map<const Scalar*,StateArrays> device_states;
	std::map<const Parameters<Scalar>*,StateArrays> host_states;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_SGDOPTIMIZER_H_ */
