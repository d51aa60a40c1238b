/*
 * neural_network/ParallelNeuralNetwork.hpp -- B200 replacement of the reference's parallel (Inception-style) container
 * (C-ATTL3/neural_network/ParallelNeuralNetwork.hpp:27-345), same class template, constructors and interface; defines
 * the reference header's include guard.
 *
 * Lanes with the same input are run and their outputs merged by concatenation along the lowest or highest rank, by
 * summation or by multiplication (:142-197); the backward pass hands every lane its share of the output gradient and
 * sums the lanes' input gradients (:199-252, 277-300).  The reference runs the lanes on pthreads over host tensors.
 * Here the parallel resource is the GPU: the lanes are enqueued one after the other on the context's stream, every
 * tensor stays in HBM, and the merges are device operations -- concatenation = strided device copies (a block per
 * index of the ranks above the concatenation rank), sum / product = one element-wise kernel per lane.  Lanes that
 * are not device networks are bridged through the host.
 */
#ifndef C_ATTL3_NEURAL_NETWORK_PARALLELNEURALNETWORK_H_
#define C_ATTL3_NEURAL_NETWORK_PARALLELNEURALNETWORK_H_

#include <array>
#include <cassert>
#include <utility>
#include <vector>

#include "neural_network/CompositeNeuralNetwork.hpp"
#include "b200/DeviceNetwork.hpp"

namespace cattle {

/**
 * The ways the outputs of the lanes may be merged.
 */
enum ParallelOutputMergeType { PARALLEL_CONCAT_LO_RANK, PARALLEL_CONCAT_HI_RANK, PARALLEL_SUM, PARALLEL_MUL };

template<typename Scalar, std::size_t Rank, ParallelOutputMergeType MergeType = PARALLEL_CONCAT_HI_RANK>
class ParallelNeuralNetwork :
		public CompositeNeuralNetwork<Scalar,Rank,false,NeuralNetwork<Scalar,Rank,false>>,
		public b200::DeviceNetwork<Scalar,Rank> {
	typedef NeuralNetwork<Scalar,Rank,false> Base;
	typedef ParallelNeuralNetwork<Scalar,Rank,MergeType> Self;
	typedef NeuralNetPtr<Scalar,Rank,false> Lane;
	typedef b200::DeviceNetwork<Scalar,Rank> DevNet;
	typedef b200::DeviceTensor<Scalar> DevTensor;
	static_assert(MergeType >= PARALLEL_CONCAT_LO_RANK && MergeType <= PARALLEL_MUL, "illegal merge type value");
	static constexpr bool CONCAT = MergeType == PARALLEL_CONCAT_HI_RANK || MergeType == PARALLEL_CONCAT_LO_RANK;
	static constexpr std::size_t CONCAT_RANK = MergeType == PARALLEL_CONCAT_HI_RANK ? Rank - 1 : 0;
public:
	/**
	 * @param lanes The lanes; all take the same input dimensions.
	 * @param foremost Whether the network is the first module of a composite (then no lane produces an input gradient).
	 */
	inline ParallelNeuralNetwork(std::vector<Lane>&& lanes, bool foremost = true) :
			lanes(std::move(lanes)),
			foremost(foremost),
			outputs(this->lanes.size()) {
		assert(this->lanes.size() > 0 && "lanes must contain at least 1 element");
		input_dims = this->lanes.front()->get_input_dims();
		output_dims = this->lanes.front()->get_output_dims();
		for (std::size_t i = 1; i < this->lanes.size(); ++i) {
			assert(this->lanes[i] != nullptr && "lanes contains null pointers");
			assert(input_dims == this->lanes[i]->get_input_dims());
			const typename Base::Dims& lane_output_dims = this->lanes[i]->get_output_dims();
			if (CONCAT) {
				for (std::size_t r = 0; r < Rank; ++r)
					assert(r == +CONCAT_RANK || output_dims(r) == lane_output_dims(r));
				output_dims(+CONCAT_RANK) += lane_output_dims(+CONCAT_RANK);
			} else {
				assert(output_dims == lane_output_dims);
			}
		}
		set_foremost(foremost);
	}
	inline ParallelNeuralNetwork(Lane&& lane, bool foremost = true) :
			ParallelNeuralNetwork(single(std::move(lane)), foremost) { }
	inline ParallelNeuralNetwork(const Self& network) :
			foremost(network.foremost),
			input_dims(network.input_dims),
			output_dims(network.output_dims),
			outputs(network.outputs) {
		for (const Lane& lane : network.lanes)
			lanes.push_back(Lane(lane->clone()));
	}
	inline ParallelNeuralNetwork(Self&& network) {
		swap(*this, network);
	}
	~ParallelNeuralNetwork() = default;
	inline Self& operator=(Self network) {
		swap(*this, network);
		return *this;
	}
	inline Base* clone() const {
		return new ParallelNeuralNetwork(*this);
	}
	inline const typename Base::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return output_dims;
	}
	inline std::vector<const Layer<Scalar,Rank>*> get_layers() const {
		std::vector<const Layer<Scalar,Rank>*> layer_ptrs;
		for (const Lane& lane : lanes) {
			for (Layer<Scalar,Rank>* layer : lane->get_layers())
				layer_ptrs.push_back(layer);
		}
		return layer_ptrs;
	}
	inline std::vector<Layer<Scalar,Rank>*> get_layers() {
		std::vector<Layer<Scalar,Rank>*> layer_ptrs;
		for (const Lane& lane : lanes) {
			for (Layer<Scalar,Rank>* layer : lane->get_layers())
				layer_ptrs.push_back(layer);
		}
		return layer_ptrs;
	}
	inline std::vector<Base*> get_modules() {
		std::vector<Base*> modules;
		for (const Lane& lane : lanes)
			modules.push_back(lane.get());
		return modules;
	}
	inline bool is_foremost() const {
		return foremost;
	}
	/** Every lane sees the network input, so every lane is foremost when the network is (:128-132). */
	inline void set_foremost(bool foremost) {
		for (const Lane& lane : lanes)
			lane->set_foremost(foremost);
		this->foremost = foremost;
	}
	inline void empty_caches() {
		for (std::size_t i = 0; i < lanes.size(); ++i) {
			lanes[i]->empty_caches();
			outputs[i] = DevTensor();
		}
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
		const std::size_t rows = input.rows;
		DevTensor out;
		std::size_t offset = 0;   // elements of a joined row filled so far (concatenation)
		if (CONCAT)
			out = DevTensor(rows, output_dims.get_volume());
		for (std::size_t i = 0; i < lanes.size(); ++i) {
			DevTensor lane_out = run_forward(*lanes[i], input, training);
			if (CONCAT) {
				const std::size_t width = block_width(rows, lanes[i]->get_output_dims());
				copy_blocks(out.data() + offset, joined_pitch(rows), lane_out.data(), width, width, outer_blocks());
				offset += width;
			} else if (MergeType == PARALLEL_SUM) {
				if (i == 0) {
					out = std::move(lane_out);
					out.make_exclusive();
				} else {
					combine(out, lane_out, false);
				}
			} else {
				// the product: every lane's output is kept for the backward pass (:170-173)
				outputs[i] = lane_out;
				if (i == 0) {
					out = DevTensor(rows, output_dims.get_volume());
					copy_blocks(out.data(), out.size(), lane_out.data(), out.size(), out.size(), 1);
				} else {
					combine(out, lane_out, true);
				}
			}
		}
		return out;
	}
	inline DevTensor backpropagate_dev(DevTensor out_grad) {
		const std::size_t rows = out_grad.rows;
		DevTensor prev_out_grad;
		std::size_t offset = 0;
		for (std::size_t i = 0; i < lanes.size(); ++i) {
			DevTensor lane_grad;
			if (CONCAT) {
				const std::size_t width = block_width(rows, lanes[i]->get_output_dims());
				lane_grad = DevTensor(rows, lanes[i]->get_output_dims().get_volume());
				copy_blocks(lane_grad.data(), width, out_grad.data() + offset, joined_pitch(rows), width, outer_blocks());
				offset += width;
			} else if (MergeType == PARALLEL_SUM) {
				lane_grad = out_grad;   // shared, read-only
			} else {
				lane_grad = DevTensor(rows, output_dims.get_volume());
				copy_blocks(lane_grad.data(), lane_grad.size(), out_grad.data(), lane_grad.size(), lane_grad.size(), 1);
				for (std::size_t j = 0; j < lanes.size(); ++j) {
					if (j == i)
						continue;
					if (outputs[j].empty() || outputs[j].rows != rows)
						throw b200::Error(CATTL3_ERR_INVALID, "ParallelNeuralNetwork: backpropagate without a matching propagate");
					combine(lane_grad, outputs[j], true);
				}
			}
			DevTensor lane_prev = run_backward(*lanes[i], std::move(lane_grad));
			if (foremost)
				continue;
			if (prev_out_grad.empty()) {
				prev_out_grad = std::move(lane_prev);
				prev_out_grad.make_exclusive();
			} else {
				combine(prev_out_grad, lane_prev, false);
			}
		}
		return prev_out_grad;
	}
	inline friend void swap(Self& network1, Self& network2) {
		using std::swap;
		swap(network1.lanes, network2.lanes);
		swap(network1.foremost, network2.foremost);
		swap(network1.input_dims, network2.input_dims);
		swap(network1.output_dims, network2.output_dims);
		swap(network1.outputs, network2.outputs);
	}
private:
	inline static std::vector<Lane> single(Lane&& lane) {
		std::vector<Lane> vec;
		vec.push_back(std::move(lane));
		return vec;
	}
	/** A joined tensor = outer_blocks() rows of joined_pitch() elements; a lane contributes block_width() of each. */
	inline std::size_t outer_blocks() const {
		std::size_t outer = 1;
		for (std::size_t r = +CONCAT_RANK + 1; r < Rank; ++r)
			outer *= output_dims(r);
		return outer;
	}
	inline std::size_t block_width(std::size_t rows, const typename Base::Dims& lane_dims) const {
		std::size_t inner = rows;
		for (std::size_t r = 0; r < +CONCAT_RANK; ++r)
			inner *= lane_dims(r);
		return inner * lane_dims(+CONCAT_RANK);
	}
	inline std::size_t joined_pitch(std::size_t rows) const {
		return block_width(rows, output_dims);
	}
	inline static void copy_blocks(Scalar* dst, std::size_t dst_pitch, const Scalar* src, std::size_t src_pitch,
			std::size_t width, std::size_t outer) {
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(cattl3_memcpy_2d(c.handle(), dst, dst_pitch * sizeof(Scalar), src, src_pitch * sizeof(Scalar),
				width * sizeof(Scalar), outer));
	}
	/** target += operand, or target *= operand. */
	inline static void combine(DevTensor& target, const DevTensor& operand, bool multiply) {
		if (target.size() != operand.size())
			throw b200::Error(CATTL3_ERR_INVALID, "ParallelNeuralNetwork: lanes disagree on the tensor size");
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		if (multiply) {
			CATTLE_B200_CHECK(b200::Api<Scalar>::mul_inplace(c.handle(), (std::int64_t) target.size(), target.data(), operand.data()));
		} else {
			CATTLE_B200_CHECK(b200::Api<Scalar>::add_inplace(c.handle(), (std::int64_t) target.size(), target.data(), operand.data()));
		}
	}
	inline DevTensor run_forward(Base& lane, const DevTensor& input, bool training) const {
		if (DevNet* dev_lane = dynamic_cast<DevNet*>(&lane))
			return dev_lane->propagate_dev(input, training);
		typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(input,
				b200::batch_extents<Rank>(input.rows, input_dims));
		return b200::to_device<Scalar,Base::DATA_RANK>(lane.propagate(std::move(host), training));
	}
	inline DevTensor run_backward(Base& lane, DevTensor out_grad) const {
		if (DevNet* dev_lane = dynamic_cast<DevNet*>(&lane))
			return dev_lane->backpropagate_dev(std::move(out_grad));
		typename Base::Data host = b200::to_host<Scalar,Base::DATA_RANK>(out_grad,
				b200::batch_extents<Rank>(out_grad.rows, lane.get_output_dims()));
		return b200::to_device<Scalar,Base::DATA_RANK>(lane.backpropagate(std::move(host)));
	}
	std::vector<Lane> lanes;
	bool foremost;
	typename Base::Dims input_dims, output_dims;
	// the lanes' outputs of the last forward pass (PARALLEL_MUL only)
	std::vector<DevTensor> outputs;
};

} /* namespace cattle */

#endif /* C_ATTL3_NEURAL_NETWORK_PARALLELNEURALNETWORK_H_ */
