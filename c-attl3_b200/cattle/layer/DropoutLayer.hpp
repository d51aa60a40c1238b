/*
 * layer/DropoutLayer.hpp -- B200 replacement of the reference's drop-out layer
 * (C-ATTL3/layer/DropoutLayer.hpp:25-108), same class template and constructor; defines the reference
 * header's include guard.
 *
 * Inverted dropout as in the reference (:74-84): in training mode every element is zeroed with probability
 * dropout_prob (a uniform draw u <= dropout_prob) and the survivors are scaled by
 * 1 / (1 - dropout_prob + epsilon); inference is the identity; pass_back multiplies by the same mask (:85-93).
 * The draws come from a counter-based generator on the device keyed by (seed, forward-pass counter, element),
 * so the layer costs one HBM-bound kernel per direction and a network containing it stays device resident.
 * The reference's own masks come from Eigen's host RNG and are not reproducible;
 * equivalence is therefore statistical, plus exact consistency between the forward mask and pass_back.
 * set_seed() makes a run repeatable.
 */
#ifndef C_ATTL3_LAYER_DROPOUTLAYER_H_
#define C_ATTL3_LAYER_DROPOUTLAYER_H_

#include <cassert>
#include <cstdint>
#include <utility>
#include <vector>

#include "core/Layer.hpp"
#include "core/NumericUtils.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class DropoutLayer : public Layer<Scalar,Rank>, public b200::DeviceLayer<Scalar,Rank> {
	typedef Layer<Scalar,Rank> Base;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * @param dims The dimensionality of the input tensor.
	 * @param dropout_prob The probability of an element being set to 0.
	 * @param epsilon A small constant for numerical stability.
	 */
	inline DropoutLayer(const typename Base::Dims& dims, Scalar dropout_prob,
			Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
				dims(dims),
				dropout_prob(dropout_prob),
				epsilon(epsilon),
				input_layer(false),
				seed(next_default_seed()),
				passes(0),
				mask_rows(0) {
		assert(dropout_prob > 0 && dropout_prob <= 1 &&
				"dropout probability must be greater than 0 and no greater than 1");
		assert(epsilon > 0 && "epsilon must be greater than 0");
	}
	inline Base* clone() const {
		return new DropoutLayer(*this);
	}
	inline Base* clone_with_shared_params() {
		return clone();
	}
	inline const Base& get_params_owner() const {
		return *this;
	}
	inline const typename Base::Dims& get_input_dims() const {
		return dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return dims;
	}
	inline bool is_input_layer() const {
		return input_layer;
	}
	inline void set_input_layer(bool input_layer) {
		this->input_layer = input_layer;
	}
	inline std::vector<const Parameters<Scalar>*> get_params() const {
		return std::vector<const Parameters<Scalar>*>();
	}
	inline std::vector<Parameters<Scalar>*> get_params() {
		return std::vector<Parameters<Scalar>*>();
	}
	inline void empty_cache() {
		mask = b200::DeviceBuffer<std::uint8_t>();
		mask_rows = 0;
	}
	/** Restarts the layer's stream of masks from `seed`. */
	inline void set_seed(std::uint64_t seed) {
		this->seed = seed;
		passes = 0;
	}
	inline typename Base::Data pass_forward(typename Base::Data in, bool training) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(in.dimensions()).template demote<>()) == dims);
		assert(in.dimension(0) > 0);
		if (!training)
			return in;
		DevTensor out = pass_forward_dev(b200::to_device<Scalar,Base::DATA_RANK>(in), training);
		return b200::to_host<Scalar,Base::DATA_RANK>(out, b200::batch_extents<Rank>(out.rows, dims));
	}
	inline typename Base::Data pass_back(typename Base::Data out_grad) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<>()) == dims);
		assert(out_grad.dimension(0) > 0 && mask_rows == (std::size_t) out_grad.dimension(0));
		if (input_layer)
			return typename Base::Data();
		DevTensor prev_out_grad = pass_back_dev(b200::to_device<Scalar,Base::DATA_RANK>(out_grad));
		return b200::to_host<Scalar,Base::DATA_RANK>(prev_out_grad, b200::batch_extents<Rank>(prev_out_grad.rows, dims));
	}
	inline DevTensor pass_forward_dev(DevTensor in, bool training) {
		if (!training)
			return in;
		DevTensor out(in.rows, dims.get_volume());
		if (mask.size() != in.size())
			mask = b200::DeviceBuffer<std::uint8_t>(in.size());
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::dropout_forward(c.handle(), (std::int64_t) in.size(), dropout_prob,
					epsilon, seed + 0x632BE59BD9B4E019ull * ++passes, in.data(), out.data(), mask.data()));
		}
		mask_rows = in.rows;
		return out;
	}
	inline DevTensor pass_back_dev(DevTensor out_grad) {
		if (mask.empty() || mask_rows != out_grad.rows)
			throw b200::Error(CATTL3_ERR_INVALID, "DropoutLayer: pass_back without a matching training pass_forward");
		if (input_layer)
			return DevTensor();
		DevTensor prev_out_grad(out_grad.rows, dims.get_volume());
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(b200::Api<Scalar>::dropout_backward(c.handle(), (std::int64_t) out_grad.size(), dropout_prob,
				epsilon, out_grad.data(), mask.data(), prev_out_grad.data()));
		return prev_out_grad;
	}
private:
	/** Distinct layers get distinct default streams; the sequence of defaults is fixed per process. */
	inline static std::uint64_t next_default_seed() {
		static std::uint64_t counter = 0;
		return 0xD1B54A32D192ED03ull * ++counter;
	}
	const typename Base::Dims dims;
	const Scalar dropout_prob, epsilon;
	bool input_layer;
	std::uint64_t seed, passes;
	// Staged computation cache: which elements survived the last training pass.
	b200::DeviceBuffer<std::uint8_t> mask;
	std::size_t mask_rows;
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_DROPOUTLAYER_H_ */
