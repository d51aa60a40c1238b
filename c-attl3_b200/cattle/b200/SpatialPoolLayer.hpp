/*
 * b200/SpatialPoolLayer.hpp -- the shared core of MaxPoolLayer and MeanPoolLayer for ranks 1-3.
 *
 * Semantics are the reference's PoolLayer (C-ATTL3/layer/PoolLayer.hpp:77-116,145-148): no padding,
 * output extent (in - receptor) / stride + 1, overlapping windows allowed.  Max pooling keeps, per
 * output element, the position of the first maximum in the reference's scan order (width outer,
 * height inner, strict '>' -- C-ATTL3/layer/pool/MaxPoolLayer.hpp:45-58) as one byte in HBM, and the
 * backward pass routes (and, for overlapping windows, sums) the gradients through it; mean pooling
 * needs no cache (C-ATTL3/layer/pool/MeanPoolLayer.hpp:35-41).  Kernels: cattl3_pool_forward /
 * cattl3_pool_backward (include/cattl3_b200.h), one HBM-bound pass each.
 */
#ifndef C_ATTL3_B200_SPATIALPOOLLAYER_H_
#define C_ATTL3_B200_SPATIALPOOLLAYER_H_

#include <cassert>
#include <utility>
#include <vector>

#include "core/Layer.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar, std::size_t Rank, int Kind>
class SpatialPoolLayer : public Layer<Scalar,Rank>, public DeviceLayer<Scalar,Rank> {
	typedef Layer<Scalar,Rank> Base;
public:
	inline Base* clone_with_shared_params() {
		return this->clone();
	}
	inline const Base& get_params_owner() const {
		return *this;
	}
	inline const typename Base::Dims& get_input_dims() const {
		return input_dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return output_dims;
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
		arg_cache = DeviceBuffer<std::uint8_t>();
		cached_rows = 0;
	}
	inline typename Base::Data pass_forward(typename Base::Data in, bool training) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(in.dimensions()).template demote<>()) == input_dims);
		assert(in.dimension(0) > 0);
		DeviceTensor<Scalar> out = pass_forward_dev(to_device<Scalar,Base::DATA_RANK>(in), training);
		return to_host<Scalar,Base::DATA_RANK>(out, batch_extents<Rank>(out.rows, output_dims));
	}
	inline typename Base::Data pass_back(typename Base::Data out_grad) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<>()) == output_dims);
		assert(out_grad.dimension(0) > 0 && (std::size_t) out_grad.dimension(0) == cached_rows);
		if (input_layer)
			return typename Base::Data();
		DeviceTensor<Scalar> prev_out_grad = pass_back_dev(to_device<Scalar,Base::DATA_RANK>(out_grad));
		return to_host<Scalar,Base::DATA_RANK>(prev_out_grad, batch_extents<Rank>(prev_out_grad.rows, input_dims));
	}
	inline DeviceTensor<Scalar> pass_forward_dev(DeviceTensor<Scalar> in, bool training) {
		cattl3_pool_geom g = geometry(in.rows);
		DeviceTensor<Scalar> out(in.rows, output_dims.get_volume());
		if (Kind == CATTL3_POOL_MAX && arg_cache.size() != out.size())
			arg_cache = DeviceBuffer<std::uint8_t>(out.size());
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(Api<Scalar>::pool_forward(c.handle(), Kind, &g, in.data(), out.data(), arg_cache.data()));
		cached_rows = in.rows;
		return out;
	}
	inline DeviceTensor<Scalar> pass_back_dev(DeviceTensor<Scalar> out_grad) {
		if (cached_rows != out_grad.rows)
			throw Error(CATTL3_ERR_INVALID, "pool layer: pass_back without a matching pass_forward");
		if (input_layer)
			return DeviceTensor<Scalar>();
		cattl3_pool_geom g = geometry(out_grad.rows);
		DeviceTensor<Scalar> prev_out_grad(out_grad.rows, input_dims.get_volume());
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(Api<Scalar>::pool_backward(c.handle(), Kind, &g, out_grad.data(), arg_cache.data(),
				prev_out_grad.data()));
		return prev_out_grad;
	}
protected:
	inline SpatialPoolLayer(const typename Base::Dims& input_dims, std::size_t receptor_height,
			std::size_t receptor_width, std::size_t vertical_stride, std::size_t horizontal_stride) :
				input_dims(input_dims),
				output_dims(pooled_dims(input_dims, receptor_height, receptor_width, vertical_stride,
						horizontal_stride)),
				receptor_height(receptor_height),
				receptor_width(receptor_width),
				vertical_stride(vertical_stride),
				horizontal_stride(horizontal_stride),
				input_layer(false),
				cached_rows(0) {
		assert(receptor_height > 0 && receptor_width > 0);
		assert(vertical_stride > 0 && horizontal_stride > 0);
		cattl3_pool_geom g = geometry(1);
		CATTLE_B200_CHECK(cattl3_pool_output_dims(&g, nullptr, nullptr));  // validates the geometry
	}
	const typename Base::Dims input_dims, output_dims;
	const std::size_t receptor_height, receptor_width, vertical_stride, horizontal_stride;
private:
	inline static std::size_t width_of(const typename Base::Dims& dims) {
		return Rank >= 2 ? dims(Rank >= 2 ? 1 : 0) : 1;
	}
	inline static std::size_t channels_of(const typename Base::Dims& dims) {
		return Rank == 3 ? dims(Rank == 3 ? 2 : 0) : 1;
	}
	inline static typename Base::Dims pooled_dims(const typename Base::Dims& in, std::size_t rh, std::size_t rw,
			std::size_t vs, std::size_t hs) {
		typename Base::Dims out = in;
		out(0) = (in(0) - rh) / vs + 1;
		if (Rank >= 2)
			out(Rank >= 2 ? 1 : 0) = (width_of(in) - rw) / hs + 1;
		return out;
	}
	inline cattl3_pool_geom geometry(std::size_t rows) const {
		cattl3_pool_geom g;
		g.n = (std::int32_t) rows;
		g.h = (std::int32_t) input_dims(0);
		g.w = (std::int32_t) width_of(input_dims);
		g.c = (std::int32_t) channels_of(input_dims);
		g.rh = (std::int32_t) receptor_height; g.rw = (std::int32_t) receptor_width;
		g.sh = (std::int32_t) vertical_stride; g.sw = (std::int32_t) horizontal_stride;
		return g;
	}
	bool input_layer;
	std::size_t cached_rows;
	// Max pooling only: position of the selected element inside each window.
	DeviceBuffer<std::uint8_t> arg_cache;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_SPATIALPOOLLAYER_H_ */
