/*
 * layer/ReshapeLayer.hpp -- B200 replacement of the reference's reshape layer
 * (C-ATTL3/layer/ReshapeLayer.hpp:22-98), same class template and constructor; defines the reference
 * header's include guard.
 *
 * "The data backing the tensor is not shifted in any way" (:20): with the batch index fastest and plain
 * column-major order behind it, a reshape of the observation dimensions is the same array, so on the device
 * the layer hands its input buffer on untouched -- a free view, no kernel, no copy -- and only the nominal
 * dimensions change.  Its purpose here is to keep a network device resident across it
 * (examples/mnist_autoencoder.cpp:40: Dense -> Reshape -> TransConv).
 */
#ifndef C_ATTL3_LAYER_RESHAPELAYER_H_
#define C_ATTL3_LAYER_RESHAPELAYER_H_

#include <array>
#include <cassert>
#include <utility>
#include <vector>

#include "core/Layer.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class ReshapeLayer : public Layer<Scalar,Rank>, public b200::DeviceLayer<Scalar,Rank> {
	typedef Layer<Scalar,Rank> Base;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	/**
	 * @param input_dims The nominal input dimensions.
	 * @param output_dims The nominal output dimensions; the volumes must agree.
	 */
	inline ReshapeLayer(const typename Base::Dims& input_dims, const typename Base::Dims& output_dims) :
			input_dims(input_dims),
			output_dims(output_dims),
			input_layer(false),
			rows(0) {
		assert(input_dims.get_volume() == output_dims.get_volume());
	}
	inline Base* clone() const {
		return new ReshapeLayer(*this);
	}
	inline Base* clone_with_shared_params() {
		return clone();
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
	inline void empty_cache() { }
	inline typename Base::Data pass_forward(typename Base::Data in, bool training) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(in.dimensions()).template demote<>()) == input_dims);
		assert(in.dimension(0) > 0);
		rows = in.dimension(0);
		return in.reshape(b200::batch_extents<Rank>(rows, output_dims));
	}
	inline typename Base::Data pass_back(typename Base::Data out_grad) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<>()) == output_dims);
		assert(out_grad.dimension(0) > 0 && rows == (std::size_t) out_grad.dimension(0));
		if (input_layer)
			return typename Base::Data();
		return out_grad.reshape(b200::batch_extents<Rank>(rows, input_dims));
	}
	inline DevTensor pass_forward_dev(DevTensor in, bool training) {
		rows = in.rows;
		return in;
	}
	inline DevTensor pass_back_dev(DevTensor out_grad) {
		if (input_layer)
			return DevTensor();
		return out_grad;
	}
private:
	const typename Base::Dims input_dims, output_dims;
	bool input_layer;
	std::size_t rows;
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_RESHAPELAYER_H_ */
