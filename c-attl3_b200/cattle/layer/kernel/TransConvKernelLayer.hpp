/*
 * layer/kernel/TransConvKernelLayer.hpp -- B200 replacement of the reference's TransConvKernelLayer
 * (C-ATTL3/layer/kernel/TransConvKernelLayer.hpp): same class template, same constructor arguments in the
 * same order with the same defaults, so code written against the reference compiles unchanged.  This
 * header defines the reference header's include guard; put this directory before the reference's on
 * the include path and the original becomes a no-op.
 *
 * y(n, ih*sh + rh*(dh+1) - ph, iw*sw + rw*(dw+1) - pw, f) += x(n,ih,iw,c) * W(c, rh + RH*(rw + RW*f)), one bias per output element: cattl3_transconv_forward / cattl3_transconv_backward.
 */
#ifndef C_ATTL3_LAYER_KERNEL_TRANSCONVKERNELLAYER_H_
#define C_ATTL3_LAYER_KERNEL_TRANSCONVKERNELLAYER_H_

#include "b200/SpatialKernelLayer.hpp"

namespace cattle {

/**
 * Rank 3: observations {height, width, channels}; rank 2: {height, width} with one implicit channel.
 * The outputs of the filters are concatenated along the last rank of the observation.
 */
template<typename Scalar, std::size_t Rank = 3>
class TransConvKernelLayer : public b200::SpatialKernelLayer<Scalar,Rank,true> {
	static_assert(Rank == 2 || Rank == 3, "rank 1 has its own specialisation");
	typedef Layer<Scalar,Rank> Root;
	typedef b200::SpatialKernelLayer<Scalar,Rank,true> Core;
public:
	inline TransConvKernelLayer(const typename Root::Dims& input_dims, std::size_t filters,
			ParamInitSharedPtr<Scalar> weight_init, std::size_t receptor_height = 3, std::size_t receptor_width = 3,
			std::size_t vertical_padding = 1, std::size_t horizontal_padding = 1, std::size_t vertical_stride = 1,
			std::size_t horizontal_stride = 1, std::size_t vertical_dilation = 0, std::size_t horizontal_dilation = 0,
			ParamRegSharedPtr<Scalar> weight_reg = nullptr, Scalar weight_clip = 0, Scalar weight_max_l1_norm = 0,
			Scalar weight_max_l2_norm = 0, Scalar weight_grad_clip = 0, Scalar weight_grad_max_l1_norm = 0,
			Scalar weight_grad_max_l2_norm = 0, ParamRegSharedPtr<Scalar> bias_reg = nullptr, Scalar bias_clip = 0,
			Scalar bias_max_l1_norm = 0, Scalar bias_max_l2_norm = 0, Scalar bias_grad_clip = 0,
			Scalar bias_grad_max_l1_norm = 0, Scalar bias_grad_max_l2_norm = 0) :
				Core(input_dims, filters, receptor_height, receptor_width, vertical_padding, horizontal_padding,
						vertical_stride, horizontal_stride, vertical_dilation, horizontal_dilation, weight_init,
						weight_reg, weight_clip, weight_max_l1_norm, weight_max_l2_norm, weight_grad_clip,
						weight_grad_max_l1_norm, weight_grad_max_l2_norm, bias_reg, bias_clip, bias_max_l1_norm,
						bias_max_l2_norm, bias_grad_clip, bias_grad_max_l1_norm, bias_grad_max_l2_norm) { }
	inline TransConvKernelLayer(const TransConvKernelLayer<Scalar,Rank>& layer, bool share_params = false) :
			Core(layer, share_params) { }
	inline Root* clone() const {
		return new TransConvKernelLayer(*this);
	}
	inline Root* clone_with_shared_params() {
		return new TransConvKernelLayer(*this, true);
	}
};

/**
 * Rank 1: observations {length}; a 1-D receptor, padding, stride and dilation.
 */
template<typename Scalar>
class TransConvKernelLayer<Scalar,1> : public b200::SpatialKernelLayer<Scalar,1,true> {
	typedef Layer<Scalar,1> Root;
	typedef b200::SpatialKernelLayer<Scalar,1,true> Core;
public:
	inline TransConvKernelLayer(const typename Root::Dims& input_dims, std::size_t filters,
			ParamInitSharedPtr<Scalar> weight_init, std::size_t receptor_length = 3, std::size_t padding = 1,
			std::size_t stride = 1, std::size_t dilation = 0, ParamRegSharedPtr<Scalar> weight_reg = nullptr,
			Scalar weight_clip = 0, Scalar weight_max_l1_norm = 0, Scalar weight_max_l2_norm = 0,
			Scalar weight_grad_clip = 0, Scalar weight_grad_max_l1_norm = 0, Scalar weight_grad_max_l2_norm = 0,
			ParamRegSharedPtr<Scalar> bias_reg = nullptr, Scalar bias_clip = 0, Scalar bias_max_l1_norm = 0,
			Scalar bias_max_l2_norm = 0, Scalar bias_grad_clip = 0, Scalar bias_grad_max_l1_norm = 0,
			Scalar bias_grad_max_l2_norm = 0) :
				Core(input_dims, filters, receptor_length, 1, padding, 0, stride, 1, dilation, 0, weight_init,
						weight_reg, weight_clip, weight_max_l1_norm, weight_max_l2_norm, weight_grad_clip,
						weight_grad_max_l1_norm, weight_grad_max_l2_norm, bias_reg, bias_clip, bias_max_l1_norm,
						bias_max_l2_norm, bias_grad_clip, bias_grad_max_l1_norm, bias_grad_max_l2_norm) { }
	inline TransConvKernelLayer(const TransConvKernelLayer<Scalar,1>& layer, bool share_params = false) :
			Core(layer, share_params) { }
	inline Root* clone() const {
		return new TransConvKernelLayer(*this);
	}
	inline Root* clone_with_shared_params() {
		return new TransConvKernelLayer(*this, true);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_KERNEL_TRANSCONVKERNELLAYER_H_ */
