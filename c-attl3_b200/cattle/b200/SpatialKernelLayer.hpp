/*
 * b200/SpatialKernelLayer.hpp -- the shared core of ConvKernelLayer and TransConvKernelLayer.
 *
 * One class covers ranks 1-3 and both directions: a rank-2 observation {H, W} is the rank-3
 * observation {H, W, 1}, a rank-1 observation {L} is {L, 1, 1}, and because the sample index is the
 * fastest dimension and the filter index the slowest (C-ATTL3/core/EigenProxy.hpp:56-57), the
 * (N, OH, OW, F) result *is* the (N, OH, OW * F) / (N, OH * F) tensor the lower-rank layers return
 * (C-ATTL3/layer/kernel/ConvKernelLayer.hpp:400-418, 494-509): no data movement is involved.
 *
 * The arithmetic is libcattl3_b200's implicit GEMM (no im2col / col2im buffers, unlike
 * ConvKernelLayer.hpp:127,159 and TransConvKernelLayer.hpp:121-141); this header only keeps the
 * geometry, the parameters and the cached input.
 */
#ifndef C_ATTL3_B200_SPATIALKERNELLAYER_H_
#define C_ATTL3_B200_SPATIALKERNELLAYER_H_

#include <cassert>
#include <memory>
#include <utility>

#include "layer/KernelLayer.hpp"
#include "parameter_initialization/ZeroParameterInitialization.hpp"
#include "parameters/B200Parameters.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar, std::size_t Rank, bool Transposed>
class SpatialKernelLayer : public KernelLayer<Scalar,Rank>, public DeviceLayer<Scalar,Rank>,
		public SplitBackwardLayer<Scalar,Rank>,
		public EpilogueProducer<Scalar> {
	typedef Layer<Scalar,Rank> Root;
	typedef KernelLayer<Scalar,Rank> Base;
	typedef Api<Scalar> Device;
public:
	inline void empty_cache() {
		in_cache = DeviceTensor<Scalar>();
	}
	inline typename Root::Data pass_forward(typename Root::Data in, bool training) {
		assert((Dimensions<std::size_t,Root::DATA_RANK>(in.dimensions()).template demote<>()) == Base::input_dims);
		assert(in.dimension(0) > 0);
		DeviceTensor<Scalar> out = pass_forward_dev(to_device<Scalar,Root::DATA_RANK>(in), training);
		return to_host<Scalar,Root::DATA_RANK>(out, batch_extents<Rank>(out.rows, Base::output_dims));
	}
	inline typename Root::Data pass_back(typename Root::Data out_grad) {
		assert((Dimensions<std::size_t,Root::DATA_RANK>(out_grad.dimensions()).template demote<>()) == Base::output_dims);
		assert(out_grad.dimension(0) > 0 && (std::size_t) out_grad.dimension(0) == in_cache.rows);
		DeviceTensor<Scalar> prev_out_grad = pass_back_dev(to_device<Scalar,Root::DATA_RANK>(out_grad));
		if (prev_out_grad.empty())
			return typename Root::Data();
		return to_host<Scalar,Root::DATA_RANK>(prev_out_grad, batch_extents<Rank>(prev_out_grad.rows, Base::input_dims));
	}
	inline DeviceTensor<Scalar> pass_forward_dev(DeviceTensor<Scalar> in, bool training) {
		cattl3_conv_geom g = geometry(in.rows);
		DeviceTensor<Scalar> out(in.rows, Base::output_dims.get_volume());
		B200Parameters<Scalar>& w = device_params(*Base::weights);
		B200Parameters<Scalar>& b = device_params(*Base::bias);
		Context& c = Context::get();
		{
			Context::Lock l = c.lock();
			if (Transposed) {
				CATTLE_B200_CHECK(Device::transconv_forward(c.handle(), &g, in.data(), w.device_values(),
						b.device_values(), out.data()));
			} else {
				CATTLE_B200_CHECK(Device::conv_forward(c.handle(), &g, in.data(), w.device_values(),
						b.device_values(), out.data()));
			}
		}
		in_cache = std::move(in);
		return out;
	}
	/** Both directions fuse a following activation; only the convolution (per-filter bias) produces column statistics. */
	inline bool can_fuse_epilogue() const {
		return true;
	}
	inline std::size_t stat_columns() const {
		return Transposed ? 0 : filters;
	}
	inline DeviceTensor<Scalar> pass_forward_dev_fused(DeviceTensor<Scalar> in, bool training, FusedEpilogue<Scalar>& ep) {
		if (Transposed && ep.want_stats)
			throw Error(CATTL3_ERR_UNSUPPORTED, "TransConvKernelLayer produces no column statistics");
		cattl3_conv_geom g = geometry(in.rows);
		const std::size_t volume = Base::output_dims.get_volume();
		const bool act = ep.act_kind != CATTL3_ACT_NONE;
		DeviceTensor<Scalar> out;
		if (!act || ep.keep_pre || ep.want_stats)
			out = DeviceTensor<Scalar>(in.rows, volume);
		B200Parameters<Scalar>& w = device_params(*Base::weights);
		B200Parameters<Scalar>& b = device_params(*Base::bias);
		cattl3_epilogue e;
		e.act_kind = ep.act_kind; e.reserved = 0; e.act_param = (double) ep.act_param;
		e.act_out = nullptr; e.col_stats = nullptr;
		if (act) {
			ep.act_out = DeviceTensor<Scalar>(in.rows, volume);
			e.act_out = ep.act_out.data();
		}
		if (ep.want_stats) {
			ep.col_stats = std::make_shared<DeviceBuffer<double>>(2 * filters + 1);  // + the element count (synchronised BN)
			e.col_stats = ep.col_stats->data();
			ep.shift = b.device_values();
		}
		Context& c = Context::get();
		{
			Context::Lock l = c.lock();
			if (Transposed) {
				CATTLE_B200_CHECK(Device::transconv_forward_fused(c.handle(), &g, in.data(), w.device_values(),
						b.device_values(), out.data(), &e));
			} else {
				CATTLE_B200_CHECK(Device::conv_forward_fused(c.handle(), &g, in.data(), w.device_values(),
						b.device_values(), out.data(), &e));
			}
		}
		in_cache = std::move(in);
		return out;
	}
	inline DeviceTensor<Scalar> pass_back_dev(DeviceTensor<Scalar> out_grad) {
		if (in_cache.empty() || in_cache.rows != out_grad.rows)
			throw Error(CATTL3_ERR_INVALID, "kernel layer: pass_back without a matching pass_forward");
		cattl3_conv_geom g = geometry(out_grad.rows);
		DeviceTensor<Scalar> prev_out_grad;
		if (!Base::is_input_layer())
			prev_out_grad = DeviceTensor<Scalar>(out_grad.rows, Base::input_dims.get_volume());
		B200Parameters<Scalar>& w = device_params(*Base::weights);
		B200Parameters<Scalar>& b = device_params(*Base::bias);
		Context& c = Context::get();
		{
			Context::Lock l = c.lock();
			if (Transposed) {
				CATTLE_B200_CHECK(Device::transconv_backward(c.handle(), &g, in_cache.data(), w.device_values(),
						out_grad.data(), w.device_grad(), b.device_grad(), prev_out_grad.data()));
			} else {
				CATTLE_B200_CHECK(Device::conv_backward(c.handle(), &g, in_cache.data(), w.device_values(),
						out_grad.data(), w.device_grad(), b.device_grad(), prev_out_grad.data()));
			}
		}
		w.grad_written_on_device();
		b.grad_written_on_device();
		return prev_out_grad;
	}
	/** b200::SplitBackwardLayer (convolutions; the transposed layer keeps its single backward call). */
	inline bool can_split_backward() const {
		return !Transposed;
	}
	inline DeviceTensor<Scalar> pass_back_input_dev(const DeviceTensor<Scalar>& out_grad) {
		if (Transposed)
			throw Error(CATTL3_ERR_UNSUPPORTED, "kernel layer: the transposed convolution's backward pass is not split");
		DeviceTensor<Scalar> prev_out_grad;
		if (Base::is_input_layer())
			return prev_out_grad;
		cattl3_conv_geom g = geometry(out_grad.rows);
		prev_out_grad = DeviceTensor<Scalar>(out_grad.rows, Base::input_dims.get_volume());
		B200Parameters<Scalar>& w = device_params(*Base::weights);
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(Device::conv_backward(c.handle(), &g, nullptr, w.device_values(), out_grad.data(), nullptr,
				nullptr, prev_out_grad.data()));
		return prev_out_grad;
	}
	inline void accumulate_param_grads_dev(const DeviceTensor<Scalar>& in, const DeviceTensor<Scalar>& out_grad) {
		if (Transposed || in.empty() || in.rows != out_grad.rows)
			throw Error(CATTL3_ERR_INVALID, "kernel layer: parameter gradients need matching input and gradient batches");
		cattl3_conv_geom g = geometry(out_grad.rows);
		B200Parameters<Scalar>& w = device_params(*Base::weights);
		B200Parameters<Scalar>& b = device_params(*Base::bias);
		Context& c = Context::get();
		{
			Context::Lock l = c.lock();
			CATTLE_B200_CHECK(Device::conv_backward(c.handle(), &g, in.data(), w.device_values(), out_grad.data(),
					w.device_grad(), b.device_grad(), nullptr));
		}
		w.grad_written_on_device();
		b.grad_written_on_device();
	}
protected:
	inline SpatialKernelLayer(const typename Root::Dims& input_dims, std::size_t filters, std::size_t receptor_height,
			std::size_t receptor_width, std::size_t vertical_padding, std::size_t horizontal_padding,
			std::size_t vertical_stride, std::size_t horizontal_stride, std::size_t vertical_dilation,
			std::size_t horizontal_dilation, ParamInitSharedPtr<Scalar> weight_init, ParamRegSharedPtr<Scalar> weight_reg,
			Scalar weight_clip, Scalar weight_max_l1_norm, Scalar weight_max_l2_norm, Scalar weight_grad_clip,
			Scalar weight_grad_max_l1_norm, Scalar weight_grad_max_l2_norm, ParamRegSharedPtr<Scalar> bias_reg,
			Scalar bias_clip, Scalar bias_max_l1_norm, Scalar bias_max_l2_norm, Scalar bias_grad_clip,
			Scalar bias_grad_max_l1_norm, Scalar bias_grad_max_l2_norm) :
				Base(input_dims, nominal_output_dims(input_dims, filters, receptor_height, receptor_width,
						vertical_padding, horizontal_padding, vertical_stride, horizontal_stride, vertical_dilation,
						horizontal_dilation),
						std::make_shared<B200Parameters<Scalar>>(
								Transposed ? channels_of(input_dims) : receptor_height * receptor_width * channels_of(input_dims),
								Transposed ? receptor_height * receptor_width * filters : filters, true, weight_init,
								weight_reg, weight_clip, weight_max_l1_norm, weight_max_l2_norm, weight_grad_clip,
								weight_grad_max_l1_norm, weight_grad_max_l2_norm),
						std::make_shared<B200Parameters<Scalar>>(1,
								Transposed ? nominal_output_dims(input_dims, filters, receptor_height, receptor_width,
										vertical_padding, horizontal_padding, vertical_stride, horizontal_stride,
										vertical_dilation, horizontal_dilation).get_volume() : filters,
								true, std::make_shared<ZeroParameterInitialization<Scalar>>(), bias_reg, bias_clip,
								bias_max_l1_norm, bias_max_l2_norm, bias_grad_clip, bias_grad_max_l1_norm,
								bias_grad_max_l2_norm)),
				filters(filters),
				receptor_height(receptor_height),
				receptor_width(receptor_width),
				vertical_padding(vertical_padding),
				horizontal_padding(horizontal_padding),
				vertical_stride(vertical_stride),
				horizontal_stride(horizontal_stride),
				vertical_dilation(vertical_dilation),
				horizontal_dilation(horizontal_dilation) {
		assert(filters > 0 && receptor_height > 0 && receptor_width > 0);
		assert(vertical_stride > 0 && horizontal_stride > 0);
		cattl3_conv_geom g = geometry(1);
		std::int32_t oh, ow;
		CATTLE_B200_CHECK(cattl3_conv_output_dims(&g, Transposed ? 1 : 0, &oh, &ow));  // validates the geometry
	}
	inline SpatialKernelLayer(const SpatialKernelLayer<Scalar,Rank,Transposed>& layer, bool share_params = false) :
			Base(layer, share_params),
			filters(layer.filters),
			receptor_height(layer.receptor_height),
			receptor_width(layer.receptor_width),
			vertical_padding(layer.vertical_padding),
			horizontal_padding(layer.horizontal_padding),
			vertical_stride(layer.vertical_stride),
			horizontal_stride(layer.horizontal_stride),
			vertical_dilation(layer.vertical_dilation),
			horizontal_dilation(layer.horizontal_dilation),
			in_cache(layer.in_cache) { }
	// The defining attributes of the layer, under the reference's names.
	const std::size_t filters, receptor_height, receptor_width, vertical_padding, horizontal_padding,
			vertical_stride, horizontal_stride, vertical_dilation, horizontal_dilation;
private:
	inline static std::size_t height_of(const typename Root::Dims& dims) {
		return dims(0);
	}
	inline static std::size_t width_of(const typename Root::Dims& dims) {
		return Rank >= 2 ? dims(Rank >= 2 ? 1 : 0) : 1;
	}
	inline static std::size_t channels_of(const typename Root::Dims& dims) {
		return Rank == 3 ? dims(Rank == 3 ? 2 : 0) : 1;
	}
	/** ConvKernelLayer.hpp:194-197 / TransConvKernelLayer.hpp:200-203. */
	inline static std::size_t spatial_output_dim(std::size_t in, std::size_t receptor, std::size_t padding,
			std::size_t dilation, std::size_t stride) {
		if (Transposed)
			return (in - 1) * stride + receptor + (receptor - 1) * dilation - 2 * padding;
		return (in + 2 * padding - receptor - (receptor - 1) * dilation) / stride + 1;
	}
	/** The filter maps are appended along the last nominal rank (ConvKernelLayer.hpp:206-217). */
	inline static typename Root::Dims nominal_output_dims(const typename Root::Dims& input_dims, std::size_t filters,
			std::size_t rh, std::size_t rw, std::size_t vp, std::size_t hp, std::size_t vs, std::size_t hs,
			std::size_t vd, std::size_t hd) {
		const std::size_t oh = spatial_output_dim(height_of(input_dims), rh, vp, vd, vs);
		const std::size_t ow = spatial_output_dim(width_of(input_dims), rw, hp, hd, hs);
		typename Root::Dims out;
		if (Rank == 3) {
			out(0) = oh; out(Rank >= 2 ? 1 : 0) = ow; out(Rank == 3 ? 2 : 0) = filters;
		} else if (Rank == 2) {
			out(0) = oh; out(Rank >= 2 ? 1 : 0) = ow * filters;
		} else {
			out(0) = oh * ow * filters;
		}
		return out;
	}
	inline cattl3_conv_geom geometry(std::size_t rows) const {
		cattl3_conv_geom g;
		g.n = (std::int32_t) rows;
		g.h = (std::int32_t) height_of(Base::input_dims);
		g.w = (std::int32_t) width_of(Base::input_dims);
		g.c = (std::int32_t) channels_of(Base::input_dims);
		g.f = (std::int32_t) filters;
		g.rh = (std::int32_t) receptor_height; g.rw = (std::int32_t) receptor_width;
		g.ph = (std::int32_t) vertical_padding; g.pw = (std::int32_t) horizontal_padding;
		g.sh = (std::int32_t) vertical_stride; g.sw = (std::int32_t) horizontal_stride;
		g.dh = (std::int32_t) vertical_dilation; g.dw = (std::int32_t) horizontal_dilation;
		return g;
	}
	inline static B200Parameters<Scalar>& device_params(Parameters<Scalar>& params) {
		return static_cast<B200Parameters<Scalar>&>(params);  // constructed above, never replaced
	}
	// The input of the last forward pass, kept in HBM for the backward pass.
	DeviceTensor<Scalar> in_cache;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_SPATIALKERNELLAYER_H_ */
