/*
 * layer/kernel/DenseKernelLayer.hpp -- B200 replacement of the reference's fully connected layer
 * (C-ATTL3/layer/kernel/DenseKernelLayer.hpp:61-115), same class template and constructor.
 * An (N, d1, d2, d3) batch is the N x (d1*d2*d3) column-major matrix without any data movement,
 * because the sample index is the fastest dimension (DenseKernelLayer.hpp:98).
 *
 *   Y = X W + 1 b,   dW += X^T dY,   db += colsum(dY),   dX = dY W^T
 *
 * through cattl3_dense_forward / cattl3_dense_backward (include/cattl3_b200.h).
 */
#ifndef C_ATTL3_LAYER_KERNEL_DENSEKERNELLAYER_H_
#define C_ATTL3_LAYER_KERNEL_DENSEKERNELLAYER_H_

#include <cassert>
#include <memory>
#include <utility>

#include "layer/KernelLayer.hpp"
#include "parameter_initialization/ZeroParameterInitialization.hpp"
#include "parameters/B200Parameters.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank = 1>
class DenseKernelLayer : public KernelLayer<Scalar,Rank>, public b200::DeviceLayer<Scalar,Rank>,
		public b200::SplitBackwardLayer<Scalar,Rank>, public b200::EpilogueProducer<Scalar> {
	typedef Layer<Scalar,Rank> Root;
	typedef KernelLayer<Scalar,Rank> Base;
	typedef b200::DeviceTensor<Scalar> DevTensor;
public:
	inline DenseKernelLayer(const typename Root::Dims& input_dims, std::size_t output_size,
			ParamInitSharedPtr<Scalar> weight_init, ParamRegSharedPtr<Scalar> weight_reg = nullptr,
			Scalar weight_clip = 0, Scalar weight_max_l1_norm = 0, Scalar weight_max_l2_norm = 0,
			Scalar weight_grad_clip = 0, Scalar weight_grad_max_l1_norm = 0, Scalar weight_grad_max_l2_norm = 0,
			ParamRegSharedPtr<Scalar> bias_reg = nullptr, Scalar bias_clip = 0, Scalar bias_max_l1_norm = 0,
			Scalar bias_max_l2_norm = 0, Scalar bias_grad_clip = 0, Scalar bias_grad_max_l1_norm = 0,
			Scalar bias_grad_max_l2_norm = 0) :
				Base(input_dims, { output_size },
						std::make_shared<B200Parameters<Scalar>>(input_dims.get_volume(), output_size, true,
								weight_init, weight_reg, weight_clip, weight_max_l1_norm, weight_max_l2_norm,
								weight_grad_clip, weight_grad_max_l1_norm, weight_grad_max_l2_norm),
						std::make_shared<B200Parameters<Scalar>>(1, output_size, true,
								std::make_shared<ZeroParameterInitialization<Scalar>>(), bias_reg, bias_clip,
								bias_max_l1_norm, bias_max_l2_norm, bias_grad_clip, bias_grad_max_l1_norm,
								bias_grad_max_l2_norm)) {
		assert(output_size > 0);
	}
	inline DenseKernelLayer(const DenseKernelLayer<Scalar,Rank>& layer, bool share_params = false) :
			Base(layer, share_params),
			in_cache(layer.in_cache) { }
	inline Root* clone() const {
		return new DenseKernelLayer(*this);
	}
	inline Root* clone_with_shared_params() {
		return new DenseKernelLayer(*this, true);
	}
	inline void empty_cache() {
		in_cache = DevTensor();
	}
	inline typename Root::Data pass_forward(typename Root::Data in, bool training) {
		assert((Dimensions<std::size_t,Root::DATA_RANK>(in.dimensions()).template demote<>()) == Base::input_dims);
		assert(in.dimension(0) > 0);
		DevTensor out = pass_forward_dev(b200::to_device<Scalar,Root::DATA_RANK>(in), training);
		return b200::to_host<Scalar,Root::DATA_RANK>(out, b200::batch_extents<Rank>(out.rows, Base::output_dims));
	}
	inline typename Root::Data pass_back(typename Root::Data out_grad) {
		assert((Dimensions<std::size_t,Root::DATA_RANK>(out_grad.dimensions()).template demote<>()) == Base::output_dims);
		assert(out_grad.dimension(0) > 0 && (std::size_t) out_grad.dimension(0) == in_cache.rows);
		DevTensor prev_out_grad = pass_back_dev(b200::to_device<Scalar,Root::DATA_RANK>(out_grad));
		if (prev_out_grad.empty())
			return typename Root::Data();
		return b200::to_host<Scalar,Root::DATA_RANK>(prev_out_grad,
				b200::batch_extents<Rank>(prev_out_grad.rows, Base::input_dims));
	}
	inline DevTensor pass_forward_dev(DevTensor in, bool training) {
		DevTensor out(in.rows, Base::output_dims.get_volume());
		B200Parameters<Scalar>& w = static_cast<B200Parameters<Scalar>&>(*Base::weights);
		B200Parameters<Scalar>& b = static_cast<B200Parameters<Scalar>&>(*Base::bias);
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::dense_forward(c.handle(), (std::int32_t) in.rows,
					(std::int32_t) Base::input_dims.get_volume(), (std::int32_t) Base::output_dims.get_volume(),
					in.data(), w.device_values(), b.device_values(), out.data()));
		}
		in_cache = std::move(in);
		return out;
	}
	inline bool can_fuse_epilogue() const {
		return true;
	}
	inline std::size_t stat_columns() const {
		return Base::output_dims.get_volume();
	}
	inline DevTensor pass_forward_dev_fused(DevTensor in, bool training, b200::FusedEpilogue<Scalar>& ep) {
		const std::size_t volume = Base::output_dims.get_volume();
		const bool act = ep.act_kind != CATTL3_ACT_NONE;
		DevTensor out;
		if (!act || ep.keep_pre || ep.want_stats)
			out = DevTensor(in.rows, volume);
		B200Parameters<Scalar>& w = static_cast<B200Parameters<Scalar>&>(*Base::weights);
		B200Parameters<Scalar>& b = static_cast<B200Parameters<Scalar>&>(*Base::bias);
		cattl3_epilogue e;
		e.act_kind = ep.act_kind; e.reserved = 0; e.act_param = (double) ep.act_param;
		e.act_out = nullptr; e.col_stats = nullptr;
		if (act) {
			ep.act_out = DevTensor(in.rows, volume);
			e.act_out = ep.act_out.data();
		}
		if (ep.want_stats) {
			ep.col_stats = std::make_shared<b200::DeviceBuffer<double>>(2 * volume + 1);  // + the element count (synchronised BN)
			e.col_stats = ep.col_stats->data();
			ep.shift = b.device_values();
		}
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::dense_forward_fused(c.handle(), (std::int32_t) in.rows,
					(std::int32_t) Base::input_dims.get_volume(), (std::int32_t) volume, in.data(), w.device_values(),
					b.device_values(), out.data(), &e));
		}
		in_cache = std::move(in);
		return out;
	}
	/** b200::SplitBackwardLayer: dX = dY W^T alone, and dW += X^T dY, db += sum dY for any batch. */
	inline bool can_split_backward() const {
		return true;
	}
	inline DevTensor pass_back_input_dev(const DevTensor& out_grad) {
		DevTensor prev_out_grad;
		if (Base::is_input_layer())
			return prev_out_grad;
		prev_out_grad = DevTensor(out_grad.rows, Base::input_dims.get_volume());
		B200Parameters<Scalar>& w = static_cast<B200Parameters<Scalar>&>(*Base::weights);
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(b200::Api<Scalar>::dense_backward(c.handle(), (std::int32_t) out_grad.rows,
				(std::int32_t) Base::input_dims.get_volume(), (std::int32_t) Base::output_dims.get_volume(), nullptr,
				w.device_values(), out_grad.data(), nullptr, nullptr, prev_out_grad.data()));
		return prev_out_grad;
	}
	inline void accumulate_param_grads_dev(const DevTensor& in, const DevTensor& out_grad) {
		if (in.empty() || in.rows != out_grad.rows)
			throw b200::Error(CATTL3_ERR_INVALID, "DenseKernelLayer: parameter gradients need matching input and gradient batches");
		B200Parameters<Scalar>& w = static_cast<B200Parameters<Scalar>&>(*Base::weights);
		B200Parameters<Scalar>& b = static_cast<B200Parameters<Scalar>&>(*Base::bias);
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::dense_backward(c.handle(), (std::int32_t) out_grad.rows,
					(std::int32_t) Base::input_dims.get_volume(), (std::int32_t) Base::output_dims.get_volume(), in.data(),
					w.device_values(), out_grad.data(), w.device_grad(), b.device_grad(), nullptr));
		}
		w.grad_written_on_device();
		b.grad_written_on_device();
	}
	inline DevTensor pass_back_dev(DevTensor out_grad) {
		if (in_cache.empty() || in_cache.rows != out_grad.rows)
			throw b200::Error(CATTL3_ERR_INVALID, "DenseKernelLayer: pass_back without a matching pass_forward");
		DevTensor prev_out_grad;
		if (!Base::is_input_layer())
			prev_out_grad = DevTensor(out_grad.rows, Base::input_dims.get_volume());
		B200Parameters<Scalar>& w = static_cast<B200Parameters<Scalar>&>(*Base::weights);
		B200Parameters<Scalar>& b = static_cast<B200Parameters<Scalar>&>(*Base::bias);
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::dense_backward(c.handle(), (std::int32_t) out_grad.rows,
					(std::int32_t) Base::input_dims.get_volume(), (std::int32_t) Base::output_dims.get_volume(),
					in_cache.data(), w.device_values(), out_grad.data(), w.device_grad(), b.device_grad(),
					prev_out_grad.data()));
		}
		w.grad_written_on_device();
		b.grad_written_on_device();
		return prev_out_grad;
	}
private:
	// The input of the last forward pass, kept in HBM for the backward pass.
	DevTensor in_cache;
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_KERNEL_DENSEKERNELLAYER_H_ */
