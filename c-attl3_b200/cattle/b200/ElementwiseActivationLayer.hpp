/*
 * b200/ElementwiseActivationLayer.hpp -- the shared core of the B200 activation layers: one
 * vectorised, HBM-bound kernel per direction (cattl3_activation_forward / _backward,
 * include/cattl3_b200.h), with the input and/or output of the forward pass kept in HBM exactly where
 * the reference keeps a cache (e.g. C-ATTL3/layer/activation/ReLUActivationLayer.hpp:45-57 keeps the
 * input, SigmoidActivationLayer the output, ELUActivationLayer.hpp:55-78 both).
 * The API contract is the reference's ActivationLayer (C-ATTL3/layer/ActivationLayer.hpp:27-79).
 */
#ifndef C_ATTL3_B200_ELEMENTWISEACTIVATIONLAYER_H_
#define C_ATTL3_B200_ELEMENTWISEACTIVATIONLAYER_H_

#include <cassert>
#include <utility>

#include "layer/ActivationLayer.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar, std::size_t Rank, int Kind>
class ElementwiseActivationLayer : public ActivationLayer<Scalar,Rank>, public DeviceLayer<Scalar,Rank>,
		public EpilogueConsumer<Scalar> {
	typedef Layer<Scalar,Rank> Root;
	typedef ActivationLayer<Scalar,Rank> Base;
	static constexpr bool KEEPS_INPUT = Kind == CATTL3_ACT_RELU || Kind == CATTL3_ACT_LEAKY_RELU ||
			Kind == CATTL3_ACT_ELU || Kind == CATTL3_ACT_SWISH || Kind == CATTL3_ACT_SOFTPLUS;
	static constexpr bool KEEPS_OUTPUT = Kind == CATTL3_ACT_ELU || Kind == CATTL3_ACT_SIGMOID ||
			Kind == CATTL3_ACT_TANH || Kind == CATTL3_ACT_SOFTMAX;
public:
	inline void empty_cache() {
		in_cache = DeviceTensor<Scalar>();
		out_cache = DeviceTensor<Scalar>();
	}
	inline typename Root::Data pass_forward(typename Root::Data in, bool training) {
		assert((Dimensions<std::size_t,Root::DATA_RANK>(in.dimensions()).template demote<>()) == Base::dims);
		assert(in.dimension(0) > 0);
		DeviceTensor<Scalar> out = pass_forward_dev(to_device<Scalar,Root::DATA_RANK>(in), training);
		return to_host<Scalar,Root::DATA_RANK>(out, batch_extents<Rank>(out.rows, Base::dims));
	}
	inline typename Root::Data pass_back(typename Root::Data out_grad) {
		assert((Dimensions<std::size_t,Root::DATA_RANK>(out_grad.dimensions()).template demote<>()) == Base::dims);
		assert(out_grad.dimension(0) > 0 && (std::size_t) out_grad.dimension(0) == cached_rows);
		if (Base::is_input_layer())
			return typename Root::Data();
		DeviceTensor<Scalar> prev_out_grad = pass_back_dev(to_device<Scalar,Root::DATA_RANK>(out_grad));
		return to_host<Scalar,Root::DATA_RANK>(prev_out_grad, batch_extents<Rank>(prev_out_grad.rows, Base::dims));
	}
	inline DeviceTensor<Scalar> pass_forward_dev(DeviceTensor<Scalar> in, bool training) {
		DeviceTensor<Scalar> out(in.rows, Base::dims.get_volume());
		Context& c = Context::get();
		{
			Context::Lock l = c.lock();
			CATTLE_B200_CHECK(Api<Scalar>::activation_forward(c.handle(), Kind, param, (std::int64_t) in.rows,
					(std::int64_t) Base::dims.get_volume(), in.data(), out.data()));
		}
		cached_rows = in.rows;
		in_cache = KEEPS_INPUT ? std::move(in) : DeviceTensor<Scalar>();
		out_cache = KEEPS_OUTPUT ? out : DeviceTensor<Scalar>();  // shares the buffer with the returned tensor
		return out;
	}
	/**
	 * An element-wise activation can be applied by the epilogue of the kernel layer (or batch-norm pass) in
	 * front of it: the producer then writes f(x) next to -- or, when this layer only caches its output, instead
	 * of -- x.  Softmax normalises over a row and is not element-wise.
	 */
	inline bool request_epilogue(FusedEpilogue<Scalar>& ep, std::size_t, bool) const {
		if (Kind == CATTL3_ACT_SOFTMAX)
			return false;
		ep.act_kind = Kind;
		ep.act_param = param;
		ep.keep_pre = KEEPS_INPUT;
		return true;
	}
	inline DeviceTensor<Scalar> accept_epilogue(DeviceTensor<Scalar> pre, FusedEpilogue<Scalar>& ep, bool,
			FusedEpilogue<Scalar>*) {
		if (ep.act_out.empty() || (KEEPS_INPUT && pre.empty()))
			throw Error(CATTL3_ERR_INVALID, "activation layer: incomplete fused epilogue");
		cached_rows = ep.act_out.rows;
		in_cache = KEEPS_INPUT ? std::move(pre) : DeviceTensor<Scalar>();
		out_cache = KEEPS_OUTPUT ? ep.act_out : DeviceTensor<Scalar>();
		return ep.act_out;
	}
	inline DeviceTensor<Scalar> pass_back_dev(DeviceTensor<Scalar> out_grad) {
		if (cached_rows != out_grad.rows || (KEEPS_INPUT && in_cache.empty()) || (KEEPS_OUTPUT && out_cache.empty()))
			throw Error(CATTL3_ERR_INVALID, "activation layer: pass_back without a matching pass_forward");
		if (Base::is_input_layer())
			return DeviceTensor<Scalar>();
		DeviceTensor<Scalar> prev_out_grad(out_grad.rows, Base::dims.get_volume());
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(Api<Scalar>::activation_backward(c.handle(), Kind, param, (std::int64_t) out_grad.rows,
				(std::int64_t) Base::dims.get_volume(), in_cache.data(), out_cache.data(), out_grad.data(),
				prev_out_grad.data()));
		return prev_out_grad;
	}
protected:
	/** `param`: alpha (LeakyReLU, ELU), beta (Swish), epsilon (Softmax); unused otherwise. */
	inline ElementwiseActivationLayer(const typename Root::Dims& dims, Scalar param) :
			Base(dims),
			param(param),
			cached_rows(0) { }
	const Scalar param;
private:
	std::size_t cached_rows;
	DeviceTensor<Scalar> in_cache, out_cache;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_ELEMENTWISEACTIVATIONLAYER_H_ */
