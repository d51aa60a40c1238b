/*
 * layer/activation/SoftmaxActivationLayer.hpp -- B200 replacement of the reference's
 * SoftmaxActivationLayer (C-ATTL3/layer/activation/SoftmaxActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * Row-wise over the whole observation: y = exp(x - max) / (sum exp(x - max) + epsilon); dx = y * (dy - <y, dy>) (SoftmaxActivationLayer.hpp:49-78).
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_SOFTMAXACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_SOFTMAXACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class SoftmaxActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SOFTMAX> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SOFTMAX> Core;
public:
	inline SoftmaxActivationLayer(const typename Root::Dims& dims, Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
			Core(dims, epsilon) { }
	inline Root* clone() const {
		return new SoftmaxActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_SOFTMAXACTIVATIONLAYER_H_ */
