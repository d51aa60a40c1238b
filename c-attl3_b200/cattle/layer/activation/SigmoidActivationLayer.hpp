/*
 * layer/activation/SigmoidActivationLayer.hpp -- B200 replacement of the reference's
 * SigmoidActivationLayer (C-ATTL3/layer/activation/SigmoidActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = 1 / (1 + exp(-x)); dx = y * (1 - y) * dy.
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_SIGMOIDACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_SIGMOIDACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class SigmoidActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SIGMOID> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SIGMOID> Core;
public:
	inline SigmoidActivationLayer(const typename Root::Dims& dims) :
			Core(dims, (Scalar) 0) { }
	inline Root* clone() const {
		return new SigmoidActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_SIGMOIDACTIVATIONLAYER_H_ */
