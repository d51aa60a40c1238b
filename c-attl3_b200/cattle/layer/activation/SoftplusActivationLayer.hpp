/*
 * layer/activation/SoftplusActivationLayer.hpp -- B200 replacement of the reference's
 * SoftplusActivationLayer (C-ATTL3/layer/activation/SoftplusActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = log(1 + exp(x)); dx = dy / (1 + exp(-x)) (SoftplusActivationLayer.hpp:40-52).
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_SOFTPLUSACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_SOFTPLUSACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class SoftplusActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SOFTPLUS> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SOFTPLUS> Core;
public:
	inline SoftplusActivationLayer(const typename Root::Dims& dims) :
			Core(dims, (Scalar) 0) { }
	inline Root* clone() const {
		return new SoftplusActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_SOFTPLUSACTIVATIONLAYER_H_ */
