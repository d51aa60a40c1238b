/*
 * layer/activation/SwishActivationLayer.hpp -- B200 replacement of the reference's
 * SwishActivationLayer (C-ATTL3/layer/activation/SwishActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = x * s, s = 1 / (1 + exp(-beta * x)); dx = s * ((1 - s) * beta * x + 1) * dy (SwishActivationLayer.hpp:45-61).
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_SWISHACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_SWISHACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class SwishActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SWISH> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_SWISH> Core;
public:
	inline SwishActivationLayer(const typename Root::Dims& dims, Scalar beta = 1) :
			Core(dims, beta) { }
	inline Root* clone() const {
		return new SwishActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_SWISHACTIVATIONLAYER_H_ */
