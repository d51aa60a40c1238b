/*
 * layer/activation/ReLUActivationLayer.hpp -- B200 replacement of the reference's
 * ReLUActivationLayer (C-ATTL3/layer/activation/ReLUActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = max(x, 0); dx = dy where x >= 0 (the derivative at 0 is 1, ReLUActivationLayer.hpp:45-57).
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_RELUACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_RELUACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class ReLUActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_RELU> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_RELU> Core;
public:
	inline ReLUActivationLayer(const typename Root::Dims& dims) :
			Core(dims, (Scalar) 0) { }
	inline Root* clone() const {
		return new ReLUActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_RELUACTIVATIONLAYER_H_ */
