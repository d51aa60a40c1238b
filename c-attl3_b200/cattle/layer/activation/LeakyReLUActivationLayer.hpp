/*
 * layer/activation/LeakyReLUActivationLayer.hpp -- B200 replacement of the reference's
 * LeakyReLUActivationLayer (C-ATTL3/layer/activation/LeakyReLUActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = max(x, alpha * x); dx = dy where x >= 0, alpha * dy elsewhere (LeakyReLUActivationLayer.hpp:50-62).
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_LEAKYRELUACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_LEAKYRELUACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class LeakyReLUActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_LEAKY_RELU> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_LEAKY_RELU> Core;
public:
	inline LeakyReLUActivationLayer(const typename Root::Dims& dims, Scalar alpha = 1e-1) :
			Core(dims, alpha) { }
	inline Root* clone() const {
		return new LeakyReLUActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_LEAKYRELUACTIVATIONLAYER_H_ */
