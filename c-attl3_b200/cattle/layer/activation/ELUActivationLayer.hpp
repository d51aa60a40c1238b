/*
 * layer/activation/ELUActivationLayer.hpp -- B200 replacement of the reference's
 * ELUActivationLayer (C-ATTL3/layer/activation/ELUActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = x where x >= 0, alpha * (exp(x) - 1) elsewhere; dx = dy or (y + alpha) * dy (ELUActivationLayer.hpp:55-78).
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_ELUACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_ELUACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class ELUActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_ELU> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_ELU> Core;
public:
	inline ELUActivationLayer(const typename Root::Dims& dims, Scalar alpha = 1e-1) :
			Core(dims, alpha) { }
	inline Root* clone() const {
		return new ELUActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_ELUACTIVATIONLAYER_H_ */
