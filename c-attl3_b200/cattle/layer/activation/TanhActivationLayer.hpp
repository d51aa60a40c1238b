/*
 * layer/activation/TanhActivationLayer.hpp -- B200 replacement of the reference's
 * TanhActivationLayer (C-ATTL3/layer/activation/TanhActivationLayer.hpp), same class template and
 * constructor; defines the reference header's include guard.
 *
 * y = tanh(x); dx = (1 - y * y) * dy.
 */
#ifndef C_ATTL3_LAYER_ACTIVATION_TANHACTIVATIONLAYER_H_
#define C_ATTL3_LAYER_ACTIVATION_TANHACTIVATIONLAYER_H_

#include "core/NumericUtils.hpp"
#include "b200/ElementwiseActivationLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank>
class TanhActivationLayer : public b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_TANH> {
	typedef Layer<Scalar,Rank> Root;
	typedef b200::ElementwiseActivationLayer<Scalar,Rank,CATTL3_ACT_TANH> Core;
public:
	inline TanhActivationLayer(const typename Root::Dims& dims) :
			Core(dims, (Scalar) 0) { }
	inline Root* clone() const {
		return new TanhActivationLayer(*this);
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_ACTIVATION_TANHACTIVATIONLAYER_H_ */
