/*
 * optimizer/RMSPropOptimizer.hpp -- B200 replacement of the reference's RMSPropOptimizer
 * (C-ATTL3/optimizer/RMSPropOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_RMSPROP); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * s <- (1 - l2_decay) * s + l2_decay * g^2, p <- p - lr * g / (sqrt(s) + epsilon) (RMSPropOptimizer.hpp:42-46).
 */
#ifndef C_ATTL3_OPTIMIZER_RMSPROPOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_RMSPROPOPTIMIZER_H_

#include <cassert>

#include "optimizer/AdaGradOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class RMSPropOptimizer : public AdaGradOptimizer<Scalar,Rank,Sequential> {
	typedef AdaGradOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline RMSPropOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar learning_rate = 1e-3, Scalar l2_decay = 1e-1, Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
				Base(loss, batch_size, learning_rate, epsilon),
				l2_decay(l2_decay) {
		assert(l2_decay >= 0 && l2_decay <= 1);
	}
protected:
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, Base::make_step(CATTL3_OPT_RMSPROP, Base::learning_rate, 0, l2_decay,
				Base::epsilon));
	}
	const Scalar l2_decay;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_RMSPROPOPTIMIZER_H_ */
