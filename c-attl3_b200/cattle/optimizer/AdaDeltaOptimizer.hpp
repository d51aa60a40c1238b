/*
 * optimizer/AdaDeltaOptimizer.hpp -- B200 replacement of the reference's AdaDeltaOptimizer
 * (C-ATTL3/optimizer/AdaDeltaOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_ADADELTA); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * s <- (1 - decay) * s + decay * g^2, u = -g * sqrt(q + epsilon) / sqrt(s + epsilon), p <- p + u, q <- (1 - decay) * q + decay * u^2 (AdaDeltaOptimizer.hpp:53-67).
 */
#ifndef C_ATTL3_OPTIMIZER_ADADELTAOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_ADADELTAOPTIMIZER_H_

#include <cassert>

#include "core/NumericUtils.hpp"
#include "optimizer/SGDOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class AdaDeltaOptimizer : public SGDOptimizer<Scalar,Rank,Sequential> {
	typedef SGDOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline AdaDeltaOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar decay = 5e-2, Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
				Base(loss, batch_size),
				decay(decay),
				epsilon(epsilon) {
		assert(decay >= 0 && decay <= 1);
		assert(epsilon > 0);
	}
protected:
	inline void _fit(const std::vector<Parameters<Scalar>*>& params_vec) { }
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, Base::make_step(CATTL3_OPT_ADADELTA, 0, decay, 0, epsilon));
	}
	const Scalar decay, epsilon;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_ADADELTAOPTIMIZER_H_ */
