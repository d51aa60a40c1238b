/*
 * optimizer/AdaGradOptimizer.hpp -- B200 replacement of the reference's AdaGradOptimizer
 * (C-ATTL3/optimizer/AdaGradOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_ADAGRAD); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * s <- s + g^2, p <- p - lr * g / (sqrt(s) + epsilon) (AdaGradOptimizer.hpp:49-71).
 */
#ifndef C_ATTL3_OPTIMIZER_ADAGRADOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_ADAGRADOPTIMIZER_H_

#include <cassert>

#include "core/NumericUtils.hpp"
#include "optimizer/SGDOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class AdaGradOptimizer : public SGDOptimizer<Scalar,Rank,Sequential> {
	typedef SGDOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline AdaGradOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar learning_rate = 1e-2, Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
				Base(loss, batch_size),
				learning_rate(learning_rate),
				epsilon(epsilon) {
		assert(learning_rate > 0);
		assert(epsilon > 0);
	}
	virtual ~AdaGradOptimizer() = default;
protected:
	inline void _fit(const std::vector<Parameters<Scalar>*>& params_vec) { }
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, Base::make_step(CATTL3_OPT_ADAGRAD, learning_rate, 0, 0, epsilon));
	}
	const Scalar learning_rate, epsilon;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_ADAGRADOPTIMIZER_H_ */
