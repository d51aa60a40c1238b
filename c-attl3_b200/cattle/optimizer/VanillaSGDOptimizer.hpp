/*
 * optimizer/VanillaSGDOptimizer.hpp -- B200 replacement of the reference's VanillaSGDOptimizer
 * (C-ATTL3/optimizer/VanillaSGDOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_VANILLA_SGD); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * p <- p - lr * g (VanillaSGDOptimizer.hpp:38-43).
 */
#ifndef C_ATTL3_OPTIMIZER_VANILLASGDOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_VANILLASGDOPTIMIZER_H_

#include <cassert>

#include "optimizer/SGDOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class VanillaSGDOptimizer : public SGDOptimizer<Scalar,Rank,Sequential> {
	typedef SGDOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline VanillaSGDOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar learning_rate = 1e-3) :
				Base(loss, batch_size),
				learning_rate(learning_rate) {
		assert(learning_rate > 0);
	}
protected:
	inline void _fit(const std::vector<Parameters<Scalar>*>& params_vec) { }
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, Base::make_step(CATTL3_OPT_VANILLA_SGD, learning_rate, 0, 0, 0));
	}
	const Scalar learning_rate;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_VANILLASGDOPTIMIZER_H_ */
