/*
 * optimizer/MomentumSGDOptimizer.hpp -- B200 replacement of the reference's MomentumSGDOptimizer
 * (C-ATTL3/optimizer/MomentumSGDOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_MOMENTUM); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * v <- momentum * v + lr_e * g, p <- p - v with lr_e = init_lr / (1 + annealing_rate * epoch) (MomentumSGDOptimizer.hpp:54-72).
 */
#ifndef C_ATTL3_OPTIMIZER_MOMENTUMSGDOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_MOMENTUMSGDOPTIMIZER_H_

#include <cassert>

#include "optimizer/SGDOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class MomentumSGDOptimizer : public SGDOptimizer<Scalar,Rank,Sequential> {
	typedef SGDOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline MomentumSGDOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar init_learning_rate = 1e-3, Scalar annealing_rate = 1e-3, Scalar momentum = .9) :
				Base(loss, batch_size),
				init_learning_rate(init_learning_rate),
				annealing_rate(annealing_rate),
				momentum(momentum) {
		assert(init_learning_rate > 0);
		assert(annealing_rate >= 0);
		assert(momentum > 0 && momentum < 1);
	}
	virtual ~MomentumSGDOptimizer() = default;
protected:
	inline void _fit(const std::vector<Parameters<Scalar>*>& params_vec) { }
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, annealed_step(CATTL3_OPT_MOMENTUM, epoch));
	}
	Scalar calculate_learning_rate(std::size_t epoch) {
		return init_learning_rate / (1 + annealing_rate * epoch);
	}
	inline cattl3_opt_step annealed_step(int kind, std::size_t epoch) {
		cattl3_opt_step step = Base::make_step(kind, init_learning_rate, annealing_rate, momentum, 0);
		step.lr_epoch = calculate_learning_rate(epoch);
		return step;
	}
	const Scalar init_learning_rate, annealing_rate, momentum;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_MOMENTUMSGDOPTIMIZER_H_ */
