/*
 * optimizer/NesterovMomentumSGDOptimizer.hpp -- B200 replacement of the reference's NesterovMomentumSGDOptimizer
 * (C-ATTL3/optimizer/NesterovMomentumSGDOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_NESTEROV); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * v' <- momentum * v - lr_e * g, p <- p - momentum * v + (1 + momentum) * v' (NesterovMomentumSGDOptimizer.hpp:43-55).
 */
#ifndef C_ATTL3_OPTIMIZER_NESTEROVMOMENTUMSGDOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_NESTEROVMOMENTUMSGDOPTIMIZER_H_

#include "optimizer/MomentumSGDOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class NesterovMomentumSGDOptimizer : public MomentumSGDOptimizer<Scalar,Rank,Sequential> {
	typedef MomentumSGDOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline NesterovMomentumSGDOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss,
			std::size_t batch_size = 1, Scalar init_learning_rate = 1e-3, Scalar annealing_rate = 1e-3,
			Scalar momentum = .9) :
				Base(loss, batch_size, init_learning_rate, annealing_rate, momentum) { }
protected:
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, Base::annealed_step(CATTL3_OPT_NESTEROV, epoch));
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_NESTEROVMOMENTUMSGDOPTIMIZER_H_ */
