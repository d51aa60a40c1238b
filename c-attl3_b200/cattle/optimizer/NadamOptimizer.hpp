/*
 * optimizer/NadamOptimizer.hpp -- B200 replacement of the reference's NadamOptimizer
 * (C-ATTL3/optimizer/NadamOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_NADAM); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * m, v as Adam, p <- p - lr * (g * d1 * c1 + m * (1 - d1) * c1_next) / sqrt(v * c2 + epsilon) (NadamOptimizer.hpp:44-63).
 */
#ifndef C_ATTL3_OPTIMIZER_NADAMOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_NADAMOPTIMIZER_H_

#include "optimizer/AdamOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class NadamOptimizer : public AdamOptimizer<Scalar,Rank,Sequential> {
	typedef AdamOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline NadamOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar learning_rate = 1e-3, Scalar l1_decay = 1e-1, Scalar l2_decay = 1e-3,
			Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
				Base(loss, batch_size, learning_rate, l1_decay, l2_decay, epsilon) { }
protected:
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, Base::corrected_step(CATTL3_OPT_NADAM, timestep));
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_NADAMOPTIMIZER_H_ */
