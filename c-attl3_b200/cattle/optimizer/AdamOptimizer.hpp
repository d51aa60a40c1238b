/*
 * optimizer/AdamOptimizer.hpp -- B200 replacement of the reference's AdamOptimizer
 * (C-ATTL3/optimizer/AdamOptimizer.hpp): same class template, constructor arguments and defaults;
 * defines the reference header's include guard.  The update rule runs as one fused device kernel per
 * parameter array (SGDOptimizer::fused_step -> cattl3_optimizer_step, kind CATTL3_OPT_ADAM); this header only
 * evaluates the step-dependent scalars, in the Scalar type and in the reference's own expression order.
 *
 * m <- (1 - d1) m + d1 g, v <- (1 - d2) v + d2 g^2, p <- p - (m * lr * c1) / sqrt(v * c2 + epsilon), with c_k = 1 / (1 - (1 - d_k)^(t+1) + epsilon): epsilon sits inside the bias corrections and inside the root (AdamOptimizer.hpp:66-82).
 */
#ifndef C_ATTL3_OPTIMIZER_ADAMOPTIMIZER_H_
#define C_ATTL3_OPTIMIZER_ADAMOPTIMIZER_H_

#include <cassert>
#include <cmath>

#include "core/NumericUtils.hpp"
#include "optimizer/SGDOptimizer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class AdamOptimizer : public SGDOptimizer<Scalar,Rank,Sequential> {
	typedef SGDOptimizer<Scalar,Rank,Sequential> Base;
public:
	inline AdamOptimizer(LossSharedPtr<Scalar,Rank,Sequential> loss, std::size_t batch_size = 1,
			Scalar learning_rate = 1e-3, Scalar l1_decay = 1e-1, Scalar l2_decay = 1e-3,
			Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
				Base(loss, batch_size),
				learning_rate(learning_rate),
				l1_decay(l1_decay),
				l2_decay(l2_decay),
				epsilon(epsilon) {
		assert(learning_rate > 0);
		assert(l1_decay >= 0 && l1_decay <= 1);
		assert(l2_decay >= 0 && l2_decay <= 1);
		assert(epsilon > 0);
	}
	virtual ~AdamOptimizer() = default;
protected:
	inline void _fit(const std::vector<Parameters<Scalar>*>& params_vec) { }
	inline void _update_params(const std::vector<Parameters<Scalar>*>& params_vec, std::size_t epoch,
			std::size_t timestep) {
		Base::fused_step(params_vec, corrected_step(CATTL3_OPT_ADAM, timestep));
	}
	/** 1 / (1 - (1 - decay)^(timestep + ahead) + epsilon). */
	inline Scalar bias_correction(Scalar decay, std::size_t timestep, std::size_t ahead) const {
		return (Scalar) 1 / (1 - pow(1 - decay, timestep + ahead) + epsilon);
	}
	inline cattl3_opt_step corrected_step(int kind, std::size_t timestep) const {
		cattl3_opt_step step = Base::make_step(kind, learning_rate, l1_decay, l2_decay, epsilon);
		step.c1 = bias_correction(l1_decay, timestep, 1);
		step.c1n = bias_correction(l1_decay, timestep, 2);
		step.c2 = bias_correction(l2_decay, timestep, 1);
		return step;
	}
	const Scalar learning_rate, l1_decay, l2_decay, epsilon;
};

} /* namespace cattle */

#endif /* C_ATTL3_OPTIMIZER_ADAMOPTIMIZER_H_ */
