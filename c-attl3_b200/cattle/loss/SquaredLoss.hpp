/*
 * loss/SquaredLoss.hpp -- B200 replacement of the reference's squared-error loss
 * (C-ATTL3/loss/SquaredLoss.hpp:20-36), same class template; defines the reference header's include guard.
 *
 * L_n = sum_j (out_nj - obj_nj)^2, dL/dout = 2 (out - obj).  The host face (UniversalLoss's
 * _function / _d_function) is kept for everything that evaluates a loss on host tensors
 * (GradientCheck, Optimizer::test, sequential networks); the device face (b200::DeviceLoss) serves the
 * batch loop.
 */
#ifndef C_ATTL3_LOSS_SQUAREDLOSS_H_
#define C_ATTL3_LOSS_SQUAREDLOSS_H_

#include <utility>

#include "loss/UniversalLoss.hpp"
#include "b200/DeviceLoss.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class SquaredLoss : public UniversalLoss<Scalar,Rank,Sequential>, public b200::DeviceLoss<Scalar> {
	typedef Loss<Scalar,Rank,Sequential> Root;
	typedef UniversalLoss<Scalar,Rank,Sequential> Base;
public:
	inline b200::DeviceTensor<Scalar> loss_and_gradient_dev(const b200::DeviceTensor<Scalar>& out,
			const b200::DeviceTensor<Scalar>& obj, Scalar grad_divisor, b200::DeviceBuffer<Scalar>& losses) const {
		return b200::DeviceLoss<Scalar>::run(CATTL3_LOSS_SQUARED, (Scalar) 0, out, obj, grad_divisor, losses);
	}
protected:
	inline ColVector<Scalar> _function(typename Root::Data out, typename Root::Data obj) const {
		const std::size_t samples = out.dimension(0), volume = out.size() / samples;
		MatrixMap<Scalar> out_mat(out.data(), samples, volume), obj_mat(obj.data(), samples, volume);
		return (out_mat - obj_mat).array().square().rowwise().sum();
	}
	inline typename Root::Data _d_function(typename Root::Data out, typename Root::Data obj,
			const typename Base::RankwiseArray& grad_dims) const {
		return (out - obj) * (Scalar) 2;
	}
};

} /* namespace cattle */

#endif /* C_ATTL3_LOSS_SQUAREDLOSS_H_ */
