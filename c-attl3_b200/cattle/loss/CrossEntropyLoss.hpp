/*
 * loss/CrossEntropyLoss.hpp -- B200 replacement of the reference's cross-entropy loss
 * (C-ATTL3/loss/CrossEntropyLoss.hpp:22-47), same class template and constructor; defines the reference
 * header's include guard.
 *
 * L_n = -sum_j ln(out_nj + epsilon) obj_nj, dL/dout = -obj / (out + epsilon): epsilon sits inside the
 * logarithm and inside the quotient (:36-41).  Host face for host-side callers, device face
 * (b200::DeviceLoss) for the batch loop.
 */
#ifndef C_ATTL3_LOSS_CROSSENTROPYLOSS_H_
#define C_ATTL3_LOSS_CROSSENTROPYLOSS_H_

#include <utility>

#include "core/NumericUtils.hpp"
#include "loss/UniversalLoss.hpp"
#include "b200/DeviceLoss.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool Sequential>
class CrossEntropyLoss : public UniversalLoss<Scalar,Rank,Sequential>, public b200::DeviceLoss<Scalar> {
	typedef Loss<Scalar,Rank,Sequential> Root;
	typedef UniversalLoss<Scalar,Rank,Sequential> Base;
public:
	/**
	 * @param epsilon A small constant that keeps the logarithm and the quotient finite.
	 */
	CrossEntropyLoss(Scalar epsilon = NumericUtils<Scalar>::EPSILON2) :
			epsilon(epsilon) { }
	inline b200::DeviceTensor<Scalar> loss_and_gradient_dev(const b200::DeviceTensor<Scalar>& out,
			const b200::DeviceTensor<Scalar>& obj, Scalar grad_divisor, b200::DeviceBuffer<Scalar>& losses) const {
		return b200::DeviceLoss<Scalar>::run(CATTL3_LOSS_CROSS_ENTROPY, epsilon, out, obj, grad_divisor, losses);
	}
protected:
	inline ColVector<Scalar> _function(typename Root::Data out, typename Root::Data obj) const {
		const std::size_t samples = out.dimension(0), volume = out.size() / samples;
		MatrixMap<Scalar> out_mat(out.data(), samples, volume), obj_mat(obj.data(), samples, volume);
		return -((out_mat.array() + epsilon).log() * obj_mat.array()).matrix().rowwise().sum();
	}
	inline typename Root::Data _d_function(typename Root::Data out, typename Root::Data obj,
			const typename Base::RankwiseArray& grad_dims) const {
		return -obj / (out + epsilon);
	}
private:
	Scalar epsilon;
};

} /* namespace cattle */

#endif /* C_ATTL3_LOSS_CROSSENTROPYLOSS_H_ */
