/*
 * b200/DeviceLoss.hpp -- the device-resident face of a loss function: per-sample losses and the
 * gradient with respect to the network output, computed where the output already is.  The batch loop
 * (optimizer/SGDOptimizer.hpp) uses it when both the network and the loss speak the device API, which
 * removes the last per-step host round trip of the reference's loop
 * (C-ATTL3/optimizer/SGDOptimizer.hpp:48,55: Loss::function / Loss::d_function on host tensors).
 */
#ifndef C_ATTL3_B200_DEVICELOSS_H_
#define C_ATTL3_B200_DEVICELOSS_H_

#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar>
class DeviceLoss {
public:
	virtual ~DeviceLoss() = default;
	/**
	 * @param out The network output, rows x volume (rows fastest).
	 * @param obj The objectives, same shape.
	 * @param grad_divisor The gradient is d loss / d out DIVIDED by this (the nominal batch size).
	 * @param losses Receives one loss per row (rows elements).
	 * @return The gradient tensor.
	 */
	virtual DeviceTensor<Scalar> loss_and_gradient_dev(const DeviceTensor<Scalar>& out, const DeviceTensor<Scalar>& obj,
			Scalar grad_divisor, DeviceBuffer<Scalar>& losses) const = 0;
protected:
	inline static DeviceTensor<Scalar> run(int kind, Scalar epsilon, const DeviceTensor<Scalar>& out,
			const DeviceTensor<Scalar>& obj, Scalar grad_divisor, DeviceBuffer<Scalar>& losses) {
		if (out.empty() || out.size() != obj.size() || out.rows != obj.rows)
			throw Error(CATTL3_ERR_INVALID, "loss: output and objective tensors differ in shape");
		const std::size_t volume = out.size() / out.rows;
		DeviceTensor<Scalar> grad(out.rows, volume);
		if (losses.size() != out.rows)
			losses = DeviceBuffer<Scalar>(out.rows);
		Context& c = Context::get();
		Context::Lock l = c.lock();
		CATTLE_B200_CHECK(Api<Scalar>::loss(c.handle(), kind, (std::int64_t) out.rows, (std::int64_t) volume, epsilon,
				grad_divisor, out.data(), obj.data(), losses.data(), grad.data()));
		return grad;
	}
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_DEVICELOSS_H_ */
