/*
 * b200/DeviceNetwork.hpp -- the device-resident face of a neural network, the network-level twin of
 * b200::DeviceLayer (the reference sketched the same split in C-ATTL3/core/gpu/GPUNeuralNetwork.hpp:18-49).
 * Networks implementing it hand activations to each other, and to the optimizer's batch loop, without
 * a round trip through host memory.
 */
#ifndef C_ATTL3_B200_DEVICENETWORK_H_
#define C_ATTL3_B200_DEVICENETWORK_H_

#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar, std::size_t Rank>
class DeviceNetwork {
public:
	virtual ~DeviceNetwork() = default;
	/** NeuralNetwork::propagate on device tensors (C-ATTL3/core/NeuralNetwork.hpp:89). */
	virtual DeviceTensor<Scalar> propagate_dev(DeviceTensor<Scalar> input, bool training) = 0;
	/** NeuralNetwork::backpropagate on device tensors (:99); empty result for a foremost network. */
	virtual DeviceTensor<Scalar> backpropagate_dev(DeviceTensor<Scalar> out_grad) = 0;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_DEVICENETWORK_H_ */
