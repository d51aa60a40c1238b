/*
 * b200/DeviceDataSource.hpp -- the device-resident face of a data provider: mini-batches cut out of a data set
 * that already lives in HBM (180 GB per B200 hold most in-memory data sets whole), so that the batch loop
 * (optimizer/SGDOptimizer.hpp) neither slices on the host nor uploads per step.  The reference's loop takes host
 * tensors from DataProvider::get_data (C-ATTL3/optimizer/SGDOptimizer.hpp:44-45).
 */
#ifndef C_ATTL3_B200_DEVICEDATASOURCE_H_
#define C_ATTL3_B200_DEVICEDATASOURCE_H_

#include <cstddef>

#include "b200/DeviceLayer.hpp"

namespace cattle {
namespace b200 {

template<typename Scalar>
class DeviceDataSource {
public:
	virtual ~DeviceDataSource() = default;
	/** Whether the data set is (or can be made) device resident; false = use DataProvider::get_data. */
	virtual bool device_resident() = 0;
	/**
	 * The next mini-batch of up to `batch_size` instances, as DataProvider::get_data would return it, but only the
	 * rows [n * rank / world, n * (rank + 1) / world) of its n instances (the data-parallel shard), on the device.
	 *
	 * Tensors that arrive with exactly the shard's shape are overwritten in place instead of being replaced (the fixed
	 * input buffers of a captured training step); pass empty tensors otherwise.
	 *
	 * @return n, the number of instances the whole mini-batch has; the tensors are empty if the shard is.
	 */
	virtual std::size_t next_batch_dev(std::size_t batch_size, std::size_t rank, std::size_t world,
			DeviceTensor<Scalar>& obs, DeviceTensor<Scalar>& obj) = 0;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_DEVICEDATASOURCE_H_ */
