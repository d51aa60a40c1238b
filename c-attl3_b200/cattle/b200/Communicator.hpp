/*
 * b200/Communicator.hpp -- the process's place in a data-parallel job: one process per GPU, WORLD_SIZE /
 * RANK / LOCAL_RANK from the environment (the torchrun convention), gradients exchanged with NCCL over
 * NVLink through the cattl3_comm_* entry points (include/cattl3_b200.h).  With WORLD_SIZE unset or 1
 * everything here is a no-op and libnccl is never loaded.  The reference has no counterpart
 * (single process).
 */
#ifndef C_ATTL3_B200_COMMUNICATOR_H_
#define C_ATTL3_B200_COMMUNICATOR_H_

#include <cstddef>
#include <cstdint>

#include "b200/Runtime.hpp"

namespace cattle {
namespace b200 {

class Communicator {
public:
	inline static Communicator& get() {
		static Communicator instance;
		return instance;
	}
	inline std::size_t world_size() const {
		return (std::size_t) cattl3_comm_world_size(comm);
	}
	inline std::size_t rank() const {
		return (std::size_t) cattl3_comm_rank(comm);
	}
	inline void group_start() {
		CATTLE_B200_CHECK(cattl3_comm_group_start(comm));
	}
	inline void group_end() {
		CATTLE_B200_CHECK(cattl3_comm_group_end(comm));
	}
	/** In-place sum over all ranks of a device array, enqueued on the context's stream. */
	inline void all_reduce_sum(float* dev, std::size_t count) {
		Context::Lock l = Context::get().lock();
		CATTLE_B200_CHECK(cattl3_comm_allreduce_sum_f32(comm, dev, (std::int64_t) count));
	}
	inline void all_reduce_sum(double* dev, std::size_t count) {
		Context::Lock l = Context::get().lock();
		CATTLE_B200_CHECK(cattl3_comm_allreduce_sum_f64(comm, dev, (std::int64_t) count));
	}
	/**
	 * The same sum on the communicator's side stream: it starts when everything enqueued so far has finished and runs
	 * beside what is enqueued next; wait() makes the context's stream wait for every exchange started this way.
	 */
	inline void all_reduce_sum_async(float* dev, std::size_t count) {
		Context::Lock l = Context::get().lock();
		CATTLE_B200_CHECK(cattl3_comm_allreduce_sum_async_f32(comm, dev, (std::int64_t) count));
	}
	inline void all_reduce_sum_async(double* dev, std::size_t count) {
		Context::Lock l = Context::get().lock();
		CATTLE_B200_CHECK(cattl3_comm_allreduce_sum_async_f64(comm, dev, (std::int64_t) count));
	}
	inline void wait() {
		Context::Lock l = Context::get().lock();
		CATTLE_B200_CHECK(cattl3_comm_wait(comm));
	}
	/** Sum over all ranks of a host scalar (loss bookkeeping); synchronises. */
	inline double all_reduce_sum(double value) {
		if (world_size() == 1)
			return value;
		DeviceBuffer<double> buf(1);
		buf.upload(&value, 1);
		all_reduce_sum(buf.data(), 1);
		buf.download(&value, 1);
		return value;
	}
	Communicator(const Communicator&) = delete;
	Communicator& operator=(const Communicator&) = delete;
private:
	inline Communicator() :
			comm(nullptr) {
		CATTLE_B200_CHECK(cattl3_comm_create_from_env(&comm, Context::get().handle()));
	}
	inline ~Communicator() {
		cattl3_comm_destroy(comm);
	}
	cattl3_comm* comm;
};

} /* namespace b200 */
} /* namespace cattle */

#endif /* C_ATTL3_B200_COMMUNICATOR_H_ */
