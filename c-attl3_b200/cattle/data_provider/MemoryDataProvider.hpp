/*
 * data_provider/MemoryDataProvider.hpp -- B200 replacement of the reference's in-memory data provider
 * (C-ATTL3/data_provider/MemoryDataProvider.hpp:33-123), same class template and constructor; defines the
 * reference header's include guard.
 *
 * Host face: DataProvider as in the reference -- rows are shuffled once, at construction, when Shuffle is set
 * (:61-62,96-107: a random row permutation from std::random_shuffle applied to observations and objectives
 * alike), get_data returns the next `batch_size` rows (:72-83).  Because the sample index is the fastest
 * dimension, a batch of rows is a run of `rows` consecutive elements in every column of the (instances x volume)
 * matrix view, so the slice here is one memcpy per column instead of a generic tensor-slice evaluation.
 *
 * Device face (b200::DeviceDataSource, non-sequential data): the whole data set is uploaded to HBM once, on the
 * first device fetch, and every mini-batch (or data-parallel shard of it) is cut out by one strided-copy kernel
 * (cattl3_slice_rows): no per-step host slicing, no per-step upload.  CATTL3_DEVICE_DATASET_GB (default 64)
 * bounds what may be made resident; larger sets fall back to get_data + the batch loop's input feed.
 */
#ifndef C_ATTL3_DATA_PROVIDER_MEMORYDATAPROVIDER_H_
#define C_ATTL3_DATA_PROVIDER_MEMORYDATAPROVIDER_H_

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#include "core/DataProvider.hpp"
#include "b200/DeviceDataSource.hpp"

namespace cattle {

/**
 * An alias for a unique pointer to a tensor.
 */
template<typename Scalar, std::size_t Rank>
using TensorPtr = std::unique_ptr<Tensor<Scalar,Rank>>;

template<typename Scalar, std::size_t Rank, bool Sequential, bool Shuffle = true>
class MemoryDataProvider : public DataProvider<Scalar,Rank,Sequential>, public b200::DeviceDataSource<Scalar> {
	typedef DataProvider<Scalar,Rank,Sequential> Base;
	typedef TensorPtr<Scalar,Base::DATA_RANK> DataPtr;
public:
	/**
	 * @param obs The observations, instances first.
	 * @param obj The objectives, one row per observation.
	 */
	inline MemoryDataProvider(DataPtr obs, DataPtr obj) :
			obs(std::move(obs)),
			obj(std::move(obj)),
			offset(0) {
		assert(this->obs != nullptr && this->obj != nullptr);
		assert(this->obs->dimension(0) == this->obj->dimension(0) && "mismatched data and obj tensor row numbers");
		obs_dims = Dimensions<std::size_t,Base::DATA_RANK>(this->obs->dimensions()).template demote<Sequential + 1>();
		obj_dims = Dimensions<std::size_t,Base::DATA_RANK>(this->obj->dimensions()).template demote<Sequential + 1>();
		instances = (std::size_t) this->obs->dimension(0);
		if (Shuffle)
			shuffle_rows();
	}
	inline const typename Base::Dims& get_obs_dims() const {
		return obs_dims;
	}
	inline const typename Base::Dims& get_obj_dims() const {
		return obj_dims;
	}
	inline bool has_more() {
		return offset < instances;
	}
	inline DataPair<Scalar,Rank,Sequential> get_data(std::size_t batch_size = std::numeric_limits<std::size_t>::max()) {
		if (!has_more())
			throw std::out_of_range("no more data left to fetch");
		const std::size_t rows = std::min(batch_size, instances - offset);
		typename Base::Data obs_batch = rows_of(*obs, offset, rows);
		typename Base::Data obj_batch = rows_of(*obj, offset, rows);
		offset += rows;
		return std::make_pair(std::move(obs_batch), std::move(obj_batch));
	}
	inline void reset() {
		offset = 0;
	}
	inline void skip(std::size_t instances) {
		offset = std::min(this->instances, offset + instances);
	}
	inline bool device_resident() {
		if (instances == 0)
			return false;   // (sequences included: samples are the fastest rank, a mini-batch is a row slice either way)
		if (!dev_obs.empty())
			return true;
		static const double budget_gb = [] {
			const char* v = std::getenv("CATTL3_DEVICE_DATASET_GB");
			return v ? std::atof(v) : 64.0;
		}();
		const double bytes = (double) sizeof(Scalar) * ((double) obs->size() + (double) obj->size());
		if (bytes > budget_gb * 1e9)
			return false;
		dev_obs = b200::DeviceBuffer<Scalar>((std::size_t) obs->size());
		dev_obs.upload(obs->data(), (std::size_t) obs->size());
		dev_obj = b200::DeviceBuffer<Scalar>((std::size_t) obj->size());
		dev_obj.upload(obj->data(), (std::size_t) obj->size());
		return true;
	}
	inline std::size_t next_batch_dev(std::size_t batch_size, std::size_t rank, std::size_t world,
			b200::DeviceTensor<Scalar>& obs_batch, b200::DeviceTensor<Scalar>& obj_batch) {
		if (!has_more())
			throw std::out_of_range("no more data left to fetch");
		if (!device_resident())
			throw b200::Error(CATTL3_ERR_UNSUPPORTED, "MemoryDataProvider: the data set is not device resident");
		const std::size_t rows = std::min(batch_size, instances - offset);
		// fewer rows than ranks (a ragged last batch): every rank takes all of them (SGDOptimizer::_train rescales)
		const std::size_t lo = rows < world ? 0 : rows * rank / world, hi = rows < world ? rows : rows * (rank + 1) / world;
		cut(dev_obs, (std::size_t) obs->size() / instances, offset + lo, hi - lo, obs_batch);
		cut(dev_obj, (std::size_t) obj->size() / instances, offset + lo, hi - lo, obj_batch);
		offset += rows;
		return rows;
	}
private:
	/** Rows [first, first + rows) of an (instances x volume) tensor: one contiguous run per column. */
	inline typename Base::Data rows_of(const typename Base::Data& data, std::size_t first, std::size_t rows) const {
		typename Base::Data::Dimensions extents = data.dimensions();
		extents[0] = rows;
		typename Base::Data batch(extents);
		const std::size_t volume = (std::size_t) data.size() / instances;
		const Scalar* src = data.data() + first;
		Scalar* dst = batch.data();
		for (std::size_t j = 0; j < volume; ++j)
			std::memcpy(dst + j * rows, src + j * instances, rows * sizeof(Scalar));
		return batch;
	}
	/** `batch` is overwritten in place if it arrives with the slice's shape (DeviceDataSource.hpp), else replaced. */
	inline void cut(const b200::DeviceBuffer<Scalar>& data, std::size_t volume, std::size_t first, std::size_t rows,
			b200::DeviceTensor<Scalar>& batch) const {
		if (rows == 0) {
			batch = b200::DeviceTensor<Scalar>();
			return;
		}
		if (batch.rows != rows || batch.size() != rows * volume)
			batch = b200::DeviceTensor<Scalar>(rows, volume);
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(b200::Api<Scalar>::slice_rows(c.handle(), (std::int64_t) instances, (std::int64_t) volume,
				(std::int64_t) first, (std::int64_t) rows, data.data(), batch.data()));
	}
	/** The reference's shuffle (:96-107): row i moves to position p[i], p = std::random_shuffle of the identity. */
	inline void shuffle_rows() {
		std::vector<int> perm(instances);
		for (std::size_t i = 0; i < instances; ++i)
			perm[i] = (int) i;
		std::random_shuffle(perm.begin(), perm.end());
		permute(*obs, perm);
		permute(*obj, perm);
	}
	inline void permute(typename Base::Data& data, const std::vector<int>& perm) const {
		const std::size_t volume = (std::size_t) data.size() / instances;
		std::vector<Scalar> column(instances);
		for (std::size_t j = 0; j < volume; ++j) {
			Scalar* col = data.data() + j * instances;
			for (std::size_t i = 0; i < instances; ++i)
				column[perm[i]] = col[i];
			std::memcpy(col, column.data(), instances * sizeof(Scalar));
		}
	}
	DataPtr obs, obj;
	typename Base::Dims obs_dims, obj_dims;
	std::size_t instances, offset;
	// the data set in HBM (uploaded on the first device fetch)
	b200::DeviceBuffer<Scalar> dev_obs, dev_obj;
};

} /* namespace cattle */

#endif /* C_ATTL3_DATA_PROVIDER_MEMORYDATAPROVIDER_H_ */
