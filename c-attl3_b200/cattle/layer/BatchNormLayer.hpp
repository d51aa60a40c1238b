/*
 * layer/BatchNormLayer.hpp -- B200 replacement of the reference's batch normalisation layer
 * (C-ATTL3/layer/BatchNormLayer.hpp:29-470), same class template (PerLastRank defaults to
 * Rank == 3), same constructor; defines the reference header's include guard.
 *
 * Semantics kept from the reference (BatchNormLayer.hpp:225-262):
 *  - training: mu = mean(x), inv_sd = 1 / sqrt(mean((x - mu)^2) + epsilon) (biased variance, epsilon
 *    inside the root), y = gamma * (x - mu) * inv_sd + beta; the running statistics are the mean and the
 *    INVERSE standard deviation, assigned on the first training batch and afterwards updated as
 *    (1 - decay) * avg + decay * new;
 *  - inference: y = gamma * (x - avg_mean) * avg_inv_sd + beta;
 *  - backward: dgamma += sum(dy * xhat), dbeta += sum(dy),
 *    dx = (L * g - sum(g) - xhat * sum(xhat * g)) * inv_sd / L with g = gamma * dy;
 *  - PerLastRank: one statistic per index of the last rank (the channel), each exposed as FOUR 1x1
 *    Parameters in the order avg_mean, avg_inv_sd, gamma, beta, channel by channel (:263-272) -- the
 *    order optimizers and parameter files rely on; otherwise one statistic per activation, exposed as
 *    four 1 x volume Parameters.
 *
 * On the B200 the four quantities are four contiguous device vectors (the 1x1 Parameters are views into
 * them), a channel is a contiguous run of N*H*W elements because the sample index is the fastest
 * dimension, and each pass is two reductions plus one apply kernel (cattl3_batchnorm_forward /
 * cattl3_batchnorm_backward, include/cattl3_b200.h) instead of a slice copy per channel (:182-188).
 *
 * Data-parallel training (b200::Communicator with more than one rank; the reference is single process): the
 * batch statistics are SYNCHRONISED -- the per-group sums of the forward pass and of the backward pass are
 * all-reduced over the ranks together with the element count (2 * groups + 1 doubles per direction, no host
 * synchronisation), so that every rank normalises with the statistics of the whole mini-batch, the running
 * averages stay identical on all replicas and G GPUs compute what one GPU computes on the full batch
 * (SURVEY.md section 8e).  CATTL3_SYNC_BN=0 turns it off (each rank then uses its shard's statistics).
 */
#ifndef C_ATTL3_LAYER_BATCHNORMLAYER_H_
#define C_ATTL3_LAYER_BATCHNORMLAYER_H_

#include <cassert>
#include <cstdlib>
#include <memory>
#include <utility>
#include <vector>

#include "core/Layer.hpp"
#include "core/NumericUtils.hpp"
#include "parameter_initialization/OneParameterInitialization.hpp"
#include "parameter_initialization/ZeroParameterInitialization.hpp"
#include "parameters/B200Parameters.hpp"
#include "b200/Communicator.hpp"
#include "b200/DeviceLayer.hpp"

namespace cattle {

template<typename Scalar, std::size_t Rank, bool PerLastRank = (Rank == 3)>
class BatchNormLayer : public Layer<Scalar,Rank>, public b200::DeviceLayer<Scalar,Rank>,
		public b200::EpilogueConsumer<Scalar> {
	typedef Layer<Scalar,Rank> Base;
	typedef BatchNormLayer<Scalar,Rank,PerLastRank> Self;
	typedef b200::DeviceTensor<Scalar> DevTensor;
	typedef b200::ParameterStorage<Scalar> Storage;
	typedef std::shared_ptr<B200Parameters<Scalar>> DevParamsSharedPtr;
public:
	inline BatchNormLayer(const typename Base::Dims& dims, Scalar norm_avg_decay = .1,
			Scalar epsilon = NumericUtils<Scalar>::EPSILON2, ParamRegSharedPtr<Scalar> gamma_reg = nullptr,
			Scalar gamma_clip = 0, Scalar gamma_max_l1_norm = 0, Scalar gamma_max_l2_norm = 0,
			Scalar gamma_grad_clip = 0, Scalar gamma_grad_max_l1_norm = 0, Scalar gamma_grad_max_l2_norm = 0,
			ParamRegSharedPtr<Scalar> beta_reg = nullptr, Scalar beta_clip = 0, Scalar beta_max_l1_norm = 0,
			Scalar beta_max_l2_norm = 0, Scalar beta_grad_clip = 0, Scalar beta_grad_max_l1_norm = 0,
			Scalar beta_grad_max_l2_norm = 0) :
				owner(*this),
				dims(dims),
				norm_avg_decay(norm_avg_decay),
				epsilon(epsilon),
				groups(PerLastRank ? dims(Rank - 1) : dims.get_volume()),
				input_layer(false),
				avgs_init(false),
				cached_rows(0) {
		assert(norm_avg_decay >= 0 && norm_avg_decay <= 1 &&
				"norm avg decay must not be less than 0 or greater than 1");
		assert(epsilon > 0 && "epsilon must be greater than 0");
		auto gamma_init = std::make_shared<OneParameterInitialization<Scalar>>();
		auto beta_init = std::make_shared<ZeroParameterInitialization<Scalar>>();
		// four device vectors of `groups` elements (+ two gradients) behind all the Parameters objects
		auto mean_store = std::make_shared<Storage>(groups);
		auto inv_sd_store = std::make_shared<Storage>(groups);
		auto gamma_store = std::make_shared<Storage>(groups);
		auto gamma_grad_store = std::make_shared<Storage>(groups);
		auto beta_store = std::make_shared<Storage>(groups);
		auto beta_grad_store = std::make_shared<Storage>(groups);
		const std::size_t views = PerLastRank ? groups : 1;
		const std::size_t view_cols = PerLastRank ? 1 : groups;
		for (std::size_t i = 0; i < views; ++i) {
			avg_means.push_back(std::make_shared<B200Parameters<Scalar>>(mean_store, nullptr, i, 1, view_cols,
					false, nullptr, nullptr));
			avg_inv_sds.push_back(std::make_shared<B200Parameters<Scalar>>(inv_sd_store, nullptr, i, 1, view_cols,
					false, nullptr, nullptr));
			gammas.push_back(std::make_shared<B200Parameters<Scalar>>(gamma_store, gamma_grad_store, i, 1,
					view_cols, true, gamma_init, gamma_reg, gamma_clip, gamma_max_l1_norm, gamma_max_l2_norm,
					gamma_grad_clip, gamma_grad_max_l1_norm, gamma_grad_max_l2_norm));
			betas.push_back(std::make_shared<B200Parameters<Scalar>>(beta_store, beta_grad_store, i, 1,
					view_cols, true, beta_init, beta_reg, beta_clip, beta_max_l1_norm, beta_max_l2_norm,
					beta_grad_clip, beta_grad_max_l1_norm, beta_grad_max_l2_norm));
		}
	}
	inline BatchNormLayer(const Self& layer, bool share_params = false) :
			owner(share_params ? layer.owner : *this),
			dims(layer.dims),
			norm_avg_decay(layer.norm_avg_decay),
			epsilon(layer.epsilon),
			groups(layer.groups),
			input_layer(layer.input_layer),
			avgs_init(layer.avgs_init),
			cached_rows(layer.cached_rows),
			in_cache(layer.in_cache),
			batch_means(layer.batch_means),
			batch_inv_sds(layer.batch_inv_sds),
			global_count(layer.global_count) {
		if (share_params) {
			avg_means = layer.avg_means;
			avg_inv_sds = layer.avg_inv_sds;
			gammas = layer.gammas;
			betas = layer.betas;
		} else {
			// an independent copy: fresh contiguous device vectors holding the original's numbers
			auto mean_store = std::make_shared<Storage>(groups);
			auto inv_sd_store = std::make_shared<Storage>(groups);
			auto gamma_store = std::make_shared<Storage>(groups);
			auto gamma_grad_store = std::make_shared<Storage>(groups);
			auto beta_store = std::make_shared<Storage>(groups);
			auto beta_grad_store = std::make_shared<Storage>(groups);
			for (std::size_t i = 0; i < layer.gammas.size(); ++i) {
				avg_means.push_back(rebased(*layer.avg_means[i], mean_store, nullptr, i));
				avg_inv_sds.push_back(rebased(*layer.avg_inv_sds[i], inv_sd_store, nullptr, i));
				gammas.push_back(rebased(*layer.gammas[i], gamma_store, gamma_grad_store, i));
				betas.push_back(rebased(*layer.betas[i], beta_store, beta_grad_store, i));
			}
		}
	}
	inline Base* clone() const {
		return new BatchNormLayer(*this);
	}
	inline Base* clone_with_shared_params() {
		return new BatchNormLayer(*this, true);
	}
	inline const Base& get_params_owner() const {
		return owner;
	}
	inline const typename Base::Dims& get_input_dims() const {
		return dims;
	}
	inline const typename Base::Dims& get_output_dims() const {
		return dims;
	}
	inline bool is_input_layer() const {
		return input_layer;
	}
	inline void set_input_layer(bool input_layer) {
		this->input_layer = input_layer;
	}
	inline std::vector<const Parameters<Scalar>*> get_params() const {
		std::vector<const Parameters<Scalar>*> params_vec;
		for (std::size_t i = 0; i < gammas.size(); ++i) {
			params_vec.push_back(avg_means[i].get());
			params_vec.push_back(avg_inv_sds[i].get());
			params_vec.push_back(gammas[i].get());
			params_vec.push_back(betas[i].get());
		}
		return params_vec;
	}
	inline std::vector<Parameters<Scalar>*> get_params() {
		std::vector<Parameters<Scalar>*> params_vec;
		for (std::size_t i = 0; i < gammas.size(); ++i) {
			params_vec.push_back(avg_means[i].get());
			params_vec.push_back(avg_inv_sds[i].get());
			params_vec.push_back(gammas[i].get());
			params_vec.push_back(betas[i].get());
		}
		return params_vec;
	}
	inline void empty_cache() {
		in_cache = DevTensor();
		cached_rows = 0;
	}
	inline typename Base::Data pass_forward(typename Base::Data in, bool training) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(in.dimensions()).template demote<>()) == dims);
		assert(in.dimension(0) > 0);
		DevTensor out = pass_forward_dev(b200::to_device<Scalar,Base::DATA_RANK>(in), training);
		return b200::to_host<Scalar,Base::DATA_RANK>(out, b200::batch_extents<Rank>(out.rows, dims));
	}
	inline typename Base::Data pass_back(typename Base::Data out_grad) {
		assert((Dimensions<std::size_t,Base::DATA_RANK>(out_grad.dimensions()).template demote<>()) == dims);
		assert(out_grad.dimension(0) > 0 && cached_rows == (std::size_t) out_grad.dimension(0));
		DevTensor prev_out_grad = pass_back_dev(b200::to_device<Scalar,Base::DATA_RANK>(out_grad));
		if (prev_out_grad.empty())
			return typename Base::Data();
		return b200::to_host<Scalar,Base::DATA_RANK>(prev_out_grad, b200::batch_extents<Rank>(prev_out_grad.rows, dims));
	}
	inline DevTensor pass_forward_dev(DevTensor in, bool training) {
		if (training && synchronised()) {
			// sums first (shifted by the running mean, which is identical on all ranks), then the shared path
			b200::FusedEpilogue<Scalar> ep;
			ep.col_stats = std::make_shared<b200::DeviceBuffer<double>>(2 * groups + 1);
			ep.shift = avg_means[0]->device_values();
			b200::Context& c = b200::Context::get();
			{
				b200::Context::Lock l = c.lock();
				CATTLE_B200_CHECK(b200::Api<Scalar>::batchnorm_stats(c.handle(), PerLastRank ? 1 : 0,
						(std::int32_t) in.rows, geom_h(), 1, geom_c(), in.data(), ep.shift, ep.col_stats->data()));
			}
			return accept_epilogue(std::move(in), ep, training, nullptr);
		}
		DevTensor out(in.rows, dims.get_volume());
		if (training && batch_means.size() != groups) {
			batch_means = b200::DeviceBuffer<Scalar>(groups);
			batch_inv_sds = b200::DeviceBuffer<Scalar>(groups);
		}
		if (!training && !avgs_init)
			throw b200::Error(CATTL3_ERR_INVALID, "BatchNormLayer: inference before any training batch");
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::batchnorm_forward(c.handle(), PerLastRank ? 1 : 0,
					(std::int32_t) in.rows, geom_h(), 1, geom_c(), training ? 1 : 0, avgs_init ? 1 : 0,
					norm_avg_decay, epsilon, in.data(), gammas[0]->device_values(), betas[0]->device_values(),
					avg_means[0]->device_values(), avg_inv_sds[0]->device_values(), batch_means.data(),
					batch_inv_sds.data(), out.data()));
		}
		if (training) {
			// the running statistics changed on the device: every view's host mirror is now stale
			avg_means[0]->values_written_on_device();
			avg_inv_sds[0]->values_written_on_device();
			avgs_init = true;
			cached_rows = in.rows;
			in_cache = std::move(in);
			global_count = b200::DeviceBuffer<double>();
		}
		return out;
	}
	/**
	 * In training mode the first of the layer's two passes over its input -- the per-group mean and variance
	 * (BatchNormLayer.hpp:225-233) -- can come out of the producing kernel layer's epilogue, when that layer's
	 * GEMM columns are this layer's groups (conv filters = channels; dense outputs = activations).
	 */
	inline bool request_epilogue(b200::FusedEpilogue<Scalar>& ep, std::size_t producer_stat_columns, bool training) const {
		if (!training || producer_stat_columns != groups)
			return false;
		ep.want_stats = true;
		ep.keep_pre = true;
		return true;
	}
	inline bool chains_epilogue() const {
		return true;
	}
	inline DevTensor accept_epilogue(DevTensor in, b200::FusedEpilogue<Scalar>& ep, bool training,
			b200::FusedEpilogue<Scalar>* next) {
		if (!training || in.empty() || !ep.col_stats || ep.col_stats->size() != 2 * groups + 1 || !ep.shift)
			throw b200::Error(CATTL3_ERR_INVALID, "BatchNormLayer: incomplete fused epilogue");
		const double* count_ptr = nullptr;
		if (synchronised()) {
			// [sums | element count] summed over the ranks in one message; the count stays on the device
			// (written by a kernel: a host -> device copy from pageable memory would synchronise with the stream)
			const double local_count = (double) in.rows * (double) geom_h();
			b200::Context& c = b200::Context::get();
			{
				b200::Context::Lock l = c.lock();
				CATTLE_B200_CHECK(b200::Api<double>::fill(c.handle(), 1, local_count, ep.col_stats->data() + 2 * groups));
			}
			b200::Communicator::get().all_reduce_sum(ep.col_stats->data(), 2 * groups + 1);
			global_count = b200::DeviceBuffer<double>(1);
			{
				b200::Context::Lock l = c.lock();
				CATTLE_B200_CHECK(cattl3_memcpy_d2d(c.handle(), global_count.data(), ep.col_stats->data() + 2 * groups,
						sizeof(double)));
			}
			count_ptr = global_count.data();
		} else {
			global_count = b200::DeviceBuffer<double>();
		}
		const bool act = next != nullptr && next->act_kind != CATTL3_ACT_NONE;
		DevTensor out;
		if (!act || next->keep_pre)
			out = DevTensor(in.rows, dims.get_volume());
		if (act)
			next->act_out = DevTensor(in.rows, dims.get_volume());
		if (batch_means.size() != groups) {
			batch_means = b200::DeviceBuffer<Scalar>(groups);
			batch_inv_sds = b200::DeviceBuffer<Scalar>(groups);
		}
		b200::Context& c = b200::Context::get();
		{
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::batchnorm_forward_stats(c.handle(), PerLastRank ? 1 : 0,
					(std::int32_t) in.rows, geom_h(), 1, geom_c(), avgs_init ? 1 : 0, norm_avg_decay, epsilon, in.data(),
					ep.col_stats->data(), count_ptr, ep.shift, gammas[0]->device_values(), betas[0]->device_values(),
					avg_means[0]->device_values(), avg_inv_sds[0]->device_values(), batch_means.data(),
					batch_inv_sds.data(), out.data(), act ? next->act_kind : CATTL3_ACT_NONE,
					act ? next->act_param : (Scalar) 0, act ? next->act_out.data() : nullptr));
		}
		avg_means[0]->values_written_on_device();
		avg_inv_sds[0]->values_written_on_device();
		avgs_init = true;
		cached_rows = in.rows;
		in_cache = std::move(in);
		return out;
	}
	inline DevTensor pass_back_dev(DevTensor out_grad) {
		if (in_cache.empty() || cached_rows != out_grad.rows)
			throw b200::Error(CATTL3_ERR_INVALID, "BatchNormLayer: pass_back without a matching training pass_forward");
		DevTensor prev_out_grad;
		if (!input_layer)
			prev_out_grad = DevTensor(out_grad.rows, dims.get_volume());
		b200::Context& c = b200::Context::get();
		if (!global_count.empty()) {
			// synchronised statistics: the two sums of the backward pass are all-reduced as well; dgamma / dbeta get
			// the local sums (the optimizer's gradient all-reduce adds the ranks' shares)
			b200::DeviceBuffer<double> sums(2 * groups);
			{
				b200::Context::Lock l = c.lock();
				CATTLE_B200_CHECK(b200::Api<Scalar>::batchnorm_backward_sums(c.handle(), PerLastRank ? 1 : 0,
						(std::int32_t) out_grad.rows, geom_h(), 1, geom_c(), in_cache.data(), batch_means.data(),
						batch_inv_sds.data(), out_grad.data(), gammas[0]->device_grad(), betas[0]->device_grad(), sums.data()));
			}
			if (!input_layer) {
				b200::Communicator::get().all_reduce_sum(sums.data(), 2 * groups);
				b200::Context::Lock l = c.lock();
				CATTLE_B200_CHECK(b200::Api<Scalar>::batchnorm_backward_apply(c.handle(), PerLastRank ? 1 : 0,
						(std::int32_t) out_grad.rows, geom_h(), 1, geom_c(), global_count.data(), in_cache.data(),
						gammas[0]->device_values(), batch_means.data(), batch_inv_sds.data(), out_grad.data(), sums.data(),
						prev_out_grad.data()));
			}
		} else {
			b200::Context::Lock l = c.lock();
			CATTLE_B200_CHECK(b200::Api<Scalar>::batchnorm_backward(c.handle(), PerLastRank ? 1 : 0,
					(std::int32_t) out_grad.rows, geom_h(), 1, geom_c(), in_cache.data(), gammas[0]->device_values(),
					batch_means.data(), batch_inv_sds.data(), out_grad.data(), gammas[0]->device_grad(),
					betas[0]->device_grad(), prev_out_grad.data()));
		}
		for (std::size_t i = 0; i < gammas.size(); ++i) {
			gammas[i]->grad_written_on_device();
			betas[i]->grad_written_on_device();
		}
		return prev_out_grad;
	}
private:
	/**
	 * The kernels see the batch as (n, h, 1, c): per last rank, c = the extent of the last rank and
	 * h = everything between the sample index and it; per activation, every element of an observation
	 * is its own group.
	 */
	inline std::int32_t geom_c() const {
		return (std::int32_t) groups;
	}
	inline std::int32_t geom_h() const {
		return (std::int32_t) (dims.get_volume() / groups);
	}
	/** Whether the batch statistics span all ranks of a data-parallel job (see the file comment). */
	inline static bool synchronised() {
		static const bool enabled = [] {
			const char* v = std::getenv("CATTL3_SYNC_BN");
			return !(v && v[0] == '0');
		}();
		return enabled && b200::Communicator::get().world_size() > 1;
	}
	inline static DevParamsSharedPtr rebased(const B200Parameters<Scalar>& original,
			std::shared_ptr<Storage> value_store, std::shared_ptr<Storage> grad_store, std::size_t index) {
		// clone() carries the hyper-parameters (initialisation, regularisation, constraints, frozen flag)
		// and the numbers; the copy is then re-seated as a view of the new shared vectors
		std::unique_ptr<B200Parameters<Scalar>> copy(static_cast<B200Parameters<Scalar>*>(original.clone()));
		return copy->as_view_of(value_store, grad_store, index * original.count());
	}
	const Self& owner;
	const typename Base::Dims dims;
	const Scalar norm_avg_decay, epsilon;
	const std::size_t groups;
	bool input_layer;
	// Dynamic batch normalization parameters and the optimizable parameters (views, see above).
	std::vector<DevParamsSharedPtr> avg_means, avg_inv_sds, gammas, betas;
	bool avgs_init;
	// Staged computation caches: the input and the statistics of the last training batch.
	std::size_t cached_rows;
	DevTensor in_cache;
	b200::DeviceBuffer<Scalar> batch_means, batch_inv_sds;
	// elements per group over all ranks in the last training pass (device scalar); empty = local statistics
	b200::DeviceBuffer<double> global_count;
};

} /* namespace cattle */

#endif /* C_ATTL3_LAYER_BATCHNORMLAYER_H_ */
