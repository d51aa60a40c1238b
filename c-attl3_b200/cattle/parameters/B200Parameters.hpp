/*
 * parameters/B200Parameters.hpp -- device-resident implementation of cattle::Parameters.
 *
 * The interface is the reference's (C-ATTL3/core/Parameters.hpp:16-94) and the constructor
 * arguments and constraint semantics are those of StandardParameters
 * (C-ATTL3/parameters/StandardParameters.hpp:61-136), so optimizers, GradientCheck
 * (C-ATTL3/core/GradientCheck.hpp:64-117) and the (de)serialisation in
 * C-ATTL3/core/NeuralNetwork.hpp:195-222 work unchanged.  What differs is where the numbers live:
 * the master copy of values and gradient is in HBM; the host matrices returned by get_values() /
 * get_grad() are mirrors refreshed lazily (one D2H per storage per change), the scheme the reference
 * sketched in C-ATTL3/parameters/gpu/StandardGPUParameters.hpp:91-116.
 *
 * Several Parameters objects may be views into one shared storage: BatchNormLayer exposes four
 * 1x1 Parameters per channel (C-ATTL3/layer/BatchNormLayer.hpp:89-99,263-272) that are backed by
 * four per-layer arrays here, so the kernels see contiguous gamma / beta / running-statistics vectors.
 */
#ifndef C_ATTL3_PARAMETERS_B200PARAMETERS_H_
#define C_ATTL3_PARAMETERS_B200PARAMETERS_H_

#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

#include "core/NumericUtils.hpp"
#include "core/ParameterInitialization.hpp"
#include "core/ParameterRegularization.hpp"
#include "core/Parameters.hpp"
#include "parameter_regularization/ElasticNetParameterRegularization.hpp"
#include "parameter_regularization/L1ParameterRegularization.hpp"
#include "parameter_regularization/L2ParameterRegularization.hpp"
#include "b200/Runtime.hpp"

namespace cattle {

template<typename Scalar>
using ParamInitSharedPtr = std::shared_ptr<ParameterInitialization<Scalar>>;

template<typename Scalar>
using ParamRegSharedPtr = std::shared_ptr<ParameterRegularization<Scalar>>;

namespace b200 {

/** A device array with a lazily refreshed host mirror, shared by the Parameters viewing it. */
template<typename Scalar>
class ParameterStorage {
public:
	inline explicit ParameterStorage(std::size_t count) :
			dev(count, true),
			host(count, (Scalar) 0),
			dev_version(1),
			host_version(1) { }
	inline std::size_t size() const {
		return host.size();
	}
	inline Scalar* device_data() {
		return dev.data();
	}
	/** The host mirror, refreshed if a kernel has written the device copy since the last look. */
	inline const Scalar* host_data() {
		if (host_version != dev_version) {
			dev.download(host.data(), host.size());
			host_version = dev_version;
		}
		return host.data();
	}
	/** To be called after a kernel wrote the device copy. */
	inline void device_written() {
		++dev_version;
	}
	inline std::uint64_t version() const {
		return dev_version;
	}
	/** Host-side write of [offset, offset + n): mirror and device copy stay coherent. */
	inline void write(std::size_t offset, std::size_t n, const Scalar* src) {
		host_data();
		std::memcpy(host.data() + offset, src, n * sizeof(Scalar));
		dev.upload(src, n, offset);
		++dev_version;
		host_version = dev_version;
	}
	/**
	 * Moves the device copy to [arena_offset, arena_offset + size()) of `arena` (contents preserved): the optimizer
	 * packs the storages of a network next to each other so that one fused kernel updates, and one message
	 * all-reduces, all of them.  The storage keeps the arena alive.
	 */
	inline void relocate(std::shared_ptr<DeviceBuffer<Scalar>> new_arena, std::size_t arena_offset) {
		if (host.empty())
			return;
		Scalar* dst = new_arena->data() + arena_offset;
		Context& c = Context::get();
		{
			Context::Lock l = c.lock();
			CATTLE_B200_CHECK(cattl3_memcpy_d2d(c.handle(), dst, dev.data(), host.size() * sizeof(Scalar)));
		}
		dev = DeviceBuffer<Scalar>::view(dst, host.size());  // releases the private array (stream ordered)
		arena = std::move(new_arena);
	}
	inline void zero(std::size_t offset, std::size_t n) {
		if (offset == 0 && n == host.size()) {
			dev.zero();
			std::fill(host.begin(), host.end(), (Scalar) 0);
			++dev_version;
			host_version = dev_version;
		} else {
			std::vector<Scalar> z(n, (Scalar) 0);
			write(offset, n, z.data());
		}
	}
private:
	std::shared_ptr<DeviceBuffer<Scalar>> arena;  // set once relocated; declared first: outlives the view in `dev`
	DeviceBuffer<Scalar> dev;
	std::vector<Scalar> host;
	std::uint64_t dev_version, host_version;
};

} /* namespace b200 */

template<typename Scalar>
class B200Parameters : public Parameters<Scalar> {
	typedef b200::ParameterStorage<Scalar> Storage;
	typedef std::shared_ptr<Storage> StorageSharedPtr;
public:
	/** Same arguments, same meaning as StandardParameters (StandardParameters.hpp:61-77). */
	inline B200Parameters(std::size_t rows, std::size_t cols, bool optimizable = true,
			ParamInitSharedPtr<Scalar> init = nullptr, ParamRegSharedPtr<Scalar> reg = nullptr,
			Scalar value_clip = 0, Scalar value_max_l1_norm = 0, Scalar value_max_l2_norm = 0,
			Scalar grad_clip = 0, Scalar grad_max_l1_norm = 0, Scalar grad_max_l2_norm = 0) :
				B200Parameters(nullptr, nullptr, 0, rows, cols, optimizable, init, reg, value_clip,
						value_max_l1_norm, value_max_l2_norm, grad_clip, grad_max_l1_norm, grad_max_l2_norm) { }
	/**
	 * A view of `rows * cols` elements at `offset` of shared storages (null storages: allocate
	 * private ones).  `grad_store` is ignored for non-optimizable parameters.
	 */
	inline B200Parameters(StorageSharedPtr value_store, StorageSharedPtr grad_store, std::size_t offset,
			std::size_t rows, std::size_t cols, bool optimizable, ParamInitSharedPtr<Scalar> init,
			ParamRegSharedPtr<Scalar> reg, Scalar value_clip = 0, Scalar value_max_l1_norm = 0,
			Scalar value_max_l2_norm = 0, Scalar grad_clip = 0, Scalar grad_max_l1_norm = 0,
			Scalar grad_max_l2_norm = 0) :
				rows(rows),
				cols(cols),
				optimizable(optimizable),
				param_init(init),
				param_reg(reg),
				value_clip(value_clip),
				value_max_l1_norm(value_max_l1_norm),
				value_max_l2_norm(value_max_l2_norm),
				grad_clip(grad_clip),
				grad_max_l1_norm(grad_max_l1_norm),
				grad_max_l2_norm(grad_max_l2_norm),
				l1_lambda(probe_lambdas(reg).first),
				l2_lambda(probe_lambdas(reg).second),
				value_store(value_store ? value_store : std::make_shared<Storage>(rows * cols)),
				grad_store(!optimizable ? nullptr : (grad_store ? grad_store : std::make_shared<Storage>(rows * cols))),
				offset(offset),
				values_seen(0),
				grad_seen(0),
				grad_known_zero(true),
				frozen(false) {
		assert(rows > 0 && cols > 0);
		assert(this->value_store->size() >= offset + rows * cols);
	}
	/** Deep copy: the clone owns private storages holding this view's current numbers. */
	inline B200Parameters(const B200Parameters<Scalar>& other) :
			rows(other.rows),
			cols(other.cols),
			optimizable(other.optimizable),
			param_init(other.param_init),
			param_reg(other.param_reg),
			value_clip(other.value_clip),
			value_max_l1_norm(other.value_max_l1_norm),
			value_max_l2_norm(other.value_max_l2_norm),
			grad_clip(other.grad_clip),
			grad_max_l1_norm(other.grad_max_l1_norm),
			grad_max_l2_norm(other.grad_max_l2_norm),
			l1_lambda(other.l1_lambda),
			l2_lambda(other.l2_lambda),
			value_store(std::make_shared<Storage>(other.rows * other.cols)),
			grad_store(other.optimizable ? std::make_shared<Storage>(other.rows * other.cols) : nullptr),
			offset(0),
			values_seen(0),
			grad_seen(0),
			grad_known_zero(other.grad_known_zero),
			frozen(other.frozen) {
		value_store->write(0, count(), other.value_store->host_data() + other.offset);
		if (optimizable && !grad_known_zero)
			grad_store->write(0, count(), other.grad_store->host_data() + other.offset);
	}
	inline Parameters<Scalar>* clone() const {
		return new B200Parameters<Scalar>(*this);
	}
	/**
	 * The same parameters (numbers, initialisation, regularisation, constraints, frozen flag)
	 * re-seated as a view at `offset` of the given shared storages.
	 */
	inline std::shared_ptr<B200Parameters<Scalar>> as_view_of(StorageSharedPtr new_value_store,
			StorageSharedPtr new_grad_store, std::size_t new_offset) const {
		auto view = std::make_shared<B200Parameters<Scalar>>(new_value_store, new_grad_store, new_offset, rows,
				cols, optimizable, param_init, param_reg, value_clip, value_max_l1_norm, value_max_l2_norm,
				grad_clip, grad_max_l1_norm, grad_max_l2_norm);
		view->value_store->write(new_offset, count(), value_store->host_data() + offset);
		if (optimizable && !grad_known_zero) {
			view->grad_store->write(new_offset, count(), grad_store->host_data() + offset);
			view->grad_known_zero = false;
		}
		view->frozen = frozen;
		return view;
	}
	/** The shared storages behind this view (the optimizer packs them into one arena). */
	inline const StorageSharedPtr& value_storage() const {
		return value_store;
	}
	inline const StorageSharedPtr& grad_storage() const {
		return grad_store;
	}
	inline bool are_optimizable() const {
		return optimizable;
	}
	inline std::size_t get_rows() const {
		return rows;
	}
	inline std::size_t get_cols() const {
		return cols;
	}
	inline void init_values() {
		Matrix<Scalar> fresh = Matrix<Scalar>::Zero(rows, cols);
		if (param_init)
			param_init->apply(fresh);  // host-side, once; then uploaded (init stays on the host)
		value_store->write(offset, count(), fresh.data());
	}
	inline void init_grad() {
		if (optimizable)
			reset_grad();
	}
	inline const Matrix<Scalar>& get_values() const {
		if (values_seen != value_store->version()) {
			const Scalar* src = value_store->host_data() + offset;
			values_host = MatrixMap<Scalar>(const_cast<Scalar*>(src), rows, cols);
			values_seen = value_store->version();
		}
		return values_host;
	}
	inline void set_values(Matrix<Scalar> values) {
		assert((std::size_t) values.rows() == rows && (std::size_t) values.cols() == cols);
		enforce_constraints(values, value_clip, value_max_l1_norm, value_max_l2_norm);
		value_store->write(offset, count(), values.data());
	}
	inline const Matrix<Scalar>& get_grad() const {
		if (!optimizable)
			return grad_host;  // empty, like StandardParameters' never-initialised gradient
		if (grad_seen != grad_store->version()) {
			const Scalar* src = grad_store->host_data() + offset;
			grad_host = MatrixMap<Scalar>(const_cast<Scalar*>(src), rows, cols);
			grad_seen = grad_store->version();
		}
		return grad_host;
	}
	/** Host-matrix accumulation (the Parameters API); kernels accumulate through device_grad(). */
	inline void accumulate_grad(const Matrix<Scalar>& grad) {
		if (!optimizable)
			return;
		assert((std::size_t) grad.rows() == rows && (std::size_t) grad.cols() == cols);
		Matrix<Scalar> sum = get_grad() + grad;
		enforce_constraints(sum, grad_clip, grad_max_l1_norm, grad_max_l2_norm);
		grad_store->write(offset, count(), sum.data());
		grad_known_zero = false;
	}
	inline void reset_grad() {
		if (!optimizable || grad_known_zero)
			return;
		grad_store->zero(offset, count());
		grad_known_zero = true;
	}
	inline Scalar get_regularization_penalty() const {
		if (optimizable && param_reg)
			return param_reg->function(get_values());
		return 0;
	}
	inline void regularize() {
		if (!optimizable || !param_reg)
			return;
		if (has_device_regularization())
			regularize_dev(nullptr);
		else
			accumulate_grad(param_reg->d_function(get_values()));
	}
	/**
	 * Whether the regularisation runs on the device: an L1, L2 or ElasticNet penalty (their derivative and value are
	 * one kernel, cattl3_regularize; gradient constraints are re-applied behind it on the device as accumulate_grad
	 * does, StandardParameters.hpp:115-123).
	 */
	inline bool has_device_regularization() const {
		return optimizable && param_reg && (l1_lambda > 0 || l2_lambda > 0);
	}
	/**
	 * regularize() and get_regularization_penalty() in one pass on the device: grad += d penalty / d values and, if
	 * `penalty` is given, *penalty += the penalty (a device double: the batch loop reads it once per epoch).
	 */
	inline void regularize_dev(double* penalty) {
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(b200::Api<Scalar>::regularize(c.handle(), (std::int64_t) count(), l1_lambda, l2_lambda,
				device_values(), device_grad(), penalty));
		grad_written_on_device();
	}
	inline bool are_frozen() const {
		return frozen;
	}
	inline void set_frozen(bool frozen) {
		this->frozen = frozen;
	}
	// ---- device side ---------------------------------------------------------------------------
	inline std::size_t count() const {
		return rows * cols;
	}
	inline const Scalar* device_values() const {
		return value_store->device_data() + offset;
	}
	inline Scalar* device_values() {
		return value_store->device_data() + offset;
	}
	/** Null for non-optimizable parameters. */
	inline Scalar* device_grad() {
		return optimizable ? grad_store->device_data() + offset : nullptr;
	}
	/** After a kernel updated the values (optimizer step, running statistics). */
	inline void values_written_on_device() {
		value_store->device_written();
		if (has_value_constraints())
			constrain_dev(device_values(), value_clip, value_max_l1_norm, value_max_l2_norm);
	}
	/**
	 * Whoever wants to know that a layer's backward pass has enqueued its gradient (the data-parallel batch loop sends
	 * finished stretches of the gradient arena on their way while the layers behind still compute).
	 */
	struct GradientListener {
		virtual ~GradientListener() = default;
		virtual void gradient_written(Scalar* dev_grad, std::size_t count) = 0;
	};
	inline static GradientListener*& gradient_listener() {
		static GradientListener* listener = nullptr;
		return listener;
	}
	/** After a kernel accumulated into the gradient (beta = 1, StandardParameters.hpp:115-123). */
	inline void grad_written_on_device() {
		if (!optimizable)
			return;
		grad_store->device_written();
		grad_known_zero = false;
		if (gradient_listener())
			gradient_listener()->gradient_written(device_grad(), count());
		if (has_grad_constraints())
			constrain_dev(device_grad(), grad_clip, grad_max_l1_norm, grad_max_l2_norm);
	}
	/** After a fused optimizer step that also cleared the gradient (SGDOptimizer.hpp:69-70). */
	inline void grad_zeroed_on_device() {
		if (!optimizable)
			return;
		grad_store->device_written();
		grad_known_zero = true;
	}
	/** Whether a regularisation is attached (its penalty and derivative are evaluated per step). */
	inline bool has_regularization() const {
		return (bool) param_reg;
	}
	/** A regularisation that is evaluated on the host (any other than L1 / L2 / ElasticNet, or one next to gradient constraints). */
	inline bool has_host_regularization() const {
		return optimizable && param_reg && !has_device_regularization();
	}
	inline bool has_value_constraints() const {
		return active(value_clip) || active(value_max_l1_norm) || active(value_max_l2_norm);
	}
	inline bool has_grad_constraints() const {
		return active(grad_clip) || active(grad_max_l1_norm) || active(grad_max_l2_norm);
	}
	/** lambda if the regularisation is an L2ParameterRegularization, else 0. */
	inline Scalar get_l2_lambda() const {
		return l2_lambda;
	}
	inline bool has_non_l2_regularization() const {
		return param_reg && !(l2_lambda > 0 && !(l1_lambda > 0));
	}
private:
	inline static bool active(Scalar limit) {
		return NumericUtils<Scalar>::decidedly_greater(limit, (Scalar) 0);
	}
	/**
	 * The constraints on a device array, in place (cattl3_constrain: the device form of enforce_constraints below) --
	 * a fused optimizer step or a layer's backward pass is followed by this instead of a round trip through the host.
	 */
	inline void constrain_dev(Scalar* dev, Scalar clip, Scalar max_l1, Scalar max_l2) const {
		b200::Context& c = b200::Context::get();
		b200::Context::Lock l = c.lock();
		CATTLE_B200_CHECK(b200::Api<Scalar>::constrain(c.handle(), (std::int64_t) count(), active(clip) ? clip : (Scalar) 0,
				active(max_l1) ? max_l1 : (Scalar) 0, active(max_l2) ? max_l2 : (Scalar) 0, dev));
	}
	/**
	 * The three constraints in the reference's order and with its definitions
	 * (StandardParameters.hpp:150-182): element clip to [-clip, clip]; the "L1" limit is compared
	 * with the Frobenius norm and rescales by max / norm; the "L2" limit is compared with the
	 * *squared* norm and rescales by max / squared-norm.
	 */
	inline static void enforce_constraints(Matrix<Scalar>& m, Scalar clip, Scalar max_l1, Scalar max_l2) {
		if (active(clip))
			m = m.cwiseMax(-clip).cwiseMin(clip);
		if (active(max_l1)) {
			const Scalar norm = m.norm();
			if (norm > max_l1)
				m *= (max_l1 / norm);
		}
		if (active(max_l2)) {
			const Scalar sq_norm = m.squaredNorm();
			if (sq_norm > max_l2)
				m *= (max_l2 / sq_norm);
		}
	}
	/** L2ParameterRegularization keeps lambda private; d_function([1]) = [lambda] reveals it. */
	/**
	 * (l1 lambda, l2 lambda) of an L1 / L2 / ElasticNet regularisation, read off its derivative (the classes keep their
	 * constants private): d(1) = l1 + l2 and d(2) = l1 + 2 l2.  (0, 0) for anything else.
	 */
	inline static std::pair<Scalar,Scalar> probe_lambdas(const ParamRegSharedPtr<Scalar>& reg) {
		if (!reg)
			return std::make_pair((Scalar) 0, (Scalar) 0);
		const Scalar d1 = reg->d_function(Matrix<Scalar>::Ones(1, 1))(0, 0);
		if (dynamic_cast<const L2ParameterRegularization<Scalar>*>(reg.get()))
			return std::make_pair((Scalar) 0, d1);
		if (dynamic_cast<const L1ParameterRegularization<Scalar>*>(reg.get()))
			return std::make_pair(d1, (Scalar) 0);
		if (dynamic_cast<const ElasticNetParameterRegularization<Scalar>*>(reg.get())) {
			const Scalar d2 = reg->d_function(Matrix<Scalar>::Constant(1, 1, (Scalar) 2))(0, 0);
			return std::make_pair(2 * d1 - d2, d2 - d1);
		}
		return std::make_pair((Scalar) 0, (Scalar) 0);
	}
	const std::size_t rows, cols;
	const bool optimizable;
	const ParamInitSharedPtr<Scalar> param_init;
	const ParamRegSharedPtr<Scalar> param_reg;
	const Scalar value_clip, value_max_l1_norm, value_max_l2_norm;
	const Scalar grad_clip, grad_max_l1_norm, grad_max_l2_norm;
	const Scalar l1_lambda, l2_lambda;
	StorageSharedPtr value_store, grad_store;
	const std::size_t offset;
	mutable Matrix<Scalar> values_host, grad_host;
	mutable std::uint64_t values_seen, grad_seen;
	bool grad_known_zero;
	bool frozen;
};

} /* namespace cattle */

#endif /* C_ATTL3_PARAMETERS_B200PARAMETERS_H_ */
