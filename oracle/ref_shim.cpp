/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A thin extern "C" wrapper around the UNMODIFIED reference library (C-ATTL3, header-only C++ on
 * Eigen) so that Python tests and bench.py can run the reference's own CPU implementation of the
 * hot path on identical inputs.  It is compiled from the headers where they lie under
 * /root/reference (see oracle/Makefile) into oracle/_ref/libcattle_ref.so; no reference source is
 * copied into this repository.  Only tests/, __graft_entry__.smoke() and bench.py's reference /
 * cpu_baseline legs may load the resulting library.
 *
 * Every entry point builds the corresponding reference layer (rank-3 form, i.e. rank-4 batch
 * tensors, Eigen column-major => N fastest), injects the caller's parameters through
 * Parameters::set_values, runs pass_forward / pass_back and copies results out.
 *   conv      -> cattle::ConvKernelLayer<S,3>       (C-ATTL3/layer/kernel/ConvKernelLayer.hpp:115-189)
 *   transconv -> cattle::TransConvKernelLayer<S,3>  (C-ATTL3/layer/kernel/TransConvKernelLayer.hpp:115-187)
 *   dense     -> cattle::DenseKernelLayer<S,1>      (C-ATTL3/layer/kernel/DenseKernelLayer.hpp:92-115)
 *   act       -> ReLU/LeakyReLU/ELU/Swish layers    (C-ATTL3/layer/activation/ *.hpp)
 *   pool      -> Max/MeanPoolLayer<S,3>             (C-ATTL3/layer/PoolLayer.hpp:77-116)
 *   batchnorm -> BatchNormLayer<S,3,true|false>     (C-ATTL3/layer/BatchNormLayer.hpp:170-262, 337-391)
 *   optimizer -> every SGDOptimizer subclass        (C-ATTL3/optimizer/ *.hpp)
 *   train     -> FeedforwardNeuralNetwork + NadamOptimizer::train (config 1 of BASELINE.json),
 *                the StackedNeuralNetwork auto-encoder of config 3 and a ResidualNeuralNetwork of
 *                conv + BatchNorm + ReLU modules (config 4)
 */
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <memory>
#include <vector>

#include "Cattle.hpp"

using namespace cattle;

namespace {

struct Geom {  // must match cattl3_conv_geom in include/cattl3_b200.h
	int n, h, w, c, f, rh, rw, ph, pw, sh, sw, dh, dw;
};

inline double now_ms() {
	return std::chrono::duration<double, std::milli>(
			std::chrono::steady_clock::now().time_since_epoch()).count();
}

template<typename S>
void set_param(Parameters<S>* p, const S* src) {
	p->init();
	Matrix<S> m = Eigen::Map<const Matrix<S>>(src, p->get_rows(), p->get_cols());
	p->set_values(std::move(m));
}

template<typename S>
void get_grad(const Parameters<S>* p, S* dst) {
	if (dst)
		std::memcpy(dst, p->get_grad().data(), sizeof(S) * p->get_rows() * p->get_cols());
}

template<typename S>
Tensor<S,4> make4(const S* src, std::size_t n, std::size_t h, std::size_t w, std::size_t c) {
	Tensor<S,4> t(n, h, w, c);
	std::memcpy(t.data(), src, sizeof(S) * t.size());
	return t;
}

// Runs forward (+ optionally `reps` backward passes, to expose gradient accumulation) on any
// rank-3 kernel layer.
template<typename S, typename LayerT>
int run_kernel_layer(LayerT& layer, const Geom& g, int ih, int iw, int ic, int oh, int ow, int oc,
		const S* x, const S* w, const S* b, const S* dy, S* y, S* dx, S* dw, S* db, int back_reps,
		double* times_ms) {
	auto params = layer.get_params();
	set_param<S>(params[0], w);
	set_param<S>(params[1], b);
	layer.set_input_layer(dx == nullptr);
	double t0 = now_ms();
	Tensor<S,4> out = layer.pass_forward(make4<S>(x, g.n, ih, iw, ic), true);
	double t1 = now_ms();
	if ((int) out.dimension(1) != oh || (int) out.dimension(2) != ow || (int) out.dimension(3) != oc)
		return -2;
	if (y)
		std::memcpy(y, out.data(), sizeof(S) * out.size());
	double t2 = t1, t3 = t1;
	if (dy) {
		Tensor<S,4> prev;
		t2 = now_ms();
		for (int r = 0; r < back_reps; ++r)
			prev = layer.pass_back(make4<S>(dy, g.n, oh, ow, oc));
		t3 = now_ms();
		if (dx)
			std::memcpy(dx, prev.data(), sizeof(S) * prev.size());
		get_grad<S>(params[0], dw);
		get_grad<S>(params[1], db);
	}
	if (times_ms) {
		times_ms[0] = t1 - t0;
		times_ms[1] = t3 - t2;
	}
	return 0;
}

template<typename S>
int conv_impl(const Geom* g, int transposed, const S* x, const S* w, const S* b, const S* dy, S* y,
		S* dx, S* dw, S* db, int back_reps, double* times_ms) {
	auto init = std::make_shared<ZeroParameterInitialization<S>>();
	Dimensions<std::size_t,3> in_dims({ (std::size_t) g->h, (std::size_t) g->w, (std::size_t) g->c });
	if (!transposed) {
		ConvKernelLayer<S,3> layer(in_dims, g->f, init, g->rh, g->rw, g->ph, g->pw, g->sh, g->sw,
				g->dh, g->dw);
		auto od = layer.get_output_dims();
		return run_kernel_layer<S>(layer, *g, g->h, g->w, g->c, od(0), od(1), od(2), x, w, b, dy, y,
				dx, dw, db, back_reps, times_ms);
	} else {
		TransConvKernelLayer<S,3> layer(in_dims, g->f, init, g->rh, g->rw, g->ph, g->pw, g->sh, g->sw,
				g->dh, g->dw);
		auto od = layer.get_output_dims();
		return run_kernel_layer<S>(layer, *g, g->h, g->w, g->c, od(0), od(1), od(2), x, w, b, dy, y,
				dx, dw, db, back_reps, times_ms);
	}
}

template<typename S>
int dense_impl(int n, int in, int out, const S* x, const S* w, const S* b, const S* dy, S* y, S* dx,
		S* dw, S* db, int back_reps, double* times_ms) {
	auto init = std::make_shared<ZeroParameterInitialization<S>>();
	DenseKernelLayer<S,1> layer(Dimensions<std::size_t,1>({ (std::size_t) in }), out, init);
	auto params = layer.get_params();
	set_param<S>(params[0], w);
	set_param<S>(params[1], b);
	layer.set_input_layer(dx == nullptr);
	Tensor<S,2> xin(n, in);
	std::memcpy(xin.data(), x, sizeof(S) * xin.size());
	double t0 = now_ms();
	Tensor<S,2> o = layer.pass_forward(std::move(xin), true);
	double t1 = now_ms();
	if (y)
		std::memcpy(y, o.data(), sizeof(S) * o.size());
	double t2 = t1, t3 = t1;
	if (dy) {
		Tensor<S,2> prev;
		t2 = now_ms();
		for (int r = 0; r < back_reps; ++r) {
			Tensor<S,2> g(n, out);
			std::memcpy(g.data(), dy, sizeof(S) * g.size());
			prev = layer.pass_back(std::move(g));
		}
		t3 = now_ms();
		if (dx)
			std::memcpy(dx, prev.data(), sizeof(S) * prev.size());
		get_grad<S>(params[0], dw);
		get_grad<S>(params[1], db);
	}
	if (times_ms) {
		times_ms[0] = t1 - t0;
		times_ms[1] = t3 - t2;
	}
	return 0;
}

// kind: 0 ReLU, 1 LeakyReLU, 2 ELU, 3 Swish (same numbering as CATTL3_ACT_* in the C ABI)
template<typename S>
int act_impl(int kind, S alpha, int n, int vol, const S* x, const S* dy, S* y, S* dx) {
	Dimensions<std::size_t,1> dims({ (std::size_t) vol });
	std::unique_ptr<Layer<S,1>> layer;
	switch (kind) {
		case 0: layer.reset(new ReLUActivationLayer<S,1>(dims)); break;
		case 1: layer.reset(new LeakyReLUActivationLayer<S,1>(dims, alpha)); break;
		case 2: layer.reset(new ELUActivationLayer<S,1>(dims, alpha)); break;
		case 3: layer.reset(new SwishActivationLayer<S,1>(dims, alpha)); break;
		case 4: layer.reset(new SigmoidActivationLayer<S,1>(dims)); break;
		case 5: layer.reset(new TanhActivationLayer<S,1>(dims)); break;
		case 6: layer.reset(new SoftplusActivationLayer<S,1>(dims)); break;
		case 7: layer.reset(new SoftmaxActivationLayer<S,1>(dims)); break;
		default: return -1;
	}
	Tensor<S,2> xin(n, vol);
	std::memcpy(xin.data(), x, sizeof(S) * xin.size());
	Tensor<S,2> o = layer->pass_forward(std::move(xin), true);
	if (y)
		std::memcpy(y, o.data(), sizeof(S) * o.size());
	if (dy && dx) {
		Tensor<S,2> g(n, vol);
		std::memcpy(g.data(), dy, sizeof(S) * g.size());
		Tensor<S,2> prev = layer->pass_back(std::move(g));
		std::memcpy(dx, prev.data(), sizeof(S) * prev.size());
	}
	return 0;
}

// kind: 0 max, 1 mean
template<typename S>
int pool_impl(int kind, int n, int h, int w, int c, int rh, int rw, int sh, int sw, const S* x,
		const S* dy, S* y, S* dx, double* times_ms) {
	Dimensions<std::size_t,3> dims({ (std::size_t) h, (std::size_t) w, (std::size_t) c });
	std::unique_ptr<Layer<S,3>> layer;
	if (kind == 0)
		layer.reset(new MaxPoolLayer<S,3>(dims, rh, rw, sh, sw));
	else
		layer.reset(new MeanPoolLayer<S,3>(dims, rh, rw, sh, sw));
	auto od = layer->get_output_dims();
	double t0 = now_ms();
	Tensor<S,4> o = layer->pass_forward(make4<S>(x, n, h, w, c), true);
	double t1 = now_ms();
	if (y)
		std::memcpy(y, o.data(), sizeof(S) * o.size());
	double t2 = t1, t3 = t1;
	if (dy && dx) {
		t2 = now_ms();
		Tensor<S,4> prev = layer->pass_back(make4<S>(dy, n, od(0), od(1), od(2)));
		t3 = now_ms();
		std::memcpy(dx, prev.data(), sizeof(S) * prev.size());
	}
	if (times_ms) {
		times_ms[0] = t1 - t0;
		times_ms[1] = t3 - t2;
	}
	return 0;
}

// Runs `steps` training forward passes on x[step] (each n*h*w*c), then one backward on dy for the
// last step, then an inference forward of x[last].  Per-channel variant exposes 4 1x1 parameters
// per channel; per-activation exposes 4 1x(h*w*c) parameters (BatchNormLayer.hpp:263-272, 326-333).
template<typename S, typename LayerT>
int bn_run(LayerT& layer, bool per_channel, int n, int h, int w, int c, int steps, const S* x,
		const S* gamma, const S* beta, const S* dy, S* y, S* dx, S* dgamma, S* dbeta, S* run_mean,
		S* run_inv_sd, S* y_infer) {
	auto params = layer.get_params();
	std::size_t groups = per_channel ? c : 1;
	std::size_t width = per_channel ? 1 : (std::size_t) h * w * c;
	for (auto p : params)
		p->init();
	for (std::size_t i = 0; i < groups; ++i) {
		set_param<S>(params[4 * i + 2], gamma + i * width);
		set_param<S>(params[4 * i + 3], beta + i * width);
	}
	layer.set_input_layer(dx == nullptr);
	std::size_t vol = (std::size_t) n * h * w * c;
	Tensor<S,4> out;
	for (int s = 0; s < steps; ++s)
		out = layer.pass_forward(make4<S>(x + s * vol, n, h, w, c), true);
	if (y)
		std::memcpy(y, out.data(), sizeof(S) * vol);
	if (dy) {
		Tensor<S,4> prev = layer.pass_back(make4<S>(dy, n, h, w, c));
		if (dx)
			std::memcpy(dx, prev.data(), sizeof(S) * vol);
		for (std::size_t i = 0; i < groups; ++i) {
			if (dgamma) std::memcpy(dgamma + i * width, params[4 * i + 2]->get_grad().data(), sizeof(S) * width);
			if (dbeta) std::memcpy(dbeta + i * width, params[4 * i + 3]->get_grad().data(), sizeof(S) * width);
		}
	}
	for (std::size_t i = 0; i < groups; ++i) {
		if (run_mean) std::memcpy(run_mean + i * width, params[4 * i]->get_values().data(), sizeof(S) * width);
		if (run_inv_sd) std::memcpy(run_inv_sd + i * width, params[4 * i + 1]->get_values().data(), sizeof(S) * width);
	}
	if (y_infer) {
		Tensor<S,4> inf = layer.pass_forward(make4<S>(x + (steps - 1) * vol, n, h, w, c), false);
		std::memcpy(y_infer, inf.data(), sizeof(S) * vol);
	}
	return 0;
}

template<typename S>
int bn_impl(int per_channel, int n, int h, int w, int c, S decay, S eps, int steps, const S* x,
		const S* gamma, const S* beta, const S* dy, S* y, S* dx, S* dgamma, S* dbeta, S* run_mean,
		S* run_inv_sd, S* y_infer) {
	Dimensions<std::size_t,3> dims({ (std::size_t) h, (std::size_t) w, (std::size_t) c });
	if (per_channel) {
		BatchNormLayer<S,3,true> layer(dims, decay, eps);
		return bn_run<S>(layer, true, n, h, w, c, steps, x, gamma, beta, dy, y, dx, dgamma, dbeta,
				run_mean, run_inv_sd, y_infer);
	} else {
		BatchNormLayer<S,3,false> layer(dims, decay, eps);
		return bn_run<S>(layer, false, n, h, w, c, steps, x, gamma, beta, dy, y, dx, dgamma, dbeta,
				run_mean, run_inv_sd, y_infer);
	}
}

// Exposes the protected _fit/_update_params of an optimizer class (SGDOptimizer.hpp:114,122).
template<typename Opt>
struct Open : public Opt {
	using Opt::Opt;
	using Opt::_fit;
	using Opt::_update_params;
};

// kind numbering == CATTL3_OPT_* in the C ABI.  hyper = {lr, a, b, eps}:
//  0 VanillaSGD(lr)  1 Momentum(lr, annealing=a, momentum=b)  2 Nesterov(lr, a, b)  3 AdaGrad(lr, eps)
//  4 RMSProp(lr, l2_decay=b, eps)  5 AdaDelta(decay=a, eps)  6 Adam(lr, l1=a, l2=b, eps)
//  7 AdaMax  8 Nadam  9 AMSGrad (same as Adam)
// Applies `steps` updates of one rows x cols parameter matrix with L2 penalty `l2_lambda` (0 = none)
// from the raw gradients grads[step]; epoch = step / steps_per_epoch, timestep = step.
// ParamsT: the Parameters implementation the optimizer updates -- StandardParameters (host matrices, the
// reference's own) or, when this file is compiled against the B200 headers (tests/cpp/Makefile), the
// device-resident B200Parameters.
template<typename S, typename ParamsT>
int opt_impl(int kind, const S* hyper, S l2_lambda, int rows, int cols, int steps, int steps_per_epoch,
		const S* p0, const S* grads, S* p_out) {
	typedef LossSharedPtr<S,1,false> LossPtr;
	LossPtr loss = std::make_shared<SquaredLoss<S,1,false>>();
	ParamRegSharedPtr<S> reg = l2_lambda > 0 ?
			std::make_shared<L2ParameterRegularization<S>>(l2_lambda) : nullptr;
	ParamsT params(rows, cols, true, nullptr, reg);
	params.init();
	params.set_values(Eigen::Map<const Matrix<S>>(p0, rows, cols));
	std::vector<Parameters<S>*> vec({ &params });
	S lr = hyper[0], a = hyper[1], b = hyper[2], eps = hyper[3];
	std::unique_ptr<SGDOptimizer<S,1,false>> opt;
	auto run = [&](auto* o) {
		o->_fit(vec);
		for (int s = 0; s < steps; ++s) {
			params.accumulate_grad(Eigen::Map<const Matrix<S>>(grads + (std::size_t) s * rows * cols, rows, cols));
			params.regularize();
			o->_update_params(vec, s / steps_per_epoch, s);
			params.reset_grad();
		}
	};
	switch (kind) {
		case 0: { Open<VanillaSGDOptimizer<S,1,false>> o(loss, 1, lr); run(&o); break; }
		case 1: { Open<MomentumSGDOptimizer<S,1,false>> o(loss, 1, lr, a, b); run(&o); break; }
		case 2: { Open<NesterovMomentumSGDOptimizer<S,1,false>> o(loss, 1, lr, a, b); run(&o); break; }
		case 3: { Open<AdaGradOptimizer<S,1,false>> o(loss, 1, lr, eps); run(&o); break; }
		case 4: { Open<RMSPropOptimizer<S,1,false>> o(loss, 1, lr, b, eps); run(&o); break; }
		case 5: { Open<AdaDeltaOptimizer<S,1,false>> o(loss, 1, a, eps); run(&o); break; }
		case 6: { Open<AdamOptimizer<S,1,false>> o(loss, 1, lr, a, b, eps); run(&o); break; }
		case 7: { Open<AdaMaxOptimizer<S,1,false>> o(loss, 1, lr, a, b, eps); run(&o); break; }
		case 8: { Open<NadamOptimizer<S,1,false>> o(loss, 1, lr, a, b, eps); run(&o); break; }
		case 9: { Open<AMSGradOptimizer<S,1,false>> o(loss, 1, lr, a, b, eps); run(&o); break; }
		default: return -1;
	}
	std::memcpy(p_out, params.get_values().data(), sizeof(S) * rows * cols);
	return 0;
}

/*
 * BASELINE.json configs[0]: the network of examples/cifar_convnet.cpp:24-37 with the two Dropout
 * layers removed (their RNG is not reproducible), CrossEntropyLoss,
 * NadamOptimizer(batch).  Observations x are total x 32 x 32 x 3 (N fastest), objectives
 * one-hot total x 10.  Trains `epochs` epochs through Optimizer::train on a MemoryDataProvider
 * WITHOUT shuffling and returns the last epoch loss; all parameters (get_all_unique_params order)
 * are copied to params_out (concatenated col-major matrices).
 */
template<typename S>
int cifar_impl(int total, int batch, int epochs, const S* x, const S* obj, const S* params_in,
		S* params_out, double* loss_out, double* train_ms) {
	typedef std::size_t sz;
	auto he = std::make_shared<HeParameterInitialization<S>>(1e-1);
	auto glorot = std::make_shared<GlorotParameterInitialization<S>>(1e-1);
	std::vector<LayerPtr<S,3>> layers;
	/* tests only: REF_SHIM_REG = "l1:<lambda>" | "l2:<lambda>" | "en:<l1>,<l2>" puts that penalty on every weight matrix */
	ParamRegSharedPtr<S> reg;
	if (const char* spec = std::getenv("REF_SHIM_REG")) {
		const std::string text(spec);
		if (text.compare(0, 3, "l1:") == 0)
			reg = std::make_shared<L1ParameterRegularization<S>>((S) std::atof(spec + 3));
		else if (text.compare(0, 3, "l2:") == 0)
			reg = std::make_shared<L2ParameterRegularization<S>>((S) std::atof(spec + 3));
		else if (text.compare(0, 3, "en:") == 0)
			reg = std::make_shared<ElasticNetParameterRegularization<S>>((S) std::atof(spec + 3),
					(S) std::atof(spec + text.find(',') + 1));
	}
	/* tests only: REF_SHIM_CONSTRAINTS = "<value clip>,<value max L1>,<value max L2>,<grad clip>,<grad max L1>,<grad max L2>"
	 * puts those constraints (StandardParameters.hpp:150-182; 0 = off) on every weight matrix */
	S con[6] = { 0, 0, 0, 0, 0, 0 };
	if (const char* spec = std::getenv("REF_SHIM_CONSTRAINTS")) {
		const char* at = spec;
		for (int i = 0; i < 6 && at; ++i) {
			con[i] = (S) std::atof(at);
			at = std::strchr(at, ',');
			if (at) ++at;
		}
	}
	layers.emplace_back(new ConvKernelLayer<S>({ 32u, 32u, 3u }, 8, he, 3, 3, 1, 1, 1, 1, 0, 0, reg,
			con[0], con[1], con[2], con[3], con[4], con[5]));
	layers.emplace_back(new ReLUActivationLayer<S,3>(layers.back()->get_output_dims()));
	layers.emplace_back(new MaxPoolLayer<S>(layers.back()->get_output_dims()));
	layers.emplace_back(new ConvKernelLayer<S>(layers.back()->get_output_dims(), 8, he, 3, 3, 1, 1, 1, 1, 0, 0, reg,
			con[0], con[1], con[2], con[3], con[4], con[5]));
	layers.emplace_back(new ReLUActivationLayer<S,3>(layers.back()->get_output_dims()));
	layers.emplace_back(new MaxPoolLayer<S>(layers.back()->get_output_dims()));
	layers.emplace_back(new DenseKernelLayer<S,3>(layers.back()->get_output_dims(), 50, glorot, reg,
			con[0], con[1], con[2], con[3], con[4], con[5]));
	layers.emplace_back(new ReLUActivationLayer<S,3>(layers.back()->get_output_dims()));
	layers.emplace_back(new DenseKernelLayer<S,3>(layers.back()->get_output_dims(), 10, glorot, reg,
			con[0], con[1], con[2], con[3], con[4], con[5]));
	layers.emplace_back(new SoftmaxActivationLayer<S,3>(layers.back()->get_output_dims()));
	FeedforwardNeuralNetwork<S,3> net(std::move(layers));
	net.init();
	auto params = net.get_all_unique_params();
	if (params_in) {
		const S* src = params_in;
		for (auto p : params) {
			p->set_values(Eigen::Map<const Matrix<S>>(src, p->get_rows(), p->get_cols()));
			src += p->get_rows() * p->get_cols();
		}
	}
	/* tests only: REF_SHIM_PRMS_LOAD / REF_SHIM_PRMS_SAVE = a directory of .prms files (NeuralNetwork.hpp:195-222,
	 * EigenProxy.hpp:147-224), read after the injection above / written after training; REF_SHIM_PRMS_TEXT=1 = the text form */
	const bool prms_binary = std::getenv("REF_SHIM_PRMS_TEXT") == nullptr;
	if (const char* dir = std::getenv("REF_SHIM_PRMS_LOAD"))
		net.load_all_unique_params_values(dir, prms_binary);
	TensorPtr<S,4> obs(new Tensor<S,4>((sz) total, 32u, 32u, 3u));
	std::memcpy(obs->data(), x, sizeof(S) * obs->size());
	TensorPtr<S,4> objs(new Tensor<S,4>((sz) total, 1u, 1u, 10u));
	std::memcpy(objs->data(), obj, sizeof(S) * objs->size());
	MemoryDataProvider<S,3,false,false> prov(std::move(obs), std::move(objs));
	auto loss = std::make_shared<CrossEntropyLoss<S,3,false>>();
	NadamOptimizer<S,3,false> opt(loss, batch);
	opt.fit(net);
	/* benchmarks only: REF_SHIM_WARMUP_EPOCHS untimed epochs first (one-time costs: allocations, data set placement) */
	if (const char* warm = std::getenv("REF_SHIM_WARMUP_EPOCHS")) {
		if (std::atoi(warm) > 0 && epochs > 0)
			opt.train(net, prov, std::atoi(warm));
	}
	double t0 = now_ms();
	S l = epochs > 0 ? opt.train(net, prov, epochs) : (S) 0;
	double t1 = now_ms();
	if (loss_out) *loss_out = (double) l;
	if (train_ms) *train_ms = t1 - t0;
	if (const char* dir = std::getenv("REF_SHIM_PRMS_SAVE"))
		net.save_all_unique_params_values(dir, prms_binary);
	if (params_out) {
		S* dst = params_out;
		for (auto p : params) {
			std::memcpy(dst, p->get_values().data(), sizeof(S) * p->get_rows() * p->get_cols());
			dst += p->get_rows() * p->get_cols();
		}
	}
	return 0;
}

/*
 * DropoutLayer<S,3> (C-ATTL3/layer/DropoutLayer.hpp:74-94): one training forward + backward and one inference
 * forward on an n x h x w x c tensor.  The masks are random (not reproducible across implementations), so callers
 * check the layer's contract: y = x * mask, dx = dy * mask with the SAME mask, mask in {0, 1 / (1 - p + eps)},
 * drop rate ~ p, inference = identity.
 */
template<typename S>
int dropout_impl(int n, int h, int w, int c, S prob, const S* x, const S* dy, S* y, S* dx, S* y_infer) {
	DropoutLayer<S,3> layer({ (std::size_t) h, (std::size_t) w, (std::size_t) c }, prob);
	Tensor<S,4> out = layer.pass_forward(make4(x, n, h, w, c), true);
	std::memcpy(y, out.data(), sizeof(S) * out.size());
	Tensor<S,4> grad = layer.pass_back(make4(dy, n, h, w, c));
	std::memcpy(dx, grad.data(), sizeof(S) * grad.size());
	Tensor<S,4> inf = layer.pass_forward(make4(x, n, h, w, c), false);
	std::memcpy(y_infer, inf.data(), sizeof(S) * inf.size());
	return 0;
}

/* Shared tail of the network trainers: inject parameters, train without shuffling, copy the parameters out. */
template<typename S, bool Seq, typename Opt>
int train_and_export(NeuralNetwork<S,3,Seq>& net, Opt& opt, TensorPtr<S,4 + Seq> obs, TensorPtr<S,4 + Seq> objs, int epochs,
		const S* params_in, S* params_out, int* n_params, double* loss_out, double* train_ms) {
	net.init();
	auto params = net.get_all_unique_params();
	int count = 0;
	for (auto p : params)
		count += (int) (p->get_rows() * p->get_cols());
	if (n_params) *n_params = count;
	if (epochs < 0)
		return 0;  /* size query only */
	if (params_in) {
		const S* src = params_in;
		for (auto p : params) {
			p->set_values(Eigen::Map<const Matrix<S>>(src, p->get_rows(), p->get_cols()));
			src += p->get_rows() * p->get_cols();
		}
	}
	MemoryDataProvider<S,3,Seq,false> prov(std::move(obs), std::move(objs));
	opt.fit(net);
	/* benchmarks only: REF_SHIM_WARMUP_EPOCHS untimed epochs first (one-time costs: allocations, data set placement) */
	if (const char* warm = std::getenv("REF_SHIM_WARMUP_EPOCHS")) {
		if (std::atoi(warm) > 0 && epochs > 0)
			opt.train(net, prov, std::atoi(warm));
	}
	double t0 = now_ms();
	S l = epochs > 0 ? opt.train(net, prov, epochs) : (S) 0;
	double t1 = now_ms();
	if (loss_out) *loss_out = (double) l;
	if (train_ms) *train_ms = t1 - t0;
	if (params_out) {
		S* dst = params_out;
		for (auto p : params) {
			std::memcpy(dst, p->get_values().data(), sizeof(S) * p->get_rows() * p->get_cols());
			dst += p->get_rows() * p->get_cols();
		}
	}
	return 0;
}

/*
 * BASELINE.json configs[2]: the auto-encoder of examples/mnist_autoencoder.cpp:26-48 (Conv 1->3 4x4 stride 2,
 * Softplus, Conv 3->3 4x4, Softplus, Dense 300->100 | Dense 100->300, Softplus, Reshape 10x10x3,
 * TransConv 3->3 4x4, Softplus, TransConv 3->1 4x4 stride 2) as a StackedNeuralNetwork, SquaredLoss against
 * the input itself, NadamOptimizer(batch).  x: total x 28 x 28 x 1.
 */
template<typename S>
int autoencoder_impl(int total, int batch, int epochs, const S* x, const S* params_in, S* params_out, int* n_params,
		double* loss_out, double* train_ms) {
	typedef std::size_t sz;
	auto init = std::make_shared<HeParameterInitialization<S>>(1e-1);
	std::vector<LayerPtr<S,3>> enc;
	enc.emplace_back(new ConvKernelLayer<S>({ 28u, 28u, 1u }, 3, init, 4, 4, 0, 0, 2, 2));
	enc.emplace_back(new SoftplusActivationLayer<S,3>(enc.back()->get_output_dims()));
	enc.emplace_back(new ConvKernelLayer<S>(enc.back()->get_output_dims(), 3, init, 4, 4, 0, 0, 1, 1));
	enc.emplace_back(new SoftplusActivationLayer<S,3>(enc.back()->get_output_dims()));
	enc.emplace_back(new DenseKernelLayer<S,3>(enc.back()->get_output_dims(), 100, init));
	NeuralNetPtr<S,3,false> encoder(new FeedforwardNeuralNetwork<S,3>(std::move(enc)));
	std::vector<LayerPtr<S,3>> dec;
	dec.emplace_back(new DenseKernelLayer<S,3>(encoder->get_output_dims(), 300, init));
	dec.emplace_back(new SoftplusActivationLayer<S,3>(dec.back()->get_output_dims()));
	dec.emplace_back(new ReshapeLayer<S,3>(dec.back()->get_output_dims(), { 10u, 10u, 3u }));
	dec.emplace_back(new TransConvKernelLayer<S>(dec.back()->get_output_dims(), 3, init, 4, 4, 0, 0, 1, 1));
	dec.emplace_back(new SoftplusActivationLayer<S,3>(dec.back()->get_output_dims()));
	dec.emplace_back(new TransConvKernelLayer<S>(dec.back()->get_output_dims(), 1, init, 4, 4, 0, 0, 2, 2));
	NeuralNetPtr<S,3,false> decoder(new FeedforwardNeuralNetwork<S,3>(std::move(dec)));
	std::vector<NeuralNetPtr<S,3,false>> modules;
	modules.push_back(std::move(encoder));
	modules.push_back(std::move(decoder));
	StackedNeuralNetwork<S,3,false> net(std::move(modules));
	TensorPtr<S,4> obs(new Tensor<S,4>((sz) (total > 0 ? total : 1), 28u, 28u, 1u));
	if (x) std::memcpy(obs->data(), x, sizeof(S) * obs->size());
	TensorPtr<S,4> objs(new Tensor<S,4>(*obs));
	auto loss = std::make_shared<SquaredLoss<S,3,false>>();
	NadamOptimizer<S,3,false> opt(loss, batch > 0 ? batch : 1);
	return train_and_export<S,false>(net, opt, std::move(obs), std::move(objs), epochs, params_in, params_out, n_params,
			loss_out, train_ms);
}

/*
 * BASELINE.json configs[3]: a ResNet-style network -- stem FeedforwardNeuralNetwork{Conv stem_r x stem_r
 * (c -> width, stride stem_s, "same"-style padding stem_r / 2), BatchNorm, ReLU[, MaxPool 2x2]}, a
 * ResidualNeuralNetwork of `blocks` modules FeedforwardNeuralNetwork{Conv 3x3 p1 (width -> width), BatchNorm,
 * ReLU, Conv 3x3 p1, BatchNorm} (module input dims == output dims, ResidualNeuralNetwork.hpp:48-49), and a
 * head {ReLU, MeanPool head_pool x head_pool, Dense -> classes, Softmax}; CrossEntropyLoss, NadamOptimizer.
 * x: total x h x w x c, obj: total x 1 x 1 x classes (one-hot).
 */
template<typename S>
int resnet_impl(int total, int batch, int epochs, int h, int w, int c, int stem_r, int stem_s, int stem_pool,
		int width, int blocks, int head_pool, int classes, const S* x, const S* obj, const S* params_in,
		S* params_out, int* n_params, double* loss_out, double* train_ms) {
	typedef std::size_t sz;
	auto he = std::make_shared<HeParameterInitialization<S>>(1e-1);
	auto glorot = std::make_shared<GlorotParameterInitialization<S>>(1e-1);
	std::vector<LayerPtr<S,3>> stem;
	stem.emplace_back(new ConvKernelLayer<S>({ (sz) h, (sz) w, (sz) c }, width, he, stem_r, stem_r, stem_r / 2,
			stem_r / 2, stem_s, stem_s));
	stem.emplace_back(new BatchNormLayer<S,3>(stem.back()->get_output_dims()));
	stem.emplace_back(new ReLUActivationLayer<S,3>(stem.back()->get_output_dims()));
	if (stem_pool)
		stem.emplace_back(new MaxPoolLayer<S>(stem.back()->get_output_dims()));
	std::vector<NeuralNetPtr<S,3,false>> stack;
	stack.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(stem)));
	std::vector<NeuralNetPtr<S,3,false>> modules;
	for (int i = 0; i < blocks; ++i) {
		std::vector<LayerPtr<S,3>> m;
		m.emplace_back(new ConvKernelLayer<S>(stack[0]->get_output_dims(), width, he));
		m.emplace_back(new BatchNormLayer<S,3>(m.back()->get_output_dims()));
		m.emplace_back(new ReLUActivationLayer<S,3>(m.back()->get_output_dims()));
		m.emplace_back(new ConvKernelLayer<S>(m.back()->get_output_dims(), width, he));
		m.emplace_back(new BatchNormLayer<S,3>(m.back()->get_output_dims()));
		modules.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(m)));
	}
	stack.emplace_back(new ResidualNeuralNetwork<S,3>(std::move(modules)));
	std::vector<LayerPtr<S,3>> head;
	head.emplace_back(new ReLUActivationLayer<S,3>(stack.back()->get_output_dims()));
	head.emplace_back(new MeanPoolLayer<S>(head.back()->get_output_dims(), head_pool, head_pool, head_pool, head_pool));
	head.emplace_back(new DenseKernelLayer<S,3>(head.back()->get_output_dims(), classes, glorot));
	head.emplace_back(new SoftmaxActivationLayer<S,3>(head.back()->get_output_dims()));
	stack.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(head)));
	StackedNeuralNetwork<S,3,false> net(std::move(stack));
	TensorPtr<S,4> obs(new Tensor<S,4>((sz) (total > 0 ? total : 1), (sz) h, (sz) w, (sz) c));
	if (x) std::memcpy(obs->data(), x, sizeof(S) * obs->size());
	TensorPtr<S,4> objs(new Tensor<S,4>((sz) (total > 0 ? total : 1), 1u, 1u, (sz) classes));
	if (obj) std::memcpy(objs->data(), obj, sizeof(S) * objs->size());
	auto loss = std::make_shared<CrossEntropyLoss<S,3,false>>();
	NadamOptimizer<S,3,false> opt(loss, batch > 0 ? batch : 1);
	return train_and_export<S,false>(net, opt, std::move(obs), std::move(objs), epochs, params_in, params_out, n_params,
			loss_out, train_ms);
}

/*
 * BASELINE.json configs[4]: a composite network over sequences of frames -- a SequentialNeuralNetwork (the time
 * steps folded into the batch, SequentialNeuralNetwork.hpp:95-124) around a StackedNeuralNetwork of
 *   an Inception-style ParallelNeuralNetwork (lanes Conv 3x3 p1 -> ReLU and Conv 1x1 -> ReLU on the same input,
 *   outputs concatenated along the channel rank, ParallelNeuralNetwork.hpp:176-194),
 *   a DenseNeuralNetwork of two modules Conv 3x3 p1 -> ReLU (each sees the input and all earlier module outputs
 *   concatenated along the channel rank, DenseNeuralNetwork.hpp:131-158) and
 *   a 2x2 MaxPool,
 * feeding a convolutional LSTMNeuralNetwork whose eight kernels are ConvKernelLayers (3x3 p1) as in
 * test/gradient_test.cpp:700-734, which emits one output frame after the last time step; SquaredLoss,
 * NadamOptimizer, both sequential.  x: total x seq x hw x hw x 3, obj: total x 1 x hw/2 x hw/2 x state.
 */
template<typename S>
int seqnet_impl(int total, int seq, int batch, int epochs, int hw, int width, int state, const S* x, const S* obj,
		const S* params_in, S* params_out, int* n_params, double* loss_out, double* train_ms) {
	typedef std::size_t sz;
	auto he = std::make_shared<HeParameterInitialization<S>>(1e-1);
	auto glorot = std::make_shared<GlorotParameterInitialization<S>>(1e-1);
	const Dimensions<sz,3> in({ (sz) hw, (sz) hw, 3u });
	std::vector<NeuralNetPtr<S,3,false>> lanes;
	{
		std::vector<LayerPtr<S,3>> a, b;
		a.emplace_back(new ConvKernelLayer<S>(in, width, he));
		a.emplace_back(new ReLUActivationLayer<S,3>(a.back()->get_output_dims()));
		b.emplace_back(new ConvKernelLayer<S>(in, width, he, 1, 1, 0, 0));
		b.emplace_back(new ReLUActivationLayer<S,3>(b.back()->get_output_dims()));
		lanes.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(a)));
		lanes.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(b)));
	}
	std::vector<NeuralNetPtr<S,3,false>> front;
	front.emplace_back(new ParallelNeuralNetwork<S,3>(std::move(lanes)));
	const Dimensions<sz,3> merged = front.back()->get_output_dims();
	std::vector<NeuralNetPtr<S,3,false>> modules;
	{
		std::vector<LayerPtr<S,3>> m0, m1;
		m0.emplace_back(new ConvKernelLayer<S>(merged, width, he));
		m0.emplace_back(new ReLUActivationLayer<S,3>(m0.back()->get_output_dims()));
		modules.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(m0)));
		m1.emplace_back(new ConvKernelLayer<S>(merged.add_along_rank(modules[0]->get_output_dims(), 2), width, he));
		m1.emplace_back(new ReLUActivationLayer<S,3>(m1.back()->get_output_dims()));
		modules.emplace_back(new FeedforwardNeuralNetwork<S,3>(std::move(m1)));
	}
	front.emplace_back(new DenseNeuralNetwork<S,3,DENSE_HIGHEST_RANK>(std::move(modules)));
	front.emplace_back(new FeedforwardNeuralNetwork<S,3>(LayerPtr<S,3>(
			new MaxPoolLayer<S>(front.back()->get_output_dims()))));
	NeuralNetPtr<S,3,false> frames(new StackedNeuralNetwork<S,3,false>(std::move(front)));
	const Dimensions<sz,3> lstm_in = frames->get_output_dims();
	const Dimensions<sz,3> lstm_out({ lstm_in(0), lstm_in(1), (sz) state });
	auto in_kernel = [&]() { return KernelPtr<S,3>(new ConvKernelLayer<S>(lstm_in, state, glorot)); };
	auto out_kernel = [&]() { return KernelPtr<S,3>(new ConvKernelLayer<S>(lstm_out, state, glorot)); };
	auto sigmoid = [&]() { return ActivationPtr<S,3>(new SigmoidActivationLayer<S,3>(lstm_out)); };
	auto tanh_act = [&]() { return ActivationPtr<S,3>(new TanhActivationLayer<S,3>(lstm_out)); };
	std::vector<NeuralNetPtr<S,3,true>> stack;
	stack.emplace_back(new SequentialNeuralNetwork<S,3>(std::move(frames)));
	stack.emplace_back(new LSTMNeuralNetwork<S,3>(in_kernel(), out_kernel(), in_kernel(), out_kernel(), in_kernel(),
			out_kernel(), in_kernel(), out_kernel(), sigmoid(), sigmoid(), tanh_act(), tanh_act(), sigmoid(),
			[](sz input_seq_length) { return std::make_pair((sz) 1, input_seq_length - 1); }));
	StackedNeuralNetwork<S,3,true> net(std::move(stack));
	const sz rows = (sz) (total > 0 ? total : 1), steps = (sz) (seq > 0 ? seq : 1);
	TensorPtr<S,5> obs(new Tensor<S,5>(rows, steps, (sz) hw, (sz) hw, 3u));
	if (x) std::memcpy(obs->data(), x, sizeof(S) * obs->size());
	TensorPtr<S,5> objs(new Tensor<S,5>(rows, 1u, lstm_out(0), lstm_out(1), lstm_out(2)));
	if (obj) std::memcpy(objs->data(), obj, sizeof(S) * objs->size());
	auto loss = std::make_shared<SquaredLoss<S,3,true>>();
	NadamOptimizer<S,3,true> opt(loss, batch > 0 ? batch : 1);
	return train_and_export<S,true>(net, opt, std::move(obs), std::move(objs), epochs, params_in, params_out, n_params,
			loss_out, train_ms);
}

} /* namespace */

#ifdef C_ATTL3_B200_CATTLE_H_
template<typename S> using DefaultParams = B200Parameters<S>;
#else
template<typename S> using DefaultParams = StandardParameters<S>;
#endif

extern "C" {

/* 1 when this driver was compiled against the B200 headers (c-attl3_b200/cattle), 0 for the reference. */
int ref_is_b200_build() {
#ifdef C_ATTL3_B200_CATTLE_H_
	return 1;
#else
	return 0;
#endif
}

int ref_num_threads() { return num_of_eval_threads(); }
void ref_set_num_threads(int n) { set_num_of_eval_threads(n); }

#define DEFINE_FOR(S, SUF) \
int ref_conv_##SUF(const Geom* g, int transposed, const S* x, const S* w, const S* b, const S* dy, S* y, \
		S* dx, S* dw, S* db, int back_reps, double* times_ms) { \
	return conv_impl<S>(g, transposed, x, w, b, dy, y, dx, dw, db, back_reps, times_ms); } \
int ref_dense_##SUF(int n, int in, int out, const S* x, const S* w, const S* b, const S* dy, S* y, S* dx, \
		S* dw, S* db, int back_reps, double* times_ms) { \
	return dense_impl<S>(n, in, out, x, w, b, dy, y, dx, dw, db, back_reps, times_ms); } \
int ref_activation_##SUF(int kind, S alpha, int n, int vol, const S* x, const S* dy, S* y, S* dx) { \
	return act_impl<S>(kind, alpha, n, vol, x, dy, y, dx); } \
int ref_pool_##SUF(int kind, int n, int h, int w, int c, int rh, int rw, int sh, int sw, const S* x, \
		const S* dy, S* y, S* dx, double* times_ms) { \
	return pool_impl<S>(kind, n, h, w, c, rh, rw, sh, sw, x, dy, y, dx, times_ms); } \
int ref_batchnorm_##SUF(int per_channel, int n, int h, int w, int c, S decay, S eps, int steps, const S* x, \
		const S* gamma, const S* beta, const S* dy, S* y, S* dx, S* dgamma, S* dbeta, S* run_mean, \
		S* run_inv_sd, S* y_infer) { \
	return bn_impl<S>(per_channel, n, h, w, c, decay, eps, steps, x, gamma, beta, dy, y, dx, dgamma, dbeta, \
			run_mean, run_inv_sd, y_infer); } \
int ref_optimizer_##SUF(int kind, const S* hyper, S l2_lambda, int rows, int cols, int steps, \
		int steps_per_epoch, const S* p0, const S* grads, S* p_out) { \
	return opt_impl<S,DefaultParams<S>>(kind, hyper, l2_lambda, rows, cols, steps, steps_per_epoch, p0, grads, p_out); } \
int ref_optimizer_hostparams_##SUF(int kind, const S* hyper, S l2_lambda, int rows, int cols, int steps, \
		int steps_per_epoch, const S* p0, const S* grads, S* p_out) { \
	return opt_impl<S,StandardParameters<S>>(kind, hyper, l2_lambda, rows, cols, steps, steps_per_epoch, p0, grads, p_out); } \
int ref_train_cifar_##SUF(int total, int batch, int epochs, const S* x, const S* obj, const S* params_in, \
		S* params_out, double* loss_out, double* train_ms) { \
	return cifar_impl<S>(total, batch, epochs, x, obj, params_in, params_out, loss_out, train_ms); } \
int ref_dropout_##SUF(int n, int h, int w, int c, S prob, const S* x, const S* dy, S* y, S* dx, S* y_infer) { \
	return dropout_impl<S>(n, h, w, c, prob, x, dy, y, dx, y_infer); } \
int ref_train_autoencoder_##SUF(int total, int batch, int epochs, const S* x, const S* params_in, S* params_out, \
		int* n_params, double* loss_out, double* train_ms) { \
	return autoencoder_impl<S>(total, batch, epochs, x, params_in, params_out, n_params, loss_out, train_ms); } \
int ref_train_resnet_##SUF(int total, int batch, int epochs, int h, int w, int c, int stem_r, int stem_s, int stem_pool, \
		int width, int blocks, int head_pool, int classes, const S* x, const S* obj, const S* params_in, S* params_out, \
		int* n_params, double* loss_out, double* train_ms) { \
	return resnet_impl<S>(total, batch, epochs, h, w, c, stem_r, stem_s, stem_pool, width, blocks, head_pool, classes, x, \
			obj, params_in, params_out, n_params, loss_out, train_ms); } \
int ref_train_seqnet_##SUF(int total, int seq, int batch, int epochs, int hw, int width, int state, const S* x, const S* obj, \
		const S* params_in, S* params_out, int* n_params, double* loss_out, double* train_ms) { \
	return seqnet_impl<S>(total, seq, batch, epochs, hw, width, state, x, obj, params_in, params_out, n_params, loss_out, \
			train_ms); }

DEFINE_FOR(float, f32)
DEFINE_FOR(double, f64)

}
