// activations.cuh -- the element-wise activation formulas (forward and derivative), shared by the stand-alone
// activation kernels (elementwise.cu) and by the kernel layers' fused epilogues (conv_tc.cu, conv_simt.cu).
#pragma once

#include "common.cuh"

namespace cattl3 {

template<typename S> struct V16;
template<> struct V16<float> { typedef float4 type; static constexpr int G = 4; };
template<> struct V16<double> { typedef double2 type; static constexpr int G = 2; };

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template<typename S> __device__ __forceinline__ S dev_exp(S v);
template<> __device__ __forceinline__ float dev_exp<float>(float v) { return expf(v); }
template<> __device__ __forceinline__ double dev_exp<double>(double v) { return exp(v); }
template<typename S> __device__ __forceinline__ S dev_log(S v);
template<> __device__ __forceinline__ float dev_log<float>(float v) { return logf(v); }
template<> __device__ __forceinline__ double dev_log<double>(double v) { return log(v); }
template<typename S> __device__ __forceinline__ S dev_tanh(S v);
template<> __device__ __forceinline__ float dev_tanh<float>(float v) { return tanhf(v); }
template<> __device__ __forceinline__ double dev_tanh<double>(double v) { return tanh(v); }
template<typename S> __device__ __forceinline__ S dev_sqrt(S v);
template<> __device__ __forceinline__ float dev_sqrt<float>(float v) { return sqrtf(v); }
template<> __device__ __forceinline__ double dev_sqrt<double>(double v) { return sqrt(v); }

// ---- activations -------------------------------------------------------------------------------
// Formulas follow the reference layer by layer (see include/cattl3_b200.h for file:line).
template<typename S, int KIND>
__device__ __forceinline__ S act_fwd(S x, S a) {
	if (KIND == CATTL3_ACT_RELU) return x > (S) 0 ? x : (S) 0;              // cwiseMax(0)
	if (KIND == CATTL3_ACT_LEAKY_RELU) { S ax = x * a; return x > ax ? x : ax; } // cwiseMax(x * alpha)
	if (KIND == CATTL3_ACT_ELU) return x >= (S) 0 ? x : a * (dev_exp<S>(x) - (S) 1);
	if (KIND == CATTL3_ACT_SWISH) return x * ((S) 1 / (dev_exp<S>(-a * x) + (S) 1));
	if (KIND == CATTL3_ACT_SIGMOID) return (S) 1 / (dev_exp<S>(-x) + (S) 1);
	if (KIND == CATTL3_ACT_TANH) return dev_tanh<S>(x);
	if (KIND == CATTL3_ACT_SOFTPLUS) return dev_log<S>(dev_exp<S>(x) + (S) 1);
	return x;
}

template<typename S, int KIND>
__device__ __forceinline__ S act_bwd(S x, S y, S g, S a) {
	if (KIND == CATTL3_ACT_RELU) return x >= (S) 0 ? g : (S) 0;             // derivative 1 at x == 0
	if (KIND == CATTL3_ACT_LEAKY_RELU) return x >= (S) 0 ? g : a * g;
	if (KIND == CATTL3_ACT_ELU) return x >= (S) 0 ? g : (y + a) * g;
	if (KIND == CATTL3_ACT_SWISH) {
		S s = (S) 1 / (dev_exp<S>(-a * x) + (S) 1);
		return s * (((S) 1 - s) * a * x + (S) 1) * g;
	}
	if (KIND == CATTL3_ACT_SIGMOID) return (y * ((S) 1 - y)) * g;
	if (KIND == CATTL3_ACT_TANH) return ((S) 1 - y * y) * g;
	if (KIND == CATTL3_ACT_SOFTPLUS) return ((S) 1 / (dev_exp<S>(-x) + (S) 1)) * g;
	return g;
}

// Run-time dispatch for fused epilogues: `kind` is uniform over the grid, so the switch does not diverge.
template<typename S>
__device__ __forceinline__ S act_fwd_rt(int kind, S x, S a) {
	switch (kind) {
		case CATTL3_ACT_RELU: return act_fwd<S, CATTL3_ACT_RELU>(x, a);
		case CATTL3_ACT_LEAKY_RELU: return act_fwd<S, CATTL3_ACT_LEAKY_RELU>(x, a);
		case CATTL3_ACT_ELU: return act_fwd<S, CATTL3_ACT_ELU>(x, a);
		case CATTL3_ACT_SWISH: return act_fwd<S, CATTL3_ACT_SWISH>(x, a);
		case CATTL3_ACT_SIGMOID: return act_fwd<S, CATTL3_ACT_SIGMOID>(x, a);
		case CATTL3_ACT_TANH: return act_fwd<S, CATTL3_ACT_TANH>(x, a);
		case CATTL3_ACT_SOFTPLUS: return act_fwd<S, CATTL3_ACT_SOFTPLUS>(x, a);
		default: return x;
	}
}

// N values at once with the switch hoisted out of the element loop (straight-line code per kind in an epilogue).
template<typename S, int N>
__device__ __forceinline__ void act_fwd_rt_n(int kind, S* v, S a) {
#define CATTL3_ACT_CASE(K) case K: _Pragma("unroll") for (int i = 0; i < N; ++i) v[i] = act_fwd<S, K>(v[i], a); break;
	switch (kind) {
		CATTL3_ACT_CASE(CATTL3_ACT_RELU) CATTL3_ACT_CASE(CATTL3_ACT_LEAKY_RELU) CATTL3_ACT_CASE(CATTL3_ACT_ELU)
		CATTL3_ACT_CASE(CATTL3_ACT_SWISH) CATTL3_ACT_CASE(CATTL3_ACT_SIGMOID) CATTL3_ACT_CASE(CATTL3_ACT_TANH)
		CATTL3_ACT_CASE(CATTL3_ACT_SOFTPLUS)
		default: break;
	}
#undef CATTL3_ACT_CASE
}

} // namespace cattl3
