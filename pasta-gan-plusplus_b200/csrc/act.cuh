// Activation arithmetic shared by the stand-alone bias_act kernel and the conv epilogue.
// Semantics follow the reference plugin (torch_utils/ops/bias_act.cu:38-146): G = 0 forward,
// G = 1 first derivative (dy * act'(.) * gain, clamp gates on yref), G = 2 second derivative.
#pragma once
#include "common.cuh"

namespace pgpp {

template <class S> __device__ __forceinline__ S act_exp(S v);
template <> __device__ __forceinline__ float act_exp<float>(float v) { return expf(v); }
template <> __device__ __forceinline__ double act_exp<double>(double v) { return exp(v); }
template <class S> __device__ __forceinline__ S act_log1p(S v);
template <> __device__ __forceinline__ float act_log1p<float>(float v) { return log1pf(v); }
template <> __device__ __forceinline__ double act_log1p<double>(double v) { return log1p(v); }
template <class S> __device__ __forceinline__ S act_expm1(S v);
template <> __device__ __forceinline__ float act_expm1<float>(float v) { return expm1f(v); }
template <> __device__ __forceinline__ double act_expm1<double>(double v) { return expm1(v); }
template <class S> __device__ __forceinline__ S act_tanh(S v);
template <> __device__ __forceinline__ float act_tanh<float>(float v) { return tanhf(v); }
template <> __device__ __forceinline__ double act_tanh<double>(double v) { return tanh(v); }

// forward activation only (conv epilogue + G == 0)
template <int A, class S>
__device__ __forceinline__ S act_forward(S x, S alpha) {
    const S one = (S)1;
    if (A == PGPP_ACT_LINEAR) return x;
    if (A == PGPP_ACT_RELU) return x > 0 ? x : (S)0;
    if (A == PGPP_ACT_LRELU) return x > 0 ? x : x * alpha;
    if (A == PGPP_ACT_TANH) return act_tanh<S>(x);
    if (A == PGPP_ACT_SIGMOID) return one / (one + act_exp<S>(-x));
    if (A == PGPP_ACT_ELU) return x >= 0 ? x : act_expm1<S>(x);
    if (A == PGPP_ACT_SELU) {
        const S sc = (S)1.0507009873554804934193349852946, al = (S)1.6732632423543772848170429916717;
        return x >= 0 ? sc * x : (sc * al) * act_expm1<S>(x);
    }
    if (A == PGPP_ACT_SOFTPLUS) return x > (S)80 ? x : (x > (S)0 ? x + act_log1p<S>(act_exp<S>(-x)) : act_log1p<S>(act_exp<S>(x)));
    if (A == PGPP_ACT_SWISH) return x / (one + act_exp<S>(-x));
    return x;
}

// One element of the plugin: returns y given the already-loaded operands.
//   G == 0: x = input (+bias added by the caller)
//   G >= 1: x = incoming gradient, xref = saved input (+bias), yref = saved output
template <int A, class S>
__device__ __forceinline__ S act_element(int G, S x, S xref, S yref, S dy, S alpha, S gain, S clamp) {
    const S one = (S)1, two = (S)2;
    S y = 0;
    if (G == 0) {
        y = act_forward<A, S>(x, alpha);
    } else {
        const S yy = (gain != 0) ? yref / gain : (S)0;
        if (A == PGPP_ACT_LINEAR) { if (G == 1) y = x; }
        if (A == PGPP_ACT_RELU)   { if (G == 1) y = yy > 0 ? x : (S)0; }
        if (A == PGPP_ACT_LRELU)  { if (G == 1) y = yy > 0 ? x : x * alpha; }
        if (A == PGPP_ACT_TANH) {
            if (G == 1) y = x * (one - yy * yy);
            if (G == 2) y = x * (one - yy * yy) * (-two * yy);
        }
        if (A == PGPP_ACT_SIGMOID) {
            if (G == 1) y = x * yy * (one - yy);
            if (G == 2) y = x * yy * (one - yy) * (one - two * yy);
        }
        if (A == PGPP_ACT_ELU) {
            if (G == 1) y = yy >= 0 ? x : x * (yy + one);
            if (G == 2) y = yy >= 0 ? (S)0 : x * (yy + one);
        }
        if (A == PGPP_ACT_SELU) {
            const S sc = (S)1.0507009873554804934193349852946, al = (S)1.6732632423543772848170429916717;
            if (G == 1) y = yy >= 0 ? x * sc : x * (yy + sc * al);
            if (G == 2) y = yy >= 0 ? (S)0 : x * (yy + sc * al);
        }
        if (A == PGPP_ACT_SOFTPLUS) {
            const S c = act_exp<S>(-yy);
            if (G == 1) y = x * (one - c);
            if (G == 2) y = x * c * (one - c);
        }
        if (A == PGPP_ACT_SWISH) {
            const bool big = xref > (S)40;
            const S c = act_exp<S>(big ? (S)40 : xref);
            const S d = c + one;
            if (G == 1) y = big ? x : x * c * (xref + d) / (d * d);
            if (G == 2) y = big ? (S)0 : x * c * (xref * (two - d) + two * d) / (d * d * d);
            yref = xref / (one + act_exp<S>(-xref)) * gain;
        }
    }
    y *= gain * dy;
    if (clamp >= 0) {
        if (G == 0) y = y > clamp ? clamp : (y < -clamp ? -clamp : y);
        else y = (yref > -clamp && yref < clamp) ? y : (S)0;
    }
    return y;
}

} // namespace pgpp
