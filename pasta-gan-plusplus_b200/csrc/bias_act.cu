// bias_act for sm_100a: HBM-bound streaming kernel.  128-bit vector loads/stores, several
// independent vectors in flight per thread, grid sized to the SM count (grid-stride loop),
// bias index computed once per vector instead of a div+mod per element.
// Replaces torch_utils/ops/bias_act.cu:23-147 + bias_act.cpp:32-90 of the reference.
#include <stdlib.h>
#include "act.cuh"

namespace pgpp {

struct BiasActArgs {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    long long size_x; int size_b; long long step_b;
    unsigned step_m, step_s, size_m, size_s;      // magic numbers: n / step_b and n / size_b for n < 2^31 (size_x <= INT_MAX)
    int grad; float alpha, gain, clamp;
    int hint;           // 1: evict-first (.cs) loads/stores, 0: default cache policy
    int bias_mode;      // 0 none, 1 same bias for the whole vector, 2 consecutive (step_b == 1), 3 per element
};

template <class T, int V> struct alignas(sizeof(T) * V) Vec { T v[V]; };

__device__ __forceinline__ unsigned fdiv(unsigned n, unsigned m, unsigned s) {
    return (unsigned)(((unsigned long long)n * m) >> (31 + s));
}
// bias index of element e: (e / step_b) % size_b without integer division
__device__ __forceinline__ int bias_index(const BiasActArgs& p, unsigned e) {
    const unsigned q = fdiv(e, p.step_m, p.step_s);
    return (int)(q - fdiv(q, p.size_m, p.size_s) * (unsigned)p.size_b);
}

template <class T, int V>
__device__ __forceinline__ Vec<T, V> ld_stream(const T* p, int hint) {
    // streaming read: every element is touched exactly once
    Vec<T, V> r;
    if (sizeof(T) * V == 16) {
        const int4 q = hint ? __ldcs(reinterpret_cast<const int4*>(p)) : *reinterpret_cast<const int4*>(p);
        r = *reinterpret_cast<const Vec<T, V>*>(&q);
    } else {
        #pragma unroll
        for (int i = 0; i < V; i++) r.v[i] = p[i];
    }
    return r;
}

template <class T, int V>
__device__ __forceinline__ void st_stream(T* p, const Vec<T, V>& r, int hint) {
    if (sizeof(T) * V == 16) { if (hint) __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&r)); else *reinterpret_cast<int4*>(p) = *reinterpret_cast<const int4*>(&r); }
    else {
        #pragma unroll
        for (int i = 0; i < V; i++) p[i] = r.v[i];
    }
}

// FWD: grad == 0 with no xref / yref / dy operands (the forward pass): only x is streamed, which keeps the register
// count low enough for 5-6 resident CTAs per SM; the gradient variants carry up to four operand streams.
template <class T, int A, int V, int UNROLL, bool FWD>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActArgs p) {
    typedef typename Acc<T>::type S;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const int G = FWD ? 0 : p.grad;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b;
    const T* xr = FWD ? nullptr : (const T*)p.xref; const T* yr = FWD ? nullptr : (const T*)p.yref;
    const T* dyp = FWD ? nullptr : (const T*)p.dy;
    T* y = (T*)p.y;
    const long long nvec = p.size_x / V;
    // each CTA walks contiguous chunks of UNROLL * 256 vectors (16 KB of every operand): DRAM-page friendly
    const long long stride = blockDim.x;
    const long long chunk = (long long)blockDim.x * UNROLL;
    long long vi = (long long)blockIdx.x * chunk + threadIdx.x;

    for (; vi - threadIdx.x < nvec; vi += (long long)gridDim.x * chunk) {
        Vec<T, V> vx[UNROLL], vxr[UNROLL], vyr[UNROLL], vdy[UNROLL];
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long w = vi + u * stride;
            if (w < nvec) {
                vx[u] = ld_stream<T, V>(x + w * V, p.hint);
                if (xr) vxr[u] = ld_stream<T, V>(xr + w * V, p.hint);
                if (yr) vyr[u] = ld_stream<T, V>(yr + w * V, p.hint);
                if (dyp) vdy[u] = ld_stream<T, V>(dyp + w * V, p.hint);
            }
        }
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long w = vi + u * stride;
            if (w >= nvec) continue;
            const long long e0 = w * V;
            S bias[V];
            if (p.bias_mode == 0) {
                #pragma unroll
                for (int i = 0; i < V; i++) bias[i] = 0;
            } else if (p.bias_mode == 1) {
                const S bv = to_acc<T>(b[bias_index(p, (unsigned)e0)]);
                #pragma unroll
                for (int i = 0; i < V; i++) bias[i] = bv;
            } else if (p.bias_mode == 2) {
                int bi = (int)((unsigned)e0 - fdiv((unsigned)e0, p.size_m, p.size_s) * (unsigned)p.size_b);
                #pragma unroll
                for (int i = 0; i < V; i++) { bias[i] = to_acc<T>(b[bi]); bi = (bi + 1 == p.size_b) ? 0 : bi + 1; }
            } else {
                #pragma unroll
                for (int i = 0; i < V; i++) bias[i] = to_acc<T>(b[bias_index(p, (unsigned)(e0 + i))]);
            }
            Vec<T, V> out;
            #pragma unroll
            for (int i = 0; i < V; i++) {
                S xv = to_acc<T>(vx[u].v[i]);
                S xrv = xr ? to_acc<T>(vxr[u].v[i]) : (S)0;
                const S yrv = yr ? to_acc<T>(vyr[u].v[i]) : (S)0;
                const S dyv = dyp ? to_acc<T>(vdy[u].v[i]) : (S)1;
                if (G == 0) xv += bias[i]; else xrv += bias[i];
                out.v[i] = from_acc<T>(act_element<A, S>(G, xv, xrv, yrv, dyv, alpha, gain, clamp));
            }
            st_stream<T, V>(y + e0, out, p.hint);
        }
    }

    // scalar tail (size_x not a multiple of V)
    const long long tail0 = nvec * V;
    const long long ti = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ti < p.size_x) {
        S bv = 0;
        if (p.bias_mode) bv = to_acc<T>(b[bias_index(p, (unsigned)ti)]);
        S xv = to_acc<T>(x[ti]);
        S xrv = xr ? to_acc<T>(xr[ti]) : (S)0;
        const S yrv = yr ? to_acc<T>(yr[ti]) : (S)0;
        const S dyv = dyp ? to_acc<T>(dyp[ti]) : (S)1;
        if (G == 0) xv += bv; else xrv += bv;
        y[ti] = from_acc<T>(act_element<A, S>(G, xv, xrv, yrv, dyv, alpha, gain, clamp));
    }
}

template <class T, int V>
static int launch_bias_act(const BiasActArgs& p, int act, cudaStream_t stream) {
    constexpr int UNROLL = 4;
    const long long nvec = p.size_x / V;
    long long blocks = (nvec + 256LL * UNROLL - 1) / (256LL * UNROLL);
    if (blocks < 1) blocks = 1;
    void (*k)(BiasActArgs) = nullptr;
    const bool fwd = p.grad == 0 && !p.xref && !p.yref && !p.dy;
#define PGPP_BA_CASE(A) case A: k = fwd ? bias_act_kernel<T, A, V, UNROLL, true> : bias_act_kernel<T, A, V, UNROLL, false>; break;
    switch (act) {
        PGPP_BA_CASE(1) PGPP_BA_CASE(2) PGPP_BA_CASE(3) PGPP_BA_CASE(4) PGPP_BA_CASE(5)
        PGPP_BA_CASE(6) PGPP_BA_CASE(7) PGPP_BA_CASE(8) PGPP_BA_CASE(9)
        default: set_error("no CUDA kernel found for the specified activation func"); return PGPP_ERR_UNSUPPORTED;
    }
#undef PGPP_BA_CASE
    const long long cap = (long long)sm_count() * occupancy_of(k, 256, 0);     // exactly one resident wave, grid-stride inside
    if (blocks > cap) blocks = cap;
    k<<<(unsigned)blocks, 256, 0, stream>>>(p);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

template <class T>
static int dispatch_bias_act(BiasActArgs p, int act, cudaStream_t stream) {
    constexpr int V = 16 / sizeof(T);
    const uintptr_t all = (uintptr_t)p.x | (uintptr_t)p.y | (uintptr_t)p.xref | (uintptr_t)p.yref | (uintptr_t)p.dy;
    const bool aligned = (all & 15) == 0;
    if (!aligned) {
        if (p.bias_mode) p.bias_mode = 3;
        return launch_bias_act<T, 1>(p, act, stream);
    }
    if (p.bias_mode) {
        if (p.step_b == 1) p.bias_mode = 2;
        else if (p.step_b % V == 0) p.bias_mode = 1;
        else p.bias_mode = 3;
    }
    return launch_bias_act<T, V>(p, act, stream);
}

} // namespace pgpp

extern "C" int pgpp_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                             int64_t size_x, int64_t size_b, int64_t step_b, int dtype, int grad, int act,
                             float alpha, float gain, float clamp, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(size_x >= 0 && size_x <= 2147483647LL, "x is too large");
    PGPP_REQUIRE(grad >= 0 && grad <= 2, "grad must be 0, 1 or 2");
    PGPP_REQUIRE(b == nullptr || (size_b >= 1 && step_b >= 1), "b has wrong number of elements");
    if (size_x == 0) return PGPP_OK;
    PGPP_REQUIRE(x != nullptr && y != nullptr, "x and y must be device pointers");
    BiasActArgs p;
    p.x = x; p.b = b; p.xref = xref; p.yref = yref; p.dy = dy; p.y = y;
    p.size_x = size_x; p.size_b = b ? (int)size_b : 1; p.step_b = b ? step_b : 1;
    p.grad = grad; p.alpha = alpha; p.gain = gain; p.clamp = clamp;
    p.bias_mode = b ? 1 : 0;
    p.hint = env_flags().ba_nostream ? 0 : 1;
    auto magic = [](unsigned dv, unsigned& m, unsigned& sft) {
        sft = 0; while ((1ull << sft) < dv) sft++;
        m = (unsigned)(((1ull << (31 + sft)) + dv - 1) / dv);
    };
    PGPP_REQUIRE(p.step_b <= 2147483647LL, "bias stride is too large");
    magic((unsigned)p.step_b, p.step_m, p.step_s);
    magic((unsigned)p.size_b, p.size_m, p.size_s);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case PGPP_F32:  return dispatch_bias_act<float>(p, act, s);
        case PGPP_F16:  return dispatch_bias_act<__half>(p, act, s);
        case PGPP_BF16: return dispatch_bias_act<__nv_bfloat16>(p, act, s);
        case PGPP_F64:  return dispatch_bias_act<double>(p, act, s);
    }
    set_error("unsupported dtype %d", dtype);
    return PGPP_ERR_UNSUPPORTED;
}
