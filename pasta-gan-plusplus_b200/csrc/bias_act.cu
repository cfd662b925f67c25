// bias_act for sm_100a: HBM-bound streaming kernel.  128-bit vector loads/stores, several
// independent vectors in flight per thread, grid sized to the SM count (grid-stride loop),
// bias index computed once per vector instead of a div+mod per element.
// Replaces torch_utils/ops/bias_act.cu:23-147 + bias_act.cpp:32-90 of the reference.
#include "act.cuh"

namespace pgpp {

struct BiasActArgs {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    long long size_x; int size_b; long long step_b;
    int grad; float alpha, gain, clamp;
    int bias_mode;      // 0 none, 1 same bias for the whole vector, 2 consecutive (step_b == 1), 3 per element
};

template <class T, int V> struct alignas(sizeof(T) * V) Vec { T v[V]; };

template <class T, int V>
__device__ __forceinline__ Vec<T, V> ld_stream(const T* p) {
    // streaming read: every element is touched exactly once
    Vec<T, V> r;
    if (sizeof(T) * V == 16) {
        const int4 q = __ldcs(reinterpret_cast<const int4*>(p));
        r = *reinterpret_cast<const Vec<T, V>*>(&q);
    } else {
        #pragma unroll
        for (int i = 0; i < V; i++) r.v[i] = p[i];
    }
    return r;
}

template <class T, int V>
__device__ __forceinline__ void st_stream(T* p, const Vec<T, V>& r) {
    if (sizeof(T) * V == 16) __stcs(reinterpret_cast<int4*>(p), *reinterpret_cast<const int4*>(&r));
    else {
        #pragma unroll
        for (int i = 0; i < V; i++) p[i] = r.v[i];
    }
}

template <class T, int A, int V, int UNROLL>
__global__ void __launch_bounds__(256) bias_act_kernel(BiasActArgs p) {
    typedef typename Acc<T>::type S;
    const S alpha = (S)p.alpha, gain = (S)p.gain, clamp = (S)p.clamp;
    const int G = p.grad;
    const T* x = (const T*)p.x; const T* b = (const T*)p.b;
    const T* xr = (const T*)p.xref; const T* yr = (const T*)p.yref; const T* dyp = (const T*)p.dy;
    T* y = (T*)p.y;
    const long long nvec = p.size_x / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long vi = (long long)blockIdx.x * blockDim.x + threadIdx.x;

    for (; vi < nvec; vi += stride * UNROLL) {
        Vec<T, V> vx[UNROLL], vxr[UNROLL], vyr[UNROLL], vdy[UNROLL];
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const long long w = vi + u * stride;
            if (w < nvec) {
                vx[u] = ld_stream<T, V>(x + w * V);
                if (xr) vxr[u] = ld_stream<T, V>(xr + w * V);
                if (yr) vyr[u] = ld_stream<T, V>(yr + w * V);
                if (dyp) vdy[u] = ld_stream<T, V>(dyp + w * V);
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
                const S bv = to_acc<T>(b[(e0 / p.step_b) % p.size_b]);
                #pragma unroll
                for (int i = 0; i < V; i++) bias[i] = bv;
            } else if (p.bias_mode == 2) {
                int bi = (int)(e0 % p.size_b);
                #pragma unroll
                for (int i = 0; i < V; i++) { bias[i] = to_acc<T>(b[bi]); bi = (bi + 1 == p.size_b) ? 0 : bi + 1; }
            } else {
                #pragma unroll
                for (int i = 0; i < V; i++) bias[i] = to_acc<T>(b[((e0 + i) / p.step_b) % p.size_b]);
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
            st_stream<T, V>(y + e0, out);
        }
    }

    // scalar tail (size_x not a multiple of V)
    const long long tail0 = nvec * V;
    const long long ti = tail0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ti < p.size_x) {
        S bv = 0;
        if (p.bias_mode) bv = to_acc<T>(b[(ti / p.step_b) % p.size_b]);
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
    const long long cap = (long long)sm_count() * 8;      // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    void (*k)(BiasActArgs) = nullptr;
    switch (act) {
        case 1: k = bias_act_kernel<T, 1, V, UNROLL>; break;
        case 2: k = bias_act_kernel<T, 2, V, UNROLL>; break;
        case 3: k = bias_act_kernel<T, 3, V, UNROLL>; break;
        case 4: k = bias_act_kernel<T, 4, V, UNROLL>; break;
        case 5: k = bias_act_kernel<T, 5, V, UNROLL>; break;
        case 6: k = bias_act_kernel<T, 6, V, UNROLL>; break;
        case 7: k = bias_act_kernel<T, 7, V, UNROLL>; break;
        case 8: k = bias_act_kernel<T, 8, V, UNROLL>; break;
        case 9: k = bias_act_kernel<T, 9, V, UNROLL>; break;
        default: set_error("no CUDA kernel found for the specified activation func"); return PGPP_ERR_UNSUPPORTED;
    }
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
