// Plane reductions of the differentiable fused modulated convolution (training/networks.py:73-82 differentiated by hand):
//   r[n,c]    = sum_{hw} a[n,c,hw] * (b[n,c,hw] - sub[n,hw])          (sub optional)
//   out[n,c,hw] = a[n,c,hw] * scale[n,c]                               (optional, may alias a)
// With (a, b, sub) = (grad_y, y, noise) this is the demodulation-coefficient gradient d(loss)/d(dcoef) * dcoef; with
// (a, b, scale) = (grad_xs, x, styles) it is the style gradient and the input gradient grad_x = grad_xs * s in the same pass.
// HBM-bound: every element is read once (a and b) and written at most once; float32 NCHW-contiguous planes, 128-bit accesses.
#include "common.cuh"

namespace pgpp {

struct MulReduceArgs {
    const float* a; const float* b; const float* sub; long long sub_stride_n;
    const float* scale; float* out; float* r;
    int c; long long hw; long long planes;
};

// G threads cooperate on one plane (G = 256: one plane per CTA; G = 32: eight planes per CTA for small images)
template <int G>
__global__ void __launch_bounds__(256) mul_reduce_kernel(MulReduceArgs p) {
    __shared__ float s_part[8];
    const int sub_id = threadIdx.x / G, lane_g = threadIdx.x % G;
    const long long plane = (long long)blockIdx.x * (256 / G) + sub_id;
    float acc = 0.f;
    if (plane < p.planes) {
        const int n = (int)(plane / p.c);
        const float* a = p.a + plane * p.hw;
        const float* b = p.b ? p.b + plane * p.hw : nullptr;
        const float* sb = p.sub ? p.sub + n * p.sub_stride_n : nullptr;
        float* o = p.out ? p.out + plane * p.hw : nullptr;
        const float sc = p.scale ? p.scale[plane] : 1.f;
        const bool vec = (p.hw % 4 == 0) && (((uintptr_t)p.a | (uintptr_t)p.b | (uintptr_t)p.sub | (uintptr_t)p.out) & 15) == 0 &&
                         (p.sub_stride_n % 4 == 0);
        if (vec) {
            const long long n4 = p.hw >> 2;
            for (long long i = lane_g; i < n4; i += G) {
                const float4 av = reinterpret_cast<const float4*>(a)[i];
                float4 bv = b ? reinterpret_cast<const float4*>(b)[i] : make_float4(1.f, 1.f, 1.f, 1.f);
                if (sb) { const float4 sv = reinterpret_cast<const float4*>(sb)[i]; bv.x -= sv.x; bv.y -= sv.y; bv.z -= sv.z; bv.w -= sv.w; }
                acc = fmaf(av.x, bv.x, acc); acc = fmaf(av.y, bv.y, acc); acc = fmaf(av.z, bv.z, acc); acc = fmaf(av.w, bv.w, acc);
                if (o) reinterpret_cast<float4*>(o)[i] = make_float4(av.x * sc, av.y * sc, av.z * sc, av.w * sc);
            }
        } else {
            for (long long i = lane_g; i < p.hw; i += G) {
                const float av = a[i];
                float bv = b ? b[i] : 1.f;
                if (sb) bv -= sb[i];
                acc = fmaf(av, bv, acc);
                if (o) o[i] = av * sc;
            }
        }
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (G == 32) {
        if (lane_g == 0 && plane < p.planes && p.r) p.r[plane] = acc;
    } else {
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0 && plane < p.planes && p.r) {
            float t = 0.f;
            #pragma unroll
            for (int w = 0; w < 8; w++) t += s_part[w];
            p.r[plane] = t;
        }
    }
}

// r[plane] = sum over the plane's hw elements (float32 accumulation), planes NCHW-contiguous; float32 / float16 / bfloat16.
// The bias gradient of bias_act / conv2d (bias_act.py:135, conv2d_gradfix.py:130) is sum_n of this: one pass over the tensor at the
// HBM rate, 16-byte loads (the library's generic reduction reaches a third of that for this access pattern).
template <class T, int G>
__global__ void __launch_bounds__(256) sum_hw_kernel(const T* __restrict__ a, float* __restrict__ r, long long hw, long long planes) {
    __shared__ float s_part[8];
    constexpr int V = 16 / sizeof(T);
    const int sub_id = threadIdx.x / G, lane_g = threadIdx.x % G;
    const long long plane = (long long)blockIdx.x * (256 / G) + sub_id;
    float acc = 0.f;
    if (plane < planes) {
        const T* ap = a + plane * hw;
        if ((hw % V == 0) && (((uintptr_t)a & 15) == 0)) {
            const long long nv = hw / V;
            for (long long i = lane_g; i < nv; i += G) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(ap) + i);
                const T* e = reinterpret_cast<const T*>(&q);
                #pragma unroll
                for (int k = 0; k < V; k++) acc += (float)to_acc<T>(e[k]);
            }
        } else {
            for (long long i = lane_g; i < hw; i += G) acc += (float)to_acc<T>(ap[i]);
        }
    }
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (G == 32) {
        if (lane_g == 0 && plane < planes) r[plane] = acc;
    } else {
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0 && plane < planes) {
            float t = 0.f;
            #pragma unroll
            for (int w = 0; w < 8; w++) t += s_part[w];
            r[plane] = t;
        }
    }
}

template <class T>
static int launch_sum_hw(const void* a, float* r, long long hw, long long planes, cudaStream_t stream) {
    if (hw >= 2048) sum_hw_kernel<T, 256><<<(unsigned)planes, 256, 0, stream>>>((const T*)a, r, hw, planes);
    else sum_hw_kernel<T, 32><<<(unsigned)((planes + 7) / 8), 256, 0, stream>>>((const T*)a, r, hw, planes);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

} // namespace pgpp

extern "C" int pgpp_sum_hw(const void* a, int dtype, float* r, int n, int c, int64_t hw, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(a && r, "a and r must be device pointers");
    PGPP_REQUIRE(n >= 1 && c >= 1 && hw >= 1, "empty problem");
    const long long planes = (long long)n * c;
    PGPP_REQUIRE(planes <= 2147483647LL, "too many planes");
    switch (dtype) {
        case PGPP_F32:  return launch_sum_hw<float>(a, r, hw, planes, (cudaStream_t)stream);
        case PGPP_F16:  return launch_sum_hw<__half>(a, r, hw, planes, (cudaStream_t)stream);
        case PGPP_BF16: return launch_sum_hw<__nv_bfloat16>(a, r, hw, planes, (cudaStream_t)stream);
    }
    set_error("sum_hw: unsupported dtype %d", dtype);
    return PGPP_ERR_UNSUPPORTED;
}

extern "C" int pgpp_mul_reduce_hw(const float* a, const float* b, const float* sub, int64_t sub_stride_n, const float* scale,
                                  float* out_scaled, float* r, int n, int c, int64_t hw, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(a && (r || out_scaled), "a and at least one of r / out_scaled must be device pointers");
    PGPP_REQUIRE(n >= 1 && c >= 1 && hw >= 1, "empty problem");
    PGPP_REQUIRE(!out_scaled || scale, "out_scaled needs scale");
    MulReduceArgs p{a, b, sub, sub_stride_n, scale, out_scaled, r, c, hw, (long long)n * c};
    if (hw >= 2048) {
        PGPP_REQUIRE(p.planes <= 2147483647LL, "too many planes");
        mul_reduce_kernel<256><<<(unsigned)p.planes, 256, 0, (cudaStream_t)stream>>>(p);
    } else {
        const long long blocks = (p.planes + 7) / 8;
        mul_reduce_kernel<32><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    }
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}
