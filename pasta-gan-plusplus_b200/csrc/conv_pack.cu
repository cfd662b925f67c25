// Operand preparation for the tensor-core convolution (sm_100a).
//   pgpp_pack_activations : any-layout x (* per-(n,c) scale) -> channels-innermost bf16 parts
//   pgpp_modconv_demod_coefs : d[n,o] = rsqrt(sum (w*s)^2 + eps)     (training/networks.py:64-68)
// Both are small HBM-bound helpers; the heavy lifting is conv_igemm.cu.
#include "common.cuh"

namespace pgpp {

struct PackArgs {
    const void* x; const float* scale; __nv_bfloat16* out;
    int n, c, h, w, c_pad, parts;
    int c_total, c_off;         // destination pixel stride and first channel
    long long s_n, s_c, s_h, s_w;
    long long part_stride;      // elements between parts = n*h*w*c_pad
};

__device__ __forceinline__ void split_store(float v, __nv_bfloat16* dst, long long part_stride, int parts) {
    // part p = bf16(v - sum of earlier parts): 8, 16, 24 significand bits for 1, 2, 3 parts
    #pragma unroll 3
    for (int p = 0; p < parts; p++) {
        const __nv_bfloat16 q = __float2bfloat16_rn(v);
        dst[p * part_stride] = q;
        v -= __bfloat162float(q);
    }
}

// x has unit stride along W (NCHW-like): transpose 64 channels x 32 pixels through shared memory so
// that both the read (along W) and the write (along C) are coalesced.
template <class T>
__global__ void __launch_bounds__(256) pack_nchw_kernel(PackArgs p, int w_tiles, int c_tiles) {
    __shared__ float tile[64][33];
    long long b = blockIdx.x;
    const int wt = (int)(b % w_tiles); b /= w_tiles;
    const int ct = (int)(b % c_tiles); b /= c_tiles;
    const int y = (int)(b % p.h);
    const int n = (int)(b / p.h);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = wt * 32, c0 = ct * 64;
    const T* src = (const T*)p.x + n * p.s_n + y * p.s_h;
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = c0 + warp * 8 + i;
        float v = 0.f;
        if (c < p.c && x0 + lane < p.w) {
            v = (float)to_acc<T>(src[c * p.s_c + (x0 + lane)]);
            if (p.scale) v *= p.scale[n * p.c + c];
        }
        tile[warp * 8 + i][lane] = v;
    }
    __syncthreads();
    // write phase: a warp covers 4 pixels x 8 groups of 8 channels, so every pixel's 128-byte channel row is written
    // by 8 consecutive lanes with one 128-bit store per part
    #pragma unroll
    for (int it = 0; it < 1; it++) {
        const int px = warp * 4 + (lane >> 3);
        const int cg = lane & 7;
        if (x0 + px < p.w && c0 + cg * 8 < p.c_pad) {
            float v[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) v[j] = tile[cg * 8 + j][px];
            __nv_bfloat16* dst = p.out + (((long long)n * p.h + y) * p.w + (x0 + px)) * p.c_total + p.c_off + c0 + cg * 8;
            for (int part = 0; part < p.parts; part++) {
                __align__(16) __nv_bfloat16 q[8];
                #pragma unroll
                for (int j = 0; j < 8; j++) { q[j] = __float2bfloat16_rn(v[j]); v[j] -= __bfloat162float(q[j]); }
                *reinterpret_cast<int4*>(dst + part * p.part_stride) = *reinterpret_cast<const int4*>(q);
            }
        }
    }
}

// any strides (channels_last inputs are coalesced here): one thread per (pixel, channel)
template <class T>
__global__ void __launch_bounds__(256) pack_generic_kernel(PackArgs p, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % p.c_pad);
        long long r = e / p.c_pad;
        const int x = (int)(r % p.w); r /= p.w;
        const int y = (int)(r % p.h);
        const int n = (int)(r / p.h);
        float v = 0.f;
        if (c < p.c) {
            v = (float)to_acc<T>(((const T*)p.x)[n * p.s_n + c * p.s_c + y * p.s_h + x * p.s_w]);
            if (p.scale) v *= p.scale[n * p.c + c];
        }
        split_store(v, p.out + (e / p.c_pad) * p.c_total + p.c_off + c, p.part_stride, p.parts);
    }
}

template <class T>
static int launch_pack(const PackArgs& p, cudaStream_t stream) {
    if (p.s_w == 1 && p.s_c != 1) {
        const int w_tiles = (p.w + 31) / 32, c_tiles = (p.c_pad + 63) / 64;
        const long long blocks = (long long)w_tiles * c_tiles * p.h * p.n;
        PGPP_REQUIRE(blocks <= 2147483647LL, "activation tensor too large to pack");
        pack_nchw_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p, w_tiles, c_tiles);
    } else {
        const long long total = (long long)p.n * p.h * p.w * p.c_pad;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)sm_count() * occupancy_of(pack_generic_kernel<T>, 256, 0);
        if (blocks > cap) blocks = cap;
        pack_generic_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p, total);
    }
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

struct Im2colArgs {
    const void* x; const float* scale; __nv_bfloat16* out;
    int n, c, h, w, hp, kw, r, pad_x, pad_y, parts;
    long long s_n, s_c, s_h, s_w, part_stride;
};

// one thread per (packed pixel, group of 8 channels): gathers up to 8 shifted input samples (L1/L2-cached reads of a tiny
// tensor) and writes one 128-bit store per part; 8 consecutive threads cover the 128-byte channel row of a pixel
template <class T>
__global__ void __launch_bounds__(256) im2col_kernel(Im2colArgs p, long long total) {
    const int used = p.r * p.kw * p.c;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(e & 7);
        long long q = e >> 3;
        const int x = (int)(q % p.w); q /= p.w;
        const int yy = (int)(q % p.hp);
        const int n = (int)(q / p.hp);
        float v[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const int ch = cg * 8 + j;
            float val = 0.f;
            if (ch < used) {
                const int c = ch % p.c;
                const int t = ch / p.c;
                const int kx = t % p.kw, ry = t / p.kw;
                const int iy = yy - p.pad_y + ry, ix = x + kx - p.pad_x;
                if (iy >= 0 && iy < p.h && ix >= 0 && ix < p.w) {
                    val = (float)to_acc<T>(((const T*)p.x)[n * p.s_n + c * p.s_c + iy * p.s_h + ix * p.s_w]);
                    if (p.scale) val *= p.scale[n * p.c + c];
                }
            }
            v[j] = val;
        }
        __nv_bfloat16* dst = p.out + (((long long)n * p.hp + yy) * p.w + x) * 64 + cg * 8;
        for (int part = 0; part < p.parts; part++) {
            __align__(16) __nv_bfloat16 qv[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) { qv[j] = __float2bfloat16_rn(v[j]); v[j] -= __bfloat162float(qv[j]); }
            *reinterpret_cast<int4*>(dst + part * p.part_stride) = *reinterpret_cast<const int4*>(qv);
        }
    }
}

struct SpadeArgs {
    const float* x; const float* mean; const float* rstd; const float* gamma; const float* beta; __nv_bfloat16* out;
    int n, c, h, w, c_pad, parts; long long gb_stride_n, part_stride; float pre_gain;
};

// same 64-channel x 32-pixel transpose tile as pack_nchw_kernel, with the SPADE arithmetic applied on the way in
__global__ void __launch_bounds__(256) spade_pack_kernel(SpadeArgs p, int w_tiles, int c_tiles) {
    __shared__ float tile[64][33];
    long long b = blockIdx.x;
    const int wt = (int)(b % w_tiles); b /= w_tiles;
    const int ct = (int)(b % c_tiles); b /= c_tiles;
    const int y = (int)(b % p.h);
    const int n = (int)(b / p.h);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = wt * 32, c0 = ct * 64;
    const long long plane = (long long)p.h * p.w;
    #pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = c0 + warp * 8 + i;
        float v = 0.f;
        if (c < p.c && x0 + lane < p.w) {
            const long long off = (long long)c * plane + (long long)y * p.w + (x0 + lane);
            const float xv = p.x[(long long)n * p.c * plane + off];
            const float g = p.gamma[n * p.gb_stride_n + off], bt = p.beta[n * p.gb_stride_n + off];
            const float nv = (xv - p.mean[n * p.c + c]) * p.rstd[n * p.c + c];
            v = fmaf(nv, 1.f + g, bt);
            if (p.pre_gain > 0.f) v = fmaxf(v, 0.f) * p.pre_gain;
        }
        tile[warp * 8 + i][lane] = v;
    }
    __syncthreads();
    const int px = warp * 4 + (lane >> 3);
    const int cg = lane & 7;
    if (x0 + px < p.w && c0 + cg * 8 < p.c_pad) {
        float v[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) v[j] = tile[cg * 8 + j][px];
        __nv_bfloat16* dst = p.out + (((long long)n * p.h + y) * p.w + (x0 + px)) * p.c_pad + c0 + cg * 8;
        for (int part = 0; part < p.parts; part++) {
            __align__(16) __nv_bfloat16 q[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) { q[j] = __float2bfloat16_rn(v[j]); v[j] -= __bfloat162float(q[j]); }
            *reinterpret_cast<int4*>(dst + part * p.part_stride) = *reinterpret_cast<const int4*>(q);
        }
    }
}

// one CTA per output channel: W2[i] = sum_t w[o,i,t]^2 in shared memory, then one warp per sample
__global__ void __launch_bounds__(256) demod_kernel(const float* __restrict__ w, const float* __restrict__ s,
                                                    float* __restrict__ d, int n, int o, int ic, int taps, float eps) {
    extern __shared__ float w2[];
    const int oc = blockIdx.x;
    const float* wp = w + (long long)oc * ic * taps;
    for (int i = threadIdx.x; i < ic; i += blockDim.x) {
        float a = 0.f;
        for (int t = 0; t < taps; t++) { const float v = wp[i * taps + t]; a = fmaf(v, v, a); }
        w2[i] = a;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = warp; b < n; b += blockDim.x / 32) {
        float a = 0.f;
        for (int i = lane; i < ic; i += 32) { const float sv = s[(long long)b * ic + i]; a = fmaf(sv * sv, w2[i], a); }
        #pragma unroll
        for (int m = 16; m > 0; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
        if (lane == 0) d[(long long)b * o + oc] = rsqrtf(a + eps);
    }
}

} // namespace pgpp

extern "C" int pgpp_pack_activations(const void* x, const int64_t size[4], const int64_t stride[4], int dtype,
                                     const float* scale, void* out, int c_pad, int parts, void* stream) {
    return pgpp_pack_activations_slice(x, size, stride, dtype, scale, out, c_pad, c_pad, 0, parts, stream);
}

namespace pgpp {
__global__ void __launch_bounds__(256) modulate_weights_kernel(const float* __restrict__ master, const float* __restrict__ s,
                                                               __nv_bfloat16* __restrict__ out, long long rows, int c_pad, int c_in,
                                                               int parts, long long per_sample_part) {
    // one thread per 8 consecutive channels of one (sample, row): 128-bit stores per part
    const int groups = c_pad / 8;
    const long long total = rows * groups;
    const int n = blockIdx.y;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long row = e / groups;
        const int c0 = (int)(e - row * groups) * 8;
        float v[8];
        #pragma unroll
        for (int j = 0; j < 8; j++) {
            const int c = c0 + j;
            v[j] = c < c_in ? master[row * c_pad + c] * s[(long long)n * c_in + c] : 0.f;
        }
        __nv_bfloat16* dst = out + (long long)n * parts * per_sample_part + row * c_pad + c0;
        for (int part = 0; part < parts; part++) {
            __align__(16) __nv_bfloat16 q[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) { q[j] = __float2bfloat16_rn(v[j]); v[j] -= __bfloat162float(q[j]); }
            *reinterpret_cast<int4*>(dst + part * per_sample_part) = *reinterpret_cast<const int4*>(q);
        }
    }
}
} // namespace pgpp

extern "C" int pgpp_modulate_weights(const float* master, const float* s, void* out, int n, int64_t rows, int c_pad, int c_in,
                                     int parts, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(master && s && out, "master, s and out must be device pointers");
    PGPP_REQUIRE(n >= 1 && n <= 65535 && rows >= 1 && c_pad % 16 == 0 && c_in >= 1 && c_in <= c_pad && parts >= 1 && parts <= 3, "bad modulate_weights arguments");
    const long long total = rows * (c_pad / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    modulate_weights_kernel<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(master, s, (__nv_bfloat16*)out, rows, c_pad, c_in,
                                                                                                  parts, rows * c_pad);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

extern "C" int pgpp_pack_activations_slice(const void* x, const int64_t size[4], const int64_t stride[4], int dtype,
                                           const float* scale, void* out, int c_pad, int c_total, int c_off, int parts, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(x && out, "x and out must be device pointers");
    PGPP_REQUIRE(parts >= 1 && parts <= 3, "parts must be 1, 2 or 3");
    PGPP_REQUIRE(c_pad >= size[1] && c_pad % 16 == 0, "c_pad must be a multiple of 16 and >= C");
    PGPP_REQUIRE(c_total >= c_off + c_pad && c_total % 8 == 0 && c_off % 8 == 0 && c_off >= 0, "bad destination channel slice");
    PackArgs p;
    p.c_total = c_total; p.c_off = c_off;
    p.x = x; p.scale = scale; p.out = (__nv_bfloat16*)out;
    p.n = (int)size[0]; p.c = (int)size[1]; p.h = (int)size[2]; p.w = (int)size[3];
    p.c_pad = c_pad; p.parts = parts;
    p.s_n = stride[0]; p.s_c = stride[1]; p.s_h = stride[2]; p.s_w = stride[3];
    p.part_stride = (long long)p.n * p.h * p.w * c_total;
    if (p.part_stride == 0) return PGPP_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case PGPP_F32:  return launch_pack<float>(p, s);
        case PGPP_F16:  return launch_pack<__half>(p, s);
        case PGPP_BF16: return launch_pack<__nv_bfloat16>(p, s);
        case PGPP_F64:  return launch_pack<double>(p, s);
    }
    set_error("unsupported dtype %d", dtype);
    return PGPP_ERR_UNSUPPORTED;
}

extern "C" int pgpp_pack_im2col(const void* x, const int64_t size[4], const int64_t stride[4], int dtype, const float* scale,
                                void* out, int kw, int r, int pad_x, int pad_y, int parts, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(x && out, "x and out must be device pointers");
    PGPP_REQUIRE(parts >= 1 && parts <= 3 && kw >= 1 && r >= 1 && pad_x >= 0 && pad_y >= 0, "bad im2col arguments");
    PGPP_REQUIRE((long long)r * kw * size[1] <= 64, "im2col packing needs r*kw*C <= 64");
    Im2colArgs p;
    p.x = x; p.scale = scale; p.out = (__nv_bfloat16*)out;
    p.n = (int)size[0]; p.c = (int)size[1]; p.h = (int)size[2]; p.w = (int)size[3]; p.hp = p.h + pad_y;
    p.kw = kw; p.r = r; p.pad_x = pad_x; p.pad_y = pad_y; p.parts = parts;
    p.s_n = stride[0]; p.s_c = stride[1]; p.s_h = stride[2]; p.s_w = stride[3];
    p.part_stride = (long long)p.n * p.hp * p.w * 64;
    const long long total = (long long)p.n * p.hp * p.w * 8;
    if (total == 0) return PGPP_OK;
    long long blocks = (total + 255) / 256;
    cudaStream_t s = (cudaStream_t)stream;
#define PGPP_IM2COL(T) { const long long cap = (long long)sm_count() * occupancy_of(im2col_kernel<T>, 256, 0); \
                         if (blocks > cap) blocks = cap; im2col_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(p, total); }
    switch (dtype) {
        case PGPP_F32:  PGPP_IM2COL(float) break;
        case PGPP_F16:  PGPP_IM2COL(__half) break;
        case PGPP_BF16: PGPP_IM2COL(__nv_bfloat16) break;
        case PGPP_F64:  PGPP_IM2COL(double) break;
        default: set_error("unsupported dtype %d", dtype); return PGPP_ERR_UNSUPPORTED;
    }
#undef PGPP_IM2COL
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

extern "C" int pgpp_spade_modulate_pack(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                        int64_t gb_stride_n, void* out, int n, int c, int h, int w, int c_pad, int parts, float pre_gain, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(x && mean && rstd && gamma && beta && out, "null pointer");
    PGPP_REQUIRE(n >= 1 && c >= 1 && h >= 1 && w >= 1 && c_pad >= c && c_pad % 16 == 0 && parts >= 1 && parts <= 3, "bad spade_modulate_pack arguments");
    SpadeArgs p;
    p.x = x; p.mean = mean; p.rstd = rstd; p.gamma = gamma; p.beta = beta; p.out = (__nv_bfloat16*)out;
    p.n = n; p.c = c; p.h = h; p.w = w; p.c_pad = c_pad; p.parts = parts; p.gb_stride_n = gb_stride_n;
    p.part_stride = (long long)n * h * w * c_pad; p.pre_gain = pre_gain;
    const int w_tiles = (w + 31) / 32, c_tiles = (c_pad + 63) / 64;
    const long long blocks = (long long)w_tiles * c_tiles * h * n;
    PGPP_REQUIRE(blocks <= 2147483647LL, "tensor too large");
    spade_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, w_tiles, c_tiles);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

extern "C" int pgpp_modconv_demod_coefs(const float* w, const float* s, float* d, int n, int o, int i, int taps,
                                        float eps, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(w && s && d, "w, s and d must be device pointers");
    PGPP_REQUIRE(n >= 1 && o >= 1 && i >= 1 && taps >= 1, "empty problem");
    PGPP_REQUIRE((size_t)i * sizeof(float) <= 200 * 1024, "too many input channels");
    const size_t smem = (size_t)i * sizeof(float);
    if (smem > 48 * 1024)
        PGPP_CUDA_OK(cudaFuncSetAttribute(demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    demod_kernel<<<o, 256, smem, (cudaStream_t)stream>>>(w, s, d, n, o, i, taps, eps);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}
