// Input / output edge of the try-on inference loop (reference test.py:126-147 and :162-166), as HBM-bound kernels:
//   pgpp_u8_to_f32    uint8 planes -> float32 `x / 127.5 - 1` (or a plain cast for masks) written into a channel slice of a
//                     wider NCHW tensor (the torch.cat of test.py:135,145,146 needs no copy), optionally composed with the
//                     retain mask: `x * m - (1 - m)` (test.py:144)
//   pgpp_image_to_u8  float32 NCHW image -> uint8 NHWC with `clip((x + 1) * 127.5, 0, 255)` truncated, channels optionally
//                     reversed (RGB -> BGR for cv2.imwrite, test.py:162-166)
// Every arithmetic step is a separately rounded IEEE operation (no FMA contraction) so the results are bit-identical to
// the reference's torch / numpy expressions.
#include "common.cuh"

namespace pgpp {

__device__ __forceinline__ float edge_value(unsigned v, int normalize, bool has_mask, float m) {
    float x = (float)v;
    if (normalize) x = __fsub_rn(__fdiv_rn(x, 127.5f), 1.0f);
    if (has_mask) x = __fsub_rn(__fmul_rn(x, m), __fsub_rn(1.0f, m));
    return x;
}

// one thread per 16 consecutive pixels of one (n, c) plane
template <bool VEC>
__global__ void __launch_bounds__(256)
u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const float* __restrict__ mask, long long planes, int c,
                 long long hw, long long dst_c_total, long long c_off, int normalize) {
    const long long groups = (hw + 15) / 16;
    const long long total = planes * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long plane = i / groups;
        const long long p0 = (i - plane * groups) * 16;
        const long long n = plane / c, ch = plane - n * c;
        const uint8_t* s = src + plane * hw + p0;
        float* d = dst + (n * dst_c_total + c_off + ch) * hw + p0;
        const float* m = mask ? mask + n * hw + p0 : nullptr;
        if (VEC) {
            const uint4 q = *reinterpret_cast<const uint4*>(s);
            const unsigned w[4] = {q.x, q.y, q.z, q.w};
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                float4 mv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m) mv = *reinterpret_cast<const float4*>(m + 4 * j);
                float4 o;
                o.x = edge_value(w[j] & 255u, normalize, m != nullptr, mv.x);
                o.y = edge_value((w[j] >> 8) & 255u, normalize, m != nullptr, mv.y);
                o.z = edge_value((w[j] >> 16) & 255u, normalize, m != nullptr, mv.z);
                o.w = edge_value(w[j] >> 24, normalize, m != nullptr, mv.w);
                *reinterpret_cast<float4*>(d + 4 * j) = o;
            }
        } else {
            for (int j = 0; j < 16 && p0 + j < hw; j++) d[j] = edge_value(s[j], normalize, m != nullptr, m ? m[j] : 0.f);
        }
    }
}

// one thread per 4 consecutive pixels: C planes in, 4 * C bytes out
template <bool VEC>
__global__ void __launch_bounds__(256)
image_to_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, long long n, int c, long long hw, int reverse) {
    const long long groups = (hw + 3) / 4;
    const long long total = n * groups;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / groups;
        const long long p0 = (i - b * groups) * 4;
        const float* src = img + b * c * hw + p0;
        uint8_t* dst = out + (b * hw + p0) * c;
        auto cvt = [](float x) -> unsigned {
            float v = __fmul_rn(__fadd_rn(x, 1.0f), 127.5f);
            v = fminf(fmaxf(v, 0.f), 255.f);
            return (unsigned)v;                     // truncation, like ndarray.astype(np.uint8) on the clipped value
        };
        if (VEC) {                                  // c == 3, hw % 4 == 0: 12 output bytes = three aligned 32-bit words
            float4 p[3];
            #pragma unroll
            for (int ch = 0; ch < 3; ch++) p[ch] = *reinterpret_cast<const float4*>(src + (reverse ? 2 - ch : ch) * hw);
            const unsigned b0[3] = {cvt(p[0].x), cvt(p[1].x), cvt(p[2].x)}, b1[3] = {cvt(p[0].y), cvt(p[1].y), cvt(p[2].y)};
            const unsigned b2[3] = {cvt(p[0].z), cvt(p[1].z), cvt(p[2].z)}, b3[3] = {cvt(p[0].w), cvt(p[1].w), cvt(p[2].w)};
            unsigned* o = reinterpret_cast<unsigned*>(dst);
            o[0] = b0[0] | (b0[1] << 8) | (b0[2] << 16) | (b1[0] << 24);
            o[1] = b1[1] | (b1[2] << 8) | (b2[0] << 16) | (b2[1] << 24);
            o[2] = b2[2] | (b3[0] << 8) | (b3[1] << 16) | (b3[2] << 24);
        } else {
            for (int j = 0; j < 4 && p0 + j < hw; j++)
                for (int ch = 0; ch < c; ch++)
                    dst[j * c + ch] = (uint8_t)cvt(src[(reverse ? c - 1 - ch : ch) * hw + j]);
        }
    }
}

} // namespace pgpp

extern "C" int pgpp_u8_to_f32(const void* src, int64_t n, int64_t c, int64_t hw, void* dst, int64_t dst_c_total, int64_t c_off,
                              int normalize, const float* mask, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(src && dst, "src and dst must be device pointers");
    PGPP_REQUIRE(n >= 0 && c >= 1 && hw >= 1, "bad tensor size");
    PGPP_REQUIRE(c_off >= 0 && c_off + c <= dst_c_total, "channel slice does not fit the destination");
    if (n == 0) return PGPP_OK;
    const long long planes = n * c;
    const long long total = planes * ((hw + 15) / 16);
    long long blocks = (total + 255) / 256;
    const long long cap = 8ll * sm_count();
    if (blocks > cap) blocks = cap;
    const bool vec = hw % 16 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0 && ((uintptr_t)mask & 15) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) u8_to_f32_kernel<true><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)src, (float*)dst, mask, planes, (int)c, hw, dst_c_total, c_off, normalize);
    else u8_to_f32_kernel<false><<<(unsigned)blocks, 256, 0, st>>>((const uint8_t*)src, (float*)dst, mask, planes, (int)c, hw, dst_c_total, c_off, normalize);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

extern "C" int pgpp_image_to_u8(const float* img, int64_t n, int64_t c, int64_t hw, void* out, int reverse_channels, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(img && out, "img and out must be device pointers");
    PGPP_REQUIRE(n >= 0 && c >= 1 && c <= 16 && hw >= 1, "bad tensor size");
    if (n == 0) return PGPP_OK;
    const long long total = n * ((hw + 3) / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = 8ll * sm_count();
    if (blocks > cap) blocks = cap;
    const bool vec = c == 3 && hw % 4 == 0 && ((uintptr_t)img & 15) == 0 && ((uintptr_t)out & 3) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec) image_to_u8_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(img, (uint8_t*)out, n, (int)c, hw, reverse_channels);
    else image_to_u8_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(img, (uint8_t*)out, n, (int)c, hw, reverse_channels);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}
