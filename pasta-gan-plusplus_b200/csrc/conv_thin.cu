// 1x1 modulated convolution with a handful of output channels on the operand format: the ToRGB layers of the generator
// (networks.py:1925-1967: 3 image channels, plus the 7-channel parsing head that reads the same x with the same styles).
//
//   y[n, o, p] (+)= clamp( act( sum_c X[n, p, c] * w[o, c] * s[n, c] + bias[o] ) * gain )        X = sum of the bf16 parts
//
// 3 + 7 outputs over 64 inputs are 20 FLOP per input byte: the layer is bound by reading x once, so it runs on the CUDA cores
// instead of padding the GEMM N dimension of the tensor-core kernel to 16 columns (which measured 0.3 of the HBM bandwidth).
// One thread per pixel; the per-sample weights w * s are staged in shared memory once per CTA and read as warp-wide broadcasts;
// two heads of up to 8 channels in total can be produced from one pass over x; outputs are float32 NCHW planes (coalesced), head 1 optionally accumulated
// into the up-sampled skip image (img.add_(y), networks.py:2190).
#include "act.cuh"

namespace pgpp {

struct ThinArgs {
    const __nv_bfloat16* x; long long part_stride; int parts; int ct;
    int n, c; long long hw;
    const float* w1; const float* w2; const float* styles; const float* b1; const float* b2;
    int o1, o2;
    int act_fn; float alpha, gain, clamp;
    float* out1; float* out2; int accumulate1;
};

constexpr int THIN_PIX_PER_CTA = 2048;

// LPP lanes share one pixel: lane j of the group owns channels [8j, 8j + 8) (+ 8 * LPP for the second chunk when CPL == 2), keeps
// its OT x 8 (x CPL) weights in registers for the whole CTA, and the group reduces the OT partial sums with a shuffle reduce-scatter
// (every lane ends up with OT / LPP outputs, or one output replicated when OT < LPP).  A warp load instruction therefore reads
// 512 contiguous bytes (4 / 2 / 1 whole pixels), and no weight is re-read per pixel.
template <int OT, int LPP, int CPL, int PARTS>
__global__ void __launch_bounds__(256) conv1x1_thin_kernel(const ThinArgs p) {
    const int n = blockIdx.y;
    const int ot = p.o1 + p.o2;
    const int lane_in = threadIdx.x % LPP;                      // position inside the pixel group
    const int group = threadIdx.x / LPP, groups = 256 / LPP;
    f32x2 w[CPL][OT][4];
    #pragma unroll
    for (int q = 0; q < CPL; q++)
        #pragma unroll
        for (int o = 0; o < OT; o++)
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                float v2[2];
                #pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = (q * LPP + lane_in) * 8 + 2 * j + e;
                    float v = 0.f;
                    if (o < ot && c < p.c) {
                        v = o < p.o1 ? __ldg(p.w1 + (long long)o * p.c + c) : __ldg(p.w2 + (long long)(o - p.o1) * p.c + c);
                        if (p.styles) v *= __ldg(p.styles + (long long)n * p.c + c);
                    }
                    v2[e] = v;
                }
                w[q][o][j] = pack2(v2[0], v2[1]);
            }
    const long long p0 = (long long)blockIdx.x * THIN_PIX_PER_CTA;
    const long long p1 = min(p0 + THIN_PIX_PER_CTA, p.hw);
    constexpr int UNROLL = 2;               // pixels per group and step
    __shared__ float so[2][OT][(256 / LPP) * UNROLL];
    int buf = 0;
    const int step = groups * UNROLL;       // pixels per CTA step (contiguous)
    typedef uint4 Stage[UNROLL][CPL][PARTS];

    auto load = [&](Stage& u, long long pb) {
        #pragma unroll
        for (int r = 0; r < UNROLL; r++) {
            const long long pix = pb + r * groups + group;
            const bool ok = pix < p1;
            const __nv_bfloat16* src = p.x + ((long long)n * p.hw + (ok ? pix : p0)) * p.ct + lane_in * 8;
            #pragma unroll
            for (int q = 0; q < CPL; q++)
                #pragma unroll
                for (int part = 0; part < PARTS; part++)
                    u[r][q][part] = (ok && (q * LPP + lane_in) * 8 < p.c) ? __ldg(reinterpret_cast<const uint4*>(src + part * p.part_stride + q * LPP * 8))
                                                                                            : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    auto process = [&](const Stage& u, long long pb) {
        #pragma unroll
        for (int r = 0; r < UNROLL; r++) {
            f32x2 acc2[OT];
            #pragma unroll
            for (int o = 0; o < OT; o++) acc2[o] = 0ull;
            #pragma unroll
            for (int q = 0; q < CPL; q++) {
                f32x2 v[4] = {bf2_to_f32x2(u[r][q][0].x), bf2_to_f32x2(u[r][q][0].y), bf2_to_f32x2(u[r][q][0].z), bf2_to_f32x2(u[r][q][0].w)};
                #pragma unroll
                for (int part = 1; part < PARTS; part++) {
                    v[0] = fadd2(v[0], bf2_to_f32x2(u[r][q][part].x)); v[1] = fadd2(v[1], bf2_to_f32x2(u[r][q][part].y));
                    v[2] = fadd2(v[2], bf2_to_f32x2(u[r][q][part].z)); v[3] = fadd2(v[3], bf2_to_f32x2(u[r][q][part].w));
                }
                #pragma unroll
                for (int o = 0; o < OT; o++)
                    #pragma unroll
                    for (int j = 0; j < 4; j++) ffma2(acc2[o], v[j], w[q][o][j]);
            }
            float acc[OT];
            #pragma unroll
            for (int o = 0; o < OT; o++) acc[o] = __uint_as_float((unsigned)acc2[o]) + __uint_as_float((unsigned)(acc2[o] >> 32));
            // reduce-scatter over the LPP lanes of the group: after the step with distance d a lane keeps the half of its values
            // selected by its bit d; once one value is left the remaining steps are plain butterfly adds
            int cnt = OT;                       // compile-time after unrolling
            int o_base = 0;
            #pragma unroll
            for (int d = LPP / 2; d >= 1; d >>= 1) {
                const bool upper = (lane_in & d) != 0;
                if (cnt > 1) {
                    const int half = cnt / 2;
                    #pragma unroll
                    for (int i = 0; i < OT / 2; i++) {
                        if (i < half) {
                            const float send = upper ? acc[i] : acc[i + half];
                            const float keep = upper ? acc[i + half] : acc[i];
                            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
                        }
                    }
                    if (upper) o_base += half;
                    cnt = half;
                } else {
                    acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], d);
                }
            }
            // the lane holds outputs o_base .. (with OT < LPP several lanes hold the same value: the one whose low bits are zero keeps
            // it).  The sums go through shared memory so that the output planes are written as coalesced rows.
            const bool writer = OT >= LPP || (lane_in & (LPP / OT - 1)) == 0;
            if (writer) {
                #pragma unroll
                for (int i = 0; i < (OT >= LPP ? OT / LPP : 1); i++) so[buf][o_base + i][r * groups + group] = acc[i];
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < ot * step; e += 256) {
            const int o = e / step, lp = e - o * step;
            const long long pix = pb + lp;
            if (pix >= p1) continue;
            const bool head1 = o < p.o1;
            float r = so[buf][o][lp];
            const float* bias = head1 ? p.b1 : p.b2;
            if (bias) r += __ldg(bias + (head1 ? o : o - p.o1));
            if (p.act_fn == PGPP_ACT_RELU) r = fmaxf(r, 0.f);
            else if (p.act_fn == PGPP_ACT_LRELU) r = r > 0.f ? r : r * p.alpha;
            r *= p.gain;
            if (p.clamp >= 0.f) r = fminf(fmaxf(r, -p.clamp), p.clamp);
            float* dst = head1 ? p.out1 + ((long long)n * p.o1 + o) * p.hw + pix : p.out2 + ((long long)n * p.o2 + (o - p.o1)) * p.hw + pix;
            if (head1 && p.accumulate1) r += *dst;
            *dst = r;
        }
        buf ^= 1;           // the next step fills the other buffer: one barrier per step is enough
    };
    // software pipeline: the loads of step k + 1 are in flight while step k is reduced and written
    Stage ua, ub;
    load(ua, p0);
    for (long long pb = p0; pb < p1; pb += 2 * step) {
        load(ub, pb + step);
        process(ua, pb);
        load(ua, pb + 2 * step);
        process(ub, pb + step);
    }
}

} // namespace pgpp

extern "C" int pgpp_conv1x1_thin(const void* x, int x_parts, int64_t x_part_stride, int c_total, int n, int c, int64_t hw,
                                 const float* w1, const float* b1, int o1, float* out1, int accumulate1,
                                 const float* w2, const float* b2, int o2, float* out2,
                                 const float* styles, int act_fn, float alpha, float gain, float clamp, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(x && w1 && out1 && o1 >= 1, "conv1x1_thin: x, w1 and out1 are required");
    PGPP_REQUIRE(o2 == 0 || (w2 && out2), "conv1x1_thin: the second head needs w2 and out2");
    PGPP_REQUIRE(o1 + o2 <= 8, "conv1x1_thin: at most 8 output channels in total per launch");
    PGPP_REQUIRE(n >= 1 && hw >= 1 && c >= 8 && c % 8 == 0 && c <= 512, "conv1x1_thin: c must be a multiple of 8, at most 512");
    PGPP_REQUIRE(x_parts >= 1 && x_parts <= 3 && c_total >= c && c_total % 8 == 0 && x_part_stride % 8 == 0 && ((uintptr_t)x & 15) == 0,
                 "conv1x1_thin: x must be the 16-byte aligned operand format with 1..3 parts");
    PGPP_REQUIRE(act_fn == PGPP_ACT_LINEAR || act_fn == PGPP_ACT_RELU || act_fn == PGPP_ACT_LRELU, "conv1x1_thin: linear, relu or lrelu");
    PGPP_REQUIRE(n <= 65535, "conv1x1_thin: batch too large");
    ThinArgs a;
    a.x = (const __nv_bfloat16*)x; a.part_stride = x_part_stride; a.parts = x_parts; a.ct = c_total;
    a.n = n; a.c = c; a.hw = hw;
    a.w1 = w1; a.w2 = w2; a.styles = styles; a.b1 = b1; a.b2 = b2; a.o1 = o1; a.o2 = o2;
    a.act_fn = act_fn; a.alpha = alpha; a.gain = gain; a.clamp = clamp;
    a.out1 = out1; a.out2 = out2; a.accumulate1 = accumulate1;
    const int ot = o1 + o2;
    const dim3 grid((unsigned)((hw + THIN_PIX_PER_CTA - 1) / THIN_PIX_PER_CTA), (unsigned)n);
    cudaStream_t st = (cudaStream_t)stream;
#define PGPP_THIN_P(OT, LPP, CPL) \
    if (x_parts == 1) conv1x1_thin_kernel<OT, LPP, CPL, 1><<<grid, 256, 0, st>>>(a);        \
    else if (x_parts == 2) conv1x1_thin_kernel<OT, LPP, CPL, 2><<<grid, 256, 0, st>>>(a);   \
    else conv1x1_thin_kernel<OT, LPP, CPL, 3><<<grid, 256, 0, st>>>(a);
#define PGPP_THIN(OT) \
    if (c <= 64) { PGPP_THIN_P(OT, 8, 1) } else if (c <= 128) { PGPP_THIN_P(OT, 16, 1) } else if (c <= 256) { PGPP_THIN_P(OT, 32, 1) } else { PGPP_THIN_P(OT, 32, 2) }
    if (ot <= 4) { PGPP_THIN(4) }
    else {
        PGPP_REQUIRE(c <= 256, "conv1x1_thin: more than 4 output channels need c <= 256");
        if (c <= 64) { PGPP_THIN_P(8, 8, 1) } else if (c <= 128) { PGPP_THIN_P(8, 16, 1) } else { PGPP_THIN_P(8, 32, 1) }
    }
#undef PGPP_THIN_P
#undef PGPP_THIN
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}
