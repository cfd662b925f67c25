// upfirdn2d on the tensor-core operand format (packed -> packed), sm_100a.
//
// The StyleGAN2 down-sampling convolution is "blur, then strided convolution" (conv2d_resample.py:119-122) and the 1x1 skip of a
// down-sampling residual block is "FIR decimation, then 1x1 convolution" (conv2d_resample.py:107-110).  When the producer of x
// already wrote the operand format (bf16 expansion, channels innermost) from its epilogue, running the FIR on that format keeps
// the whole encoder chain free of float32 NCHW intermediates and of packing passes:
//
//   out[part][n][oy][ox][c] = part-th bf16 term of   gain * sum_{jy,jx} X[n][oy*down + jy - pady0][ox*down + jx - padx0][c] * k[jy][jx]
//   X = sum of the input parts (zero outside the image), k = f flipped unless `flip`  (upfirdn2d.py:168-208 with up = 1)
//
// One thread owns 8 channels (one 128-bit vector per part) of R consecutive output rows of one output column: the rows of the
// input window slide through registers, so an input vector is loaded once per (column tap) instead of once per (row tap, column
// tap); neighbouring lanes read neighbouring channel groups / pixels (512 contiguous bytes per warp load) and the column overlap
// between neighbouring pixels is served by L1.  HBM-bound: one read and one write of the tensor.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace pgpp {

struct FirPackedArgs {
    const __nv_bfloat16* in; long long in_part_stride; int in_parts; int in_ct;
    __nv_bfloat16* out; long long out_part_stride; int out_parts; int out_ct;
    int n, h, w, oh, ow, cg, c_out;
    int fw, fh, padx0, pady0;
    float k[16];            // k[jy * 4 + jx]: flipped, gain folded in, zero beyond (fh, fw)
    // optional epilogue (pgpp_fir_packed_act, tiled kernel only): v = clamp(act(v + noise[n?, oy, ox] + bias[c]) * act_gain)
    const float* noise; long long noise_stride_n; const float* bias;
    int act; float alpha, act_gain, clamp;
    float* out_nchw;        // pgpp_fir_packed_act: float32 [N, C, oh, ow] contiguous result instead of the operand format (out == NULL)
};

template <int D, int R>
__global__ void __launch_bounds__(256, 2) fir_packed_kernel(const FirPackedArgs p, long long total, int strips) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    unsigned t = (unsigned)idx;                         // total <= INT_MAX (checked on the host)
    unsigned q = t / (unsigned)p.cg; const int cgi = (int)(t - q * p.cg); t = q;
    q = t / (unsigned)p.ow; const int ox = (int)(t - q * p.ow); t = q;
    q = t / (unsigned)strips; const int strip = (int)(t - q * strips);
    const int n = (int)q;
    const int oy0 = strip * R;
    const int rows_valid = min(R, p.oh - oy0);
    constexpr int ROWS = (R - 1) * D + 4;

    float acc[R][8];
    #pragma unroll
    for (int r = 0; r < R; r++)
        #pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.f;

    const __nv_bfloat16* const in_n = p.in + ((long long)n * p.h * p.w) * p.in_ct + cgi * 8;
    const int iy_base = oy0 * D - p.pady0, ix_base = ox * D - p.padx0;
    // branch-free window walk: every load goes to a clamped (always valid) address and is zeroed by a select when the tap lies
    // outside the image or the filter, so the loads of a whole row (and of the next rows) can be in flight together
    int xoff[4]; bool xok[4];
    #pragma unroll
    for (int jx = 0; jx < 4; jx++) {
        const int ix = ix_base + jx;
        xok[jx] = jx < p.fw && ix >= 0 && ix < p.w;
        xoff[jx] = min(max(ix, 0), p.w - 1) * p.in_ct;
    }
    const int last_row = (rows_valid - 1) * D + p.fh - 1;
    const bool two = p.in_parts > 1, three = p.in_parts > 2;
    #pragma unroll
    for (int i = 0; i < ROWS; i++) {
        const int iy = iy_base + i;
        const bool yok = iy >= 0 && iy < p.h && i <= last_row;
        const __nv_bfloat16* const row = in_n + (long long)min(max(iy, 0), p.h - 1) * p.w * p.in_ct;
        uint4 u0[4], u1[4], u2[4];
        #pragma unroll
        for (int jx = 0; jx < 4; jx++) {
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            const bool ok = yok && xok[jx];
            const __nv_bfloat16* src = row + xoff[jx];
            u0[jx] = ok ? __ldg(reinterpret_cast<const uint4*>(src)) : z;
            u1[jx] = (ok && two) ? __ldg(reinterpret_cast<const uint4*>(src + p.in_part_stride)) : z;
            u2[jx] = (ok && three) ? __ldg(reinterpret_cast<const uint4*>(src + 2 * p.in_part_stride)) : z;
        }
        #pragma unroll
        for (int jx = 0; jx < 4; jx++) {
            const uint32_t a[4] = {u0[jx].x, u0[jx].y, u0[jx].z, u0[jx].w};
            const uint32_t b[4] = {u1[jx].x, u1[jx].y, u1[jx].z, u1[jx].w};
            const uint32_t c2[4] = {u2[jx].x, u2[jx].y, u2[jx].z, u2[jx].w};
            float v[8];
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                v[2 * j] = __uint_as_float(a[j] << 16) + (__uint_as_float(b[j] << 16) + __uint_as_float(c2[j] << 16));
                v[2 * j + 1] = __uint_as_float(a[j] & 0xffff0000u) + (__uint_as_float(b[j] & 0xffff0000u) + __uint_as_float(c2[j] & 0xffff0000u));
            }
            #pragma unroll
            for (int r = 0; r < R; r++) {
                const int jy = i - r * D;
                if (jy < 0 || jy >= 4) continue;
                const float kk = p.k[jy * 4 + jx];
                #pragma unroll
                for (int c = 0; c < 8; c++) acc[r][c] = fmaf(v[c], kk, acc[r][c]);
            }
        }
    }
    __nv_bfloat16* const out_n = p.out + ((long long)n * p.oh * p.ow) * p.out_ct + cgi * 8;
    #pragma unroll
    for (int r = 0; r < R; r++) {
        if (r >= rows_valid) break;
        __nv_bfloat16* dst = out_n + ((long long)(oy0 + r) * p.ow + ox) * p.out_ct;
        for (int part = 0; part < p.out_parts; part++) {
            const bool more = part + 1 < p.out_parts;
            uint32_t w4[4];
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(acc[r][2 * j], acc[r][2 * j + 1]);
                w4[j] = *reinterpret_cast<const uint32_t*>(&hh);
                if (more) {
                    acc[r][2 * j] -= __uint_as_float(w4[j] << 16);
                    acc[r][2 * j + 1] -= __uint_as_float(w4[j] & 0xffff0000u);
                }
            }
            *reinterpret_cast<uint4*>(dst + part * p.out_part_stride) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Shared-memory version for down = 1 and separable filters (every filter setup_filter() builds from a 1-D tap list is one): the
// register kernel above converts every input vector once per (column tap) of every thread that needs it and runs 16 FMAs per
// output, which makes it instruction-bound at about a third of the HBM bandwidth.  Here a CTA first converts its halo tile ONCE
// (sum of the bf16 parts as float32) into shared memory, then every thread filters a column of FT_H outputs for 8 channels:
// horizontal taps from 128-bit shared loads, vertical taps in registers, packed float32x2 FMAs (FFMA2) throughout.
// 128 threads and 53.5 KB per CTA: four CTAs per SM overlap each other's load and filter phases.
constexpr int FT_W = 16, FT_H = 8, FT_HALO = 3;
constexpr int FT_SW = FT_W + FT_HALO, FT_SH = FT_H + FT_HALO;
constexpr int FT_THREADS = FT_W * 8;
constexpr int FT_SMEM = FT_SH * FT_SW * 64 * 4;         // float32 [row][col][half][cg][4]

struct FirTileArgs {
    FirPackedArgs a;
    float kx[4], ky[4];
    int tiles_x, tiles_y, cblocks;
};


template <int PARTS>
__device__ __forceinline__ void fir_tile_stage(const FirPackedArgs& p, const __nv_bfloat16* src, bool ok, uint4 (&u)[PARTS]) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    #pragma unroll
    for (int part = 0; part < PARTS; part++) u[part] = ok ? __ldg(reinterpret_cast<const uint4*>(src + part * p.in_part_stride)) : z;
}

template <int PARTS>
__device__ __forceinline__ void fir_tile_store(float* dst, const uint4 (&u)[PARTS]) {
    f32x2 v[4] = {bf2_to_f32x2(u[0].x), bf2_to_f32x2(u[0].y), bf2_to_f32x2(u[0].z), bf2_to_f32x2(u[0].w)};
    #pragma unroll
    for (int part = 1; part < PARTS; part++) {
        v[0] = fadd2(v[0], bf2_to_f32x2(u[part].x)); v[1] = fadd2(v[1], bf2_to_f32x2(u[part].y));
        v[2] = fadd2(v[2], bf2_to_f32x2(u[part].z)); v[3] = fadd2(v[3], bf2_to_f32x2(u[part].w));
    }
    *reinterpret_cast<ulonglong2*>(dst) = make_ulonglong2(v[0], v[1]);          // half 0: channels 0..3 of the group
    *reinterpret_cast<ulonglong2*>(dst + 32) = make_ulonglong2(v[2], v[3]);     // half 1: channels 4..7, 32 floats further
}

template <int PARTS, bool EPI = false>
__global__ void __launch_bounds__(FT_THREADS, 4) fir_tile_packed_kernel(const FirTileArgs t, long long total_tiles) {
    extern __shared__ __align__(16) float sm[];
    const FirPackedArgs& p = t.a;
    const int tid = threadIdx.x;
    const int cg = tid & 7, px = tid >> 3;
    f32x2 kx2[4], ky2[4];
    #pragma unroll
    for (int j = 0; j < 4; j++) { kx2[j] = pack2(t.kx[j], t.kx[j]); ky2[j] = pack2(t.ky[j], t.ky[j]); }
    const long long row_pitch = (long long)p.w * p.in_ct;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        unsigned q = (unsigned)tile, r;
        r = q / (unsigned)t.tiles_x; const int bx = (int)(q - r * t.tiles_x); q = r;
        r = q / (unsigned)t.tiles_y; const int by = (int)(q - r * t.tiles_y); q = r;
        r = q / (unsigned)t.cblocks; const int cb = (int)(q - r * t.cblocks);
        const int n = (int)r;
        const int ox0 = bx * FT_W, oy0 = by * FT_H;
        const int ix0 = ox0 - p.padx0, iy0 = oy0 - p.pady0;
        const int cgs = min(8, p.cg - cb * 8);              // 8-channel groups of this channel block
        const __nv_bfloat16* const in_n = p.in + ((long long)n * p.h * p.w) * p.in_ct + cb * 64;
        __syncthreads();                                    // previous tile fully consumed
        // phase 1: halo tile -> float32 sum of the parts, zero outside the image.  Loads go to clamped (always valid) addresses and
        // are zeroed by a select, so the loads of a batch of rows are in flight together.
        {
            // main columns: (row sy, column px, group cg) for every row
            const int ix = ix0 + px;
            const bool col_ok = cg < cgs && ix >= 0 && ix < p.w;
            const __nv_bfloat16* const col = in_n + (long long)min(max(ix, 0), p.w - 1) * p.in_ct + min(cg, cgs - 1) * 8;
            float* const sdst = sm + px * 64 + cg * 4;
            #pragma unroll
            for (int b0 = 0; b0 < FT_SH; b0 += 6) {
                uint4 u[6][PARTS];
                #pragma unroll
                for (int k = 0; k < 6; k++) {
                    const int sy = b0 + k, iy = iy0 + sy;
                    if (sy < FT_SH) fir_tile_stage<PARTS>(p, col + (long long)min(max(iy, 0), p.h - 1) * row_pitch, col_ok && iy >= 0 && iy < p.h, u[k]);
                }
                #pragma unroll
                for (int k = 0; k < 6; k++)
                    if (b0 + k < FT_SH) fir_tile_store<PARTS>(sdst + (b0 + k) * FT_SW * 64, u[k]);
            }
            // halo columns FT_W .. FT_W + 2: FT_SH * 3 * 8 items over the CTA
            constexpr int HITEMS = FT_SH * FT_HALO * 8, HROUNDS = (HITEMS + FT_THREADS - 1) / FT_THREADS;
            uint4 u[HROUNDS][PARTS];
            #pragma unroll
            for (int k = 0; k < HROUNDS; k++) {
                const int item = tid + k * FT_THREADS;
                const int g = item & 7, hc = (item >> 3) % FT_HALO, sy = (item >> 3) / FT_HALO;
                const int iy = iy0 + sy, ixh = ix0 + FT_W + hc;
                const bool ok = item < HITEMS && g < cgs && iy >= 0 && iy < p.h && ixh >= 0 && ixh < p.w;
                fir_tile_stage<PARTS>(p, in_n + (long long)min(max(iy, 0), p.h - 1) * row_pitch + (long long)min(max(ixh, 0), p.w - 1) * p.in_ct + min(g, cgs - 1) * 8,
                                      ok, u[k]);
            }
            #pragma unroll
            for (int k = 0; k < HROUNDS; k++) {
                const int item = tid + k * FT_THREADS;
                const int g = item & 7, hc = (item >> 3) % FT_HALO, sy = (item >> 3) / FT_HALO;
                if (item < HITEMS) fir_tile_store<PARTS>(sm + (sy * FT_SW + FT_W + hc) * 64 + g * 4, u[k]);
            }
        }
        __syncthreads();
        // phase 2: column of FT_H outputs for 8 channels (4 float32x2 accumulators per output row)
        f32x2 acc[FT_H][4];
        #pragma unroll
        for (int rr = 0; rr < FT_H; rr++)
            #pragma unroll
            for (int j = 0; j < 4; j++) acc[rr][j] = 0ull;
        #pragma unroll
        for (int i = 0; i < FT_SH; i++) {
            f32x2 h[4] = {0ull, 0ull, 0ull, 0ull};
            const float* rowp = sm + (i * FT_SW + px) * 64 + cg * 4;
            #pragma unroll
            for (int jx = 0; jx < 4; jx++) {
                const ulonglong2 lo = *reinterpret_cast<const ulonglong2*>(rowp + jx * 64);
                const ulonglong2 hi = *reinterpret_cast<const ulonglong2*>(rowp + jx * 64 + 32);
                ffma2(h[0], lo.x, kx2[jx]); ffma2(h[1], lo.y, kx2[jx]);
                ffma2(h[2], hi.x, kx2[jx]); ffma2(h[3], hi.y, kx2[jx]);
            }
            #pragma unroll
            for (int rr = 0; rr < FT_H; rr++) {
                const int jy = i - rr;
                if (jy < 0 || jy >= 4) continue;
                #pragma unroll
                for (int j = 0; j < 4; j++) ffma2(acc[rr][j], h[j], ky2[jy]);
            }
        }
        const int ox = ox0 + px;
        if (cg < cgs && ox < p.ow) {
            float bias8[8];
            #pragma unroll
            for (int j = 0; j < 8; j++) bias8[j] = (EPI && p.bias) ? __ldg(p.bias + cb * 64 + cg * 8 + j) : 0.f;
            __nv_bfloat16* const out_n = p.out + ((long long)n * p.oh * p.ow) * p.out_ct + cb * 64 + cg * 8;
            #pragma unroll
            for (int rr = 0; rr < FT_H; rr++) {
                const int oy = oy0 + rr;
                if (oy >= p.oh) break;
                __nv_bfloat16* dst = out_n + ((long long)oy * p.ow + ox) * p.out_ct;
                float v[8];
                #pragma unroll
                for (int j = 0; j < 4; j++) { v[2 * j] = __uint_as_float((unsigned)acc[rr][j]); v[2 * j + 1] = __uint_as_float((unsigned)(acc[rr][j] >> 32)); }
                if (EPI) {
                    const float nz = p.noise ? __ldg(p.noise + n * p.noise_stride_n + (long long)oy * p.ow + ox) : 0.f;
                    #pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float r = v[j] + nz + bias8[j];
                        if (p.act == PGPP_ACT_RELU) r = fmaxf(r, 0.f);
                        else if (p.act == PGPP_ACT_LRELU) r = r > 0.f ? r : r * p.alpha;
                        r *= p.act_gain;
                        if (p.clamp >= 0.f) r = fminf(fmaxf(r, -p.clamp), p.clamp);
                        v[j] = r;
                    }
                }
                if (EPI && p.out_nchw) {
                    // float32 NCHW result (the low-resolution blocks hand tensors over): 8 channel planes, one element each
                    const int c0 = cb * 64 + cg * 8, c_real = p.cg * 8;
                    float* o = p.out_nchw + (((long long)n * p.c_out + c0) * p.oh + oy) * p.ow + ox;
                    #pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (c0 + j < p.c_out && c0 + j < c_real) o[(long long)j * p.oh * p.ow] = v[j];
                    continue;
                }
                for (int part = 0; part < p.out_parts; part++) {
                    const bool more = part + 1 < p.out_parts;
                    uint32_t w4[4];
                    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        w4[j] = *reinterpret_cast<const uint32_t*>(&hh);
                        if (more) {
                            v[2 * j] -= __uint_as_float(w4[j] << 16);
                            v[2 * j + 1] -= __uint_as_float(w4[j] & 0xffff0000u);
                        }
                    }
                    *reinterpret_cast<uint4*>(dst + part * p.out_part_stride) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
            }
        }
    }
}

// channel-slice copy (fw = fh = 1, down = 1, gain 1): 128-bit vectors, every part
__global__ void __launch_bounds__(256) copy_packed_kernel(const FirPackedArgs p, long long pixels, long long total) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long pix = idx / p.cg;
        const int g = (int)(idx - pix * p.cg);
        for (int part = 0; part < p.out_parts; part++) {
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (part < p.in_parts) u = __ldg(reinterpret_cast<const uint4*>(p.in + part * p.in_part_stride + pix * p.in_ct + g * 8));
            *reinterpret_cast<uint4*>(p.out + part * p.out_part_stride + pix * p.out_ct + g * 8) = u;
        }
    }
}

} // namespace pgpp

namespace pgpp {
struct FirEpilogue { const float* noise; long long noise_stride_n; const float* bias; int act; float alpha, gain, clamp; float* out_nchw; int c_out; };
static int fir_packed_launch(const void* in, int in_parts, int64_t in_part_stride, int n, int h, int w, int c, int in_c_total,
                             const float* f_host, int fw, int fh, int down, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                             void* out, int out_parts, int64_t out_part_stride, int out_c_total, const FirEpilogue* epi, void* stream);
}

extern "C" int pgpp_fir_packed(const void* in, int in_parts, int64_t in_part_stride, int n, int h, int w, int c, int in_c_total,
                               const float* f_host, int fw, int fh, int down, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                               void* out, int out_parts, int64_t out_part_stride, int out_c_total, void* stream) {
    return pgpp::fir_packed_launch(in, in_parts, in_part_stride, n, h, w, c, in_c_total, f_host, fw, fh, down, padx0, padx1, pady0, pady1, flip, gain,
                                   out, out_parts, out_part_stride, out_c_total, nullptr, stream);
}

extern "C" int pgpp_fir_packed_act(const void* in, int in_parts, int64_t in_part_stride, int n, int h, int w, int c, int in_c_total,
                                   const float* f_host, int fw, int fh, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                                   const float* noise, int64_t noise_stride_n, const float* bias, int act_fn, float alpha, float act_gain, float clamp,
                                   void* out, int out_parts, int64_t out_part_stride, int out_c_total, float* out_nchw, int c_out, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(act_fn == PGPP_ACT_LINEAR || act_fn == PGPP_ACT_RELU || act_fn == PGPP_ACT_LRELU, "fir_packed_act: linear, relu or lrelu");
    PGPP_REQUIRE((out != nullptr) != (out_nchw != nullptr), "fir_packed_act: exactly one of out (operand format) and out_nchw (float32 NCHW)");
    PGPP_REQUIRE(!out_nchw || (c_out >= 1 && c_out <= c), "fir_packed_act: c_out must be in [1, c]");
    const FirEpilogue epi{noise, noise_stride_n, bias, act_fn, alpha, act_gain, clamp, out_nchw, c_out};
    return fir_packed_launch(in, in_parts, in_part_stride, n, h, w, c, in_c_total, f_host, fw, fh, 1, padx0, padx1, pady0, pady1, flip, gain,
                             out ? out : (void*)in, out ? out_parts : in_parts, out ? out_part_stride : in_part_stride, out ? out_c_total : in_c_total, &epi, stream);
}

static int pgpp::fir_packed_launch(const void* in, int in_parts, int64_t in_part_stride, int n, int h, int w, int c, int in_c_total,
                                   const float* f_host, int fw, int fh, int down, int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                                   void* out, int out_parts, int64_t out_part_stride, int out_c_total, const FirEpilogue* epi, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(in && out, "in and out must be device pointers");
    PGPP_REQUIRE(n >= 1 && h >= 1 && w >= 1 && c >= 8 && c % 8 == 0, "fir_packed: empty tensor or channel count not a multiple of 8");
    PGPP_REQUIRE(in_parts >= 1 && in_parts <= 3 && out_parts >= 1 && out_parts <= 3, "fir_packed: 1..3 parts");
    PGPP_REQUIRE(in_c_total >= c && in_c_total % 8 == 0 && out_c_total >= c && out_c_total % 8 == 0, "fir_packed: pixel strides must cover c and be multiples of 8");
    PGPP_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && in_part_stride % 8 == 0 && out_part_stride % 8 == 0,
                 "fir_packed: operands must be 16-byte aligned");
    PGPP_REQUIRE(fw >= 1 && fw <= 4 && fh >= 1 && fh <= 4 && (f_host != nullptr || fw * fh == 1), "fir_packed: filter of at most 4 x 4 taps (host array)");
    PGPP_REQUIRE(down == 1 || down == 2, "fir_packed: down must be 1 or 2");
    const int oh = (h + pady0 + pady1 - fh) / down + 1, ow = (w + padx0 + padx1 - fw) / down + 1;
    PGPP_REQUIRE(h + pady0 + pady1 >= fh && w + padx0 + padx1 >= fw && oh >= 1 && ow >= 1, "fir_packed: output must be at least 1 x 1");
    FirPackedArgs a;
    a.in = (const __nv_bfloat16*)in; a.in_part_stride = in_part_stride; a.in_parts = in_parts; a.in_ct = in_c_total;
    a.out = (__nv_bfloat16*)out; a.out_part_stride = out_part_stride; a.out_parts = out_parts; a.out_ct = out_c_total;
    a.n = n; a.h = h; a.w = w; a.oh = oh; a.ow = ow; a.cg = c / 8;
    a.fw = fw; a.fh = fh; a.padx0 = padx0; a.pady0 = pady0;
    a.noise = epi ? epi->noise : nullptr; a.noise_stride_n = epi ? epi->noise_stride_n : 0; a.bias = epi ? epi->bias : nullptr;
    a.act = epi ? epi->act : PGPP_ACT_LINEAR; a.alpha = epi ? epi->alpha : 0.f; a.act_gain = epi ? epi->gain : 1.f; a.clamp = epi ? epi->clamp : -1.f;
    a.out_nchw = epi ? epi->out_nchw : nullptr; a.c_out = epi ? epi->c_out : c;
    for (int i = 0; i < 16; i++) a.k[i] = 0.f;
    for (int jy = 0; jy < fh; jy++)
        for (int jx = 0; jx < fw; jx++) {
            const int sy = flip ? jy : fh - 1 - jy, sx = flip ? jx : fw - 1 - jx;
            a.k[jy * 4 + jx] = (f_host ? f_host[sy * fw + sx] : 1.f) * gain;
        }
    if (!epi && down == 1 && fw * fh == 1 && a.k[0] == 1.f && padx0 == 0 && pady0 == 0 && in_parts == out_parts) {
        const long long pixels = (long long)n * h * w, total_v = pixels * a.cg;
        long long blocks = (total_v + 255) / 256;
        const long long cap = (long long)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        copy_packed_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, pixels, total_v);
        count_launch();
        PGPP_CUDA_OK(cudaGetLastError());
        return PGPP_OK;
    }
    if (down == 1 && (epi || !env_flags().fir_packed_no_tile)) {
        // separable?  f = outer(ky, kx) with the pivot at the largest tap
        int pj = 0;
        for (int i = 1; i < 16; i++) if (fabsf(a.k[i]) > fabsf(a.k[pj])) pj = i;
        const int py = pj / 4, pxx = pj % 4;
        const float piv = a.k[pj];
        bool sep = piv != 0.f;
        FirTileArgs ta;
        for (int j = 0; j < 4; j++) { ta.kx[j] = a.k[py * 4 + j]; ta.ky[j] = sep ? a.k[j * 4 + pxx] / piv : 0.f; }
        for (int jy = 0; jy < 4 && sep; jy++)
            for (int jx = 0; jx < 4; jx++)
                if (fabsf(ta.ky[jy] * ta.kx[jx] - a.k[jy * 4 + jx]) > 1e-6f * fabsf(piv)) { sep = false; break; }
        if (sep) {
            ta.a = a;
            ta.tiles_x = (ow + FT_W - 1) / FT_W; ta.tiles_y = (oh + FT_H - 1) / FT_H; ta.cblocks = (c + 63) / 64;
            const long long tiles = (long long)ta.tiles_x * ta.tiles_y * ta.cblocks * n;
            PGPP_REQUIRE(tiles < (1ll << 31), "fir_packed: tensor too large");
            static std::atomic<bool> attr_done[64];
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev < 0 || dev >= 64 || !attr_done[dev].load(std::memory_order_acquire)) {
                PGPP_CUDA_OK(cudaFuncSetAttribute(fir_tile_packed_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                PGPP_CUDA_OK(cudaFuncSetAttribute(fir_tile_packed_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                PGPP_CUDA_OK(cudaFuncSetAttribute(fir_tile_packed_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                PGPP_CUDA_OK(cudaFuncSetAttribute(fir_tile_packed_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                PGPP_CUDA_OK(cudaFuncSetAttribute(fir_tile_packed_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                PGPP_CUDA_OK(cudaFuncSetAttribute(fir_tile_packed_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM));
                if (dev >= 0 && dev < 64) attr_done[dev].store(true, std::memory_order_release);
            }
            long long blocks = tiles;
            const long long cap = (long long)sm_count() * 4;
            if (blocks > cap) blocks = cap;
            if (epi) {
                if (in_parts == 1) fir_tile_packed_kernel<1, true><<<(unsigned)blocks, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(ta, tiles);
                else if (in_parts == 2) fir_tile_packed_kernel<2, true><<<(unsigned)blocks, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(ta, tiles);
                else fir_tile_packed_kernel<3, true><<<(unsigned)blocks, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(ta, tiles);
            }
            else if (in_parts == 1) fir_tile_packed_kernel<1><<<(unsigned)blocks, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(ta, tiles);
            else if (in_parts == 2) fir_tile_packed_kernel<2><<<(unsigned)blocks, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(ta, tiles);
            else fir_tile_packed_kernel<3><<<(unsigned)blocks, FT_THREADS, FT_SMEM, (cudaStream_t)stream>>>(ta, tiles);
            count_launch();
            PGPP_CUDA_OK(cudaGetLastError());
            return PGPP_OK;
        }
    }
    PGPP_REQUIRE(!epi, "fir_packed_act needs a separable filter (outer product of two 1-D tap lists)");
    const int R = down == 1 ? 8 : 4;
    const int strips = (oh + R - 1) / R;
    const long long total = (long long)n * strips * ow * a.cg;
    PGPP_REQUIRE(total < (1ll << 31), "fir_packed: tensor too large");
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (down == 1) fir_packed_kernel<1, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(a, total, strips);
    else fir_packed_kernel<2, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(a, total, strips);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}
