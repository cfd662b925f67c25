// upfirdn2d for sm_100a.  Replaces torch_utils/ops/upfirdn2d.cu:29-341 + upfirdn2d.cpp:16-94.
//
//   out[n,c,oy,ox] = gain * sum_{jy,jx} z[oy*downy + jy - pady0, ox*downx + jx - padx0] * k[jy,jx]
//   z = x with (up-1) zeros inserted after every sample, zero outside; k = f flipped unless `flip`.
//
// Kernels:
//   * fir_tile_kernel  - the shapes that carry the traffic in the generator (up = down = 1, filter
//     up to 4x4, unit stride along W): shared-memory halo tile, each thread produces a 4x4 patch of
//     outputs from a sliding window of 128-bit shared loads with the taps held in registers, and
//     writes 128-bit rows.  HBM-bound; one read and one write per element.
//   * fir_down2_kernel - down = 2 from the same kind of halo tile (1x1-skip / encoder downsampling).
//   * fir_up2_kernel   - up = 2 (image-skip upsampling, and the gradient of every down = 2 call): the
//     zero-stuffed signal is never formed; output parity selects 2 of the 4 taps per axis at compile time.
//   Rows whose pitch is not a multiple of 4 elements (the 513-wide blurred images) leave through a
//   shared-memory transpose so that every store instruction of a warp covers one contiguous row segment.
//   * generic_kernel   - any up/down/pad/filter size/strides (channels_last included); one thread per
//     4 consecutive outputs, taps from shared memory, input through the read-only path.
#include "common.cuh"

namespace pgpp {

struct UpfirdnArgs {
    const void* x; const float* f; void* y;
    int n, c, ih, iw, oh, ow;
    long long xs_n, xs_c, xs_h, xs_w, ys_n, ys_c, ys_h, ys_w;
    int fw, fh; long long fs_x, fs_y;
    int upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
};

__device__ __forceinline__ int floor_div_i(int a, int b) {   // b > 0
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

// ------------------------------------------------------------------------------------------
constexpr int MAX_TAPS = 32;     // per axis, generic kernel (shared-memory filter)

template <class T>
__global__ void __launch_bounds__(256) generic_kernel(UpfirdnArgs p, long long total_quads, int quads_per_row) {
    typedef typename Acc<T>::type S;
    __shared__ S sk[MAX_TAPS * MAX_TAPS];
    for (int t = threadIdx.x; t < p.fw * p.fh; t += blockDim.x) {
        const int jy = t / p.fw, jx = t - jy * p.fw;
        const int sy = p.flip ? jy : p.fh - 1 - jy, sx = p.flip ? jx : p.fw - 1 - jx;
        sk[t] = (S)p.f[sy * p.fs_y + sx * p.fs_x] * (S)p.gain;
    }
    __syncthreads();
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += (long long)gridDim.x * blockDim.x) {
        const unsigned q32 = (unsigned)q;               // total_quads <= output numel <= INT_MAX
        unsigned r = q32 / (unsigned)quads_per_row;
        const int qx = (int)(q32 - r * (unsigned)quads_per_row);
        unsigned r2 = r / (unsigned)p.oh;
        const int oy = (int)(r - r2 * (unsigned)p.oh);
        const int n = (int)(r2 / (unsigned)p.c);
        const int c = (int)(r2 - (unsigned)n * (unsigned)p.c);
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        T* yp = y + n * p.ys_n + c * p.ys_c + oy * p.ys_h;
        // rows: taps jy with (base_y + jy) % upy == 0
        const int base_y = oy * p.downy - p.pady0;
        int jy0 = (-base_y) % p.upy; if (jy0 < 0) jy0 += p.upy;
        const int iy0 = (base_y + jy0) / p.upy;     // exact division
        #pragma unroll 1
        for (int v = 0; v < 4; v++) {
            const int ox = qx * 4 + v;
            if (ox >= p.ow) break;
            const int base_x = ox * p.downx - p.padx0;
            int jx0 = (-base_x) % p.upx; if (jx0 < 0) jx0 += p.upx;
            const int ix0 = (base_x + jx0) / p.upx;
            S acc = 0;
            for (int jy = jy0, iy = iy0; jy < p.fh; jy += p.upy, iy++) {
                if (iy < 0 || iy >= p.ih) continue;
                const T* row = xp + iy * p.xs_h;
                for (int jx = jx0, ix = ix0; jx < p.fw; jx += p.upx, ix++) {
                    if (ix < 0 || ix >= p.iw) continue;
                    acc += to_acc<T>(__ldg(row + ix * p.xs_w)) * sk[jy * p.fw + jx];
                }
            }
            yp[ox * p.ys_w] = from_acc<T>(acc);
        }
    }
}


// ------------------------------------------------------------------------------------------
// Output helpers shared by the tiled kernels.  A thread owns 4 consecutive outputs of a row.
template <class T> __device__ __forceinline__ void store4(T* dst, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float* dst, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4*>(dst), make_float4(a, b, c, d));
}
template <> __device__ __forceinline__ void store4<__half>(__half* dst, float a, float b, float c, float d) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 v; v.x = *reinterpret_cast<const uint32_t*>(&lo); v.y = *reinterpret_cast<const uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(dst), v);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float c, float d) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v; v.x = *reinterpret_cast<const uint32_t*>(&lo); v.y = *reinterpret_cast<const uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(dst), v);
}

// every row of y starts on a 4-element boundary: a thread's 4 outputs go out as one 16-byte (8-byte for 16-bit types) store
template <class T> __device__ __forceinline__ bool rows_aligned4(const UpfirdnArgs& p) {
    return ((p.ys_h & 3) == 0) && ((p.ys_c & 3) == 0) && ((p.ys_n & 3) == 0) && (((uintptr_t)p.y & (4 * sizeof(T) - 1)) == 0);
}

// Unaligned rows: the CTA's TH x TW output tile goes through shared memory (`stage`, rows of PITCH floats, written as float4) and leaves as whole row
// segments, consecutive lanes writing consecutive elements (a warp store covers 128 contiguous bytes instead of 32 4-byte
// pieces 16 bytes apart).  Called by all 256 threads; `stage` must not alias live input data (caller syncs before).
template <class T, int TW, int TH, int PITCH>
__device__ __forceinline__ void store_tile_staged(const float* stage, T* yp, long long ys_h, int ox0, int oy0, int ow, int oh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    #pragma unroll 1
    for (int ry = warp; ry < TH; ry += 8) {
        const int oy = oy0 + ry;
        if (oy >= oh) break;
        T* row = yp + (long long)oy * ys_h + ox0;
        #pragma unroll
        for (int k = 0; k < TW / 32; k++) {
            const int rx = lane + 32 * k;
            if (ox0 + rx < ow) row[rx] = from_acc<T>(stage[ry * PITCH + rx]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// up = down = 1, fw, fh <= 4, x/y unit stride along W.
constexpr int TILE_W = 128, TILE_H = 32, HALO = 3;
constexpr int SM_W = TILE_W + 4;        // 131 needed; 132 keeps rows 16-byte aligned
constexpr int SM_H = TILE_H + HALO;

template <class T>
__global__ void __launch_bounds__(256, 4) fir_tile_kernel(UpfirdnArgs p, int tiles_x, int tiles_y, long long total_tiles) {
    typedef float S;
    __shared__ __align__(16) S sx[SM_H][SM_W];
    // taps in registers, zero-padded to 4x4, already flipped and scaled by gain.  (A packed FFMA2 formulation - two adjacent outputs
    // per instruction - measured 10 % slower: FFMA2 issues at half rate and the unaligned input pairs cost register moves.)
    S k[4][4];
    #pragma unroll
    for (int jy = 0; jy < 4; jy++)
        #pragma unroll
        for (int jx = 0; jx < 4; jx++) {
            S v = 0;
            if (jy < p.fh && jx < p.fw) {
                const int sy = p.flip ? jy : p.fh - 1 - jy, sxx = p.flip ? jx : p.fw - 1 - jx;
                v = p.f[sy * p.fs_y + sxx * p.fs_x] * p.gain;
            }
            k[jy][jx] = v;
        }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads, 4 x 4 outputs each
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const bool vec_store = rows_aligned4<T>(p);

    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        unsigned r = (unsigned)t;
        const int bx = divmod_u32(r, (unsigned)tiles_x);
        const int by = divmod_u32(r, (unsigned)tiles_y);
        const int c = divmod_u32(r, (unsigned)p.c);
        const int n = (int)r;
        const int ox0 = bx * TILE_W, oy0 = by * TILE_H;
        const int ix0 = ox0 - p.padx0, iy0 = oy0 - p.pady0;
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        __syncthreads();        // previous tile fully consumed
        // halo load: a warp walks one tile row at a time (coalesced 128-byte requests), 8 warps interleave rows.
        // All of a thread's loads are issued before the first shared-memory store so ~25 requests per thread are in flight.
        // Tiles whose halo lies inside the image (all but the border ring) skip the per-element bounds tests.
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            constexpr int ROWS = (SM_H + 7) / 8, COLS = (SM_W + 31) / 32;
            T v[ROWS][COLS];        // kept in the source type until the staging store: a conversion between two loads makes the later load wait for the earlier one's data
            const T zero = from_acc<T>(0.f);
            const bool interior = iy0 >= 0 && iy0 + SM_H <= p.ih && ix0 >= 0 && ix0 + SM_W <= p.iw;
            if (interior) {
                #pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    const int ry = warp + 8 * r;
                    const T* row = xp + (long long)(iy0 + ry) * p.xs_h + ix0 + lane;
                    const bool row_ok = ry < SM_H;
                    #pragma unroll
                    for (int k = 0; k < COLS; k++) {
                        const bool ok = row_ok && (k < COLS - 1 || lane + 32 * k < SM_W);
                        v[r][k] = ok ? __ldg(row + 32 * k) : zero;
                    }
                }
            } else {
                #pragma unroll
                for (int r = 0; r < ROWS; r++) {
                    const int ry = warp + 8 * r;
                    const int iy = iy0 + ry;
                    const bool row_ok = ry < SM_H && iy >= 0 && iy < p.ih;
                    const T* row = xp + (long long)iy * p.xs_h + ix0;
                    #pragma unroll
                    for (int k = 0; k < COLS; k++) {
                        const int rx = lane + 32 * k;
                        const int ix = ix0 + rx;
                        v[r][k] = (row_ok && rx < SM_W && ix >= 0 && ix < p.iw) ? __ldg(row + rx) : zero;
                    }
                }
            }
            #pragma unroll
            for (int r = 0; r < ROWS; r++) {
                const int ry = warp + 8 * r;
                #pragma unroll
                for (int k = 0; k < COLS; k++) {
                    const int rx = lane + 32 * k;
                    if (ry < SM_H && rx < SM_W) sx[ry][rx] = to_acc<T>(v[r][k]);
                }
            }
        }
        __syncthreads();
        S acc[4][4];
        #pragma unroll
        for (int a = 0; a < 4; a++)
            #pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0;
        const int cx = tx * 4, cy = ty * 4;
        #pragma unroll
        for (int ry = 0; ry < 4 + HALO; ry++) {
            const float4 lo = *reinterpret_cast<const float4*>(&sx[cy + ry][cx]);
            const float4 hi = *reinterpret_cast<const float4*>(&sx[cy + ry][cx + 4]);
            const S in[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            #pragma unroll
            for (int a = 0; a < 4; a++) {
                const int jy = ry - a;          // filter row feeding output row a
                if (jy < 0 || jy > 3) continue;
                #pragma unroll
                for (int b = 0; b < 4; b++)
                    #pragma unroll
                    for (int jx = 0; jx < 4; jx++) acc[a][b] = fmaf(in[b + jx], k[jy][jx], acc[a][b]);
            }
        }
        T* yp = y + n * p.ys_n + c * p.ys_c;
        if (vec_store) {
            #pragma unroll
            for (int a = 0; a < 4; a++) {
                const int oy = oy0 + cy + a, ox = ox0 + cx;
                if (oy >= p.oh || ox >= p.ow) continue;
                T* dst = yp + (long long)oy * p.ys_h + ox;
                if (ox + 3 < p.ow) store4<T>(dst, acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
                else {
                    #pragma unroll
                    for (int b = 0; b < 4; b++) if (ox + b < p.ow) dst[b] = from_acc<T>(acc[a][b]);
                }
            }
        } else {
            __syncthreads();        // every thread has read its window: the halo tile becomes the output stage
            #pragma unroll
            for (int a = 0; a < 4; a++)
                *reinterpret_cast<float4*>(&sx[cy + a][cx]) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
            __syncthreads();
            store_tile_staged<T, TILE_W, TILE_H, SM_W>(&sx[0][0], yp, p.ys_h, ox0, oy0, p.ow, p.oh);
        }
    }
}

// up = 1, down = 2 in both axes, fw, fh <= 4, unit stride along W: 64 x 16 output tile from a 132 x 35 input halo tile
// (the FIR-downsampling of the 1x1-skip and encoder paths, conv2d_resample.py:107-110).  Each thread produces 4
// consecutive outputs of one row: 10 inputs per filter row through two 128-bit and one 64-bit shared load.
constexpr int D2_TILE_W = 64, D2_TILE_H = 16;
constexpr int D2_SM_W = 2 * D2_TILE_W + 4, D2_SM_H = 2 * D2_TILE_H + 3;

template <class T>
__global__ void __launch_bounds__(256) fir_down2_kernel(UpfirdnArgs p, int tiles_x, int tiles_y, long long total_tiles) {
    typedef float S;
    __shared__ __align__(16) S sx[D2_SM_H][D2_SM_W];
    S k[4][4];
    #pragma unroll
    for (int jy = 0; jy < 4; jy++)
        #pragma unroll
        for (int jx = 0; jx < 4; jx++) {
            S v = 0;
            if (jy < p.fh && jx < p.fw) {
                const int sy = p.flip ? jy : p.fh - 1 - jy, sxx = p.flip ? jx : p.fw - 1 - jx;
                v = p.f[sy * p.fs_y + sxx * p.fs_x] * p.gain;
            }
            k[jy][jx] = v;
        }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4 x 1 outputs each
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const bool vec_store = rows_aligned4<T>(p);
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        unsigned r = (unsigned)t;
        const int bx = divmod_u32(r, (unsigned)tiles_x);
        const int by = divmod_u32(r, (unsigned)tiles_y);
        const int c = divmod_u32(r, (unsigned)p.c);
        const int n = (int)r;
        const int ox0 = bx * D2_TILE_W, oy0 = by * D2_TILE_H;
        const int ix0 = ox0 * 2 - p.padx0, iy0 = oy0 * 2 - p.pady0;
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        __syncthreads();
        {
            constexpr int ROWS = (D2_SM_H + 7) / 8, COLS = (D2_SM_W + 31) / 32;
            T v[ROWS][COLS];        // source type until the staging store (see fir_tile_kernel)
            const T zero = from_acc<T>(0.f);
            #pragma unroll
            for (int rr = 0; rr < ROWS; rr++) {
                const int ry = warp + 8 * rr;
                const int iy = iy0 + ry;
                const bool row_ok = ry < D2_SM_H && iy >= 0 && iy < p.ih;
                const T* row = xp + (long long)iy * p.xs_h + ix0;
                #pragma unroll
                for (int kk = 0; kk < COLS; kk++) {
                    const int rx = lane + 32 * kk;
                    const int ix = ix0 + rx;
                    v[rr][kk] = (row_ok && ix >= 0 && ix < p.iw) ? __ldg(row + rx) : zero;
                }
            }
            #pragma unroll
            for (int rr = 0; rr < ROWS; rr++) {
                const int ry = warp + 8 * rr;
                #pragma unroll
                for (int kk = 0; kk < COLS; kk++) {
                    const int rx = lane + 32 * kk;
                    if (ry < D2_SM_H && rx < D2_SM_W) sx[ry][rx] = to_acc<T>(v[rr][kk]);
                }
            }
        }
        __syncthreads();
        S acc[4] = {0, 0, 0, 0};
        #pragma unroll
        for (int jy = 0; jy < 4; jy++) {
            const S* rowp = &sx[2 * ty + jy][8 * tx];
            const float4 a = *reinterpret_cast<const float4*>(rowp);
            const float4 b = *reinterpret_cast<const float4*>(rowp + 4);
            const float2 cc = *reinterpret_cast<const float2*>(rowp + 8);
            const S in[10] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, cc.x, cc.y};
            #pragma unroll
            for (int o = 0; o < 4; o++)
                #pragma unroll
                for (int jx = 0; jx < 4; jx++) acc[o] = fmaf(in[2 * o + jx], k[jy][jx], acc[o]);
        }
        const int oy = oy0 + ty, ox = ox0 + 4 * tx;
        T* yp = y + n * p.ys_n + c * p.ys_c;
        if (vec_store) {
            if (oy < p.oh && ox < p.ow) {
                T* dst = yp + (long long)oy * p.ys_h + ox;
                if (ox + 3 < p.ow) store4<T>(dst, acc[0], acc[1], acc[2], acc[3]);
                else {
                    #pragma unroll
                    for (int o = 0; o < 4; o++) if (ox + o < p.ow) dst[o] = from_acc<T>(acc[o]);
                }
            }
        } else {
            __syncthreads();
            *reinterpret_cast<float4*>(&sx[ty][4 * tx]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            __syncthreads();
            store_tile_staged<T, D2_TILE_W, D2_TILE_H, D2_SM_W>(&sx[0][0], yp, p.ys_h, ox0, oy0, p.ow, p.oh);
        }
    }
}

// up = 2 in both axes, down = 1, fw, fh <= 4, unit stride along W (upsample2d of the image skip, networks.py ToRGB path /
// upfirdn2d.py:326-348, and the gradient of every down = 2 call, upfirdn2d.py:232-247).  out[o] = sum_j z[o + j - pad0] k[j] with
// z[2i] = x[i], z[odd] = 0: an output of parity q = (o - pad0) & 1 sees taps q and q + 2 on inputs (o - pad0 + q) / 2 and the
// next one.  PX / PY = parity of padx0 / pady0, so a thread's 4 x 4 outputs (tile origins are even) have compile-time tap
// sets: 4 FMAs per output from a 4 x 4 input window read as two 64-bit shared loads per row.
constexpr int U2_TILE_W = 128, U2_TILE_H = 32;
constexpr int U2_SM_W = 68, U2_SM_H = 18;         // 66 x 18 inputs used
constexpr int U2_STAGE_W = 132;

template <class T, int PX, int PY>
__global__ void __launch_bounds__(256) fir_up2_kernel(UpfirdnArgs p, int tiles_x, int tiles_y, long long total_tiles) {
    typedef float S;
    __shared__ __align__(16) S sx[U2_SM_H][U2_SM_W];
    __shared__ __align__(16) S dyn_stage[U2_TILE_H * U2_STAGE_W];     // output stage, used only when rows are unaligned
    S k[4][4];
    #pragma unroll
    for (int jy = 0; jy < 4; jy++)
        #pragma unroll
        for (int jx = 0; jx < 4; jx++) {
            S v = 0;
            if (jy < p.fh && jx < p.fw) {
                const int sy = p.flip ? jy : p.fh - 1 - jy, sxx = p.flip ? jx : p.fw - 1 - jx;
                v = p.f[sy * p.fs_y + sxx * p.fs_x] * p.gain;
            }
            k[jy][jx] = v;
        }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads, 4 x 4 outputs each
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const bool vec_store = rows_aligned4<T>(p);
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        unsigned r = (unsigned)t;
        const int bx = divmod_u32(r, (unsigned)tiles_x);
        const int by = divmod_u32(r, (unsigned)tiles_y);
        const int c = divmod_u32(r, (unsigned)p.c);
        const int n = (int)r;
        const int ox0 = bx * U2_TILE_W, oy0 = by * U2_TILE_H;
        // first input column / row any output of the tile can touch: (o0 - pad0 + P) / 2 (exact: the numerator is even)
        const int ix0 = (ox0 - p.padx0 + PX) >> 1, iy0 = (oy0 - p.pady0 + PY) >> 1;
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        __syncthreads();
        {
            constexpr int USED_W = 66, PER = (U2_SM_H * USED_W + 255) / 256;
            T v[PER];               // source type until the staging store (see fir_tile_kernel)
            const T zero = from_acc<T>(0.f);
            #pragma unroll
            for (int i = 0; i < PER; i++) {
                const int idx = threadIdx.x + 256 * i;
                const int ry = idx / USED_W, rx = idx - ry * USED_W;
                const int iy = iy0 + ry, ix = ix0 + rx;
                const bool ok = ry < U2_SM_H && iy >= 0 && iy < p.ih && ix >= 0 && ix < p.iw;
                v[i] = ok ? __ldg(xp + (long long)iy * p.xs_h + ix) : zero;
            }
            #pragma unroll
            for (int i = 0; i < PER; i++) {
                const int idx = threadIdx.x + 256 * i;
                const int ry = idx / USED_W, rx = idx - ry * USED_W;
                if (ry < U2_SM_H) sx[ry][rx] = to_acc<T>(v[i]);
            }
        }
        __syncthreads();
        S in[4][4];
        #pragma unroll
        for (int ry = 0; ry < 4; ry++) {
            const float2 lo = *reinterpret_cast<const float2*>(&sx[2 * ty + ry][2 * tx]);
            const float2 hi = *reinterpret_cast<const float2*>(&sx[2 * ty + ry][2 * tx + 2]);
            in[ry][0] = lo.x; in[ry][1] = lo.y; in[ry][2] = hi.x; in[ry][3] = hi.y;
        }
        S acc[4][4];
        #pragma unroll
        for (int a = 0; a < 4; a++) {
            const int qy = (PY + a) & 1, ry = (a + qy - PY) / 2;
            #pragma unroll
            for (int b = 0; b < 4; b++) {
                const int qx = (PX + b) & 1, rx = (b + qx - PX) / 2;
                S s = in[ry][rx] * k[qy][qx];
                s = fmaf(in[ry][rx + 1], k[qy][qx + 2], s);
                s = fmaf(in[ry + 1][rx], k[qy + 2][qx], s);
                s = fmaf(in[ry + 1][rx + 1], k[qy + 2][qx + 2], s);
                acc[a][b] = s;
            }
        }
        const int cx = tx * 4, cy = ty * 4;
        T* yp = y + n * p.ys_n + c * p.ys_c;
        if (vec_store) {
            #pragma unroll
            for (int a = 0; a < 4; a++) {
                const int oy = oy0 + cy + a, ox = ox0 + cx;
                if (oy >= p.oh || ox >= p.ow) continue;
                T* dst = yp + (long long)oy * p.ys_h + ox;
                if (ox + 3 < p.ow) store4<T>(dst, acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
                else {
                    #pragma unroll
                    for (int b = 0; b < 4; b++) if (ox + b < p.ow) dst[b] = from_acc<T>(acc[a][b]);
                }
            }
        } else {
            #pragma unroll
            for (int a = 0; a < 4; a++)
                *reinterpret_cast<float4*>(&dyn_stage[(cy + a) * U2_STAGE_W + cx]) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
            __syncthreads();
            store_tile_staged<T, U2_TILE_W, U2_TILE_H, U2_STAGE_W>(dyn_stage, yp, p.ys_h, ox0, oy0, p.ow, p.oh);
        }
    }
}

// ------------------------------------------------------------------------------------------
// One-dimensional filters of up to 16 taps with up = 2, down = 2 or neither along the filter axis: the two passes of a separable
// upfirdn2d (upfirdn2d.py:239-240), i.e. the 12-tap sym6 resampling of the ADA pipeline (augment.py:290,301).  Unit stride along W.
// A thread owns 4 consecutive outputs of a row.
//   fir_sep_x_kernel: the CTA (32 x 8 threads) stages the input span of its 128 x 8 output tile in shared memory;
//   fir_sep_y_kernel: taps run along H, so a thread reads 4 consecutive pixels of every tap row straight from global memory (coalesced).
constexpr int SEP_TAPS = 16, SEP_TW = 128, SEP_TH = 8;
constexpr int SEP_SPAN = 2 * (SEP_TW - 1) + SEP_TAPS;      // widest input span of a tile row (down = 2)

__device__ __forceinline__ void sep_load_taps(const UpfirdnArgs& p, int taps, long long stride, float* sk) {
    if (threadIdx.x < SEP_TAPS) {
        const int j = threadIdx.x;
        float v = 0.f;
        if (j < taps) v = p.f[(p.flip ? j : taps - 1 - j) * stride] * p.gain;
        sk[j] = v;
    }
}

template <class T, int UP, int DOWN>
__global__ void __launch_bounds__(256) fir_sep_x_kernel(UpfirdnArgs p, int tiles_x, int tiles_y, long long total_tiles) {
    __shared__ float sk[SEP_TAPS];
    __shared__ float sx[SEP_TH][SEP_SPAN + 2];
    sep_load_taps(p, p.fw, p.fs_x, sk);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    const long long planes = (long long)p.n * p.c;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        unsigned r = (unsigned)t;
        const int bx = divmod_u32(r, (unsigned)tiles_x);
        const int by = divmod_u32(r, (unsigned)tiles_y);
        const int c = divmod_u32(r, (unsigned)p.c);
        const int n = (int)r;
        (void)planes;
        const int ox0 = bx * SEP_TW, oy0 = by * SEP_TH;
        // first input column any output of the tile can touch, and the span length
        const int t0 = ox0 * DOWN - p.padx0;                                  // position of tap 0 of output ox0 on the (up-sampled) grid
        const int ix0 = UP == 1 ? t0 : (t0 >= 0 ? (t0 + 1) / 2 : -((-t0) / 2));     // ceil(t0 / UP)
        const int span = UP == 1 ? DOWN * (SEP_TW - 1) + p.fw : (SEP_TW - 1 + p.fw) / 2 + 1;
        const T* xp = x + n * p.xs_n + c * p.xs_c;
        __syncthreads();
        for (int ry = 0; ry < SEP_TH; ry++) {
            const int iy = oy0 + ry;                                          // rows are not resampled by this pass
            const T* row = xp + (long long)iy * p.xs_h;
            for (int k = threadIdx.x; k < span; k += 256) {
                const int ix = ix0 + k;
                sx[ry][k] = (iy < p.ih && ix >= 0 && ix < p.iw) ? to_acc<T>(__ldg(row + ix)) : 0.f;
            }
        }
        __syncthreads();
        const int oy = oy0 + ty;
        if (oy >= p.oh) continue;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        #pragma unroll
        for (int v = 0; v < 4; v++) {
            const int e = tx * 4 + v;                                         // output column inside the tile
            if (UP == 1) {
                const float* src = &sx[ty][e * DOWN];
                for (int j = 0; j < p.fw; j++) acc[v] = fmaf(src[j], sk[j], acc[v]);
            } else {
                const int tt = t0 + e;                                        // tap 0 position; taps j with (tt + j) even hit input (tt + j) / 2
                const int j0 = tt & 1;
                const int rel = ((tt + j0) >> 1) - ix0;
                const float* src = &sx[ty][rel];
                for (int j = j0, i = 0; j < p.fw; j += 2, i++) acc[v] = fmaf(src[i], sk[j], acc[v]);
            }
        }
        T* dst = y + n * p.ys_n + c * p.ys_c + (long long)oy * p.ys_h + ox0 + tx * 4;
        #pragma unroll
        for (int v = 0; v < 4; v++) if (ox0 + tx * 4 + v < p.ow) dst[v] = from_acc<T>(acc[v]);
    }
}

template <class T, int UP, int DOWN>
__global__ void __launch_bounds__(256) fir_sep_y_kernel(UpfirdnArgs p, long long total_quads, int quads_per_row) {
    __shared__ float sk[SEP_TAPS];
    sep_load_taps(p, p.fh, p.fs_y, sk);
    __syncthreads();
    const T* x = (const T*)p.x;
    T* y = (T*)p.y;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += (long long)gridDim.x * blockDim.x) {
        unsigned r = (unsigned)q;
        const int qx = divmod_u32(r, (unsigned)quads_per_row);
        const int oy = divmod_u32(r, (unsigned)p.oh);
        const int c = divmod_u32(r, (unsigned)p.c);
        const int n = (int)r;
        const int ox = qx * 4;
        const T* xp = x + n * p.xs_n + c * p.xs_c + ox;
        const int t0 = oy * DOWN - p.pady0;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const int j0 = UP == 1 ? 0 : (t0 & 1);
        for (int j = j0; j < p.fh; j += UP) {
            const int iy = UP == 1 ? t0 + j : (t0 + j) >> 1;
            if (iy < 0 || iy >= p.ih) continue;
            const T* row = xp + (long long)iy * p.xs_h;
            const float kk = sk[j];
            #pragma unroll
            for (int v = 0; v < 4; v++)
                if (ox + v < p.iw) acc[v] = fmaf(to_acc<T>(__ldg(row + v)), kk, acc[v]);
        }
        T* dst = y + n * p.ys_n + c * p.ys_c + (long long)oy * p.ys_h + ox;
        #pragma unroll
        for (int v = 0; v < 4; v++) if (ox + v < p.ow) dst[v] = from_acc<T>(acc[v]);
    }
}

template <class T>
static int launch_upfirdn(const UpfirdnArgs& p, cudaStream_t stream) {
    const bool tile_ok = p.upx == 1 && p.upy == 1 && p.downx == 1 && p.downy == 1 && p.fw <= 4 && p.fh <= 4 &&
                         p.xs_w == 1 && p.ys_w == 1 && sizeof(T) <= 4;
    const bool down2_ok = p.upx == 1 && p.upy == 1 && p.downx == 2 && p.downy == 2 && p.fw <= 4 && p.fh <= 4 &&
                          p.xs_w == 1 && p.ys_w == 1 && sizeof(T) <= 4;
    const bool up2_ok = p.upx == 2 && p.upy == 2 && p.downx == 1 && p.downy == 1 && p.fw <= 4 && p.fh <= 4 &&
                        p.xs_w == 1 && p.ys_w == 1 && sizeof(T) <= 4;
    const bool unit = p.xs_w == 1 && p.ys_w == 1 && sizeof(T) <= 4;
    const bool sep_x = unit && p.fh == 1 && p.fw > 1 && p.fw <= SEP_TAPS && p.upy == 1 && p.downy == 1 && p.pady0 == 0 && p.oh == p.ih &&
                       ((p.upx == 1 && p.downx <= 2) || (p.upx == 2 && p.downx == 1)) && !(p.upx == 1 && p.downx == 1 && p.fw <= 4);
    const bool sep_y = unit && p.fw == 1 && p.fh > 1 && p.fh <= SEP_TAPS && p.upx == 1 && p.downx == 1 && p.padx0 == 0 && p.ow == p.iw &&
                       ((p.upy == 1 && p.downy <= 2) || (p.upy == 2 && p.downy == 1)) && !(p.upy == 1 && p.downy == 1 && p.fh <= 4);
    if (sep_x) {
        const int tiles_x = (p.ow + SEP_TW - 1) / SEP_TW, tiles_y = (p.oh + SEP_TH - 1) / SEP_TH;
        const long long total = (long long)tiles_x * tiles_y * p.c * p.n;
        void (*kern)(UpfirdnArgs, int, int, long long) = p.upx == 2 ? fir_sep_x_kernel<T, 2, 1> : (p.downx == 2 ? fir_sep_x_kernel<T, 1, 2> : fir_sep_x_kernel<T, 1, 1>);
        long long blocks = total;
        const long long cap = (long long)sm_count() * occupancy_of(kern, 256, 0);
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, 256, 0, stream>>>(p, tiles_x, tiles_y, total);
    } else if (sep_y) {
        const int quads_per_row = (p.ow + 3) / 4;
        const long long total = (long long)quads_per_row * p.oh * p.c * p.n;
        void (*kern)(UpfirdnArgs, long long, int) = p.upy == 2 ? fir_sep_y_kernel<T, 2, 1> : (p.downy == 2 ? fir_sep_y_kernel<T, 1, 2> : fir_sep_y_kernel<T, 1, 1>);
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)sm_count() * occupancy_of(kern, 256, 0);
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, 256, 0, stream>>>(p, total, quads_per_row);
    } else if (up2_ok) {
        const int tiles_x = (p.ow + U2_TILE_W - 1) / U2_TILE_W, tiles_y = (p.oh + U2_TILE_H - 1) / U2_TILE_H;
        const long long total = (long long)tiles_x * tiles_y * p.c * p.n;
        void (*kern)(UpfirdnArgs, int, int, long long) =
            (p.padx0 & 1) ? ((p.pady0 & 1) ? fir_up2_kernel<T, 1, 1> : fir_up2_kernel<T, 1, 0>)
                          : ((p.pady0 & 1) ? fir_up2_kernel<T, 0, 1> : fir_up2_kernel<T, 0, 0>);
        long long blocks = total;
        const long long cap = (long long)sm_count() * occupancy_of(kern, 256, 0);
        if (blocks > cap) blocks = cap;
        kern<<<(unsigned)blocks, 256, 0, stream>>>(p, tiles_x, tiles_y, total);
    } else if (down2_ok) {
        const int tiles_x = (p.ow + D2_TILE_W - 1) / D2_TILE_W, tiles_y = (p.oh + D2_TILE_H - 1) / D2_TILE_H;
        const long long total = (long long)tiles_x * tiles_y * p.c * p.n;
        long long blocks = total;
        const long long cap = (long long)sm_count() * occupancy_of(fir_down2_kernel<T>, 256, 0);
        if (blocks > cap) blocks = cap;
        fir_down2_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p, tiles_x, tiles_y, total);
    } else if (tile_ok) {
        const int tiles_x = (p.ow + TILE_W - 1) / TILE_W, tiles_y = (p.oh + TILE_H - 1) / TILE_H;
        const long long total = (long long)tiles_x * tiles_y * p.c * p.n;
        long long blocks = total;
        const long long cap = (long long)sm_count() * occupancy_of(fir_tile_kernel<T>, 256, 0);
        if (blocks > cap) blocks = cap;
        fir_tile_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p, tiles_x, tiles_y, total);
    } else {
        const int quads_per_row = (p.ow + 3) / 4;
        const long long total = (long long)quads_per_row * p.oh * p.c * p.n;
        long long blocks = (total + 255) / 256;
        const long long cap = (long long)sm_count() * occupancy_of(generic_kernel<T>, 256, 0);
        if (blocks > cap) blocks = cap;
        generic_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(p, total, quads_per_row);
    }
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

template <>
int launch_upfirdn<double>(const UpfirdnArgs& p, cudaStream_t stream) {
    const int quads_per_row = (p.ow + 3) / 4;
    const long long total = (long long)quads_per_row * p.oh * p.c * p.n;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * occupancy_of(generic_kernel<double>, 256, 0);
    if (blocks > cap) blocks = cap;
    generic_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(p, total, quads_per_row);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

} // namespace pgpp

extern "C" int pgpp_upfirdn2d(const void* x, const float* f, void* y,
                              const int64_t in_size[4], const int64_t in_stride[4],
                              const int64_t out_size[4], const int64_t out_stride[4],
                              int fw, int fh, int64_t f_stride_x, int64_t f_stride_y,
                              int upx, int upy, int downx, int downy, int padx0, int pady0,
                              int flip, float gain, int dtype, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(x && f && y, "x, f and y must be device pointers");
    PGPP_REQUIRE(fw >= 1 && fh >= 1, "f must be at least 1x1");
    PGPP_REQUIRE(fw <= MAX_TAPS && fh <= MAX_TAPS, "f is too large (at most %d taps per axis)", MAX_TAPS);
    PGPP_REQUIRE(upx >= 1 && upy >= 1, "upsampling factor must be at least 1");
    PGPP_REQUIRE(downx >= 1 && downy >= 1, "downsampling factor must be at least 1");
    PGPP_REQUIRE(out_size[2] >= 1 && out_size[3] >= 1, "output must be at least 1x1");
    PGPP_REQUIRE(in_size[0] == out_size[0] && in_size[1] == out_size[1], "batch/channel mismatch between x and y");
    long long in_numel = 1, out_numel = 1;
    for (int i = 0; i < 4; i++) { in_numel *= in_size[i]; out_numel *= out_size[i]; }
    PGPP_REQUIRE(in_numel <= 2147483647LL, "x is too large");
    PGPP_REQUIRE(out_numel <= 2147483647LL, "output is too large");
    if (in_numel == 0 || out_numel == 0) return PGPP_OK;
    UpfirdnArgs p;
    p.x = x; p.f = f; p.y = y;
    p.n = (int)in_size[0]; p.c = (int)in_size[1]; p.ih = (int)in_size[2]; p.iw = (int)in_size[3];
    p.oh = (int)out_size[2]; p.ow = (int)out_size[3];
    p.xs_n = in_stride[0]; p.xs_c = in_stride[1]; p.xs_h = in_stride[2]; p.xs_w = in_stride[3];
    p.ys_n = out_stride[0]; p.ys_c = out_stride[1]; p.ys_h = out_stride[2]; p.ys_w = out_stride[3];
    p.fw = fw; p.fh = fh; p.fs_x = f_stride_x; p.fs_y = f_stride_y;
    p.upx = upx; p.upy = upy; p.downx = downx; p.downy = downy; p.padx0 = padx0; p.pady0 = pady0;
    p.flip = flip ? 1 : 0; p.gain = gain;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case PGPP_F32:  return launch_upfirdn<float>(p, s);
        case PGPP_F16:  return launch_upfirdn<__half>(p, s);
        case PGPP_BF16: return launch_upfirdn<__nv_bfloat16>(p, s);
        case PGPP_F64:  return launch_upfirdn<double>(p, s);
    }
    set_error("unsupported dtype %d", dtype);
    return PGPP_ERR_UNSUPPORTED;
}
