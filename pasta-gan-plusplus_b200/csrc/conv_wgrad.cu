// Weight gradient of the (transposed) convolution for sm_100a: tcgen05.mma with TMEM accumulators, TMA operand loads,
// split-K over the pixels with fp32 atomics.  Replaces the cuDNN call of torch_utils/ops/conv2d_gradfix.py:135-142
// (Conv2dGradWeight.forward, `aten::cudnn_convolution_backward_weight` / `..._transpose_backward_weight`).
//
//   G[a, b, ky, kx] = sum_{n, y, x} S[n, a, y, x] * L[n, b, y*s + ky - pad_y, x*s + kx - pad_x]
//
//   conv2d:            S = grad_output (a = out channel), L = input       (b = in channel)   -> G = dW[O, I, kh, kw]
//   conv_transpose2d:  S = input       (a = in channel),  L = grad_output (b = out channel)  -> G = dW[I, O, kh, kw]
//
// GEMM view: M = 128 channels a (TMEM lanes), N = block_n channels b for each of the T taps of a tap group (T * block_n
// TMEM columns), K = the pixels (n, y, x), walked in blocks of 64 = a TW x TH x TN box of S.  Both operands are the
// channels-innermost packed activation format of the forward kernel ([parts][N][H][W][c], pgpp_pack_activations), so
// here the reduction dimension is the OUTER one of a shared-memory tile: the operands are fed to the tensor core as
// MN-major matrices (instruction-descriptor bits 15 / 16).  A TMA box {64 channels, TW, TH, TN} lands as 64 pixel rows
// of 128 bytes = eight 8-row swizzle atoms, which is the canonical MN-major SWIZZLE_128B layout with stride byte offset
// 1024 (next 8 pixels) and leading byte offset = box size (next 64 channels).  The B tile of tap (ky, kx) is the same
// box of L shifted by the tap with element strides (s, s) - pixels outside the image come back as zeros (TMA
// out-of-bounds fill) = the convolution's zero padding.  One A stage is shared by all taps of the group and by all
// split-precision products (as in conv_igemm.cu).
//
// Work decomposition: job = (a block, b block, tap group, K split); persistent CTAs loop over jobs, each job ends with
// an atomicAdd epilogue into the fp32 [Ca][Cb][kh][kw] gradient (zeroed by this call).
//
// Pair mode (at most 64 channels a, stride 1, kh > 1, kw <= 4): 64 channels would fill half of the 128 accumulator lanes, so the
// taps move into the tile instead.  With the pixel index (y', x') = (L row, S column),
//   G[a, b, ky, kx] = sum S[a, y' - ky + pad_y, x'] * L[b, y', x' + kx - pad_x]:
// the vertical tap shifts S, the horizontal tap shifts L.  One S slab of th + 2 * ceil(kh / 2) - 1 rows serves all ky as
// row-shifted views, and the A descriptor's leading byte offset (distance of the "next 64 channels") is ONE ROW of that slab:
// lanes 0-63 of an MMA see the view of tap ky, lanes 64-127 the view of tap ky - 1, from the same bytes.  The B tile is kw boxes
// of L, one per kx (N = kw * 64 columns).  A 3x3 filter is 2 MMAs of 128 x 192 per K step instead of 9 of 128 x 64 with half of
// the lanes idle: 1.5x less tensor time, 4.5x fewer instructions, 2.2x less shared-memory fill per K block.
//
// Warp roles (320 threads, 1 CTA per SM): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-9 epilogue.
#include <stdlib.h>
#include "tc_ptx.cuh"

namespace pgpp {

constexpr int kWgThreads = 320;
constexpr unsigned kChunkBytes = 64u * 128u;    // one TMA box: 64 pixels x 64 channels of bf16

// shared-memory matrix descriptor of an MN-major SWIZZLE_128B operand (cute/atom/mma_traits_sm100.hpp, make_umma_desc:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 channels contiguous, next 64 channels LBO bytes further,
// K rows (pixels) 128 bytes apart, next group of 8 pixels SBO bytes further
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct WgradParams {
    int n, ca, cb;
    int taps, kw, pad_y, pad_x, stride, dil_y;
    int tw, th, tn, tiles_x, tiles_y, tiles_n;
    long long kblocks;
    int parts;
    int block_n, t_group, n_groups;
    int a_blks, b_blks, ksplit;
    long long total_jobs;
    int a_stages, b_stages;
    int b_chunks;                       // 64-channel boxes per B tile (block_n / 64)
    int reuse;                          // 1: one L slab of th + kh - 1 rows per (kx, K block), vertical taps are row-shifted views
    int pair, kh, npairs;               // pair mode (see above): npairs = ceil(kh / 2) accumulators of kw * 64 columns
    unsigned b_chunk_bytes;             // bytes of one L box (= leading byte offset of the B descriptor)
    unsigned ky_step_bytes;             // tw * 128: shift of the slab view per vertical tap
    unsigned a_part_bytes, a_stage_bytes, b_bytes;
    unsigned idesc, tmem_cols;
    int cb_pad;
    float* ws;
};

struct WgJob { int ab, bb, tap0, tap_step, nt; long long k_begin, k_end; };   // accumulator t holds tap0 + t * tap_step

__device__ __forceinline__ WgJob decode_job(const WgradParams& p, long long job) {
    WgJob j;
    const int ks = (int)(job % p.ksplit); job /= p.ksplit;
    const int tg = (int)(job % p.n_groups); job /= p.n_groups;
    j.bb = (int)(job % p.b_blks);
    j.ab = (int)(job / p.b_blks);
    if (p.pair) { j.tap0 = 0; j.tap_step = 0; j.nt = p.npairs; }              // all taps: accumulator t = the ky pair (kh-1-2t, kh-2-2t), every kx
    else if (p.reuse) { j.tap0 = tg; j.tap_step = p.kw; j.nt = p.t_group; } // group = one kx, all ky
    else { j.tap0 = tg * p.t_group; j.tap_step = 1; j.nt = min(p.t_group, p.taps - j.tap0); }
    j.k_begin = p.kblocks * ks / p.ksplit;
    j.k_end = p.kblocks * (ks + 1) / p.ksplit;
    return j;
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_s, const __grid_constant__ CUtensorMap map_l, const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_base = smem_base + p.a_stages * p.a_stage_bytes;
    const uint32_t bar_base = b_base + p.b_stages * p.b_bytes;
    const int SA = p.a_stages, SB = p.b_stages;
    auto afull_bar = [&](int s) { return bar_base + 8u * s; };
    auto aempty_bar = [&](int s) { return bar_base + 8u * (SA + s); };
    auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * SA + s); };
    auto bempty_bar = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
    const uint32_t tfull_bar = bar_base + 8u * (2 * SA + 2 * SB);
    const uint32_t tempty_bar = tfull_bar + 8u;
    const uint32_t tmem_slot = tfull_bar + 16u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_s) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_l) : "memory");
        for (int s = 0; s < SA; s++) { mbar_init(afull_bar(s), 1); mbar_init(aempty_bar(s), 1); }
        for (int s = 0; s < SB; s++) { mbar_init(bfull_bar(s), 1); mbar_init(bempty_bar(s), 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int sa = 0; uint32_t pha = 0;
            int sb = 0; uint32_t phb = 0;
            for (long long job = blockIdx.x; job < p.total_jobs; job += gridDim.x) {
                const WgJob j = decode_job(p, job);
                int tx = (int)(j.k_begin % p.tiles_x);
                int ty = (int)((j.k_begin / p.tiles_x) % p.tiles_y);
                int tz = (int)(j.k_begin / ((long long)p.tiles_x * p.tiles_y));
                for (long long kb = j.k_begin; kb < j.k_end; kb++) {
                    const int x0 = tx * p.tw, y0 = ty * p.th, n0 = tz * p.tn;
                    mbar_wait(aempty_bar(sa), pha ^ 1);
                    mbar_expect_tx(afull_bar(sa), p.parts * p.a_part_bytes);
                    if (p.pair) {
                        for (int pa = 0; pa < p.parts; pa++)        // S slab: the view of tap ky starts kh - 1 - ky rows into it
                            tma_load_5d(smem_base + sa * p.a_stage_bytes + pa * p.a_part_bytes, &map_s, afull_bar(sa),
                                        0, x0, y0 + p.pad_y - (p.kh - 1), n0, pa);
                    } else {
                        for (int pa = 0; pa < p.parts; pa++)
                            for (int ch = 0; ch < 2; ch++)
                                tma_load_5d(smem_base + sa * p.a_stage_bytes + pa * p.a_part_bytes + ch * kChunkBytes, &map_s, afull_bar(sa),
                                            j.ab * 128 + ch * 64, x0, y0, n0, pa);
                    }
                    if (++sa == SA) { sa = 0; pha ^= 1; }
                    if (p.pair) {
                        for (int pb = 0; pb < p.parts; pb++) {
                            mbar_wait(bempty_bar(sb), phb ^ 1);
                            mbar_expect_tx(bfull_bar(sb), p.b_bytes);
                            for (int kx = 0; kx < p.kw; kx++)       // one L box per horizontal tap
                                tma_load_5d(b_base + sb * p.b_bytes + kx * p.b_chunk_bytes, &map_l, bfull_bar(sb),
                                            j.bb * 64, x0 + kx - p.pad_x, y0, n0, pb);
                            if (++sb == SB) { sb = 0; phb ^= 1; }
                        }
                    } else if (p.reuse) {
                        const int lx = x0 + j.tap0 - p.pad_x, ly = y0 - p.pad_y;
                        for (int pb = 0; pb < p.parts; pb++) {
                            mbar_wait(bempty_bar(sb), phb ^ 1);
                            mbar_expect_tx(bfull_bar(sb), p.b_bytes);
                            for (int ch = 0; ch < p.b_chunks; ch++)
                                tma_load_5d(b_base + sb * p.b_bytes + ch * p.b_chunk_bytes, &map_l, bfull_bar(sb),
                                            j.bb * p.block_n + ch * 64, lx, ly, n0, pb);
                            if (++sb == SB) { sb = 0; phb ^= 1; }
                        }
                    } else {
                        for (int t = 0; t < j.nt; t++) {
                            const int tap = j.tap0 + t;
                            const int ky = tap / p.kw, kx = tap - ky * p.kw;
                            const int lx = x0 * p.stride + kx - p.pad_x;
                            const int ly = y0 * p.stride + ky * p.dil_y - p.pad_y;
                            for (int pb = 0; pb < p.parts; pb++) {
                                mbar_wait(bempty_bar(sb), phb ^ 1);
                                mbar_expect_tx(bfull_bar(sb), p.b_bytes);
                                for (int ch = 0; ch < p.b_chunks; ch++)
                                    tma_load_5d(b_base + sb * p.b_bytes + ch * p.b_chunk_bytes, &map_l, bfull_bar(sb),
                                                j.bb * p.block_n + ch * 64, lx, ly, n0, pb);
                                if (++sb == SB) { sb = 0; phb ^= 1; }
                            }
                        }
                    }
                    if (++tx == p.tiles_x) { tx = 0; if (++ty == p.tiles_y) { ty = 0; tz++; } }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        const uint64_t desc_hi = make_smem_desc_mn(p.pair ? p.ky_step_bytes : kChunkBytes, 1024u);
        const uint64_t desc_b_hi = make_smem_desc_mn(p.b_chunk_bytes, 1024u);
        const uint32_t ky16 = p.ky_step_bytes >> 4;
        const uint32_t part16 = p.a_part_bytes >> 4;
        const int parts = p.parts;
        int sa = 0; uint32_t pha = 0;
        int sb = 0; uint32_t phb = 0;
        uint32_t acc_phase = 0;
        for (long long job = blockIdx.x; job < p.total_jobs; job += gridDim.x) {
            const WgJob j = decode_job(p, job);
            mbar_wait(tempty_bar, acc_phase ^ 1);                   // epilogue has drained the accumulators
            tc_fence_after();
            for (long long kb = j.k_begin; kb < j.k_end; kb++) {
                mbar_wait(afull_bar(sa), pha);
                tc_fence_after();
                const uint32_t a16 = (smem_base + sa * p.a_stage_bytes) >> 4;
                const uint32_t not_first = kb > j.k_begin ? 1u : 0u;
                if (p.pair) {
                    #pragma unroll
                    for (int pb = 0; pb < 3; pb++) {
                        if (pb >= parts) break;
                        mbar_wait(bfull_bar(sb), phb);
                        tc_fence_after();
                        const uint64_t db = desc_b_hi | (uint64_t)(((b_base + sb * p.b_bytes) >> 4) & 0x3FFF);
                        for (int t = 0; t < j.nt; t++) {
                            const uint32_t tmem_d = tmem_base + (uint32_t)(t * p.block_n);
                            #pragma unroll
                            for (int pa = 0; pa < 3; pa++) {
                                if (pa + pb >= parts) break;
                                const uint64_t da = desc_hi | (uint64_t)((a16 + pa * part16 + 2 * t * ky16) & 0x3FFF);
                                if (leader) {
                                    umma_bf16(tmem_d, da, db, p.idesc, (pb | pa) ? 1u : not_first);
                                    umma_bf16(tmem_d, da + 128, db + 128, p.idesc, 1);
                                    umma_bf16(tmem_d, da + 256, db + 256, p.idesc, 1);
                                    umma_bf16(tmem_d, da + 384, db + 384, p.idesc, 1);
                                }
                            }
                        }
                        if (leader) umma_commit(bempty_bar(sb));
                        if (++sb == SB) { sb = 0; phb ^= 1; }
                    }
                } else if (p.reuse) {
                    #pragma unroll
                    for (int pb = 0; pb < 3; pb++) {
                        if (pb >= parts) break;
                        mbar_wait(bfull_bar(sb), phb);
                        tc_fence_after();
                        const uint32_t b16 = (b_base + sb * p.b_bytes) >> 4;
                        for (int t = 0; t < j.nt; t++) {
                            const uint32_t tmem_d = tmem_base + (uint32_t)(t * p.block_n);
                            const uint64_t db = desc_b_hi | (uint64_t)((b16 + t * ky16) & 0x3FFF);
                            #pragma unroll
                            for (int pa = 0; pa < 3; pa++) {
                                if (pa + pb >= parts) break;
                                const uint64_t da = desc_hi | (uint64_t)((a16 + pa * part16) & 0x3FFF);
                                if (leader) {
                                    umma_bf16(tmem_d, da, db, p.idesc, (pb | pa) ? 1u : not_first);
                                    umma_bf16(tmem_d, da + 128, db + 128, p.idesc, 1);
                                    umma_bf16(tmem_d, da + 256, db + 256, p.idesc, 1);
                                    umma_bf16(tmem_d, da + 384, db + 384, p.idesc, 1);
                                }
                            }
                        }
                        if (leader) umma_commit(bempty_bar(sb));
                        if (++sb == SB) { sb = 0; phb ^= 1; }
                    }
                } else {
                    for (int t = 0; t < j.nt; t++) {
                        const uint32_t tmem_d = tmem_base + (uint32_t)(t * p.block_n);
                        uint32_t acc = not_first;
                        #pragma unroll
                        for (int pb = 0; pb < 3; pb++) {
                            if (pb >= parts) break;
                            mbar_wait(bfull_bar(sb), phb);
                            tc_fence_after();
                            const uint64_t db = desc_b_hi | (uint64_t)(((b_base + sb * p.b_bytes) >> 4) & 0x3FFF);
                            #pragma unroll
                            for (int pa = 0; pa < 3; pa++) {
                                if (pa + pb >= parts) break;
                                const uint64_t da = desc_hi | (uint64_t)((a16 + pa * part16) & 0x3FFF);
                                if (leader) {
                                    // one MMA covers 16 pixels = two 8-row atoms: advance the start address by 2048 bytes
                                    umma_bf16(tmem_d, da, db, p.idesc, acc);
                                    umma_bf16(tmem_d, da + 128, db + 128, p.idesc, 1);
                                    umma_bf16(tmem_d, da + 256, db + 256, p.idesc, 1);
                                    umma_bf16(tmem_d, da + 384, db + 384, p.idesc, 1);
                                }
                                acc = 1;
                            }
                            if (leader) umma_commit(bempty_bar(sb));
                            if (++sb == SB) { sb = 0; phb ^= 1; }
                        }
                    }
                }
                if (leader) umma_commit(aempty_bar(sa));
                if (++sa == SA) { sa = 0; pha ^= 1; }
            }
            if (leader) umma_commit(tfull_bar);
            acc_phase ^= 1;
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps =====================
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        uint32_t acc_phase = 0;
        for (long long job = blockIdx.x; job < p.total_jobs; job += gridDim.x) {
            const WgJob j = decode_job(p, job);
            const int a = j.ab * 128 + quarter * 32 + lane;
            const int chunks = j.nt * p.block_n / 16;
            const int c_mid = (chunks + 1) / 2;
            const int c_begin = half ? c_mid : 0, c_end = half ? chunks : c_mid;
            mbar_wait(tfull_bar, acc_phase);
            tc_fence_after();
            for (int c = c_begin; c < c_end; c++) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 16), v);
                const int col = c * 16;
                const int t = col / p.block_n;
                if (p.pair) {
                    // lanes 0-63: tap row kh-1-2t, lanes 64-127: the row above it; columns = kx * 64 + b
                    const int within = col - t * p.block_n;
                    const int kx = within >> 6;
                    const int ky = p.kh - 1 - 2 * t - (quarter >> 1);
                    const int ap = (quarter & 1) * 32 + lane;
                    const int b0 = j.bb * 64 + (within & 63);
                    if (ky >= 0 && ap < p.ca) {
                        float* dst = p.ws + ((long long)(ky * p.kw + kx) * p.ca + ap) * p.cb_pad + b0;
                        #pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                         ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
                    }
                    continue;
                }
                const int b0 = j.bb * p.block_n + (col - t * p.block_n);
                if (a < p.ca && b0 < p.cb_pad) {
                    // workspace [tap][ca][cb_pad]: the 16 columns of this chunk are 64 contiguous bytes -> 4 vector reductions
                    float* dst = p.ws + ((long long)(j.tap0 + t * j.tap_step) * p.ca + a) * p.cb_pad + b0;
                    #pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                     ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            acc_phase ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// workspace [taps][ca][cb_pad] -> G [ca][cb][taps]: one thread per (a, b) reads its taps (coalesced along b) and writes
// taps consecutive floats
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ out, int ca, int cb, int cb_pad, int taps, float scale) {
    const long long total = (long long)ca * cb;
    const long long plane = (long long)ca * cb_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i % cb);
        const long long a = i / cb;
        const float* src = ws + a * cb_pad + b;
        float* dst = out + i * taps;
        for (int t = 0; t < taps; t++) dst[t] = src[t * plane] * scale;
    }
}

} // namespace pgpp

extern "C" int pgpp_conv2d_wgrad(const pgpp_wgrad_desc* d, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(d != nullptr, "desc is NULL");
    PGPP_REQUIRE(d->small && d->large && d->out && d->workspace, "small, large, out and workspace must be device pointers");
    PGPP_REQUIRE(((uintptr_t)d->workspace & 15) == 0, "workspace must be 16-byte aligned");
    PGPP_REQUIRE(d->n >= 1 && d->ca >= 1 && d->cb >= 1 && d->hs >= 1 && d->ws >= 1 && d->hl >= 1 && d->wl >= 1, "empty problem");
    PGPP_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->kh * d->kw <= 64, "filter must be between 1x1 and 64 taps");
    PGPP_REQUIRE(d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
    PGPP_REQUIRE(d->products == 1 || d->products == 3 || d->products == 6, "products must be 1, 3 or 6");
    const int parts = d->products == 1 ? 1 : (d->products == 3 ? 2 : 3);
    PGPP_REQUIRE(d->s_parts >= parts && d->l_parts >= parts, "operand parts do not cover the requested products");
    PGPP_REQUIRE(d->ca_pad >= d->ca && d->ca_pad % 64 == 0 && d->cb_pad >= d->cb && d->cb_pad % 64 == 0,
                 "ca_pad / cb_pad must be multiples of 64 covering ca / cb (channels beyond ca / cb zero)");
    const int s_stride = d->s_pixel_stride > 0 ? d->s_pixel_stride : d->ca_pad;
    const int l_stride = d->l_pixel_stride > 0 ? d->l_pixel_stride : d->cb_pad;
    PGPP_REQUIRE(s_stride >= d->ca_pad && s_stride % 8 == 0 && l_stride >= d->cb_pad && l_stride % 8 == 0,
                 "pixel strides must be >= the padded channel counts and multiples of 8");
    PGPP_REQUIRE(((uintptr_t)d->small & 15) == 0 && ((uintptr_t)d->large & 15) == 0, "operands must be 16-byte aligned");

    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return PGPP_ERR_CUDA; }

    WgradParams p;
    p.n = d->n; p.ca = d->ca; p.cb = d->cb;
    p.taps = d->kh * d->kw; p.kw = d->kw; p.pad_y = d->pad_y; p.pad_x = d->pad_x; p.stride = d->stride;
    p.dil_y = d->dil_y > 1 ? d->dil_y : 1;
    PGPP_REQUIRE(p.dil_y == 1 || d->stride == 1, "wgrad: dil_y > 1 needs stride 1");
    int bn = d->cb_pad >= 128 ? 128 : 64;
    if (p.taps == 1 && d->cb_pad >= 256) bn = 256;
    p.block_n = bn;
    p.b_chunks = bn / 64;
    // K block = 64 pixels of S.  Stride-1 filters with kh > 1 on images of at least 64 pixels take a 16 (or 8) wide box so
    // that one L slab of th + kh - 1 rows serves all vertical taps of a filter column (the ky shift is a whole number of
    // 8-pixel swizzle atoms); everything else takes the widest box that fits the image, then rows, then samples.
    p.reuse = (d->stride == 1 && d->kh > 1 && d->kh * bn <= 512 && d->ws >= 8 && (long long)d->ws * d->hs >= 64 && p.dil_y == 1) ? 1 : 0;
    if (env_flags().wgrad_no_reuse) p.reuse = 0;
    p.kh = d->kh; p.npairs = (d->kh + 1) / 2;
    p.pair = (d->ca_pad == 64 && d->stride == 1 && d->kh > 1 && p.dil_y == 1 && d->kw * 64 <= 256 && p.npairs * d->kw * 64 <= 512 &&
              d->ws >= 8 && (long long)d->ws * d->hl >= 64 && !env_flags().wgrad_no_pair) ? 1 : 0;
    if (p.pair) {
        p.reuse = 0;
        bn = d->kw * 64; p.block_n = bn; p.b_chunks = d->kw;
        p.tw = d->ws >= 16 ? 16 : 8; p.th = 64 / p.tw; p.tn = 1;
        p.t_group = p.npairs; p.n_groups = 1;
    } else if (p.reuse) {
        p.tw = d->ws >= 16 ? 16 : 8; p.th = 64 / p.tw; p.tn = 1;
        p.t_group = d->kh; p.n_groups = d->kw;
    } else {
        p.tw = pow2_ceil(d->ws); if (p.tw > 64) p.tw = 64;
        p.th = pow2_ceil(d->hs); if (p.th > 64 / p.tw) p.th = 64 / p.tw;
        p.tn = 64 / (p.tw * p.th);
        int t_max = 512 / bn; if (t_max > p.taps) t_max = p.taps;
        p.n_groups = (p.taps + t_max - 1) / t_max;
        p.t_group = (p.taps + p.n_groups - 1) / p.n_groups;
        p.n_groups = (p.taps + p.t_group - 1) / p.t_group;
    }
    p.tiles_x = (d->ws + p.tw - 1) / p.tw;
    p.tiles_y = ((p.pair ? d->hl : d->hs) + p.th - 1) / p.th;       // pair mode walks the rows of L (see the header)
    p.tiles_n = (d->n + p.tn - 1) / p.tn;
    p.kblocks = (long long)p.tiles_n * p.tiles_x * p.tiles_y;
    p.parts = parts;
    const int l_rows = p.reuse ? p.th + d->kh - 1 : p.th;       // rows of one L box (after the element stride)
    const int s_rows = p.pair ? p.th + 2 * p.npairs - 1 : p.th; // rows of one S box
    p.b_chunk_bytes = (unsigned)(l_rows * p.tw * p.tn) * 128u;
    p.ky_step_bytes = (unsigned)p.tw * 128u;
    p.a_blks = (d->ca + 127) / 128;
    p.b_blks = p.pair ? (d->cb + 63) / 64 : (d->cb + bn - 1) / bn;
    const long long base_jobs = (long long)p.a_blks * p.b_blks * p.n_groups;
    const int sms = sm_count();
    {
        // split K so that the jobs fill whole waves: minimise waves(ks) * ceil(kblocks / ks), ties to the smaller split
        long long ks_max = (4ll * sms + base_jobs - 1) / base_jobs;
        if (ks_max > p.kblocks) ks_max = p.kblocks;
        if (ks_max < 1) ks_max = 1;
        int best = 1; double best_cost = 1e30;
        for (int ks = 1; ks <= ks_max; ks++) {
            const long long waves = (base_jobs * ks + sms - 1) / sms;
            const double cost = (double)waves * (double)((p.kblocks + ks - 1) / ks);
            if (cost < best_cost * 0.999) { best_cost = cost; best = ks; }
        }
        p.ksplit = best;
    }
    p.total_jobs = base_jobs * p.ksplit;
    p.a_part_bytes = p.pair ? (unsigned)(s_rows * p.tw) * 128u : 2u * kChunkBytes;
    p.a_stage_bytes = (unsigned)parts * p.a_part_bytes;
    p.b_bytes = (unsigned)p.b_chunks * p.b_chunk_bytes;
    // instruction descriptor: fp32 accumulate, bf16 A / B, both MN-major (bits 15, 16), N = block_n, M = 128
    // a_format / b_format: 1 = bf16, 0 = f16 (fp16 layers); both operands MN-major (bits 15 / 16)
    const unsigned ab_fmt = d->operand_f16 ? 0u : ((1u << 7) | (1u << 10));
    PGPP_REQUIRE(!d->operand_f16 || d->products == 1, "fp16 operands are a single part (products must be 1)");
    p.idesc = (1u << 4) | ab_fmt | (1u << 15) | (1u << 16) | ((unsigned)(bn >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    unsigned cols = (unsigned)pow2_ceil(p.t_group * bn); if (cols < 32) cols = 32;
    PGPP_REQUIRE(cols <= 512, "internal: tap group does not fit TMEM");
    p.tmem_cols = cols;
    auto smem_need = [&](long long a_st, long long b_st) -> long long {
        return 1024 + a_st * p.a_stage_bytes + b_st * p.b_bytes + 8 * (2 * a_st + 2 * b_st + 2) + 16;
    };
    const long long smem_max = 227 * 1024;
    p.a_stages = 2; p.b_stages = 2;
    if (smem_need(2, 2) > smem_max) { set_error("tile does not fit shared memory"); return PGPP_ERR_UNSUPPORTED; }
    const int b_want = 2 * ((p.reuse || p.pair) ? 1 : p.t_group) * parts;      // two K blocks of L tiles in flight
    while (p.b_stages < b_want && p.b_stages < 24 && smem_need(p.a_stages, p.b_stages + 1) <= smem_max) p.b_stages++;
    while (p.a_stages < 4 && smem_need(p.a_stages + 1, p.b_stages) <= smem_max) p.a_stages++;
    while (p.b_stages < 24 && smem_need(p.a_stages, p.b_stages + 1) <= smem_max) p.b_stages++;
    const size_t smem_bytes = (size_t)smem_need(p.a_stages, p.b_stages);
    p.cb_pad = d->cb_pad;
    p.ws = d->workspace;

    CUtensorMap map_s, map_l;
    {
        const cuuint64_t sp = (cuuint64_t)s_stride * 2;
        const cuuint64_t dims[5] = {(cuuint64_t)d->ca_pad, (cuuint64_t)d->ws, (cuuint64_t)d->hs, (cuuint64_t)d->n, (cuuint64_t)d->s_parts};
        const cuuint64_t strides[4] = {sp, sp * d->ws, sp * d->ws * d->hs, sp * d->ws * d->hs * d->n};
        const cuuint32_t box[5] = {64, (cuuint32_t)p.tw, (cuuint32_t)s_rows, (cuuint32_t)p.tn, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&map_s, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->small), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(small operand) failed with CUresult %d", (int)r); return PGPP_ERR_CUDA; }
        const cuuint64_t lp = (cuuint64_t)l_stride * 2;
        const cuuint64_t ldims[5] = {(cuuint64_t)d->cb_pad, (cuuint64_t)d->wl, (cuuint64_t)d->hl, (cuuint64_t)d->n, (cuuint64_t)d->l_parts};
        const cuuint64_t lstrides[4] = {lp, lp * d->wl, lp * d->wl * d->hl, lp * d->wl * d->hl * d->n};
        const cuuint32_t lbox[5] = {64, (cuuint32_t)(p.tw * d->stride), (cuuint32_t)(l_rows * d->stride), (cuuint32_t)p.tn, 1};
        const cuuint32_t lestr[5] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1, 1};
        r = encode(&map_l, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->large), ldims, lstrides, lbox, lestr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(large operand) failed with CUresult %d", (int)r); return PGPP_ERR_CUDA; }
    }
    {
        static bool done[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !done[dev]) {
            PGPP_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
    }
    PGPP_CUDA_OK(cudaMemsetAsync(d->workspace, 0, sizeof(float) * (size_t)p.taps * d->ca * d->cb_pad, (cudaStream_t)stream));
    long long grid = p.total_jobs;
    if (grid > sms) grid = sms;
    wgrad_kernel<<<(unsigned)grid, kWgThreads, smem_bytes, (cudaStream_t)stream>>>(map_s, map_l, p);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    {
        const long long total = (long long)d->ca * d->cb;
        long long blocks = (total + 255) / 256;
        if (blocks > 8ll * sms) blocks = 8ll * sms;
        wgrad_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d->workspace, d->out, d->ca, d->cb, d->cb_pad, p.taps,
                                                                                    d->out_scale != 0.f ? d->out_scale : 1.f);
        count_launch();
        PGPP_CUDA_OK(cudaGetLastError());
    }
    return PGPP_OK;
}
