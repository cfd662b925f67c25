// Probe: does tcgen05.mma read a K-major SWIZZLE_128B operand correctly when the descriptor start address is shifted by whole
// 128-byte rows (not a multiple of the 1024-byte swizzle atom) and the 8-row group stride (SBO) is 1280 bytes instead of 1024?
// That is what "one activation slab with an x halo, filter columns as row-shifted views" needs (known gaps).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I pasta-gan-plusplus_b200/csrc tools/umma_probe.cu -o /tmp/umma_probe && /tmp/umma_probe
//
// The slab is written the way TMA writes it (16-byte chunk index XOR (row & 7), rows anchored at a 1024-byte aligned base);
// B is a 64 x 64 identity, so D[m][n] = A_row(m)[n] shows exactly which bytes the tensor core fetched for row m.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

using namespace pgpp;

constexpr int kRows = 256;          // pixels in the slab

__host__ __device__ inline float a_value(int q, int k) { return (float)(((q * 7 + k * 3) % 13) - 6); }

__global__ void __launch_bounds__(128, 1) probe_kernel(float* out, int q0, unsigned sbo_bytes, unsigned base_offset) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    __nv_bfloat16* slab = reinterpret_cast<__nv_bfloat16*>(gen);                          // kRows x 128 B
    __nv_bfloat16* bt = reinterpret_cast<__nv_bfloat16*>(gen + kRows * 128);              // 64 x 128 B, 1024-aligned
    uint64_t* bar = reinterpret_cast<uint64_t*>(gen + kRows * 128 + 64 * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < kRows * 64; i += blockDim.x) {
        const int q = i / 64, k = i % 64;
        const int chunk = k / 8, within = k % 8;
        slab[q * 64 + ((chunk ^ (q & 7)) * 8) + within] = __float2bfloat16(a_value(q, k));
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        const int chunk = k / 8, within = k % 8;
        bt[n * 64 + ((chunk ^ (n & 7)) * 8) + within] = __float2bfloat16(n == k ? 1.f : 0.f);
    }
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t da = make_smem_desc(base + (uint32_t)q0 * 128u, 2u, sbo_bytes) | ((uint64_t)(base_offset & 7) << 49);
        uint64_t db = make_smem_desc(base + kRows * 128, 2u, 1024u);
        for (int k = 0; k < 4; k++) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
        umma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; j++) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

static int run(int q0, unsigned sbo, unsigned base_offset, int rows_per_group_pitch) {
    float* d_out;
    cudaMalloc(&d_out, 128 * 64 * sizeof(float));
    const size_t smem = 1024 + kRows * 128 + 64 * 128 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(d_out, q0, sbo, base_offset);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("q0=%d sbo=%u base_offset=%u: CUDA error %s\n", q0, sbo, base_offset, cudaGetErrorString(e)); return -1; }
    std::vector<float> h(128 * 64);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    int bad = 0, first_bad_m = -1;
    for (int m = 0; m < 128; m++) {
        const int q = q0 + (m / 8) * rows_per_group_pitch + (m % 8);
        for (int n = 0; n < 64; n++)
            if (h[m * 64 + n] != a_value(q, n)) { bad++; if (first_bad_m < 0) first_bad_m = m; }
    }
    printf("q0=%2d sbo=%4u base_offset=%u: %s (%d mismatches%s)\n", q0, sbo, base_offset, bad ? "MISMATCH" : "exact", bad,
           bad ? "" : "");
    if (bad) printf("    first mismatching row m=%d\n", first_bad_m);
    return bad;
}

int main() {
    printf("reference case (aligned start, SBO 1024):\n");
    run(0, 1024, 0, 8);
    run(8, 1024, 0, 8);
    printf("row-shifted start, SBO 1024 (filter-column view inside one 8-aligned slab):\n");
    for (int q0 : {1, 2, 3, 9}) { run(q0, 1024, 0, 8); run(q0, 1024, q0 & 7, 8); }
    printf("row-shifted start, SBO 1280 (10-pixel slab rows, 8-pixel tiles):\n");
    for (int q0 : {0, 1, 2, 10, 11, 21}) { run(q0, 1280, 0, 10); run(q0, 1280, q0 & 7, 10); }
    return 0;
}
