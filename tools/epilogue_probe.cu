// SASS instruction count of one epilogue chunk pass (development aid, no GPU needed):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pasta-gan-plusplus_b200/csrc -cubin -o /tmp/epi.cubin tools/epilogue_probe.cu
//   cuobjdump -sass -fun 'epi_probe_bf16' /tmp/epi.cubin | grep -c '^ *\/\*[0-9a-f]*\*\/'
#include "../pasta-gan-plusplus_b200/csrc/conv_igemm.cu"

namespace pgpp {
template <int A, class OT, bool CLAMP, bool ACC>
__device__ __forceinline__ void probe_body(const IgemmParams& p, const float2* s_cs) {
    TileCoord tc{0, (int)blockIdx.y, (int)blockIdx.x, 0};
    PixelCoord pc{(int)threadIdx.x & 7, (int)threadIdx.x >> 3, 0};
    epilogue_fast<A, OT, CLAMP, ACC>(p, tc, 0u, pc, 0, 32, s_cs, 0.25f);
}
}
extern "C" __global__ void epi_probe_bf16(const __grid_constant__ pgpp::IgemmParams p) {
    extern __shared__ float2 s_cs[];
    pgpp::probe_body<PGPP_ACT_LRELU, __nv_bfloat16, true, false>(p, s_cs);
}
extern "C" __global__ void epi_probe_f32(const __grid_constant__ pgpp::IgemmParams p) {
    extern __shared__ float2 s_cs[];
    pgpp::probe_body<PGPP_ACT_LRELU, float, true, false>(p, s_cs);
}
extern "C" __global__ void epi_probe_f32_acc(const __grid_constant__ pgpp::IgemmParams p) {
    extern __shared__ float2 s_cs[];
    pgpp::probe_body<PGPP_ACT_LINEAR, float, false, true>(p, s_cs);
}
