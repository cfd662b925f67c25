// grid_sample for the ADA augmentation pipeline (torch_utils/ops/grid_sample_gradfix.py:27-83; augment.py:290-301):
// 2-D, mode = bilinear, padding_mode = zeros, align_corners = False - the only configuration the reference uses.
//   forward   out[n,c,i,j] = sum over the 4 neighbours of (x, y) = unnormalise(grid[n,i,j]) of w_k * input[n,c,y_k,x_k]
//   backward  grad_input (scatter-add of w_k * grad_out) and grad_grid (d out / d x, d out / d y times W/2, H/2)
// Gather / scatter kernels bound by HBM (and L2 atomics in the backward); one thread per output pixel, channels in a loop so
// that the index arithmetic is done once and every channel plane is read with the same (coalesced along j) pattern.
// Replaces aten::grid_sampler_2d / aten::grid_sampler_2d_backward at the reference's two call sites.
#include "common.cuh"

namespace pgpp {

struct GridArgs {
    const float* input; const float* grid; const float* grad_out; float* out; float* grad_input; float* grad_grid;
    int n, c, h, w, ho, wo;
};

struct Taps { int x0, y0; float wx1, wy1; bool in_x0, in_x1, in_y0, in_y1; };   // wx1 = x - x0 (weight of the right neighbour)

__device__ __forceinline__ Taps grid_taps(float gx, float gy, int w, int h, float& x, float& y) {
    // grid_sampler_unnormalize with align_corners = False: ((coord + 1) * size - 1) / 2
    x = ((gx + 1.f) * w - 1.f) * 0.5f;
    y = ((gy + 1.f) * h - 1.f) * 0.5f;
    Taps t;
    const float xf = floorf(x), yf = floorf(y);
    t.x0 = (int)xf; t.y0 = (int)yf;
    t.wx1 = x - xf; t.wy1 = y - yf;
    t.in_x0 = t.x0 >= 0 && t.x0 < w; t.in_x1 = t.x0 + 1 >= 0 && t.x0 + 1 < w;
    t.in_y0 = t.y0 >= 0 && t.y0 < h; t.in_y1 = t.y0 + 1 >= 0 && t.y0 + 1 < h;
    return t;
}

__global__ void __launch_bounds__(256) grid_sample_fwd_kernel(GridArgs p) {
    const long long total = (long long)p.n * p.ho * p.wo;
    const long long plane = (long long)p.h * p.w, oplane = (long long)p.ho * p.wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / oplane, pix = i - n * oplane;
        const float2 g = *reinterpret_cast<const float2*>(p.grid + 2 * i);
        float x, y;
        const Taps t = grid_taps(g.x, g.y, p.w, p.h, x, y);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float nw = wx0 * wy0, ne = t.wx1 * wy0, sw = wx0 * t.wy1, se = t.wx1 * t.wy1;
        const bool finite = isfinite(x) && isfinite(y);            // NaN / Inf coordinates sample nothing (all taps out of bounds)
        const bool b_nw = finite && t.in_x0 && t.in_y0, b_ne = finite && t.in_x1 && t.in_y0;
        const bool b_sw = finite && t.in_x0 && t.in_y1, b_se = finite && t.in_x1 && t.in_y1;
        const long long o_nw = (long long)t.y0 * p.w + t.x0;
        const float* src = p.input + n * p.c * plane;
        float* dst = p.out + n * p.c * oplane + pix;
        for (int c = 0; c < p.c; c++, src += plane, dst += oplane) {
            float acc = 0.f;
            if (b_nw) acc += __ldg(src + o_nw) * nw;
            if (b_ne) acc += __ldg(src + o_nw + 1) * ne;
            if (b_sw) acc += __ldg(src + o_nw + p.w) * sw;
            if (b_se) acc += __ldg(src + o_nw + p.w + 1) * se;
            *dst = acc;
        }
    }
}

__global__ void __launch_bounds__(256) grid_sample_bwd_kernel(GridArgs p) {
    const long long total = (long long)p.n * p.ho * p.wo;
    const long long plane = (long long)p.h * p.w, oplane = (long long)p.ho * p.wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / oplane, pix = i - n * oplane;
        const float2 g = *reinterpret_cast<const float2*>(p.grid + 2 * i);
        float x, y;
        const Taps t = grid_taps(g.x, g.y, p.w, p.h, x, y);
        const float wx0 = 1.f - t.wx1, wy0 = 1.f - t.wy1;
        const float nw = wx0 * wy0, ne = t.wx1 * wy0, sw = wx0 * t.wy1, se = t.wx1 * t.wy1;
        const bool finite = isfinite(x) && isfinite(y);
        const bool b_nw = finite && t.in_x0 && t.in_y0, b_ne = finite && t.in_x1 && t.in_y0;
        const bool b_sw = finite && t.in_x0 && t.in_y1, b_se = finite && t.in_x1 && t.in_y1;
        const long long o_nw = (long long)t.y0 * p.w + t.x0;
        const float* src = p.input + n * p.c * plane;
        float* gin = p.grad_input ? p.grad_input + n * p.c * plane : nullptr;
        const float* go = p.grad_out + n * p.c * oplane + pix;
        float gix = 0.f, giy = 0.f;
        for (int c = 0; c < p.c; c++, src += plane, go += oplane) {
            const float d = *go;
            if (gin) {
                if (b_nw) atomicAdd(gin + o_nw, nw * d);
                if (b_ne) atomicAdd(gin + o_nw + 1, ne * d);
                if (b_sw) atomicAdd(gin + o_nw + p.w, sw * d);
                if (b_se) atomicAdd(gin + o_nw + p.w + 1, se * d);
                gin += plane;
            }
            if (p.grad_grid) {
                if (b_nw) { const float v = __ldg(src + o_nw); gix -= v * wy0 * d; giy -= v * wx0 * d; }
                if (b_ne) { const float v = __ldg(src + o_nw + 1); gix += v * wy0 * d; giy -= v * t.wx1 * d; }
                if (b_sw) { const float v = __ldg(src + o_nw + p.w); gix -= v * t.wy1 * d; giy += v * wx0 * d; }
                if (b_se) { const float v = __ldg(src + o_nw + p.w + 1); gix += v * t.wy1 * d; giy += v * t.wx1 * d; }
            }
        }
        if (p.grad_grid)        // d unnormalise / d coord = size / 2
            *reinterpret_cast<float2*>(p.grad_grid + 2 * i) = make_float2(gix * (0.5f * p.w), giy * (0.5f * p.h));
    }
}

static int grid_blocks(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = 16ll * sm_count();
    return (int)(blocks > cap ? cap : blocks);
}

} // namespace pgpp

extern "C" int pgpp_grid_sample_2d(const float* input, const float* grid, float* out, int n, int c, int h, int w, int ho, int wo, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(n >= 0 && c >= 0 && h >= 1 && w >= 1 && ho >= 0 && wo >= 0, "bad grid_sample sizes");
    if ((long long)n * c * ho * wo == 0) return PGPP_OK;
    PGPP_REQUIRE(input && grid && out, "input, grid and out must be device pointers");
    PGPP_REQUIRE(((uintptr_t)grid & 7) == 0, "grid must be 8-byte aligned");
    GridArgs p{input, grid, nullptr, out, nullptr, nullptr, n, c, h, w, ho, wo};
    grid_sample_fwd_kernel<<<grid_blocks((long long)n * ho * wo), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}

extern "C" int pgpp_grid_sample_2d_backward(const float* grad_out, const float* input, const float* grid, float* grad_input, float* grad_grid,
                                            int n, int c, int h, int w, int ho, int wo, void* stream) {
    using namespace pgpp;
    PGPP_REQUIRE(n >= 0 && c >= 0 && h >= 1 && w >= 1 && ho >= 0 && wo >= 0, "bad grid_sample sizes");
    PGPP_REQUIRE(grad_input || grad_grid, "at least one of grad_input / grad_grid must be requested");
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_input && (long long)n * c * h * w > 0) PGPP_CUDA_OK(cudaMemsetAsync(grad_input, 0, sizeof(float) * (size_t)n * c * h * w, st));
    if ((long long)n * ho * wo == 0) return PGPP_OK;
    PGPP_REQUIRE(grad_out && input && grid, "grad_out, input and grid must be device pointers");
    PGPP_REQUIRE(((uintptr_t)grid & 7) == 0 && ((uintptr_t)grad_grid & 7) == 0, "grid and grad_grid must be 8-byte aligned");
    if (c == 0) {
        if (grad_grid) PGPP_CUDA_OK(cudaMemsetAsync(grad_grid, 0, sizeof(float) * 2 * (size_t)n * ho * wo, st));
        return PGPP_OK;
    }
    GridArgs p{input, grid, grad_out, nullptr, grad_input, grad_grid, n, c, h, w, ho, wo};
    grid_sample_bwd_kernel<<<grid_blocks((long long)n * ho * wo), 256, 0, st>>>(p);
    count_launch();
    PGPP_CUDA_OK(cudaGetLastError());
    return PGPP_OK;
}
