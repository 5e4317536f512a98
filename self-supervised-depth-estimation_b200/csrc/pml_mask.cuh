// Predictive-mask ablation (trainer.py:571-583, options.py:190-193):
//   mask = F.interpolate(outputs["predictive_mask"]["disp", s], [H, W], "bilinear", align_corners=False)
//   reprojection_losses *= mask                       -> handled inside the fused sweep (PassDev::fw)
//   loss += 0.2 * nn.BCELoss()(mask, ones)            -> bce_ones kernels below
// The bilinear resize is also the layer-level drop-in of the reference's F.interpolate call on the
// disparities (trainer.py:474-475, row a1).
#pragma once
#include "pml_common.cuh"

namespace pml {

// ATen upsample_bilinear2d, align_corners=False: src = scale * (dst + 0.5) - 0.5 clamped at 0,
// i1 = min(i0 + 1, n - 1).  x: [planes,h,w] -> out: [planes,H,W].
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int h, int w, int H, int W) {
    const int X = blockIdx.x * 256 + threadIdx.x, Y = blockIdx.y, pl = blockIdx.z;
    if (X >= W) return;
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;
    const float sy = fmaxf(fmaf(sh, (float)Y + 0.5f, -0.5f), 0.f);
    const float sx = fmaxf(fmaf(sw, (float)X + 0.5f, -0.5f), 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float* p = x + (size_t)pl * h * w;
    const float v00 = __ldg(p + y0 * w + x0), v01 = __ldg(p + y0 * w + x1);
    const float v10 = __ldg(p + y1 * w + x0), v11 = __ldg(p + y1 * w + x1);
    out[((size_t)pl * H + Y) * W + X] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// Transposed resize as a gather (deterministic, no atomics): low-res cell (i, j) collects every
// full-res pixel whose two-tap footprint contains it.  One thread per low-res cell.
__device__ __forceinline__ float upsample_tap_weight(int dst, int cell, float scale, int n_src) {
    const float s = fmaxf(fmaf(scale, (float)dst + 0.5f, -0.5f), 0.f);
    const int i0 = (int)s, i1 = min(i0 + 1, n_src - 1);
    const float l = s - (float)i0;
    return (i0 == cell ? 1.f - l : 0.f) + (i1 == cell ? l : 0.f);
}
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float* __restrict__ g_out, float* __restrict__ g_x, int h, int w, int H, int W) {
    const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y, pl = blockIdx.z;
    if (j >= w) return;
    const float sh = (float)h / (float)H, sw = (float)w / (float)W;
    // full-res rows / columns whose source coordinate lies in (i - 1, i + 1)
    const float ky = (float)H / (float)h, kx = (float)W / (float)w;
    const int ya = max((int)floorf(((float)i - 0.5f) * ky - 0.5f) - 1, 0), yb = min((int)ceilf(((float)i + 1.5f) * ky - 0.5f) + 1, H - 1);
    const int xa = max((int)floorf(((float)j - 0.5f) * kx - 0.5f) - 1, 0), xb = min((int)ceilf(((float)j + 1.5f) * kx - 0.5f) + 1, W - 1);
    const float* g = g_out + (size_t)pl * H * W;
    float acc = 0.f;
    for (int y = ya; y <= yb; ++y) {
        const float wy = upsample_tap_weight(y, i, sh, h);
        if (wy == 0.f) continue;
        float row = 0.f;
        for (int x = xa; x <= xb; ++x) {
            const float wx = upsample_tap_weight(x, j, sw, w);
            if (wx != 0.f) row = fmaf(wx, __ldg(g + y * W + x), row);
        }
        acc = fmaf(wy, row, acc);
    }
    g_x[((size_t)pl * h + i) * w + j] = acc;
}

// nn.BCELoss()(mask, ones) = mean(-max(log(mask), -100))   (ATen binary_cross_entropy clamps the log)
constexpr int kBceBlocks = 296;   // 2 per SM
__global__ void __launch_bounds__(256)
bce_ones_partial_kernel(const float* __restrict__ mask, long long n, float* __restrict__ part) {
    __shared__ float s_red[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        acc -= fmaxf(logf(__ldg(mask + i)), -100.f);
    const float v = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += s_red[k];
        part[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(32)
bce_ones_finish_kernel(const float* __restrict__ part, int nblk, long long n, float* __restrict__ out) {
    if (threadIdx.x != 0) return;
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += (double)part[b];
    out[0] = (float)(t / (double)n);
}
// ATen binary_cross_entropy_backward: g * (x - y) / max((1 - x) * x, 1e-12) / n with y = 1
__global__ void __launch_bounds__(256)
bce_ones_bwd_kernel(const float* __restrict__ mask, const float* __restrict__ g_out, float* __restrict__ g_mask, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float x = __ldg(mask + i);
    const float g = __ldg(g_out) / (float)n;
    g_mask[i] = g * (x - 1.f) / fmaxf((1.f - x) * x, 1e-12f);
}

}  // namespace pml
