// Edge-aware smoothness term of the loss: trainer.py:612-616 over layers.get_smooth_loss
// (layers.py:202-215), forward and adjoint in one sweep, all scales in one launch.
#pragma once
#include "pml_common.cuh"

namespace pml {

struct SmoothPass {
    const float* disp;    // [B,1,h,w]
    const float* color;   // [B,3,h,w]
    float* grad_disp;     // [B,1,h,w] or null
    int h, w, blocks;     // blocks per image (256 pixels each)
    int block_off;        // first global block of this pass
    float weight;         // disparity_smoothness / 2^s
};
struct SmoothParams {
    int B, n_pass;
    SmoothPass pass[PML_MAX_PASSES];
    float* disp_mean;     // [n_pass][B]
    float* mean_part;     // [n_pass][B][max_chunks] partial sums of disp (4096 px each)
    int max_chunks;
    float* part;          // [total_blocks][3]
    int n_seg, seg_size;  // chunked batch (pml_segments): colour of pass i, chunk j = color_c[i].p[j]
    ChunkPtrs color_c[PML_MAX_PASSES];
};

// per-image mean of the disparity (trainer.py:612), two stages: 4096-pixel partial sums here,
// combined (in fixed order) by every consumer through image_mean().  grid = (chunks, B, n_pass).
constexpr int kMeanChunk = 4096;
__global__ void __launch_bounds__(256)
disp_sum_kernel(const SmoothParams q) {
    __shared__ float s_w[8];
    const int chunk = blockIdx.x, b = blockIdx.y, pi = blockIdx.z;
    const SmoothPass& ps = q.pass[pi];
    const int n = ps.h * ps.w;
    const int lo = chunk * kMeanChunk;
    if (lo >= n) return;
    const int hi = min(lo + kMeanChunk, n);
    const float* d = ps.disp + (size_t)b * n;
    float v = 0.f;
    for (int i = lo + threadIdx.x; i < hi; i += 256) v += __ldg(d + i);
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        q.mean_part[((size_t)pi * q.B + b) * q.max_chunks + chunk] = t;
    }
}
__device__ __forceinline__ float image_mean(const SmoothParams& q, int pi, int b) {
    const SmoothPass& ps = q.pass[pi];
    const int n = ps.h * ps.w, nc = (n + kMeanChunk - 1) / kMeanChunk;
    const float* mp = q.mean_part + ((size_t)pi * q.B + b) * q.max_chunks;
    float t = 0.f;
    for (int i = 0; i < nc; ++i) t += mp[i];
    return t / (float)n;
}

__device__ __forceinline__ float edge_weight(const float* img, size_t plane, int a, int bidx) {
    float s = fabsf(__ldg(img + a) - __ldg(img + bidx)) + fabsf(__ldg(img + plane + a) - __ldg(img + plane + bidx)) +
              fabsf(__ldg(img + 2 * plane + a) - __ldg(img + 2 * plane + bidx));
    return expf(-s * (1.0f / 3.0f));   // layers.py:209-213
}
__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

template <bool GRAD>
__global__ void __launch_bounds__(256)
smooth_kernel(const SmoothParams q) {
    __shared__ float s_red[8][3];
    __shared__ float s_mean;
    // locate (pass, image, block)
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < q.n_pass; ++i) if ((int)blockIdx.x >= q.pass[i].block_off) pi = i;
    const SmoothPass& ps = q.pass[pi];
    int rel = blockIdx.x - ps.block_off;
    const int b = rel / ps.blocks, blk = rel - b * ps.blocks;
    const int h = ps.h, w = ps.w, n = h * w;
    const int idx = blk * 256 + threadIdx.x;
    if (threadIdx.x == 0) {
        s_mean = image_mean(q, pi, b);
        if (blk == 0) q.disp_mean[pi * q.B + b] = s_mean;
    }
    __syncthreads();
    float ex = 0.f, ey = 0.f, gd = 0.f;
    if (idx < n) {
        const int y = idx / w, x = idx - y * w;
        const float* d = ps.disp + (size_t)b * n;
        int bl;
        const float* imgc = chunk_of(ps.color, q.color_c[pi], q.n_seg, q.seg_size, b, bl);
        const float* img = imgc + (size_t)bl * 3 * n;
        const float inv = __fdiv_rn(1.0f, s_mean + 1e-7f);
        const float nx_ = 1.0f / ((float)q.B * (float)h * (float)(w - 1));
        const float ny_ = 1.0f / ((float)q.B * (float)(h - 1) * (float)w);
        const float dc = __ldg(d + idx);
        const float nc = dc * inv;
        float g = 0.f;
        if (x + 1 < w) {
            float wgt = edge_weight(img, n, idx, idx + 1);
            float df = nc - __ldg(d + idx + 1) * inv;
            ex = fabsf(df) * wgt;
            g += sgnf(df) * wgt * nx_;
        }
        if (y + 1 < h) {
            float wgt = edge_weight(img, n, idx, idx + w);
            float df = nc - __ldg(d + idx + w) * inv;
            ey = fabsf(df) * wgt;
            g += sgnf(df) * wgt * ny_;
        }
        if (GRAD) {
            if (x > 0) {
                float wgt = edge_weight(img, n, idx - 1, idx);
                float df = __ldg(d + idx - 1) * inv - nc;
                g -= sgnf(df) * wgt * nx_;
            }
            if (y > 0) {
                float wgt = edge_weight(img, n, idx - w, idx);
                float df = __ldg(d + idx - w) * inv - nc;
                g -= sgnf(df) * wgt * ny_;
            }
            gd = g * dc;
            if (ps.grad_disp != nullptr) ps.grad_disp[(size_t)b * n + idx] = ps.weight * g * inv;
        }
    }
    ex = warp_sum(ex); ey = warp_sum(ey); gd = warp_sum(gd);
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_red[warp][0] = ex; s_red[warp][1] = ey; s_red[warp][2] = gd; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_red[i][threadIdx.x];
        q.part[(size_t)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

}  // namespace pml
