// Depth monitoring metrics: Trainer.compute_depth_losses (trainer.py:624-652) over
// layers.compute_depth_errors (layers.py:251-269).  SURVEY.md section 8 row f3.
//
//   prepare : bilinear resize of depth_pred to the ground-truth resolution (F.interpolate,
//             align_corners=False), clamp to [min_d, max_d], mask = (gt > 0) & Garg/Eigen crop;
//             valid pixels are written to dense arrays, invalid ones as +inf (so that a sort puts
//             them last and torch.median's "lower middle" is element (n_valid-1)/2), count n_valid.
//   reduce  : pred *= median(gt) / median(pred); clamp; the seven error sums in one pass with a
//             fixed-order two-level reduction (deterministic); finish divides by n_valid.
#pragma once
#include "pml_common.cuh"

namespace pml {

struct MetricsPrepParams {
    const float* depth;   // [B,1,H,W]
    const float* gt;      // [B,1,Hg,Wg]
    float* pred_out;      // [B*Hg*Wg]
    float* gt_out;        // [B*Hg*Wg]
    int* count;           // [1], zeroed by the caller
    int B, H, W, Hg, Wg;
    int cy0, cy1, cx0, cx1;   // crop [cy0,cy1) x [cx0,cx1)
    float min_d, max_d;
};

__global__ void __launch_bounds__(256)
depth_metrics_prepare_kernel(const MetricsPrepParams q) {
    __shared__ int s_cnt[8];
    const long long n = (long long)q.B * q.Hg * q.Wg;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    int valid = 0;
    if (i < n) {
        const int x = (int)(i % q.Wg);
        const int y = (int)((i / q.Wg) % q.Hg);
        const int b = (int)(i / ((long long)q.Wg * q.Hg));
        const float g = __ldg(q.gt + i);
        valid = (g > 0.f) && (y >= q.cy0) && (y < q.cy1) && (x >= q.cx0) && (x < q.cx1);
        float pv = __int_as_float(0x7f800000), gv = __int_as_float(0x7f800000);   // +inf
        if (valid) {
            // ATen upsample_bilinear2d, align_corners=False: src = scale * (dst + 0.5) - 0.5, clamped at 0
            const float sh = (float)q.H / (float)q.Hg, sw = (float)q.W / (float)q.Wg;
            const float sy = fmaxf(fmaf(sh, (float)y + 0.5f, -0.5f), 0.f);
            const float sx = fmaxf(fmaf(sw, (float)x + 0.5f, -0.5f), 0.f);
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = min(y0 + 1, q.H - 1), x1 = min(x0 + 1, q.W - 1);
            const float ly = sy - (float)y0, lx = sx - (float)x0;
            const float* d = q.depth + (size_t)b * q.H * q.W;
            const float v00 = __ldg(d + y0 * q.W + x0), v01 = __ldg(d + y0 * q.W + x1);
            const float v10 = __ldg(d + y1 * q.W + x0), v11 = __ldg(d + y1 * q.W + x1);
            const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
            pv = fminf(fmaxf(v, q.min_d), q.max_d);   // trainer.py:631-632
            gv = g;
        }
        q.pred_out[i] = pv;
        q.gt_out[i] = gv;
    }
    int c = valid;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_cnt[w];
        if (t) atomicAdd(q.count, t);
    }
}

constexpr int kMetricsBlocks = 296;   // 2 per SM

// ---- medians (trainer.py:645: torch.median(gt) / torch.median(pred) over the valid pixels) by radix selection:
// torch.median returns the lower middle element, i.e. element k = (n_valid - 1) / 2 of the sorted values; the arrays
// hold non-negative floats (+inf where masked), whose bit patterns order like the values.  Three histogram passes
// over 11 + 11 + 10 key bits narrow the k-th element down exactly; no sort, no host synchronisation.
constexpr int kSelBins = 2048;
constexpr int kSelBlocks = 296;
struct SelectParams {
    const float* a[2];      // gt, pred (dense, +inf where masked)
    long long n;
    const int* count;       // n_valid
    unsigned* hist;         // [2][kSelBins]
    unsigned* state;        // [2][2]: key prefix found so far, rank k inside it
    float* ratio;           // out: median(a[0]) / median(a[1])
    int pass;               // 0, 1, 2
};
__device__ __forceinline__ int sel_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }
__device__ __forceinline__ int sel_bits(int pass) { return pass == 2 ? 10 : 11; }

__global__ void __launch_bounds__(256)
select_init_kernel(const SelectParams q) {
    for (int i = threadIdx.x; i < 2 * kSelBins; i += 256) q.hist[i] = 0u;
    if (threadIdx.x < 2) {
        const int c = *q.count;
        q.state[threadIdx.x * 2 + 0] = 0u;
        q.state[threadIdx.x * 2 + 1] = (unsigned)(c > 0 ? (c - 1) / 2 : 0);
    }
}

// grid = (kSelBlocks, 2)
__global__ void __launch_bounds__(256)
select_hist_kernel(const SelectParams q) {
    __shared__ unsigned s_h[kSelBins];
    const int arr = blockIdx.y;
    for (int i = threadIdx.x; i < kSelBins; i += 256) s_h[i] = 0u;
    __syncthreads();
    const int shift = sel_shift(q.pass), bits = sel_bits(q.pass);
    const unsigned prefix = q.state[arr * 2];
    const float* a = q.a[arr];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < q.n; i += (long long)gridDim.x * 256) {
        const unsigned key = __float_as_uint(__ldg(a + i));
        // pass 0 looks at every element; later passes only at those inside the prefix found so far
        if (q.pass == 0 || (key >> (shift + bits)) == prefix) atomicAdd(&s_h[(key >> shift) & ((1u << bits) - 1u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += 256)
        if (s_h[i]) atomicAdd(q.hist + arr * kSelBins + i, s_h[i]);
}

// grid = 2 (one block per array): locate the bin that holds rank k, extend the prefix, clear the histogram
__global__ void __launch_bounds__(256)
select_scan_kernel(const SelectParams q) {
    __shared__ unsigned s_bin, s_before;
    const int arr = blockIdx.x;
    unsigned* h = q.hist + arr * kSelBins;
    const int nb = 1 << sel_bits(q.pass);
    if (threadIdx.x == 0) {
        const unsigned k = q.state[arr * 2 + 1];
        unsigned cum = 0, bin = (unsigned)(nb - 1), before = 0;
        for (int i = 0; i < nb; ++i) {
            const unsigned c = h[i];
            if (cum + c > k) { bin = (unsigned)i; before = cum; break; }
            cum += c;
            before = cum;
        }
        s_bin = bin; s_before = before;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += 256) h[i] = 0u;
    if (threadIdx.x == 0) {
        q.state[arr * 2 + 0] = (q.state[arr * 2 + 0] << sel_bits(q.pass)) | s_bin;
        q.state[arr * 2 + 1] -= s_before;
    }
}

__global__ void select_ratio_kernel(const SelectParams q) {
    if (threadIdx.x == 0) q.ratio[0] = __uint_as_float(q.state[0]) / __uint_as_float(q.state[2]);   // trainer.py:645
}

// part: [kMetricsBlocks][8]; sums of abs_rel, sq_rel, sq err, sq log err, a1, a2, a3
__global__ void __launch_bounds__(256)
depth_metrics_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ gt, long long n,
                            const float* __restrict__ ratio, float min_d, float max_d, float* __restrict__ part) {
    __shared__ float s_red[8][7];
    const float r = __ldg(ratio);
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float g = __ldg(gt + i);
        if (g < 3.0e38f) {   // finite => valid
            const float p = fminf(fmaxf(__ldg(pred + i) * r, min_d), max_d);   // trainer.py:645-647
            const float th = fmaxf(g / p, p / g);                              // layers.py:254
            const float d = g - p;
            const float dl = logf(g) - logf(p);
            acc[0] += fabsf(d) / g;
            acc[1] += d * d / g;
            acc[2] += d * d;
            acc[3] += dl * dl;
            acc[4] += (th < 1.25f) ? 1.f : 0.f;
            acc[5] += (th < 1.25f * 1.25f) ? 1.f : 0.f;
            acc[6] += (th < 1.25f * 1.25f * 1.25f) ? 1.f : 0.f;
        }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const float v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
        part[blockIdx.x * 8 + threadIdx.x] = t;
    }
}

// out[7] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3  (trainer.py depth_metric_names order)
__global__ void __launch_bounds__(32)
depth_metrics_finish_kernel(const float* __restrict__ part, int nblk, const int* __restrict__ count, float* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= 7) return;
    double t = 0.0;
    for (int b = 0; b < nblk; ++b) t += (double)part[b * 8 + k];
    const double m = t / (double)max(*count, 1);
    out[k] = (k == 2 || k == 3) ? (float)sqrt(m) : (float)m;
}

}  // namespace pml
