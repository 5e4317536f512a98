// Layer-level drop-in kernels: one forward and one backward per layers.py symbol, so the
// reference's own (unfused) trainer code path can run on libpml.so too.  The fused path in
// pml_photometric.cuh is the fast one; these keep the layers.py call signatures alive.
#pragma once
#include "pml_common.cuh"
#include "pml_smooth.cuh"

namespace pml {

// ---------------------------------------------------------------- autograd-node backward
struct ScaleParams {
    int n_pass, B, S;
    float* g[PML_MAX_PASSES];
    int per_image[PML_MAX_PASSES];
    long long off[PML_MAX_PASSES + 1];
    const float* gconst;   // [n_pass][B]
    const float* gT;       // [n_pass][S][B][16]
    const float* up;       // [n_pass] gradient arriving at the per-scale losses, or null
    const float* up_total; // [1] gradient arriving at the total (mean over scales), or null
    float inv_div;         // 1 / divisor of the total
    float* gT_out;         // [S][B][16]
};
__device__ __forceinline__ float scale_upstream(const ScaleParams& q, int k) {
    return (q.up ? __ldg(q.up + k) : 0.f) + (q.up_total ? __ldg(q.up_total) * q.inv_div : 0.f);
}

// grid = (chunks, B, n_pass + 1): planes z < n_pass scale one image of one scale (float4 where the
// image size allows), plane z == n_pass reduces the pose gradients over the scales.
constexpr int kScaleChunk = 256 * 4 * 4;   // elements per block: 256 threads x 4 float4
__global__ void __launch_bounds__(256)
scale_grads_kernel(const ScaleParams q) {
    const int pi = blockIdx.z, b = blockIdx.y;
    if (pi == q.n_pass) {
        const int e = (blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        if (e < q.S * q.B * 16) {
            float acc = 0.f;
            for (int k = 0; k < q.n_pass; ++k)
                acc = fmaf(scale_upstream(q, k), q.gT[(size_t)k * q.S * q.B * 16 + e], acc);
            q.gT_out[e] = acc;
        }
        return;
    }
    const int n = q.per_image[pi];
    const int base = blockIdx.x * kScaleChunk;
    if (base >= n) return;
    const float up = scale_upstream(q, pi), c = __ldg(q.gconst + pi * q.B + b);
    float* g = q.g[pi] + (size_t)b * n;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = base + (k * 256 + threadIdx.x) * 4;
            if (i < n) {
                float4 v = *reinterpret_cast<float4*>(g + i);
                v.x = up * (v.x + c); v.y = up * (v.y + c); v.z = up * (v.z + c); v.w = up * (v.w + c);
                *reinterpret_cast<float4*>(g + i) = v;
            }
        }
    } else {
        for (int i = base + threadIdx.x; i < min(base + kScaleChunk, n); i += 256) g[i] = up * (g[i] + c);
    }
}

// ---------------------------------------------------------------- identity_selection (trainer.py:606-608)
struct SelectionParams {
    const uint8_t* argmin[PML_MAX_PASSES];
    float* out[PML_MAX_PASSES];
    long long n_pix;
    int n_id;
};
// grid = (chunks, n_pass); 16 pixels per thread where the pointers allow (uint4 in, 4 x float4 out)
__global__ void __launch_bounds__(256)
selection_mask_kernel(const SelectionParams q) {
    const uint8_t* a = q.argmin[blockIdx.y];
    float* o = q.out[blockIdx.y];
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 16;
    if (i0 >= q.n_pix) return;
    const unsigned thr = (unsigned)(q.n_id - 1);   // idx > n_id - 1 (trainer.py:607); the caller guarantees n_id >= 1
    if (i0 + 16 <= q.n_pix && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(a + i0);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 r;
            r.x = ((w[k] & 0xffu) > thr) ? 1.f : 0.f;
            r.y = (((w[k] >> 8) & 0xffu) > thr) ? 1.f : 0.f;
            r.z = (((w[k] >> 16) & 0xffu) > thr) ? 1.f : 0.f;
            r.w = ((w[k] >> 24) > thr) ? 1.f : 0.f;
            *reinterpret_cast<float4*>(o + i0 + 4 * k) = r;
        }
    } else {
        for (long long i = i0; i < min(i0 + 16, q.n_pix); ++i) o[i] = ((unsigned)a[i] > thr) ? 1.f : 0.f;
    }
}

// ---------------------------------------------------------------- disp_to_depth (layers.py:16-25)
__global__ void __launch_bounds__(256)
disp_to_depth_fwd_kernel(const float* __restrict__ disp, float* __restrict__ scaled, float* __restrict__ depth,
                         long long n, float lo, float range) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = fmaf(range, disp[i], lo);
    if (scaled) scaled[i] = s;
    if (depth) depth[i] = __fdiv_rn(1.0f, s);
}
__global__ void __launch_bounds__(256)
disp_to_depth_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ g_scaled,
                         const float* __restrict__ g_depth, float* __restrict__ g_disp, long long n, float lo, float range) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = g_scaled ? g_scaled[i] : 0.f;
    if (g_depth) {
        float d = __fdiv_rn(1.0f, fmaf(range, disp[i], lo));
        g = fmaf(-d * d, g_depth[i], g);
    }
    g_disp[i] = range * g;
}

// ---------------------------------------------------------------- BackprojectDepth (layers.py:163-168)
__global__ void __launch_bounds__(256)
backproject_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ invK, float* __restrict__ cam,
                       int H, int W) {
    const int b = blockIdx.y, n = H * W;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* k = invK + (size_t)b * 16;
    const float x = (float)(i % W), y = (float)(i / W), d = depth[(size_t)b * n + i];
    float* o = cam + (size_t)b * 4 * n + i;
    o[0] = d * fmaf(k[0], x, fmaf(k[1], y, k[2]));
    o[n] = d * fmaf(k[4], x, fmaf(k[5], y, k[6]));
    o[2 * (size_t)n] = d * fmaf(k[8], x, fmaf(k[9], y, k[10]));
    o[3 * (size_t)n] = 1.0f;
}
__global__ void __launch_bounds__(256)
backproject_bwd_kernel(const float* __restrict__ g_cam, const float* __restrict__ invK, float* __restrict__ g_depth,
                       int H, int W) {
    const int b = blockIdx.y, n = H * W;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* k = invK + (size_t)b * 16;
    const float x = (float)(i % W), y = (float)(i / W);
    const float* g = g_cam + (size_t)b * 4 * n + i;
    float r0 = fmaf(k[0], x, fmaf(k[1], y, k[2])), r1 = fmaf(k[4], x, fmaf(k[5], y, k[6])),
          r2 = fmaf(k[8], x, fmaf(k[9], y, k[10]));
    g_depth[(size_t)b * n + i] = fmaf(g[0], r0, fmaf(g[n], r1, g[2 * (size_t)n] * r2));
}

// ---------------------------------------------------------------- Project3D (layers.py:182-193)
__device__ __forceinline__ void load_P(const float* K, const float* T, int b, float* sPm) {
    // P = (K @ T)[:3, :] -- 12 threads fill shared memory
    if (threadIdx.x < 12) {
        int i = threadIdx.x >> 2, j = threadIdx.x & 3;
        const float* Kb = K + (size_t)b * 16;
        const float* Tb = T + (size_t)b * 16;
        float a = 0.f;
        for (int k = 0; k < 4; ++k) a = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], a);
        sPm[threadIdx.x] = a;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256)
project_fwd_kernel(const float* __restrict__ pts, const float* __restrict__ K, const float* __restrict__ T,
                   float* __restrict__ grid, int H, int W, float eps) {
    __shared__ float sPm[12];
    const int b = blockIdx.y, n = H * W;
    load_P(K, T, b, sPm);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = pts + (size_t)b * 4 * n + i;
    const float X = q[0], Y = q[n], Z = q[2 * (size_t)n], Wh = q[3 * (size_t)n];
    float c0 = fmaf(sPm[0], X, fmaf(sPm[1], Y, fmaf(sPm[2], Z, sPm[3] * Wh)));
    float c1 = fmaf(sPm[4], X, fmaf(sPm[5], Y, fmaf(sPm[6], Z, sPm[7] * Wh)));
    float c2 = fmaf(sPm[8], X, fmaf(sPm[9], Y, fmaf(sPm[10], Z, sPm[11] * Wh)));
    float z = c2 + eps;
    float u = __fdiv_rn(c0, z), v = __fdiv_rn(c1, z);
    float* o = grid + ((size_t)b * n + i) * 2;
    o[0] = (__fdiv_rn(u, (float)(W - 1)) - 0.5f) * 2.0f;
    o[1] = (__fdiv_rn(v, (float)(H - 1)) - 0.5f) * 2.0f;
}
// g_points + per-block dL/dP partials [B][nblk][12]
__global__ void __launch_bounds__(256)
project_bwd_kernel(const float* __restrict__ pts, const float* __restrict__ K, const float* __restrict__ T,
                   const float* __restrict__ g_grid, float* __restrict__ g_pts, float* __restrict__ part,
                   int H, int W, float eps) {
    __shared__ float sPm[12];
    __shared__ float s_red[8][12];
    const int b = blockIdx.y, n = H * W;
    load_P(K, T, b, sPm);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float gp[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) gp[e] = 0.f;
    if (i < n) {
        const float* q = pts + (size_t)b * 4 * n + i;
        const float Xh[4] = {q[0], q[n], q[2 * (size_t)n], q[3 * (size_t)n]};
        float c[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            c[r] = fmaf(sPm[r * 4], Xh[0], fmaf(sPm[r * 4 + 1], Xh[1], fmaf(sPm[r * 4 + 2], Xh[2], sPm[r * 4 + 3] * Xh[3])));
        const float invz = __fdiv_rn(1.0f, c[2] + eps);
        const float u = c[0] * invz, v = c[1] * invz;
        const float* g = g_grid + ((size_t)b * n + i) * 2;
        const float gu = g[0] * 2.0f / (float)(W - 1), gv = g[1] * 2.0f / (float)(H - 1);
        const float dc[3] = {gu * invz, gv * invz, -(u * gu + v * gv) * invz};
        float* o = g_pts + (size_t)b * 4 * n + i;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            o[(size_t)j * n] = fmaf(sPm[j], dc[0], fmaf(sPm[4 + j], dc[1], sPm[8 + j] * dc[2]));
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 4; ++j) gp[r * 4 + j] = dc[r] * Xh[j];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        float v = warp_sum(gp[e]);
        if (lane == 0) s_red[warp][e] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += s_red[w][threadIdx.x];
        part[((size_t)b * gridDim.x + blockIdx.x) * 12 + threadIdx.x] = t;
    }
}
__global__ void __launch_bounds__(64)
project_bwd_reduce_kernel(const float* __restrict__ part, const float* __restrict__ K, float* __restrict__ g_T, int nblk) {
    __shared__ float gP[12];
    const int b = blockIdx.x;
    if (threadIdx.x < 12) {
        float t = 0.f;
        for (int i = 0; i < nblk; ++i) t += part[((size_t)b * nblk + i) * 12 + threadIdx.x];
        gP[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        const int k = threadIdx.x >> 2, j = threadIdx.x & 3;
        const float* Kb = K + (size_t)b * 16;
        g_T[(size_t)b * 16 + threadIdx.x] = fmaf(Kb[k], gP[j], fmaf(Kb[4 + k], gP[4 + j], Kb[8 + k] * gP[8 + j]));
    }
}

// ---------------------------------------------------------------- SSIM (layers.py:234-248)
__device__ __forceinline__ void ssim_sums(const float* xp, const float* yp, int y, int x, int H, int W,
                                          float& Sx, float& Sy, float& Sxx, float& Syy, float& Sxy) {
    Sx = Sy = Sxx = Syy = Sxy = 0.f;
#pragma unroll
    for (int j = -1; j <= 1; ++j) {
        const int yy = reflect1(y + j, H);
#pragma unroll
        for (int i = -1; i <= 1; ++i) {
            const int xx = reflect1(x + i, W);
            const float a = __ldg(xp + (size_t)yy * W + xx), bq = __ldg(yp + (size_t)yy * W + xx);
            Sx += a; Sy += bq;
            Sxx = fmaf(a, a, Sxx); Syy = fmaf(bq, bq, Syy); Sxy = fmaf(a, bq, Sxy);
        }
    }
}
__global__ void __launch_bounds__(256)
ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int H, int W) {
    const int pl = blockIdx.z, yy = blockIdx.y;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= W) return;
    const size_t off = (size_t)pl * H * W;
    float Sx, Sy, Sxx, Syy, Sxy, d0, d1, d2;
    ssim_sums(x + off, y + off, yy, xx, H, W, Sx, Sy, Sxx, Syy, Sxy);
    out[off + (size_t)yy * W + xx] = ssim_window<false>(Sx, Sy, Sxx, Syy, Sxy, d0, d1, d2);
}
// gather form of the adjoint: every window p that contains q (with the reflection multiplicity)
__global__ void __launch_bounds__(256)
ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g_out,
                float* __restrict__ g_x, float* __restrict__ g_y, int H, int W) {
    const int pl = blockIdx.z, qy = blockIdx.y;
    const int qx = blockIdx.x * blockDim.x + threadIdx.x;
    if (qx >= W) return;
    const size_t off = (size_t)pl * H * W;
    const float* xp = x + off;
    const float* yp = y + off;
    const float xq = xp[(size_t)qy * W + qx], yq = yp[(size_t)qy * W + qx];
    float gx = 0.f, gy = 0.f;
    for (int j = -1; j <= 1; ++j) {
        const int py = qy + j;
        if (py < 0 || py >= H) continue;
        const float wy = 1.f + ((qy == 1 && py == 0) ? 1.f : 0.f) + ((qy == H - 2 && py == H - 1) ? 1.f : 0.f);
        for (int i = -1; i <= 1; ++i) {
            const int px = qx + i;
            if (px < 0 || px >= W) continue;
            const float wx = 1.f + ((qx == 1 && px == 0) ? 1.f : 0.f) + ((qx == W - 2 && px == W - 1) ? 1.f : 0.f);
            const float go = g_out[off + (size_t)py * W + px] * wx * wy;
            float Sx, Sy, Sxx, Syy, Sxy, a, b2, c;
            ssim_sums(xp, yp, py, px, H, W, Sx, Sy, Sxx, Syy, Sxy);
            if (g_x) {
                ssim_window<true>(Sx, Sy, Sxx, Syy, Sxy, a, b2, c);
                gx = fmaf(go, fmaf(2.f * xq, b2, fmaf(yq, c, a)), gx);
            }
            if (g_y) {   // SSIM is symmetric in its arguments
                ssim_window<true>(Sy, Sx, Syy, Sxx, Sxy, a, b2, c);
                gy = fmaf(go, fmaf(2.f * yq, b2, fmaf(xq, c, a)), gy);
            }
        }
    }
    if (g_x) g_x[off + (size_t)qy * W + qx] = gx;
    if (g_y) g_y[off + (size_t)qy * W + qx] = gy;
}

// ---------------------------------------------------------------- get_smooth_loss (layers.py:202-215)
__device__ __forceinline__ float img_edge(const float* img, size_t plane, int C, int a, int b2) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += fabsf(__ldg(img + c * plane + a) - __ldg(img + c * plane + b2));
    return s / (float)C;
}
__global__ void __launch_bounds__(256)
smooth_layer_fwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, float* __restrict__ part,
                        int C, int H, int W) {
    __shared__ float s_red[8][2];
    const int b = blockIdx.y, n = H * W;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    float ex = 0.f, ey = 0.f;
    if (idx < n) {
        const int y = idx / W, x = idx - y * W;
        const float* d = disp + (size_t)b * n;
        const float* im = img + (size_t)b * C * n;
        if (x + 1 < W) ex = fabsf(d[idx] - d[idx + 1]) * expf(-img_edge(im, n, C, idx, idx + 1));
        if (y + 1 < H) ey = fabsf(d[idx] - d[idx + W]) * expf(-img_edge(im, n, C, idx, idx + W));
    }
    ex = warp_sum(ex); ey = warp_sum(ey);
    if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5][0] = ex; s_red[threadIdx.x >> 5][1] = ey; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_red[i][threadIdx.x];
        part[((size_t)b * gridDim.x + blockIdx.x) * 2 + threadIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256)
smooth_layer_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int n_part, float inv_nx, float inv_ny) {
    __shared__ float s0[256], s1[256];
    float a = 0.f, b2 = 0.f;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) { a += part[i * 2]; b2 += part[i * 2 + 1]; }
    s0[threadIdx.x] = a; s1[threadIdx.x] = b2;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) { s0[threadIdx.x] += s0[threadIdx.x + st]; s1[threadIdx.x] += s1[threadIdx.x + st]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s0[0] * inv_nx + s1[0] * inv_ny;
}
__global__ void __launch_bounds__(256)
smooth_layer_bwd_kernel(const float* __restrict__ disp, const float* __restrict__ img, const float* __restrict__ g_out,
                        float* __restrict__ g_disp, float* __restrict__ g_img, int C, int H, int W, float inv_nx, float inv_ny) {
    const int b = blockIdx.y, n = H * W;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int y = idx / W, x = idx - y * W;
    const float* d = disp + (size_t)b * n;
    const float* im = img + (size_t)b * C * n;
    const float go = g_out[0];
    const float gx = go * inv_nx, gy = go * inv_ny;
    // the four edges incident to this pixel: (other index, normaliser, this pixel is first?)
    const int other[4] = {idx + 1, idx - 1, idx + W, idx - W};
    const bool have[4] = {x + 1 < W, x > 0, y + 1 < H, y > 0};
    const float nrm[4] = {gx, gx, gy, gy};
    float gd = 0.f;
    for (int e = 0; e < 4; ++e) {
        if (!have[e]) continue;
        const int o = other[e];
        const float wgt = expf(-img_edge(im, n, C, idx, o));
        const float df = d[idx] - d[o];
        gd += sgnf(df) * wgt * nrm[e];
    }
    if (g_disp) g_disp[(size_t)b * n + idx] = gd;
    if (g_img) {
        for (int c = 0; c < C; ++c) {
            float gi = 0.f;
            for (int e = 0; e < 4; ++e) {
                if (!have[e]) continue;
                const int o = other[e];
                const float wgt = expf(-img_edge(im, n, C, idx, o));
                const float ad = fabsf(d[idx] - d[o]);
                const float di = im[(size_t)c * n + idx] - im[(size_t)c * n + o];
                gi -= ad * wgt * sgnf(di) * nrm[e] / (float)C;
            }
            g_img[((size_t)b * C + c) * n + idx] = gi;
        }
    }
}

// ---------------------------------------------------------------- transformation_from_parameters
// (layers.py:28-103).  The backward runs the same closed form on dual numbers (forward-mode AD,
// six seeds) -- twelve-output Jacobian of a six-input function, one thread per pose.
struct Dual {
    float v, d;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, fmaf(a.v, b.d, a.d * b.v)}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    float q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ Dual dconst(float c) { return {c, 0.f}; }
__device__ __forceinline__ Dual dsqrt(Dual a) {
    float s = sqrtf(a.v);
    return {s, a.v > 0.f ? a.d / (2.f * s) : 0.f};   // torch.norm backward is 0 at the origin
}
__device__ __forceinline__ Dual dsin(Dual a) { return {sinf(a.v), cosf(a.v) * a.d}; }
__device__ __forceinline__ Dual dcos(Dual a) { return {cosf(a.v), -sinf(a.v) * a.d}; }
__device__ __forceinline__ float dsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float dsin(float a) { return sinf(a); }
__device__ __forceinline__ float dcos(float a) { return cosf(a); }
__device__ __forceinline__ float dconst_like(float, float c) { return c; }
__device__ __forceinline__ Dual dconst_like(Dual, float c) { return {c, 0.f}; }

template <class R>
__device__ __forceinline__ void pose_matrix(const R aa[3], const R tr[3], int invert, R M[16]) {
    const R one = dconst_like(aa[0], 1.0f), zero = dconst_like(aa[0], 0.0f), eps = dconst_like(aa[0], 1e-7f);
    R angle = dsqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
    R den = angle + eps;
    R x = aa[0] / den, y = aa[1] / den, z = aa[2] / den;
    R ca = dcos(angle), sa = dsin(angle), C = one - ca;
    R xs = x * sa, ys = y * sa, zs = z * sa, xC = x * C, yC = y * C, zC = z * C;
    R xyC = x * yC, yzC = y * zC, zxC = z * xC;
    R Rm[9] = {x * xC + ca, xyC - zs, zxC + ys,
               xyC + zs, y * yC + ca, yzC - xs,
               zxC - ys, yzC + xs, z * zC + ca};
    R t[3] = {tr[0], tr[1], tr[2]};
    if (invert) {
        R tmp;
        tmp = Rm[1]; Rm[1] = Rm[3]; Rm[3] = tmp;
        tmp = Rm[2]; Rm[2] = Rm[6]; Rm[6] = tmp;
        tmp = Rm[5]; Rm[5] = Rm[7]; Rm[7] = tmp;
        for (int i = 0; i < 3; ++i) t[i] = zero - t[i];
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) M[i * 4 + j] = Rm[i * 3 + j];
        // invert: M = R @ T  -> last column R t ;  else: M = T @ R -> last column t
        M[i * 4 + 3] = invert ? (Rm[i * 3] * t[0] + Rm[i * 3 + 1] * t[1] + Rm[i * 3 + 2] * t[2]) : t[i];
    }
    M[12] = zero; M[13] = zero; M[14] = zero; M[15] = one;
}
__global__ void __launch_bounds__(128)
pose_fwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, float* __restrict__ T, int B, int invert) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float a[3] = {aa[b * 3], aa[b * 3 + 1], aa[b * 3 + 2]}, t[3] = {tr[b * 3], tr[b * 3 + 1], tr[b * 3 + 2]}, M[16];
    pose_matrix<float>(a, t, invert, M);
    for (int e = 0; e < 16; ++e) T[(size_t)b * 16 + e] = M[e];
}
__global__ void __launch_bounds__(128)
pose_bwd_kernel(const float* __restrict__ aa, const float* __restrict__ tr, const float* __restrict__ gT,
                float* __restrict__ g_aa, float* __restrict__ g_tr, int B, int invert) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int k = 0; k < 6; ++k) {
        Dual a[3], t[3], M[16];
        for (int i = 0; i < 3; ++i) {
            a[i] = {aa[b * 3 + i], (k == i) ? 1.f : 0.f};
            t[i] = {tr[b * 3 + i], (k == 3 + i) ? 1.f : 0.f};
        }
        pose_matrix<Dual>(a, t, invert, M);
        float g = 0.f;
        for (int e = 0; e < 12; ++e) g = fmaf(gT[(size_t)b * 16 + e], M[e].d, g);
        if (k < 3) g_aa[b * 3 + k] = g;
        else g_tr[b * 3 + k - 3] = g;
    }
}

}  // namespace pml
