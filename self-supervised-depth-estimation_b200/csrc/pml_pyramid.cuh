// Input colour pyramid on the device.
//
// datasets/mono_dataset.py:84-111 builds, per frame and in DataLoader workers on the CPU,
//     ("color", f, i) = transforms.Resize((H >> i, W >> i), Image.ANTIALIAS)(("color", f, i - 1))   i = 1..3
// on PIL uint8 images, then ToTensor (uint8 HWC -> float32 CHW / 255), and the trainer uploads all
// scales of all frames as fp32 (trainer.py:233-237): 4 x the bytes of the uint8 scale-0 frames plus
// the pyramid.  Here the host uploads the uint8 scale-0 frames once and the pyramid is produced by
// these kernels, BIT-EXACT with Pillow + torchvision:
//   * Pillow ImagingResample, 8 bits per channel, Lanczos (support 3), in = 2 * out: every output
//     sample is clip8((2^21 + sum_k u8 * coeff_k) >> 22) over <= 12 taps, horizontal pass first into a
//     uint8 intermediate, then the vertical pass (src/libImaging/Resample.c; restated in
//     oracle/pyramid_oracle.py and pinned there against the real Pillow);
//   * the integer coefficients depend on the output index only near the borders (3 outputs each
//     side, where the tap window is cut and renormalised); the host computes those 7 sets per axis
//     in double precision exactly like precompute_coeffs / normalize_coeffs_8bpc and passes them by
//     value (no allocation, no host-to-device copy, CUDA-graph capturable);
//   * ToTensor: IEEE fp32 division by 255.
// Byte / integer work, bound by HBM: per level the kernel reads the parent level once (uint8) and
// writes the child as uint8 (parent of the next level) and as fp32 planes (what the loss reads).
#pragma once
#include "pml_common.cuh"

namespace pml {

constexpr int kResampleTaps = 13;      // ksize = ceil(3 * 2) * 2 + 1 for a 2x reduction
constexpr int kResampleBits = 22;      // Resample.c PRECISION_BITS = 32 - 8 - 2
constexpr int kPyrTOH = 8, kPyrTOW = 32;                       // output tile of one CTA
constexpr int kPyrIH = 2 * kPyrTOH + 10, kPyrIW = 2 * kPyrTOW + 10;   // parent pixels under it

// Coefficient sets of one axis: set 0..2 = outputs 0..2, set 3 = every interior output
// (xmin = 2 * xx - 5, 12 taps), set 4..6 = outputs out-3 .. out-1.
struct ResampleAxis {
    int out, in;
    int xmin[7];   // first tap (absolute for the border sets, unused for set 3)
    int cnt[7];
    int k[7][kResampleTaps];
};
struct PyramidLevelParams {
    const uint8_t* parent;   // [N, 2h, 2w, 3]
    uint8_t* child;          // [N, h, w, 3] (nullable for the last level)
    float* child_f;          // [N, 3, h, w]
    int N;
    ResampleAxis ax, ay;     // ax.out = w, ay.out = h
};

__device__ __forceinline__ int resample_set(int xx, int out) { return xx < 3 ? xx : (xx > out - 4 ? 4 + xx - (out - 3) : 3); }
__device__ __forceinline__ int clip8(int acc) { return min(max(acc >> kResampleBits, 0), 255); }

__global__ void __launch_bounds__(256)
pyramid_level_kernel(const PyramidLevelParams q) {
    __shared__ uint8_t s_in[kPyrIH][kPyrIW * 3 + 2];
    __shared__ uint8_t s_tmp[kPyrIH][kPyrTOW * 3];
    const int n = blockIdx.z, oy0 = blockIdx.y * kPyrTOH, ox0 = blockIdx.x * kPyrTOW;
    const int h = q.ay.out, w = q.ax.out, ph = q.ay.in, pw = q.ax.in;
    // parent window under this tile (clamped to the image: the border coefficient sets never
    // reach outside, so the clamped rows / columns are simply never read)
    const int iy0 = max(2 * oy0 - 5, 0), ix0 = max(2 * ox0 - 5, 0);
    const int iy1 = min(2 * (oy0 + kPyrTOH - 1) + 7, ph), ix1 = min(2 * (ox0 + kPyrTOW - 1) + 7, pw);
    const int nrow = iy1 - iy0, nbyte = (ix1 - ix0) * 3;
    const uint8_t* src = q.parent + ((size_t)n * ph * pw) * 3;
    for (int i = threadIdx.x; i < nrow * nbyte; i += 256) {
        const int r = i / nbyte, c = i - r * nbyte;
        s_in[r][c] = __ldg(src + ((size_t)(iy0 + r) * pw + ix0) * 3 + c);
    }
    __syncthreads();
    // horizontal pass: (parent rows of the window) x (tile columns) x 3 channels -> uint8
    const int tw = min(kPyrTOW, w - ox0);
    for (int i = threadIdx.x; i < nrow * tw * 3; i += 256) {
        const int r = i / (tw * 3), rem = i - r * (tw * 3), ox = rem / 3, ch = rem - ox * 3;
        const int xx = ox0 + ox, set = resample_set(xx, w);
        const int xmin = (set == 3) ? 2 * xx - 5 : q.ax.xmin[set];
        const int cnt = q.ax.cnt[set];
        int acc = 1 << (kResampleBits - 1);
        for (int t = 0; t < cnt; ++t) acc += (int)s_in[r][(xmin - ix0 + t) * 3 + ch] * q.ax.k[set][t];
        s_tmp[r][ox * 3 + ch] = (uint8_t)clip8(acc);
    }
    __syncthreads();
    // vertical pass + ToTensor
    const int th = min(kPyrTOH, h - oy0);
    for (int i = threadIdx.x; i < th * tw * 3; i += 256) {
        const int ch = i / (th * tw), rem = i - ch * (th * tw), oy = rem / tw, ox = rem - oy * tw;
        const int yy = oy0 + oy, set = resample_set(yy, h);
        const int ymin = (set == 3) ? 2 * yy - 5 : q.ay.xmin[set];
        const int cnt = q.ay.cnt[set];
        int acc = 1 << (kResampleBits - 1);
        for (int t = 0; t < cnt; ++t) acc += (int)s_tmp[ymin - iy0 + t][ox * 3 + ch] * q.ay.k[set][t];
        const int v = clip8(acc);
        const int xx = ox0 + ox;
        if (q.child) q.child[(((size_t)n * h + yy) * w + xx) * 3 + ch] = (uint8_t)v;
        q.child_f[(((size_t)n * 3 + ch) * h + yy) * w + xx] = __fdiv_rn((float)v, 255.0f);
    }
}

// scale 0: ToTensor only.  frames [N,H,W,3] uint8 -> out [N,3,H,W] fp32
__global__ void __launch_bounds__(256)
to_tensor_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int HW, long long total) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;   // one thread per pixel
    if (i >= total) return;
    const long long n = i / HW;
    const int p = (int)(i - n * HW);
    const uint8_t* s = frames + i * 3;
    float* d = out + n * 3 * HW + p;
    d[0] = __fdiv_rn((float)__ldg(s), 255.0f);
    d[HW] = __fdiv_rn((float)__ldg(s + 1), 255.0f);
    d[2 * HW] = __fdiv_rn((float)__ldg(s + 2), 255.0f);
}

}  // namespace pml
