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
constexpr int kPyrTOH = 16, kPyrTOW = 32;                      // output tile of one CTA (256 threads)
constexpr int kPyrIH = 2 * kPyrTOH + 10, kPyrIW = 2 * kPyrTOW + 10;   // parent pixels under it: 42 x 74
constexpr int kPyrRowWords = 64;       // staged parent row: <= 74 * 3 + 3 bytes, word aligned
constexpr int kPyrTmpStride = kPyrTOW * 3 + 4;

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

// One output sample of one pass: `src` points at tap 0 of the interior window (which may lie
// outside the staged data for the cut windows at the image border -- those use their own first tap).
template <int STRIDE>
__device__ __forceinline__ int resample_sample(const uint8_t* src, int xx, int first_interior, const ResampleAxis& a,
                                               const int (&kin)[12]) {
    const int set = resample_set(xx, a.out);
    int acc = 1 << (kResampleBits - 1);
    if (set == 3) {
#pragma unroll
        for (int t = 0; t < 12; ++t) acc += (int)src[t * STRIDE] * kin[t];
    } else {
        const uint8_t* sb = src + (a.xmin[set] - first_interior) * STRIDE;
        const int cnt = a.cnt[set];
        for (int t = 0; t < cnt; ++t) acc += (int)sb[t * STRIDE] * a.k[set][t];
    }
    return clip8(acc);
}

__global__ void __launch_bounds__(256)
pyramid_level_kernel(const PyramidLevelParams q) {
    __shared__ uint32_t s_in[kPyrIH][kPyrRowWords];          // parent window, HWC bytes, rows word aligned
    __shared__ uint8_t s_tmp[kPyrIH][kPyrTmpStride];         // after the horizontal pass
    __shared__ float s_div[256];                             // ToTensor: v / 255, IEEE division
    const int tid = threadIdx.x;
    const int n = blockIdx.z, oy0 = blockIdx.y * kPyrTOH, ox0 = blockIdx.x * kPyrTOW;
    const int h = q.ay.out, w = q.ax.out, ph = q.ay.in, pw = q.ax.in;
    s_div[tid] = __fdiv_rn((float)tid, 255.0f);
    // parent window under this tile, clamped to the image (the cut border windows never reach outside)
    const int iy0 = max(2 * oy0 - 5, 0), ix0 = max(2 * ox0 - 5, 0);
    const int iy1 = min(2 * (oy0 + kPyrTOH - 1) + 7, ph), ix1 = min(2 * (ox0 + kPyrTOW - 1) + 7, pw);
    const int nrow = iy1 - iy0;
    const uint8_t* src = q.parent + ((size_t)n * ph * pw) * 3;
    {   // aligned 32-bit loads: 4 rows per pass, 64 words per row (tensor bases are 16-byte aligned and
        // their sizes multiples of 4, so an aligned word never leaves the tensor)
        const int wi = tid & 63;
        for (int r = tid >> 6; r < nrow; r += 4) {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(src + ((size_t)(iy0 + r) * pw + ix0) * 3);
            const uintptr_t a1 = reinterpret_cast<uintptr_t>(src + ((size_t)(iy0 + r) * pw + ix1) * 3);
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(a0 & ~(uintptr_t)3) + wi;
            if (reinterpret_cast<uintptr_t>(wp) < a1) s_in[r][wi] = __ldg(wp);
        }
    }
    __syncthreads();
    int kin[12];
#pragma unroll
    for (int t = 0; t < 12; ++t) kin[t] = q.ax.k[3][t];      // interior set (identical for both axes)
    const int tw = min(kPyrTOW, w - ox0), th = min(kPyrTOH, h - oy0);
    const int ox = tid & 31;
    // horizontal pass: every staged parent row x tile columns x 3 channels -> uint8
    if (ox < tw) {
        const int xx = ox0 + ox;
        for (int r = tid >> 5; r < nrow; r += 8) {
            const uintptr_t a0 = reinterpret_cast<uintptr_t>(src + ((size_t)(iy0 + r) * pw + ix0) * 3);
            const uint8_t* row = reinterpret_cast<const uint8_t*>(&s_in[r][0]) + (a0 & 3);
            const uint8_t* p0 = row + (2 * xx - 5 - ix0) * 3;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                s_tmp[r][ox * 3 + ch] = (uint8_t)resample_sample<3>(p0 + ch, xx, 2 * xx - 5, q.ax, kin);
        }
    }
    __syncthreads();
    // vertical pass + ToTensor
    if (ox < tw) {
        const int xx = ox0 + ox;
        for (int oy = tid >> 5; oy < th; oy += 8) {
            const int yy = oy0 + oy;
            const uint8_t* p0 = &s_tmp[0][ox * 3] + (2 * yy - 5 - iy0) * kPyrTmpStride;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int v = resample_sample<kPyrTmpStride>(p0 + ch, yy, 2 * yy - 5, q.ay, kin);
                if (q.child) q.child[(((size_t)n * h + yy) * w + xx) * 3 + ch] = (uint8_t)v;
                q.child_f[(((size_t)n * 3 + ch) * h + yy) * w + xx] = s_div[v];
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// Packed version of the level kernel (used when the parent width is a multiple of 4, i.e. every
// level of the reference's sizes): the parent window is de-interleaved into one byte plane per
// channel while it is staged, so the 12 taps of an output are 12 CONSECUTIVE bytes = 3 words, and a
// pass is 9 DP4A: the 22-bit coefficients are split into three 8-bit limbs (k = k0 + 2^8 k1 + 2^16 k2,
// k0, k1 unsigned, k2 signed) -- exact integer arithmetic, same result as the scalar sum.  The cut
// windows at the image border use the same code with zero-padded 12-tap coefficient vectors (the
// host shifts each border set to the interior window origin), positions outside the image hold 0.
// -------------------------------------------------------------------------------------------------
struct ResampleLimbs { uint32_t w[7][3][3]; };   // [set][limb][word]: 12 taps, 4 per word
struct PyramidPackedParams {
    const uint8_t* parent;   // [N, 2h, 2w, 3]
    uint8_t* child;          // [N, h, w, 3] (nullable for the last level)
    float* child_f;          // [N, 3, h, w]
    float* parent_f;         // [N, 3, 2h, 2w] or null: ToTensor of the parent written from the staged window (level 1:
                             // saves the separate pass over the scale-0 frames); needs 16-byte aligned planes
    int N, h, w;
    ResampleLimbs lx, ly;
};

constexpr int kPkIW = 80;                    // staged parent columns: 2 * 32 + 16 (origin 2 * ox0 - 8)
constexpr int kPkIH = 2 * kPyrTOH + 16;      // staged parent rows: origin 2 * oy0 - 8
constexpr int kPkTmpStride = 52;             // bytes per (channel, column) in the transposed intermediate: 13 words

__device__ __forceinline__ uint32_t pk_dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
#ifdef PML_HOST_EMU
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
    return c;
#else
    return __dp4a(a, b, c);
#endif
}
__device__ __forceinline__ int pk_dp4a_us(uint32_t a, uint32_t b, int c) {   // unsigned bytes of a x signed bytes of b
#ifdef PML_HOST_EMU
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 255u) * (int)(int8_t)((b >> (8 * i)) & 255u);
    return c;
#else
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}
__device__ __forceinline__ uint32_t pk_shift(uint32_t lo, uint32_t hi, int bytes) {   // bytes in 1..3
#ifdef PML_HOST_EMU
    return (lo >> (8 * bytes)) | (hi << (32 - 8 * bytes));
#else
    return __funnelshift_r(lo, hi, 8 * bytes);
#endif
}
// 12 taps starting at byte `off` (off & 3 in {1, 3}) of a word-aligned byte row
__device__ __forceinline__ int pk_sample(const uint32_t* row_words, int off, const uint32_t (&k)[3][3]) {
    const uint32_t* p = row_words + (off >> 2);
    const int sh = off & 3;
    const uint32_t a = p[0], b = p[1], c = p[2], d = p[3];
    const uint32_t t0 = pk_shift(a, b, sh), t1 = pk_shift(b, c, sh), t2 = pk_shift(c, d, sh);
    uint32_t s0 = pk_dp4a_uu(t0, k[0][0], 0u); s0 = pk_dp4a_uu(t1, k[0][1], s0); s0 = pk_dp4a_uu(t2, k[0][2], s0);
    uint32_t s1 = pk_dp4a_uu(t0, k[1][0], 0u); s1 = pk_dp4a_uu(t1, k[1][1], s1); s1 = pk_dp4a_uu(t2, k[1][2], s1);
    int s2 = pk_dp4a_us(t0, k[2][0], 0); s2 = pk_dp4a_us(t1, k[2][1], s2); s2 = pk_dp4a_us(t2, k[2][2], s2);
    const int acc = (1 << (kResampleBits - 1)) + (int)s0 + (int)(s1 << 8) + (int)((uint32_t)s2 << 16);
    return clip8(acc);
}

__global__ void __launch_bounds__(256)
pyramid_level_packed_kernel(const PyramidPackedParams q) {
    __shared__ uint32_t s_pl[3][kPkIH][kPkIW / 4];                         // de-interleaved parent window
    __shared__ uint32_t s_tmp[3][kPyrTOW][kPkTmpStride / 4];                // horizontal pass, transposed: [ch][ox][row]
    __shared__ float s_div[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z, oy0 = blockIdx.y * kPyrTOH, ox0 = blockIdx.x * kPyrTOW;
    const int h = q.h, w = q.w, ph = 2 * q.h, pw = 2 * q.w;
    s_div[tid] = __fdiv_rn((float)tid, 255.0f);
    const int wy0 = 2 * oy0 - 8, wx0 = 2 * ox0 - 8;                        // window origin (may be negative)
    const uint8_t* src = q.parent + ((size_t)n * ph * pw) * 3;
    // ---- stage + de-interleave: one task = 4 pixels = 3 aligned words in, one word per channel out
    for (int t = tid; t < kPkIH * (kPkIW / 4); t += 256) {
        const int r = t / (kPkIW / 4), g = t - r * (kPkIW / 4);
        const int y = wy0 + r, x = wx0 + 4 * g;
        uint32_t cr = 0u, cg = 0u, cb = 0u;
        if (y >= 0 && y < ph && x >= 0 && x < pw) {                        // pw % 4 == 0: a group is inside or outside as a whole
            const uint32_t* pwd = reinterpret_cast<const uint32_t*>(src + ((size_t)y * pw + x) * 3);
            const uint32_t a = __ldg(pwd), b = __ldg(pwd + 1), c = __ldg(pwd + 2);
            // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
            cr = (a & 255u) | ((a >> 24) << 8) | (((b >> 16) & 255u) << 16) | (((c >> 8) & 255u) << 24);
            cg = ((a >> 8) & 255u) | ((b & 255u) << 8) | ((b >> 24) << 16) | (((c >> 16) & 255u) << 24);
            cb = ((a >> 16) & 255u) | (((b >> 8) & 255u) << 8) | ((c & 255u) << 16) | ((c >> 24) << 24);
            // the parent pixels under this tile's outputs (every parent pixel belongs to exactly one tile)
            if (q.parent_f != nullptr && r >= 8 && r < 8 + 2 * kPyrTOH && g >= 2 && g < 2 + kPyrTOW / 2) {
                float* d = q.parent_f + ((size_t)n * 3 * ph + y) * pw + x;
                const float k = 255.0f;
                const size_t pl = (size_t)ph * pw;
                *reinterpret_cast<float4*>(d) = make_float4(__fdiv_rn((float)(cr & 255u), k), __fdiv_rn((float)((cr >> 8) & 255u), k),
                                                            __fdiv_rn((float)((cr >> 16) & 255u), k), __fdiv_rn((float)(cr >> 24), k));
                *reinterpret_cast<float4*>(d + pl) = make_float4(__fdiv_rn((float)(cg & 255u), k), __fdiv_rn((float)((cg >> 8) & 255u), k),
                                                                 __fdiv_rn((float)((cg >> 16) & 255u), k), __fdiv_rn((float)(cg >> 24), k));
                *reinterpret_cast<float4*>(d + 2 * pl) = make_float4(__fdiv_rn((float)(cb & 255u), k), __fdiv_rn((float)((cb >> 8) & 255u), k),
                                                                     __fdiv_rn((float)((cb >> 16) & 255u), k), __fdiv_rn((float)(cb >> 24), k));
            }
        }
        s_pl[0][r][g] = cr; s_pl[1][r][g] = cg; s_pl[2][r][g] = cb;
    }
    __syncthreads();
    const int xx = ox0 + lane;
    // ---- horizontal pass: lane = output column, warps stride over (row, channel)
    {
        const int setx = resample_set(min(xx, w - 1), w);
        uint32_t kx[3][3];
#pragma unroll
        for (int l = 0; l < 3; ++l)
#pragma unroll
            for (int j = 0; j < 3; ++j) kx[l][j] = q.lx.w[setx][l][j];
        uint8_t* tmp_b = reinterpret_cast<uint8_t*>(&s_tmp[0][0][0]);
        for (int t = warp; t < kPkIH * 3; t += 8) {
            const int r = t / 3, ch = t - r * 3;
            const int y = wy0 + r;
            int v = 0;
            if (xx < w && y >= 0 && y < ph) v = pk_sample(&s_pl[ch][r][0], 2 * lane + 3, kx);   // taps 2xx-5.. = window byte 2*lane+3
            tmp_b[(ch * kPyrTOW + lane) * kPkTmpStride + r] = (uint8_t)v;
        }
    }
    __syncthreads();
    // ---- vertical pass + ToTensor: lane = output column, warps stride over (output row, channel)
    if (xx < w) {
        for (int t = warp; t < kPyrTOH * 3; t += 8) {
            const int oy = t / 3, ch = t - oy * 3;
            const int yy = oy0 + oy;
            if (yy >= h) continue;
            const int sety = resample_set(yy, h);
            uint32_t ky[3][3];
#pragma unroll
            for (int l = 0; l < 3; ++l)
#pragma unroll
                for (int j = 0; j < 3; ++j) ky[l][j] = q.ly.w[sety][l][j];
            const int v = pk_sample(&s_tmp[ch][lane][0], 2 * oy + 3, ky);
            if (q.child) q.child[(((size_t)n * h + yy) * w + xx) * 3 + ch] = (uint8_t)v;
            q.child_f[(((size_t)n * 3 + ch) * h + yy) * w + xx] = s_div[v];
        }
    }
}

// scale 0: ToTensor only.  frames [N,H,W,3] uint8 -> out [N,3,H,W] fp32; one thread per 4 pixels
// (three aligned words in, three float4 out) when H*W is a multiple of 4, else one per pixel.
__global__ void __launch_bounds__(256)
to_tensor_kernel(const uint8_t* __restrict__ frames, float* __restrict__ out, int HW, long long total, int vec) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (vec) {
        const long long p4 = i * 4;
        if (p4 >= total) return;
        const long long n = p4 / HW;
        const int p = (int)(p4 - n * HW);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(frames + p4 * 3);
        const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);
        // bytes: r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3
        float* d = out + n * 3 * HW + p;
        const float k = 255.0f;
        *reinterpret_cast<float4*>(d) = make_float4(__fdiv_rn((float)(w0 & 255u), k), __fdiv_rn((float)(w0 >> 24), k),
                                                    __fdiv_rn((float)((w1 >> 16) & 255u), k), __fdiv_rn((float)((w2 >> 8) & 255u), k));
        *reinterpret_cast<float4*>(d + HW) = make_float4(__fdiv_rn((float)((w0 >> 8) & 255u), k), __fdiv_rn((float)(w1 & 255u), k),
                                                         __fdiv_rn((float)(w1 >> 24), k), __fdiv_rn((float)((w2 >> 16) & 255u), k));
        *reinterpret_cast<float4*>(d + 2 * HW) = make_float4(__fdiv_rn((float)((w0 >> 16) & 255u), k), __fdiv_rn((float)((w1 >> 8) & 255u), k),
                                                             __fdiv_rn((float)(w2 & 255u), k), __fdiv_rn((float)(w2 >> 24), k));
    } else {
        if (i >= total) return;
        const long long n = i / HW;
        const int p = (int)(i - n * HW);
        const uint8_t* s = frames + i * 3;
        float* d = out + n * 3 * HW + p;
        d[0] = __fdiv_rn((float)__ldg(s), 255.0f);
        d[HW] = __fdiv_rn((float)__ldg(s + 1), 255.0f);
        d[2 * HW] = __fdiv_rn((float)__ldg(s + 2), 255.0f);
    }
}

}  // namespace pml
