// Shared device helpers for libpml.so (sm_100a).  Compiles with nvcc for the product and, with
// tests/emu/cuda_emu.h force-included, with g++ for the kernel-logic emulator (test only).
#pragma once
#include <stdint.h>

#ifndef PML_HOST_EMU
#include <cuda_runtime.h>
#define PML_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define PML_DYN_SMEM(type, name) \
    extern __shared__ __align__(16) unsigned char pml_dyn_smem_raw[]; \
    type* name = reinterpret_cast<type*>(pml_dyn_smem_raw)
#endif

#include "../../include/pml.h"

namespace pml {

constexpr float kSsimC1 = 0.0001f;   // 0.01^2, layers.py:231
constexpr float kSsimC2 = 0.0009f;   // 0.03^2, layers.py:232
constexpr float kTieNoise = 0.00001f;  // trainer.py:595

// A batch that arrives as n_seg separately allocated chunks of seg_size images (pml_segments: the per-timestep
// tensors of trainer_gru.py:890-899,943-957).  n_seg == 0: one tensor over the whole batch.
struct ChunkPtrs { const float* p[PML_MAX_SEGMENTS]; };
// base pointer of the tensor that holds image b, and b's index inside it
__device__ __forceinline__ const float* chunk_of(const float* whole, const ChunkPtrs& c, int n_seg, int seg_size, int b, int& local) {
    if (n_seg <= 0) { local = b; return whole; }
    const int sb = b / seg_size;
    local = b - sb * seg_size;
    return c.p[sb];
}

// ReflectionPad2d(1) index map restricted to the one-pixel ring (layers.py:229): -1 -> 1, n -> n-2.
__device__ __forceinline__ int reflect1(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * n - 2 - i : i;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// One SSIM window (layers.py:238-248) from its five 3x3 sums.  Returns the dissimilarity
// clamp((1 - n/d)/2, 0, 1); if WITH_GRAD also the partials of that value with respect to the
// three x-dependent *sums* (Sx, Sxx, Sxy), zero where the clamp is active.
//
// LOWER_GATE: the reference clamps at 0 and 1 (layers.py:248) and autograd zeroes the gradient outside.
// In exact arithmetic n/d <= 1 always (2 mu_x mu_y <= mu_x^2 + mu_y^2, 2 sigma_xy <= sigma_x + sigma_y), so
// the clamp at 0 can only trigger through fp32 round-off on windows where x ~ y; the layer-level drop-in
// keeps ATen's behaviour bit for bit (LOWER_GATE), the fused loss follows the mathematical function -- and
// the float64 reference -- and lets the (tiny) gradient through there.
template <bool WITH_GRAD, bool LOWER_GATE = true>
__device__ __forceinline__ float ssim_window(float Sx, float Sy, float Sxx, float Syy, float Sxy,
                                             float& dSx, float& dSxx, float& dSxy) {
    const float k9 = 1.0f / 9.0f;
    float mx = Sx * k9, my = Sy * k9;
    // Products are rounded separately (never contracted into the sums) so that the reference's
    // symmetry survives: when x == y bit for bit, numerator == denominator bit for bit and the
    // dissimilarity is exactly 0, as in layers.py:245-248 (static frames, mono_dataset.py:165-170).
    float mxx = __fmul_rn(mx, mx), myy = __fmul_rn(my, my), mxy = __fmul_rn(mx, my);
    float sx = fmaf(Sxx, k9, -mxx);
    float sy = fmaf(Syy, k9, -myy);
    float sxy = fmaf(Sxy, k9, -mxy);
    float A1 = fmaf(2.0f, mxy, kSsimC1);
    float A2 = fmaf(2.0f, sxy, kSsimC2);
    float B1 = __fadd_rn(__fadd_rn(mxx, myy), kSsimC1);
    float B2 = __fadd_rn(__fadd_rn(sx, sy), kSsimC2);
    float num = __fmul_rn(A1, A2), den = __fmul_rn(B1, B2);
    float inv = __frcp_rn(den);
    float ratio = num * inv;
    ratio = fmaf(fmaf(-den, ratio, num), inv, ratio);   // one Newton step: correctly rounded n/d
    float raw = fmaf(-0.5f, ratio, 0.5f);
    float val = fminf(fmaxf(raw, 0.0f), 1.0f);
    if (WITH_GRAD) {
        // clamp backward passes the gradient on the closed interval [0,1]
        float gate = ((!LOWER_GATE || raw >= 0.0f) && raw <= 1.0f) ? 1.0f : 0.0f;
        float dA1 = -0.5f * A2 * inv;
        float dA2 = -0.5f * A1 * inv;
        float dB1 = 0.5f * ratio * (inv * B2);  // ratio / B1
        float dB2 = 0.5f * ratio * (inv * B1);  // ratio / B2
        float gmu = 2.0f * my * (dA1 - dA2) + 2.0f * mx * (dB1 - dB2);
        // mu_x = Sx/9, E[xx] = Sxx/9, E[xy] = Sxy/9
        dSx = gate * gmu * k9;
        dSxx = gate * dB2 * k9;
        dSxy = gate * 2.0f * dA2 * k9;
    }
    return val;
}

// Philox4x32-10 (counter-based RNG) -> four standard normals by Box-Muller.  Used only when the
// caller does not supply the tie-break noise tensor (trainer.py:594 draws it on the host).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t idx_lo, uint32_t idx_hi,
                                               uint32_t stream_id, float n[4]) {
    uint32_t r[4];
    philox4x32_10(idx_lo, idx_hi, stream_id, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const float k = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float u1 = ((float)r[2 * i] + 1.0f) * k;      // (0,1]
        float u2 = (float)r[2 * i + 1] * k;           // [0,1)
        float rad = sqrtf(-2.0f * __logf(u1));
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        n[2 * i] = rad * c;
        n[2 * i + 1] = rad * s;
    }
}


// Philox2x32-10 -> two standard normals by one Box-Muller step: the tie-break noise of one pixel
// (two identity candidates, trainer.py:594) from ~45 instructions.  The noise only orders candidates
// that are closer than ~1e-4, so MUFU-precision log / sin / cos are ample.
__device__ __forceinline__ void philox2_normal2(uint32_t key, uint32_t c0, uint32_t c1, float& n0, float& n1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi = __umulhi(0xD256D193u, c0), lo = 0xD256D193u * c0;
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += 0x9E3779B9u;
    }
    const float k = 2.3283064365386963e-10f;            // 2^-32
    const float u1 = ((float)c0 + 1.0f) * k;            // (0,1], never subnormal
    const float ang = ((float)c1 * k - 0.5f) * 6.283185307179586f;   // [-pi, pi)
#ifdef PML_HOST_EMU
    const float rad = sqrtf(-2.0f * logf(u1));
#else
    // MUFU.LG2 / MUFU.SQRT directly: the libm-style wrappers carry subnormal and slow-path handling
    // (a CALL in the hot loop) that this argument range never needs
    float lg, rad;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(lg * -1.3862943611198906f));   // -2 ln 2 * log2(u1)
#endif
    float sn, cs;
    __sincosf(ang, &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

}  // namespace pml
