// Fused view-synthesis + reprojection-loss sweep, second generation ("warp strips").
//
// Same mathematics as pml_photometric.cuh (trainer.py:469-515 + :546-610 and their adjoint in one
// pass; cam_points / pix_coords / warped images never exist in HBM), re-shaped around what bounds
// this path on sm_100a: it is fp32-pipe / issue bound, not HBM bound (the 67 MB working set of the
// headline config lives in L2), so the design minimises executed instructions and keeps every warp
// free of block-wide barriers:
//
//  * one WARP owns a strip of 28 image columns (+2 halo columns each side = 32 lanes) x TH rows of
//    one (scale, image) and marches down the rows on its own; horizontal neighbours of the 3x3 SSIM
//    windows and of the transposed windows travel by warp shuffles -- no __syncthreads anywhere;
//  * the two source frames of a pixel are evaluated as one packed fp32x2 value (FFMA2 on sm_100a):
//    projection, bilinear taps, SSIM statistics, SSIM adjoint, pose / depth adjoint are written once
//    and executed for both frames per instruction (S = 1 aliases frame 0 into the second half);
//  * per-pixel geometry needed by the adjoint two rows later sits in a lane-private shared-memory
//    ring (7 x 128-bit per row, conflict free), rolling 3x3 sums stay in registers and rotate by
//    renaming (the row loop is unrolled by two with the roles of the two register sets swapped);
//  * streaming inputs (target, disparity, identity losses, noise) are loaded one row step ahead of
//    their use; all index arithmetic is 32 bit, one IMAD.WIDE per address, +1 taps are immediates.
//
// Layout of the work: grid = (n_chunks * n_strips, B, n_pass), one warp per CTA, so that image and
// pass are warp-uniform by construction.  item = ((pass * B + b) * n_chunks + chunk) * n_strips +
// strip writes one row of partials (loss, 12 x S pose-gradient sums) that finalize_image_kernel
// reduces in fixed order.
#pragma once
#include "pml_common.cuh"
#include "pml_photometric.cuh"

namespace pml {

constexpr int kSweepWarps = 1;    // warps (= items) per CTA
constexpr int kSweepTW = 28;      // owned columns per strip
constexpr int kSweepRingQ = 7;    // float4 per lane per ring slot
constexpr int kSweepRingQ1 = 4;   // ... of a single-frame launch
constexpr int kSweepRingSlots = 4;   // rows r .. r-3 (the adjoint of row r-3 runs while row r's taps are in flight)
constexpr int kSweepWarpFloats = 48 + 32 + kSweepRingSlots * kSweepRingQ * 4 * 32;   // P/IK, staging row, ring
constexpr int kSweepFwdCtas = 16;  // forward-only sweeps: registers capped for 16 resident warps per SM (they have no adjoint work to hide the gathers behind)
constexpr int kSweepLoneCtas = 12; // single-frame sweeps with adjoint: scalar arithmetic fits 168 registers without spills
// register budget of the forward+backward sweep of a frame pair and the resident warps per SM it allows
#ifndef PML_SWEEP_REGS
#define PML_SWEEP_REGS 255
#endif
constexpr int kSweepPairCtas = 65536 / (((PML_SWEEP_REGS * 32 + 511) / 512) * 512);

// ---- packed fp32x2 helpers (FFMA2 / FMUL2 / FADD2 on sm_100a) -----------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 shfl_up2(float2 v) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float2 shfl_down2(float2 v) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
// The same arithmetic for a launch that sweeps ONE frame (the last "pair" of an odd frame count): only the first
// half is live, the second is the constant 0, so everything computed for it (taps, clamps, shuffles, address
// arithmetic) is dead code to the compiler.  The kernel and ssim_pair shadow the helpers above with these.
template <bool PAIR>
struct Packed {
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) {
        if constexpr (PAIR) return fma2(a, b, c); else return make_float2(fmaf(a.x, b.x, c.x), 0.f);
    }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) {
        if constexpr (PAIR) return mul2(a, b); else return make_float2(a.x * b.x, 0.f);
    }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) {
        if constexpr (PAIR) return add2(a, b); else return make_float2(a.x + b.x, 0.f);
    }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b) {
        if constexpr (PAIR) return sub2(a, b); else return make_float2(a.x - b.x, 0.f);
    }
    static __device__ __forceinline__ float2 up(float2 v) {
        if constexpr (PAIR) return shfl_up2(v); else return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), 0.f);
    }
    static __device__ __forceinline__ float2 down(float2 v) {
        if constexpr (PAIR) return shfl_down2(v); else return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), 0.f);
    }
};
#define PML_PACKED_OPS(PAIR)                                                                              \
    auto fma2 = [](float2 a, float2 b, float2 c) { return Packed<PAIR>::fma(a, b, c); };                  \
    auto mul2 = [](float2 a, float2 b) { return Packed<PAIR>::mul(a, b); };                               \
    auto add2 = [](float2 a, float2 b) { return Packed<PAIR>::add(a, b); };                               \
    auto sub2 = [](float2 a, float2 b) { return Packed<PAIR>::sub(a, b); };                               \
    auto shfl_up2 = [](float2 v) { return Packed<PAIR>::up(v); };                                         \
    auto shfl_down2 = [](float2 v) { return Packed<PAIR>::down(v); };                                     \
    (void)fma2; (void)mul2; (void)add2; (void)sub2; (void)shfl_up2; (void)shfl_down2

// base + 32-bit element index as ONE instruction (IMAD.WIDE); tensors have < 2^31 elements
__device__ __forceinline__ const float* at(const float* base, int idx) {
#ifdef PML_HOST_EMU
    return base + idx;
#else
    const float* r;
    asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(r) : "r"(idx), "l"(base));
    return r;
#endif
}
__device__ __forceinline__ float* at(float* base, int idx) {
    return const_cast<float*>(at(static_cast<const float*>(base), idx));
}
// MUFU.RCP (1 ulp); callers add the Newton step where the quotient decides something
__device__ __forceinline__ float rcp_approx(float x) {
#ifdef PML_HOST_EMU
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
__device__ __forceinline__ float rcp_nr(float x) {   // reciprocal, one Newton step: < 1 ulp
    float r = rcp_approx(x);
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
template <bool PAIR = true>
__device__ __forceinline__ float2 rcp_nr2(float2 x) {
    if constexpr (!PAIR) return make_float2(rcp_nr(x.x), 0.f);
    float2 r = f2(rcp_approx(x.x), rcp_approx(x.y));
    float2 e = fma2(f2(-x.x, -x.y), r, splat(1.0f));
    return fma2(r, e, r);
}

// SSIM dissimilarity of one window for both frames (layers.py:238-248).  x statistics are packed
// over the frame pair, the target statistics (my, myy + C1, sy + C2) are shared scalars.
// If WITH_GRAD, pa / pb / pe receive d val / d(Sx, x_q-multiplier, y_q-multiplier) * 9, i.e. the
// three adjoint coefficients of ssim_window() before the common factor kssim * wgt / 9.
template <bool WITH_GRAD, bool PAIR = true>
__device__ __forceinline__ float2 ssim_pair(float2 Sx, float2 Sxx, float2 Sxy, float my, float myyC1, float syC2,
                                            float2& pa, float2& pb, float2& pe) {
    PML_PACKED_OPS(PAIR);
    const float k9 = 1.0f / 9.0f;
    const float2 mx = mul2(Sx, splat(k9));
    const float2 mxx = mul2(mx, mx);
    const float2 mxy = mul2(mx, splat(my));
    const float2 sx = fma2(Sxx, splat(k9), f2(-mxx.x, -mxx.y));
    const float2 sxy = fma2(Sxy, splat(k9), f2(-mxy.x, -mxy.y));
    const float2 A1 = fma2(mxy, splat(2.0f), splat(kSsimC1));
    const float2 A2 = fma2(sxy, splat(2.0f), splat(kSsimC2));
    const float2 B1 = add2(mxx, splat(myyC1));
    const float2 B2 = add2(sx, splat(syC2));
    const float2 num = mul2(A1, A2), den = mul2(B1, B2);
    const float2 q = f2(rcp_approx(den.x), PAIR ? rcp_approx(den.y) : 0.f);
    // MUFU reciprocal (1 ulp) is enough here: the dissimilarity of a *warped* frame is never an
    // exact tie (the identity candidates, which can tie exactly, come from identity_kernel)
    const float2 ratio = mul2(num, q);
    const float2 val = f2(__saturatef(fmaf(-0.5f, ratio.x, 0.5f)), __saturatef(fmaf(-0.5f, ratio.y, 0.5f)));
    if (WITH_GRAD) {
        // clamp backward (layers.py:248) zeroes the gradient where raw > 1  <=>  ratio < -1.  The clamp at 0
        // (ratio > 1) cannot trigger in exact arithmetic (see ssim_window): it is left open, like float64
        const float2 gq = f2(ratio.x >= -1.0f ? q.x : 0.f, ratio.y >= -1.0f ? q.y : 0.f);
        pe = mul2(f2(-gq.x, -gq.y), A1);                         // d/dSxy * 9
        pb = mul2(mul2(gq, ratio), B1);                          // 2 * d/dSxx * 9
        const float2 w = mul2(mul2(mx, ratio), sub2(B2, B1));
        pa = mul2(gq, fma2(splat(-my), sub2(A2, A1), w));        // d/dSx * 9
    }
    return val;
}

// Transposed horizontal upsample of one finished low-res row of disparity gradients (the adjoint
// of trainer.py:474 along x), added to grad_disp.  `h` holds this lane's full-res column sum (0 for
// halo lanes); low-res column jbase + lane collects the <= 2k full-res columns that touch it.
// Runs once per k rows and strip, hence out of line.
__device__ __noinline__ void sweep_flush_row(float* sG, float* gd_row, float h, int lane, int jbase, int low_cols,
                                             int wd, int kk, int x0, int x1, float rscale) {
    sG[lane] = h;
    __syncwarp();
    const int j = jbase + lane;
    if (lane < low_cols && j < wd) {
        const int xa = max(kk * j - kk / 2, x0), xb = min(kk * j + (3 * kk) / 2 - 1, x1 - 1);
        float s = 0.f;
        for (int x = xa; x <= xb; ++x) {
            const float sx = fmaxf(fmaf(rscale, (float)x + 0.5f, -0.5f), 0.f);
            const int jj0 = (int)sx, jj1 = min(jj0 + 1, wd - 1);
            const float l = sx - (float)jj0;
            const float w = (jj0 == j ? 1.f - l : 0.f) + (jj1 == j ? l : 0.f);
            s = fmaf(w, sG[x - x0 + 2], s);
        }
        if (s != 0.f) atomicAdd(gd_row + j, s);
    }
    __syncwarp();
}

// MODE (see PhotoParams::mode): 0 selection in the sweep (all frames in this pair); 1 reprojection losses
// only (stored); 2 selection given (adjoint of an earlier pair); 3 last pair of several: selection over the
// identity candidates, the stored losses of the earlier pairs and this pair, per-frame weights applied.
// EMIT: some pass wants the by-products outputs[("depth",0,s)] / outputs[("color",f,s)] (trainer.py:480,
// :508) written.  The hot loop is two unrolled row steps of ~1000 instructions; keeping it under the 32 KB
// instruction cache matters (measured: 0.408 -> 0.371 ms), so everything optional is compiled out.
// COMMON: the default training configuration (a full pair of source frames, automask, per-frame min,
// in-kernel tie-break noise, no predictive mask) with its run-time flags folded into constants; the
// generic instantiation serves the rest.
// PAIR = false: the launch sweeps a single frame (the last one of an odd frame count) with scalar arithmetic.
template <bool GRAD, bool SSIM, int MODE, bool EMIT, bool COMMON = false, bool PAIR = true>
__global__ void __maxnreg__(GRAD ? (PAIR ? PML_SWEEP_REGS : 168) : 128)
sweep_kernel(const PhotoParams p) {
    PML_DYN_SMEM(float, smem);
    PML_PACKED_OPS(PAIR);
    // grid = (n_chunks * n_strips, B, n_pass): image and pass come straight from blockIdx, so every
    // base pointer derived from them is warp-uniform by construction (no division on that path)
    const int lane = threadIdx.x;
    const int b = blockIdx.y, pass_i = blockIdx.z;
    const int chunk = blockIdx.x / p.n_strips;
    const int strip = blockIdx.x - chunk * p.n_strips;
    const int item = (pass_i * p.B + b) * (p.n_chunks * p.n_strips) + blockIdx.x;
    const PassDev& ps = p.pass[pass_i];

    const int H = p.H, W = p.W, S = p.S;
    const int x0 = strip * kSweepTW, x1 = min(x0 + kSweepTW, W);
    const int y0 = chunk * p.TH, y1 = min(y0 + p.TH, H);
    const int cx = x0 - 2 + lane;                         // column on the reflect-padded grid
    const int rx = reflect1(clampi(cx, -1, W), W);        // image column actually evaluated
    const bool col_in_image = (cx >= 0) && (cx < W);
    const bool col_owned = (cx >= x0) && (cx < x1);
    const bool lane_inner = (lane >= 1) && (lane <= 30);  // lanes whose 3x3 window is complete

    const bool automask = COMMON ? true : !(p.flags & PML_FLAG_NO_AUTOMASK);
    const bool avg = COMMON ? false : (p.flags & PML_FLAG_AVG_REPROJ) != 0;
    constexpr int mode = MODE;
    const int n_sel = COMMON ? (MODE == 0 ? 2 : S) : (automask ? (avg ? 1 : S) : 0);   // identity candidates of the selection
    const int n_id = (mode == 0) ? n_sel : 0;             // ... evaluated by this launch
    const int fa = (COMMON && MODE == 0) ? 0 : p.f_base;  // frames in the two halves of every pair
    const bool two = !PAIR ? false : ((COMMON && MODE == 0) ? true : p.pair_n > 1);   // modes 1-3: the last pair of an odd frame count has one frame
    const int fb = two ? fa + 1 : fa;                     // one frame: it is aliased into the second half

    // ---- per-warp shared memory ----------------------------------------------------------------
    float* wsm = smem;
    const float4* sP4 = reinterpret_cast<const float4*>(wsm);        // [6]: 12 x (frame0, frame1) of P = (K T)[:3]
    const float4* sIK4 = reinterpret_cast<const float4*>(wsm + 24);  // inv_K: [0][1] [0][2] [1][1] [1][2] | [2][1] [2][2]
    float* sG = wsm + 48;                                  // [32] staging row of the transposed upsample
    float4* sRing = reinterpret_cast<float4*>(wsm + 80);   // [kSweepRingSlots][RQ][32]
    constexpr int RQ = PAIR ? kSweepRingQ : kSweepRingQ1;

    // chunked batch (pml_segments): image b is image `bl` of the tensors of chunk b / seg_size
    int bl = b;
    const float* __restrict__ tgt_g = p.target;
    const float* __restrict__ src0_g = p.src[fa];
    const float* __restrict__ src1_g = p.src[fb];
    const float* K_g = p.K;
    const float* iK_g = p.invK;
    if (p.n_seg > 0) {
        const int sb = b / p.seg_size;
        bl = b - sb * p.seg_size;
        tgt_g = p.target_c.p[sb]; src0_g = p.src_c[fa].p[sb]; src1_g = p.src_c[fb].p[sb];
        K_g = p.K_c.p[sb]; iK_g = p.invK_c.p[sb];
    }
    float rc0, rc1, rc2;    // column part of the back-projection ray r = inv_K[:3,:3] @ (x, y, 1)
    {
        if (lane < 24) {
            const int e = lane >> 1, f = (lane & 1) ? fb : fa, i = e >> 2, j = e & 3;
            const float* Kb = K_g + bl * 16;
            const float* Tb = p.T[f] + b * 16;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) a = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], a);   // layers.py:183
            wsm[lane] = a;
        }
        const float* ikb = iK_g + bl * 16;                                           // layers.py:164
        if (lane == 0) {
            wsm[24] = ikb[1]; wsm[25] = ikb[2]; wsm[26] = ikb[5]; wsm[27] = ikb[6];
            wsm[28] = ikb[9]; wsm[29] = ikb[10]; wsm[30] = 0.f; wsm[31] = 0.f;
        }
        const float fxc = (float)rx;
        rc0 = ikb[0] * fxc; rc1 = ikb[4] * fxc; rc2 = ikb[8] * fxc;
        // The adjoint block reads ring slots before the first rows have been written (and for lanes / rows that own
        // nothing) and masks du / dv, not the factors it multiplies them with: whatever shared memory held before
        // this CTA (e.g. the byte patterns of the pyramid kernel = NaN as floats) must not reach 0 * garbage
        if (GRAD) {
#pragma unroll
            for (int k = 0; k < kSweepRingSlots * RQ; ++k) sRing[k * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
    }

    // horizontal part of the disparity upsample (trainer.py:474): fixed per lane
    const int kk = ps.k, wd = ps.wd, hd = ps.hd;
    const float rscale = ps.rscale;
    int j0 = rx, j1 = rx;
    float lam = 0.f;
    if (kk > 1) {
        const float sx = fmaxf(fmaf(rscale, (float)rx + 0.5f, -0.5f), 0.f);
        j0 = (int)sx;
        j1 = min(j0 + 1, wd - 1);
        lam = sx - (float)j0;
    }

    const int plane = H * W;
    // All global addressing is `parameter pointer [32-bit element index]`: every tensor of a call has
    // fewer than 2^31 elements (checked on the host).
    const float* __restrict__ disp_g = ps.disp;
    const float* __restrict__ id_g = p.identity;
    const float* __restrict__ nz_g = COMMON ? nullptr : ps.noise;
    const float* __restrict__ fw_g = COMMON ? nullptr : ps.fw;      // predictive mask and its gradient
    float* __restrict__ gfw_g = COMMON ? nullptr : ps.gfw;
    const int b3p = bl * 3 * plane;         // image offset in a [B,3,H,W] input tensor (of its chunk)
    const int bdp = b * hd * wd;            // ... in disp_s / grad_disp_s
    const int bip = b * n_id * plane;       // ... in the identity-loss / noise tensors
    const int bp = b * plane;               // ... in a [B,1,H,W] tensor
    const float wscale = (float)W / (float)(W - 1), hscale = (float)H / (float)(H - 1);
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);
    const float wmax1 = (float)(W - 2), hmax1 = (float)(H - 2);
    const bool emit_any = EMIT && (mode != 2) && ((ps.depth != nullptr && fa == 0) || (ps.warped != nullptr));

    // ---- rolling state ---------------------------------------------------------------------------
    float hy1[3], hy2[3], hyy1[3], hyy2[3];
    float2 hx1[3], hx2[3], hxx1[3], hxx2[3], hxy1[3], hxy2[3];
    float2 hc1[GRAD ? 9 : 1], hc2[GRAD ? 9 : 1];
    float2 gP[GRAD ? 12 : 1];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        hy1[c] = hy2[c] = hyy1[c] = hyy2[c] = 0.f;
        hx1[c] = hx2[c] = hxx1[c] = hxx2[c] = hxy1[c] = hxy2[c] = splat(0.f);
    }
#pragma unroll
    for (int m = 0; m < (GRAD ? 9 : 1); ++m) hc1[m] = hc2[m] = splat(0.f);
#pragma unroll
    for (int e = 0; e < (GRAD ? 12 : 1); ++e) gP[e] = splat(0.f);
    float loss_acc = 0.f;
    float2 l1_prev = splat(0.f);     // sum_c |target - pred| of the previous row
    float2 wq1 = splat(0.f), wq2 = splat(0.f);   // winner weight / N of window rows r-2 and r-3
    float2 cfc[GRAD ? 9 : 1];        // adjoint coefficients of window row r-2 (consumed one step later)
#pragma unroll
    for (int m = 0; m < (GRAD ? 9 : 1); ++m) cfc[m] = splat(0.f);
    // transposed vertical upsample: low-res rows `cur` and `cur + 1` accumulate in registers
    float acc0 = 0.f, acc1 = 0.f;
    int cur = 0, jbase = 0;
    if (GRAD && kk > 1) {
        cur = (int)fmaxf(fmaf(rscale, (float)y0 + 0.5f, -0.5f), 0.f);
        jbase = (int)fmaxf(fmaf(rscale, (float)x0 + 0.5f, -0.5f), 0.f);
    }
    const int low_cols = kSweepTW / kk + 3;

    const float wl = (cx == 1) ? 2.f : 1.f, wr = (cx == W - 2) ? 2.f : 1.f;   // reflection fold (columns)
    const float kssim9 = SSIM ? (0.85f / 27.0f) : 0.f;      // 0.85 / 3 channels / 9 window taps
    const float kl1 = SSIM ? (0.15f / 3.0f) : (1.0f / 3.0f);

    auto flush_row = [&](int irow, float h) {
        sweep_flush_row(sG, ps.grad_disp + (bdp + irow * wd), h, lane, jbase, low_cols, wd, kk, x0, x1, rscale);
    };

    // Streaming inputs of one row, issued one row step before they are consumed: target colours and
    // the (up to four) disparity values around (rx, ry).
    float ynA[3], dnA[4], ynB[3], dnB[4];
    auto load_row = [&](int r, float (&yn)[3], float (&dn)[4]) {
        const int ry = reflect1(clampi(r, -1, H), H);
        const float* tq = at(tgt_g, b3p + ry * W + rx);
        yn[0] = __ldg(tq); yn[1] = __ldg(at(tq, plane)); yn[2] = __ldg(at(tq, 2 * plane));
        if (kk > 1) {
            const float sy = fmaxf(fmaf(rscale, (float)ry + 0.5f, -0.5f), 0.f);
            const int i0 = (int)sy, i1 = min(i0 + 1, hd - 1);
            const float* d0 = at(disp_g, bdp + i0 * wd);
            const float* d1 = at(disp_g, bdp + i1 * wd);
            dn[0] = __ldg(at(d0, j0)); dn[1] = __ldg(at(d0, j1)); dn[2] = __ldg(at(d1, j0)); dn[3] = __ldg(at(d1, j1));
        } else {
            dn[0] = __ldg(at(disp_g, bdp + ry * W + rx));
        }
    };
    load_row(y0 - 2, ynA, dnA);

    int slotA = 0;   // ring slot of row r; slot (slotA + 1) & 3 holds row r-3
    const int r_end = GRAD ? (y1 + 2) : y1;
    // One row step.  The rolling 3x3 sums are passed as (previous row, row before): the step reads
    // both and overwrites the older one, so calling it with the two sets swapped on alternate rows
    // (loop unrolled by two below) rotates the window without a single register move.
    auto step = [&](const int r, const float (&yv)[3], const float (&dn)[4], float (&ynN)[3], float (&dnN)[4],
                    float (&hyA)[3], float (&hyB)[3], float (&hyyA)[3], float (&hyyB)[3],
                    float2 (&hxA)[3], float2 (&hxB)[3], float2 (&hxxA)[3], float2 (&hxxB)[3],
                    float2 (&hxyA)[3], float2 (&hxyB)[3], float2 (&hcA)[GRAD ? 9 : 1], float2 (&hcB)[GRAD ? 9 : 1]) {
        // =================================== (A) warp row r ======================================
        const int ry = reflect1(clampi(r, -1, H), H);
        float d;
        if (kk > 1) {   // bilinear upsample of disp_s, align_corners=False (trainer.py:474)
            const float sy = fmaxf(fmaf(rscale, (float)ry + 0.5f, -0.5f), 0.f);
            const float mu = sy - floorf(sy);
            const float top = fmaf(lam, dn[1], (1.f - lam) * dn[0]), bot = fmaf(lam, dn[3], (1.f - lam) * dn[2]);
            d = fmaf(mu, bot, (1.f - mu) * top);
        } else {
            d = dn[0];
        }
        load_row(r + 1, ynN, dnN);   // consumed by the next step
        // identity losses / noise of the window row r-1 (used at the end of (B)); clamped so that the
        // early, unconditional loads stay inside the tensors
        const int py = r - 1;
        float idv0 = 0.f, idv1 = 0.f, nzv0 = 0.f, nzv1 = 0.f;
        if (n_id > 0) {
            const float* ic = at(id_g, bip + clampi(py, 0, H - 1) * W + rx);
            idv0 = __ldg(ic);
            if (n_id > 1) idv1 = __ldg(at(ic, plane));
            if (nz_g != nullptr) {
                const float* nc = at(nz_g, bip + clampi(py, 0, H - 1) * W + rx);
                nzv0 = __ldg(nc);
                if (n_id > 1) nzv1 = __ldg(at(nc, plane));
            }
        }
        // multi-frame selection, default flags: the best candidate ahead of this pair (select_prepass_kernel)
        float ob = 0.f;
        int oi = 0;
        if (COMMON && MODE == 3) {
            const int o = clampi(py, 0, H - 1) * W + rx;
            ob = __ldg(p.rp + ((size_t)(pass_i * S + fa) * p.B + b) * plane + o);
            oi = __ldg(p.presel + (size_t)(pass_i * p.B + b) * plane + o);
        }
        if (MODE == 2) oi = ps.argmin[bp + clampi(py, 0, H - 1) * W + rx];   // the selection made by the mode-3 sweep

        const float sigma = fmaf(p.disp_range, d, p.min_disp);   // layers.py:23
        const float D = rcp_nr(sigma);                            // layers.py:24
        const float fy = (float)ry;
        const float4 ika = sIK4[0], ikb4 = sIK4[1];
        const float r0 = rc0 + fmaf(ika.x, fy, ika.y);
        const float r1 = rc1 + fmaf(ika.z, fy, ika.w);
        const float r2 = rc2 + fmaf(ikb4.x, fy, ikb4.y);
        const float X0 = D * r0, X1 = D * r1, X2 = D * r2;       // layers.py:165

        // projection of both frames (layers.py:183-187); P comes from shared memory (broadcast)
        float2 c0, c1, c2;
        {
            const float4 pa0 = sP4[0], pa1 = sP4[1], pb0 = sP4[2], pb1 = sP4[3], pc0 = sP4[4], pc1 = sP4[5];
            c0 = fma2(f2(pa0.x, pa0.y), splat(X0), fma2(f2(pa0.z, pa0.w), splat(X1), fma2(f2(pa1.x, pa1.y), splat(X2), f2(pa1.z, pa1.w))));
            c1 = fma2(f2(pb0.x, pb0.y), splat(X0), fma2(f2(pb0.z, pb0.w), splat(X1), fma2(f2(pb1.x, pb1.y), splat(X2), f2(pb1.z, pb1.w))));
            c2 = fma2(f2(pc0.x, pc0.y), splat(X0), fma2(f2(pc0.z, pc0.w), splat(X1), fma2(f2(pc1.x, pc1.y), splat(X2), f2(pc1.z, pc1.w))));
        }
        const float2 invz = rcp_nr2<PAIR>(add2(c2, splat(p.eps)));
        const float2 u = mul2(c0, invz), v = mul2(c1, invz);
        // layers.py:190-192 + grid_sample unnormalise (align_corners=False): ix = u*W/(W-1) - 0.5
        const float2 ixr = fma2(u, splat(wscale), splat(-0.5f)), iyr = fma2(v, splat(hscale), splat(-0.5f));
        const float2 ix = f2(fminf(fmaxf(ixr.x, 0.f), wmax), fminf(fmaxf(ixr.y, 0.f), wmax));
        const float2 iy = f2(fminf(fmaxf(iyr.x, 0.f), hmax), fminf(fmaxf(iyr.y, 0.f), hmax));
        // base tap clamped to W-2 / H-2: the +1 tap then always exists and carries the weight that
        // grid_sample gives to the last column / row (its out-of-range tap has weight 0)
        const float2 fx0 = f2(fminf(floorf(ix.x), wmax1), fminf(floorf(ix.y), wmax1));
        const float2 fy0 = f2(fminf(floorf(iy.x), hmax1), fminf(floorf(iy.y), hmax1));
        const float2 tx = sub2(ix, fx0), ty = sub2(iy, fy0);
        // clip backward (zero outside the open interval) times d ix / d u
        const float2 mx = f2((ixr.x > 0.f && ixr.x < wmax) ? wscale : 0.f, (ixr.y > 0.f && ixr.y < wmax) ? wscale : 0.f);
        const float2 my = f2((iyr.x > 0.f && iyr.x < hmax) ? hscale : 0.f, (iyr.y > 0.f && iyr.y < hmax) ? hscale : 0.f);
        // address of the north-west tap; the other taps are warp-uniform offsets (plane, W) from it,
        // or the immediate +1
        const int o0 = b3p + (int)fy0.x * W + (int)fx0.x, o1 = b3p + (int)fy0.y * W + (int)fx0.y;

        // ---- issue the 24 taps of row r; they are consumed after the adjoint below ----
        float2 nw[3], ne[3], sw[3], se[3];
        {
            // one address from the (uniform) tensor base per frame, the other five chained from it in
            // registers: base + index from a uniform base costs two instructions, register + index one
            const float* a0 = at(src0_g, o0);
            const float* a1 = at(src1_g, o1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* w0 = at(a0, W);
                const float* w1 = at(a1, W);
                nw[c] = f2(__ldg(a0), __ldg(a1)); ne[c] = f2(__ldg(a0 + 1), __ldg(a1 + 1));
                sw[c] = f2(__ldg(w0), __ldg(w1)); se[c] = f2(__ldg(w0 + 1), __ldg(w1 + 1));
                if (c < 2) { a0 = at(a0, plane); a1 = at(a1, plane); }
            }
        }

        // ============ (C) adjoint for the pixels of row r-3, while the taps are in flight ==========
        // cfc: coefficients of window row r-2 (previous step); wq2: winner weight of row r-3
        if (GRAD) {
            const int qy = r - 3;
            const bool do_q = (qy >= y0) && (qy < y1);
            const float wt = (qy == 1) ? 2.f : 1.f, wb = (qy == H - 2) ? 2.f : 1.f;   // reflection fold (rows)
            float2 V[9];
            if (SSIM) {
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    const float2 cf = cfc[m];
                    const float2 cl = shfl_up2(cf), cr = shfl_down2(cf);
                    const float2 hn = fma2(splat(wl), cl, fma2(splat(wr), cr, cf));
                    V[m] = fma2(splat(wt), hcB[m], fma2(splat(wb), hn, hcA[m]));
                    hcB[m] = hn;
                }
            }
            float g_d;
            {   // straight-line for every lane (no divergent join): lanes / rows that own nothing read
                // whatever the ring holds and are masked where du / dv enter the sums
                const bool own_q = do_q && col_owned;
                const float4* rc = sRing + (((slotA + 1) & 3) * RQ) * 32 + lane;
                float yq[3], Dq;
                float2 xq[3], dxq[3], dyq[3], invzq, uq, vq;
                if constexpr (PAIR) {
                    const float4 q0r = rc[0 * 32], q1r = rc[1 * 32], q2r = rc[2 * 32], q3r = rc[3 * 32];
                    const float4 q4r = rc[4 * 32], q5r = rc[5 * 32], q6r = rc[6 * 32];
                    yq[0] = q0r.x; yq[1] = q0r.y; yq[2] = q0r.z; Dq = q0r.w;
                    xq[0] = f2(q1r.x, q1r.y); xq[1] = f2(q1r.z, q1r.w); xq[2] = f2(q2r.x, q2r.y);
                    dxq[0] = f2(q2r.z, q2r.w); dxq[1] = f2(q3r.x, q3r.y); dxq[2] = f2(q3r.z, q3r.w);
                    dyq[0] = f2(q4r.x, q4r.y); dyq[1] = f2(q4r.z, q4r.w); dyq[2] = f2(q5r.x, q5r.y);
                    invzq = f2(q5r.z, q5r.w); uq = f2(q6r.x, q6r.y); vq = f2(q6r.z, q6r.w);
                } else {
                    const float4 q0r = rc[0 * 32], q1r = rc[1 * 32], q2r = rc[2 * 32], q3r = rc[3 * 32];
                    yq[0] = q0r.x; yq[1] = q0r.y; yq[2] = q0r.z; Dq = q0r.w;
                    xq[0] = f2(q1r.x, 0.f); xq[1] = f2(q1r.y, 0.f); xq[2] = f2(q1r.z, 0.f);
                    dxq[0] = f2(q1r.w, 0.f); dxq[1] = f2(q2r.x, 0.f); dxq[2] = f2(q2r.y, 0.f);
                    dyq[0] = f2(q2r.z, 0.f); dyq[1] = f2(q2r.w, 0.f); dyq[2] = f2(q3r.x, 0.f);
                    invzq = f2(q3r.y, 0.f); uq = f2(q3r.z, 0.f); vq = f2(q3r.w, 0.f);
                }
                const float2 kw = mul2(wq2, splat(kl1));
                float2 du = splat(0.f), dv = splat(0.f);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float2 df = sub2(xq[c], splat(yq[c]));
                    // d|x - y| / dx = sign(x - y).  sign(0) is taken as +1 instead of the reference's 0:
                    // exact equality of a warped value and the target only happens on locally constant
                    // images, where the bilinear slopes that multiply this term are exactly 0
                    const float2 sg = f2(copysignf(kw.x, df.x), copysignf(kw.y, df.y));
                    float2 g = sg;
                    if (SSIM) g = add2(g, fma2(xq[c], V[3 + c], fma2(splat(yq[c]), V[6 + c], V[c])));
                    du = fma2(g, dxq[c], du);
                    dv = fma2(g, dyq[c], dv);
                }
                du = f2(own_q ? du.x : 0.f, own_q ? du.y : 0.f);
                dv = f2(own_q ? dv.x : 0.f, own_q ? dv.y : 0.f);
                const float2 dc0 = mul2(du, invzq), dc1 = mul2(dv, invzq);
                const float2 t = fma2(uq, du, mul2(vq, dv));
                const float2 dc2 = mul2(f2(-t.x, -t.y), invzq);
                const float fq = (float)qy;
                const float4 ikq = sIK4[0], ikq2 = sIK4[1];
                const float rq0 = rc0 + fmaf(ikq.x, fq, ikq.y), rq1 = rc1 + fmaf(ikq.z, fq, ikq.w);
                const float rq2 = rc2 + fmaf(ikq2.x, fq, ikq2.y);
                const float Xq0 = Dq * rq0, Xq1 = Dq * rq1, Xq2 = Dq * rq2;
                gP[0] = fma2(dc0, splat(Xq0), gP[0]); gP[1] = fma2(dc0, splat(Xq1), gP[1]);
                gP[2] = fma2(dc0, splat(Xq2), gP[2]); gP[3] = add2(gP[3], dc0);
                gP[4] = fma2(dc1, splat(Xq0), gP[4]); gP[5] = fma2(dc1, splat(Xq1), gP[5]);
                gP[6] = fma2(dc1, splat(Xq2), gP[6]); gP[7] = add2(gP[7], dc1);
                gP[8] = fma2(dc2, splat(Xq0), gP[8]); gP[9] = fma2(dc2, splat(Xq1), gP[9]);
                gP[10] = fma2(dc2, splat(Xq2), gP[10]); gP[11] = add2(gP[11], dc2);
                // gX = P[:, :3]^T dc
                const float4 pa0 = sP4[0], pa1 = sP4[1], pb0 = sP4[2], pb1 = sP4[3], pc0 = sP4[4], pc1 = sP4[5];
                const float2 gX0 = fma2(f2(pa0.x, pa0.y), dc0, fma2(f2(pb0.x, pb0.y), dc1, mul2(f2(pc0.x, pc0.y), dc2)));
                const float2 gX1 = fma2(f2(pa0.z, pa0.w), dc0, fma2(f2(pb0.z, pb0.w), dc1, mul2(f2(pc0.z, pc0.w), dc2)));
                const float2 gX2 = fma2(f2(pa1.x, pa1.y), dc0, fma2(f2(pb1.x, pb1.y), dc1, mul2(f2(pc1.x, pc1.y), dc2)));
                const float2 gD = fma2(splat(rq0), gX0, fma2(splat(rq1), gX1, mul2(splat(rq2), gX2)));
                g_d = own_q ? -p.disp_range * Dq * Dq * (gD.x + gD.y) : 0.f;   // d(1/sigma)/d disp, both frames
            }
            if (do_q) {
                if (kk > 1) {
                    const float sy = fmaxf(fmaf(rscale, (float)qy + 0.5f, -0.5f), 0.f);
                    const int i0 = (int)sy;
                    const float mu = sy - (float)i0;
                    if (i0 > cur) {            // low-res row `cur` is complete (warp-uniform)
                        flush_row(cur, acc0);
                        acc0 = acc1; acc1 = 0.f; cur = i0;
                    }
                    if (i0 + 1 <= hd - 1) { acc0 = fmaf(1.f - mu, g_d, acc0); acc1 = fmaf(mu, g_d, acc1); }
                    else acc0 += g_d;
                } else if (col_owned && ps.grad_disp != nullptr) {
                    atomicAdd(at(ps.grad_disp, bp + qy * W + cx), g_d);   // RED: fire and forget
                }
            }
        }
        // ---- bilinear interpolation of row r (trainer.py:508) and its slopes ----
        float2 xv[3], dpx[3], dpy[3];
        float2 l1_cur = splat(0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float2 dt = sub2(ne[c], nw[c]), db = sub2(se[c], sw[c]);
            const float2 top = fma2(tx, dt, nw[c]), bot = fma2(tx, db, sw[c]);
            const float2 dvert = sub2(bot, top);
            xv[c] = fma2(ty, dvert, top);
            if (GRAD) {
                dpx[c] = mul2(mx, fma2(ty, sub2(db, dt), dt));
                dpy[c] = mul2(my, dvert);
            }
            const float2 df = sub2(xv[c], splat(yv[c]));
            l1_cur.x += fabsf(df.x);
            l1_cur.y += fabsf(df.y);
        }
        if (EMIT && emit_any && col_owned && (r >= y0) && (r < y1)) {   // trainer.py:480, :508 (on request)
            const int o = r * W + cx;
            if (ps.depth != nullptr && fa == 0) ps.depth[bp + o] = D;
            if (ps.warped != nullptr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    ps.warped[(size_t)fa * p.B * 3 * plane + 3 * bp + c * plane + o] = xv[c].x;
                    if (two) ps.warped[(size_t)fb * p.B * 3 * plane + 3 * bp + c * plane + o] = xv[c].y;
                }
            }
        }
        if (GRAD) {
            float4* ra = sRing + (slotA * RQ) * 32 + lane;
            ra[0 * 32] = make_float4(yv[0], yv[1], yv[2], D);
            if constexpr (PAIR) {
                ra[1 * 32] = make_float4(xv[0].x, xv[0].y, xv[1].x, xv[1].y);
                ra[2 * 32] = make_float4(xv[2].x, xv[2].y, dpx[0].x, dpx[0].y);
                ra[3 * 32] = make_float4(dpx[1].x, dpx[1].y, dpx[2].x, dpx[2].y);
                ra[4 * 32] = make_float4(dpy[0].x, dpy[0].y, dpy[1].x, dpy[1].y);
                ra[5 * 32] = make_float4(dpy[2].x, dpy[2].y, invz.x, invz.y);
                ra[6 * 32] = make_float4(u.x, u.y, v.x, v.y);
            } else {
                ra[1 * 32] = make_float4(xv[0].x, xv[1].x, xv[2].x, dpx[0].x);
                ra[2 * 32] = make_float4(dpx[1].x, dpx[2].x, dpy[0].x, dpy[1].x);
                ra[3 * 32] = make_float4(dpy[2].x, invz.x, u.x, v.x);
            }
        }

        // ========================= (B) close the windows centred on row r-1 ======================
        const bool p_valid = (r >= y0) && (py >= 0) && (py < H) && col_in_image && lane_inner;
        float2 rp = splat(0.f);
        float2 pa[3], pb[3], pe[3];
        {
            float2 ssim_sum = splat(0.f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (SSIM) {
                    const float yl = __shfl_up_sync(0xffffffffu, yv[c], 1), yr = __shfl_down_sync(0xffffffffu, yv[c], 1);
                    const float hyn = yl + yv[c] + yr;
                    const float hyyn = fmaf(yl, yl, fmaf(yv[c], yv[c], yr * yr));
                    const float2 xl = shfl_up2(xv[c]), xr = shfl_down2(xv[c]);
                    const float2 hxn = add2(add2(xl, xv[c]), xr);
                    const float2 hxxn = fma2(xl, xl, fma2(xv[c], xv[c], mul2(xr, xr)));
                    const float2 hxyn = fma2(xl, splat(yl), fma2(xv[c], splat(yv[c]), mul2(xr, splat(yr))));
                    const float Sy = hyB[c] + hyA[c] + hyn;
                    const float Syy = hyyB[c] + hyyA[c] + hyyn;
                    const float2 Sx = add2(add2(hxB[c], hxA[c]), hxn);
                    const float2 Sxx = add2(add2(hxxB[c], hxxA[c]), hxxn);
                    const float2 Sxy = add2(add2(hxyB[c], hxyA[c]), hxyn);
                    hyB[c] = hyn; hyyB[c] = hyyn; hxB[c] = hxn; hxxB[c] = hxxn; hxyB[c] = hxyn;
                    const float k9 = 1.0f / 9.0f;
                    const float my_ = Sy * k9;
                    const float myy = my_ * my_;
                    const float sy_ = fmaf(Syy, k9, -myy);
                    ssim_sum = add2(ssim_sum, ssim_pair<GRAD, PAIR>(Sx, Sxx, Sxy, my_, myy + kSsimC1, sy_ + kSsimC2,
                                                              pa[c], pb[c], pe[c]));
                }
            }
            // trainer.py:527 (0.85 * SSIM.mean(1) + 0.15 * L1.mean(1)) or :523 (L1 only)
            rp = SSIM ? fma2(ssim_sum, splat(0.85f / 3.0f), mul2(l1_prev, splat(0.15f / 3.0f)))
                      : mul2(l1_prev, splat(1.0f / 3.0f));
        }
        l1_prev = l1_cur;

        // In-kernel tie-break noise (no noise tensor given).  Box-Muller on 32-bit uniforms is bounded,
        // |n| <= sqrt(-2 ln 2^-32) = 6.66, so noise * 1e-5 cannot change the selection where the
        // candidates are further apart than 1.4e-4; the generator only runs for row steps in which
        // some pixel of the strip is that close (always on static frames, whose identity losses are 0).
        if (mode == 0 && n_id > 0 && nz_g == nullptr) {
            const float m_id = (n_id > 1) ? fminf(idv0, idv1) : idv0;
            const float m_rp = avg ? (two ? (rp.x + rp.y) * 0.5f : rp.x) : (two ? fminf(rp.x, rp.y) : rp.x);
            const bool close = (n_id > 1 && fabsf(idv0 - idv1) < 1.4e-4f) || (fabsf(m_id - m_rp) < 1.4e-4f);
            if (__any_sync(0xffffffffu, close && p_valid))
            {
                const unsigned long long sd = p.seed_dev ? (p.seed ^ __ldg(p.seed_dev)) : p.seed;
                philox2_normal2((uint32_t)sd ^ ((uint32_t)(sd >> 32) * 0x9E3779B9u),
                                (uint32_t)(bp + py * W + cx), (uint32_t)pass_i, nzv0, nzv1);
            }
        }
        float2 wgt = splat(0.f);
        if (mode == 1) {
            // more than two source frames, first sweep: only the reprojection losses of this pair
            if (p_valid && col_owned && py >= y0 && py < y1) {
                float* rq = p.rp + ((size_t)(pass_i * S + fa) * p.B + b) * plane + (py * W + cx);
                rq[0] = rp.x;
                if (two) rq[(size_t)p.B * plane] = rp.y;
            }
        } else if (mode == 2) {
            // ... adjoint sweep of an earlier pair: the selection was made by the last pair's sweep (mode 3)
            if (p_valid) {
                const int idx = oi;
                if (avg) wgt = (idx == n_sel) ? f2(1.0f / (float)S, two ? 1.0f / (float)S : 0.f) : splat(0.f);
                else wgt = f2(idx == n_sel + fa ? 1.f : 0.f, (two && idx == n_sel + fb) ? 1.f : 0.f);
                if (fw_g != nullptr) {   // predictive mask (trainer.py:579): d (rp * m) / d rp = m
                    const float* mq = at(fw_g, (b * S + fa) * plane + py * W + cx);
                    wgt = mul2(wgt, f2(__ldg(mq), two ? __ldg(at(mq, plane)) : 0.f));
                }
            }
        } else if (COMMON && mode == 3) {
            // identity candidates and the frames of the earlier pairs were reduced by select_prepass_kernel (in the
            // reference's order); this pair's frames come last, so they win on strictly smaller only
            if (p_valid) {
                float best = ob;
                int best_i = oi;
                if (rp.x < best) { best = rp.x; best_i = n_sel + fa; }
                if (two && rp.y < best) { best = rp.y; best_i = n_sel + fb; }
                wgt = f2(best_i == n_sel + fa ? 1.f : 0.f, (two && best_i == n_sel + fb) ? 1.f : 0.f);
                if (col_owned && py >= y0 && py < y1) {
                    loss_acc += best;
                    ps.argmin[bp + py * W + cx] = (uint8_t)best_i;
                }
            }
        } else if (mode == 3) {
            if (p_valid) {
                const int pix = py * W + cx;
                float best = 3.0e38f;
                int best_i = 0;
                // reprojection losses of the pairs swept before (mode 1), in frame order, then this pair;
                // the predictive mask weighs every frame (trainer.py:579)
                const float* rq = p.rp + ((size_t)(pass_i * S) * p.B + b) * plane + pix;   // frame stride: B * plane
                const float* mq = (fw_g != nullptr) ? at(fw_g, b * S * plane + pix) : nullptr;
                float rsum = 0.f;
                for (int f = 0; f < fa; ++f) {
                    float r = __ldg(at(rq, f * p.B * plane));
                    if (mq != nullptr) r *= __ldg(at(mq, f * plane));
                    if (avg) rsum += r;
                    else if (r < best) { best = r; best_i = n_sel + f; }
                }
                const float2 m = (mq != nullptr) ? f2(__ldg(at(mq, fa * plane)), two ? __ldg(at(mq, fb * plane)) : 0.f) : splat(1.f);
                const float2 rw = mul2(rp, m);
                if (avg) {
                    best = (rsum + rw.x + (two ? rw.y : 0.f)) / (float)S;     // trainer.py:585-586
                    best_i = n_sel;
                } else {
                    if (rw.x < best) { best = rw.x; best_i = n_sel + fa; }
                    if (two && rw.y < best) { best = rw.y; best_i = n_sel + fb; }
                }
                // identity candidates + tie-break noise (trainer.py:592-597), any number of them.  They precede the
                // reprojection candidates in the reference's order (torch.min returns the first minimum), hence `<=` below.
                // In-kernel noise is bounded (|n| <= 6.66, see above): an identity candidate further than 1.4e-4 above
                // the best reprojection candidate cannot win whatever the noise, and the generator is skipped.
                if (n_sel > 0) {
                    float m_id = 3.0e38f;
                    for (int i = 0; i < n_sel; ++i) m_id = fminf(m_id, __ldg(at(id_g, (b * n_sel + i) * plane + pix)));
                    if (nz_g != nullptr || m_id - best < 1.4e-4f) {
                        float ib = 3.0e38f;
                        int ib_i = 0;
                        for (int i = 0; i < n_sel; i += 2) {
                            float n0 = 0.f, n1 = 0.f;
                            const int o = (b * n_sel + i) * plane + pix;
                            if (nz_g != nullptr) {
                                n0 = __ldg(at(nz_g, o));
                                if (i + 1 < n_sel) n1 = __ldg(at(nz_g, o + plane));
                            } else {
                                const unsigned long long sd = p.seed_dev ? (p.seed ^ __ldg(p.seed_dev)) : p.seed;
                                philox2_normal2((uint32_t)sd ^ ((uint32_t)(sd >> 32) * 0x9E3779B9u), (uint32_t)(bp + pix),
                                                (uint32_t)pass_i + 0x10000u * (uint32_t)(i >> 1), n0, n1);
                            }
                            const float c0 = fmaf(n0, kTieNoise, __ldg(at(id_g, o)));
                            if (c0 < ib) { ib = c0; ib_i = i; }
                            if (i + 1 < n_sel) {
                                const float c1 = fmaf(n1, kTieNoise, __ldg(at(id_g, o + plane)));
                                if (c1 < ib) { ib = c1; ib_i = i + 1; }
                            }
                        }
                        if (ib <= best) { best = ib; best_i = ib_i; }
                    }
                }
                if (avg) wgt = (best_i == n_sel) ? f2(1.0f / (float)S, two ? 1.0f / (float)S : 0.f) : splat(0.f);
                else wgt = f2(best_i == n_sel + fa ? 1.f : 0.f, (two && best_i == n_sel + fb) ? 1.f : 0.f);
                if (col_owned && py >= y0 && py < y1) {
                    loss_acc += best;
                    if (ps.argmin != nullptr) ps.argmin[bp + pix] = (uint8_t)best_i;
                    if (gfw_g != nullptr) {   // d mean(to_optimise) / d mask_f = [f selected] * rp_f / N
                        float* gq = at(gfw_g, b * S * plane + pix);
                        for (int f = 0; f < fa; ++f) {
                            const float sel = avg ? (best_i == n_sel ? 1.0f / (float)S : 0.f) : (best_i == n_sel + f ? 1.f : 0.f);
                            *at(gq, f * plane) = sel * __ldg(at(rq, f * p.B * plane)) * p.inv_n;
                        }
                        *at(gq, fa * plane) = wgt.x * rp.x * p.inv_n;
                        if (two) *at(gq, fb * plane) = wgt.y * rp.y * p.inv_n;
                    }
                }
                wgt = mul2(wgt, m);     // d (rp * m) / d rp
            }
        } else if (p_valid) {
            // candidates in the reference's order: identity (+noise) first, then reprojection
            // (trainer.py:597); torch.min returns the first minimum.
            float best = 3.0e38f;
            int best_i = 0;
            const int pix = py * W + cx;
            if (n_id > 0) {
                best = fmaf(nzv0, kTieNoise, idv0);
                if (n_id > 1) {
                    const float cand = fmaf(nzv1, kTieNoise, idv1);
                    if (cand < best) { best = cand; best_i = 1; }
                }
            }
            if (avg) {
                const float m = two ? (rp.x + rp.y) / 2.0f : rp.x;
                if (m < best) { best = m; best_i = n_id; }
                if (best_i == n_id) wgt = two ? splat(0.5f) : f2(1.f, 0.f);
            } else {
                if (rp.x < best) { best = rp.x; best_i = n_id; }
                if (two && rp.y < best) { best = rp.y; best_i = n_id + 1; }
                wgt = f2(best_i == n_id ? 1.f : 0.f, best_i == n_id + 1 ? 1.f : 0.f);
            }
            if (col_owned && py >= y0 && py < y1) {
                loss_acc += best;
                if (ps.argmin != nullptr) ps.argmin[bp + pix] = (uint8_t)best_i;
            }
        }

        // coefficients of window row r-1, consumed by the adjoint of the next step
        if (GRAD) {
            const float2 wsc = mul2(wgt, splat(p.inv_n));
            if (SSIM) {
                const float2 base = mul2(wsc, splat(kssim9));
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    const int c = m % 3;
                    // d rp / d x_q = (0.85/27) * (pa + x_q pb + y_q pe) for q in the window of p
                    cfc[m] = mul2(base, m < 3 ? pa[c] : (m < 6 ? pb[c] : pe[c]));
                }
            }
            wq2 = wq1;
            wq1 = wsc;
        }
        slotA = (slotA + 1) & 3;
    };
    // Always an even number of steps: a possible extra step past r_end only evaluates rows nobody
    // owns (no emission, do_q false), and keeps the unrolled loop free of a conditional join.
#pragma unroll 1
    for (int r = y0 - 2; r <= r_end; r += 2) {
        step(r, ynA, dnA, ynB, dnB, hy1, hy2, hyy1, hyy2, hx1, hx2, hxx1, hxx2, hxy1, hxy2, hc1, hc2);
        step(r + 1, ynB, dnB, ynA, dnA, hy2, hy1, hyy2, hyy1, hx2, hx1, hxx2, hxx1, hxy2, hxy1, hc2, hc1);
    }

    // ------------------------------------ epilogue ------------------------------------------------
    if (GRAD && kk > 1) {
        flush_row(cur, acc0);
        if (cur + 1 <= hd - 1) flush_row(cur + 1, acc1);
    }
    float* out = p.part + (size_t)item * p.part_stride;
    {
        const float v = warp_sum(loss_acc);
        if (lane == 0 && (mode == 0 || mode == 3)) out[0] = v;
    }
    if (GRAD) {
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const float a = warp_sum(gP[e].x), c = warp_sum(gP[e].y);
            if (lane == 0) {
                out[1 + 12 * fa + e] = a;
                if (two) out[1 + 12 * fb + e] = c;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// More than two source frames, default flags: the candidates that do not belong to the last frame pair
// (identity losses + tie-break noise, trainer.py:592-597, and the reprojection losses the mode-1 sweeps
// stored) are reduced here, one thread per pixel, to the first minimum in the reference's candidate order.
// The mode-3 sweep then reads one value and one index per pixel a row step ahead instead of
// S + f_base dependent loads at the point of use.  value -> rp slot f_base (unused by mode 1), index -> presel.
// ---------------------------------------------------------------------------------------------
// NF = f_base (frames swept before the last pair: 2, 4 or 6), S = NF + 1 or NF + 2.  One thread per pixel does all
// passes (scales): the identity candidates and their two smallest values are the same for every pass.
template <int NF>
__global__ void __launch_bounds__(256) select_prepass_kernel(const PhotoParams p) {
    const int plane = p.H * p.W;
    const int b = blockIdx.y;
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= plane) return;
    const int S = p.S;
    const float* iq = p.identity + (size_t)b * S * plane + pix;
    float idv[NF + 2];
#pragma unroll
    for (int i = 0; i < NF + 1; ++i) idv[i] = __ldg(at(iq, i * plane));
    idv[NF + 1] = (S > NF + 1) ? __ldg(at(iq, (NF + 1) * plane)) : 3.0e38f;
    float m1 = 3.0e38f, m2 = 3.0e38f;   // smallest and second smallest identity candidate
    int i1 = 0;
#pragma unroll
    for (int i = 0; i < NF + 2; ++i) {
        const float v = idv[i];
        m2 = fminf(m2, fmaxf(v, m1));
        if (v < m1) { m1 = v; i1 = i; }
    }
    const size_t fstride = (size_t)p.B * plane;     // frame stride of rp
    float* rq = p.rp + (size_t)b * plane + pix;
    uint8_t* sq = p.presel + (size_t)b * plane + pix;
#pragma unroll 1
    for (int pass_i = 0; pass_i < p.n_pass; ++pass_i, rq += S * fstride, sq += fstride) {
        // every load is issued before the first comparison: one round trip to memory per pass
        float rv[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) rv[f] = rq[f * fstride];
        float best = 3.0e38f;
        int best_i = 0;
#pragma unroll
        for (int f = 0; f < NF; ++f)
            if (rv[f] < best) { best = rv[f]; best_i = S + f; }
        // The in-kernel noise is bounded (|n| * 1e-5 <= 6.66e-5), so it can only decide between candidates closer than
        // 1.4e-4: two identity candidates, or the smallest of them and `best`.  Elsewhere the generator is skipped and
        // the tie-break noise is 0, exactly like the gate of the two-frame sweep (sweep_kernel, mode 0); against the
        // last pair's frames (compared in the mode-3 sweep) an unambiguous identity winner enters with noise 0 as well.
        if ((m1 - best < 1.4e-4f) && ((m2 - m1 < 1.4e-4f) || (best - m1 < 1.4e-4f))) {
            float ib = 3.0e38f;
            int ib_i = 0;
            const unsigned long long sd = p.seed_dev ? (p.seed ^ __ldg(p.seed_dev)) : p.seed;
#pragma unroll
            for (int i = 0; i < NF + 2; i += 2) {
                float n0, n1;
                philox2_normal2((uint32_t)sd ^ ((uint32_t)(sd >> 32) * 0x9E3779B9u), (uint32_t)(b * plane + pix),
                                (uint32_t)pass_i + 0x10000u * (uint32_t)(i >> 1), n0, n1);
                const float c0 = fmaf(n0, kTieNoise, idv[i]);
                if (c0 < ib) { ib = c0; ib_i = i; }
                const float c1 = fmaf(n1, kTieNoise, idv[i + 1]);     // a missing last candidate stays at 3e38
                if (c1 < ib) { ib = c1; ib_i = i + 1; }
            }
            if (ib <= best) { best = ib; best_i = ib_i; }   // identity candidates precede the reprojection ones
        } else if (m1 < best) {
            best = m1; best_i = i1;
        }
        rq[NF * fstride] = best;
        *sq = (uint8_t)best_i;
    }
}

// forward-only sweeps use neither the adjoint ring nor the upsample staging row
inline size_t sweep_smem_bytes(bool grad, bool pair = true) {
    const int ring = kSweepRingSlots * (pair ? kSweepRingQ : kSweepRingQ1) * 4 * 32;
    return (size_t)kSweepWarps * (grad ? 80 + ring : 80) * sizeof(float) + 16;
}

}  // namespace pml
