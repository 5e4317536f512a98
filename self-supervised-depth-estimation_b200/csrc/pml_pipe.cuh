// Fused view-synthesis + reprojection-loss sweep, third generation: the row step of pml_sweep.cuh
// cut into a three-stage, warp-specialised pipeline.
//
// Why: the single-warp sweep needs ~250 registers per thread (rolling 3x3 sums of the forward and
// of the adjoint, pose-gradient accumulators, 24 taps in flight), which caps an SM at 8 warps; the
// measured limiter is then not a pipe but issue latency (two warps per scheduler, ~45 % issue
// utilisation).  Here a CTA of three warps owns one strip (28 columns + halo) of one (scale, image)
// and each warp keeps only its own stage's state:
//
//   stage A "gather"  : disparity -> depth -> back-projection -> K.T projection of both source frames
//                       -> 24 bilinear taps -> warped colours + tap slopes          (trainer.py:469-511)
//   stage B "windows" : rolling 3x3 SSIM statistics, L1, identity automask + tie-break noise, per
//                       pixel min / argmin, loss partial, SSIM adjoint coefficients  (trainer.py:517-610)
//   stage C "adjoint" : transposed 3x3 sums, tap slopes -> (du, dv) -> pose / depth gradients,
//                       transposed disparity upsample                               (backward of all above)
//
// Rows flow A -> B -> C through shared-memory queues (lane-private 128-bit slots, conflict free)
// guarded by mbarriers (full / empty per slot, producer-consumer parity protocol); within a stage the
// horizontal neighbours still travel by warp shuffles.  Nothing is re-read from HBM and no block-wide
// barrier exists after the prologue.  grid = (n_chunks * n_strips, B, n_pass), 96 threads.
#pragma once
#include "pml_common.cuh"
#include "pml_photometric.cuh"
#include "pml_sweep.cuh"

namespace pml {

constexpr int kPipeNAB = 3;   // queue depths (rows)
constexpr int kPipeNAC = 6;
constexpr int kPipeNBC = 3;
constexpr int kPipeQAB = 3;   // float4 per lane per row
constexpr int kPipeQAC = 7;
constexpr int kPipeQBC = 5;
// shared memory (floats): P/IK 48 | staging row 32 | mbarriers 2 x (3 + 6 + 3) x 2 floats (64) | queues
constexpr int kPipeQueueFloats = (kPipeNAB * kPipeQAB + kPipeNAC * kPipeQAC + kPipeNBC * kPipeQBC) * 4 * 32;
constexpr int kPipeFloats = 48 + 32 + 64 + kPipeQueueFloats;

// ---- mbarrier primitives (shared::cta, phase parity protocol) -----------------------------------
#ifdef PML_HOST_EMU
struct EmuMbar { uint16_t count, pending; uint32_t phase; };
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    m->count = (uint16_t)count; m->pending = (uint16_t)count; m->phase = 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    if (--m->pending == 0) { m->pending = m->count; ++m->phase; }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    EmuMbar* m = reinterpret_cast<EmuMbar*>(bar);
    while ((m->phase & 1u) == parity) emu::yield_to_scheduler();
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "PML_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra PML_DONE;\n"
        " bra PML_WAIT;\n"
        "PML_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif
// producer side of one slot: all lanes have written -> make it visible -> one arrival
__device__ __forceinline__ void slot_publish(uint64_t* full, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(full);
}
// consumer side: all lanes have read their data into registers -> one arrival on `empty`
__device__ __forceinline__ void slot_release(uint64_t* empty, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(empty);
}

template <bool GRAD, bool SSIM>
__global__ void __launch_bounds__(96)
pipe_kernel(const PhotoParams p) {
    PML_DYN_SMEM(float, smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const bool lane_inner = (lane >= 1) && (lane <= 30);

    const bool automask = !(p.flags & PML_FLAG_NO_AUTOMASK);
    const bool avg = (p.flags & PML_FLAG_AVG_REPROJ) != 0;
    const int n_id = automask ? (avg ? 1 : S) : 0;
    const int f1 = (S > 1) ? 1 : 0;
    const int plane = H * W;
    const int b3p = b * 3 * plane, bp = b * plane;
    const int kk = ps.k, wd = ps.wd, hd = ps.hd;
    const int bdp = b * hd * wd;
    const float rscale = ps.rscale;

    // ---- shared memory ---------------------------------------------------------------------------
    const float4* sP4 = reinterpret_cast<const float4*>(smem);          // 12 x (frame0, frame1) of P = (K T)[:3]
    const float4* sIK4 = reinterpret_cast<const float4*>(smem + 24);    // inv_K entries used per row
    float* sG = smem + 48;                                               // staging row (transposed upsample)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 80);
    uint64_t* fullAB = bars, *emptyAB = fullAB + kPipeNAB;
    uint64_t* fullAC = emptyAB + kPipeNAB, *emptyAC = fullAC + kPipeNAC;
    uint64_t* fullBC = emptyAC + kPipeNAC, *emptyBC = fullBC + kPipeNBC;
    float4* qAB = reinterpret_cast<float4*>(smem + 144);
    float4* qAC = qAB + kPipeNAB * kPipeQAB * 32;
    float4* qBC = qAC + kPipeNAC * kPipeQAC * 32;

    if (tid < 24) {
        const int e = tid >> 1, f = (tid & 1) ? f1 : 0, i = e >> 2, j = e & 3;
        const float* Kb = p.K + b * 16;
        const float* Tb = p.T[f] + b * 16;
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) a = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], a);   // layers.py:183
        smem[tid] = a;
    }
    const float* ikb = p.invK + b * 16;                                          // layers.py:164
    if (tid == 32) {
        smem[24] = ikb[1]; smem[25] = ikb[2]; smem[26] = ikb[5]; smem[27] = ikb[6];
        smem[28] = ikb[9]; smem[29] = ikb[10]; smem[30] = 0.f; smem[31] = 0.f;
    }
    if (tid == 64) {
        for (int i = 0; i < 2 * (kPipeNAB + kPipeNAC + kPipeNBC); ++i) mbar_init(bars + i, 1);
    }
    __syncthreads();   // the only block-wide barrier

    const float fxc = (float)rx;
    const float rc0 = ikb[0] * fxc, rc1 = ikb[4] * fxc, rc2 = ikb[8] * fxc;   // column part of the ray
    const int r_begin = y0 - 2, r_end = GRAD ? (y1 + 1) : y1;               // rows warped (inclusive)
    float* out = p.part + (size_t)item * p.part_stride;

    if (warp == 0) {
        // ======================================= stage A ==========================================
        int j0 = rx, j1 = rx;
        float lam = 0.f;
        if (kk > 1) {   // horizontal part of the disparity upsample (trainer.py:474): fixed per lane
            const float sx = fmaxf(fmaf(rscale, (float)rx + 0.5f, -0.5f), 0.f);
            j0 = (int)sx;
            j1 = min(j0 + 1, wd - 1);
            lam = sx - (float)j0;
        }
        const float* __restrict__ tgt_g = p.target;
        const float* __restrict__ src0_g = p.src[0];
        const float* __restrict__ src1_g = p.src[f1];
        const float* __restrict__ disp_g = ps.disp;
        const float wscale = (float)W / (float)(W - 1), hscale = (float)H / (float)(H - 1);
        const float wmax = (float)(W - 1), hmax = (float)(H - 1);
        const float wmax1 = (float)(W - 2), hmax1 = (float)(H - 2);
        const bool emit_any = (ps.depth != nullptr) || (ps.warped != nullptr);

        // per-row geometry, then the 24 taps; the taps of row r+1 are issued before row r is
        // interpolated so that a full row of arithmetic hides their latency
        struct Geo { float yv[3]; float D; float2 invz, u, v, tx, ty, mx, my; };
        struct Taps { float2 nw[3], ne[3], sw[3], se[3]; };
        auto issue = [&](int r, Geo& g, Taps& t) {
            const int ry = reflect1(clampi(r, -1, H), H);
            const float* tq = at(tgt_g, b3p + ry * W + rx);
            g.yv[0] = __ldg(tq); g.yv[1] = __ldg(at(tq, plane)); g.yv[2] = __ldg(at(tq, 2 * plane));
            float d;
            if (kk > 1) {   // bilinear upsample of disp_s, align_corners=False (trainer.py:474)
                const float sy = fmaxf(fmaf(rscale, (float)ry + 0.5f, -0.5f), 0.f);
                const int i0 = (int)sy, i1 = min(i0 + 1, hd - 1);
                const float mu = sy - (float)i0;
                const float* d0 = at(disp_g, bdp + i0 * wd);
                const float* d1 = at(disp_g, bdp + i1 * wd);
                const float v00 = __ldg(at(d0, j0)), v01 = __ldg(at(d0, j1)), v10 = __ldg(at(d1, j0)), v11 = __ldg(at(d1, j1));
                const float top = fmaf(lam, v01, (1.f - lam) * v00), bot = fmaf(lam, v11, (1.f - lam) * v10);
                d = fmaf(mu, bot, (1.f - mu) * top);
            } else {
                d = __ldg(at(disp_g, bdp + ry * W + rx));
            }
            const float sigma = fmaf(p.disp_range, d, p.min_disp);   // layers.py:23
            const float D = rcp_nr(sigma);                            // layers.py:24
            g.D = D;
            const float fy = (float)ry;
            const float4 ika = sIK4[0], ikc = sIK4[1];
            const float X0 = D * (rc0 + fmaf(ika.x, fy, ika.y));      // layers.py:165
            const float X1 = D * (rc1 + fmaf(ika.z, fy, ika.w));
            const float X2 = D * (rc2 + fmaf(ikc.x, fy, ikc.y));
            const float4 pa0 = sP4[0], pa1 = sP4[1], pb0 = sP4[2], pb1 = sP4[3], pc0 = sP4[4], pc1 = sP4[5];
            const float2 c0 = fma2(f2(pa0.x, pa0.y), splat(X0), fma2(f2(pa0.z, pa0.w), splat(X1), fma2(f2(pa1.x, pa1.y), splat(X2), f2(pa1.z, pa1.w))));
            const float2 c1 = fma2(f2(pb0.x, pb0.y), splat(X0), fma2(f2(pb0.z, pb0.w), splat(X1), fma2(f2(pb1.x, pb1.y), splat(X2), f2(pb1.z, pb1.w))));
            const float2 c2 = fma2(f2(pc0.x, pc0.y), splat(X0), fma2(f2(pc0.z, pc0.w), splat(X1), fma2(f2(pc1.x, pc1.y), splat(X2), f2(pc1.z, pc1.w))));
            g.invz = rcp_nr2(add2(c2, splat(p.eps)));                 // layers.py:187
            g.u = mul2(c0, g.invz); g.v = mul2(c1, g.invz);
            // layers.py:190-192 + grid_sample unnormalise (align_corners=False): ix = u*W/(W-1) - 0.5
            const float2 ixr = fma2(g.u, splat(wscale), splat(-0.5f)), iyr = fma2(g.v, splat(hscale), splat(-0.5f));
            const float2 ix = f2(fminf(fmaxf(ixr.x, 0.f), wmax), fminf(fmaxf(ixr.y, 0.f), wmax));
            const float2 iy = f2(fminf(fmaxf(iyr.x, 0.f), hmax), fminf(fmaxf(iyr.y, 0.f), hmax));
            // base tap clamped to W-2 / H-2: the +1 tap always exists and carries the weight that
            // grid_sample gives to the last column / row (its out-of-range tap has weight 0)
            const float2 fx0 = f2(fminf(floorf(ix.x), wmax1), fminf(floorf(ix.y), wmax1));
            const float2 fy0 = f2(fminf(floorf(iy.x), hmax1), fminf(floorf(iy.y), hmax1));
            g.tx = sub2(ix, fx0); g.ty = sub2(iy, fy0);
            // clip backward (zero outside the open interval) times d ix / d u
            g.mx = f2((ixr.x > 0.f && ixr.x < wmax) ? wscale : 0.f, (ixr.y > 0.f && ixr.y < wmax) ? wscale : 0.f);
            g.my = f2((iyr.x > 0.f && iyr.x < hmax) ? hscale : 0.f, (iyr.y > 0.f && iyr.y < hmax) ? hscale : 0.f);
            const int o0 = b3p + (int)fy0.x * W + (int)fx0.x, o1 = b3p + (int)fy0.y * W + (int)fx0.y;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* a0 = at(src0_g, o0 + c * plane);
                const float* a1 = at(src1_g, o1 + c * plane);
                const float* w0 = at(src0_g, o0 + (c * plane + W));
                const float* w1 = at(src1_g, o1 + (c * plane + W));
                t.nw[c] = f2(__ldg(a0), __ldg(a1)); t.ne[c] = f2(__ldg(a0 + 1), __ldg(a1 + 1));
                t.sw[c] = f2(__ldg(w0), __ldg(w1)); t.se[c] = f2(__ldg(w0 + 1), __ldg(w1 + 1));
            }
        };
        auto finish = [&](int r, const Geo& g, const Taps& t) {
            float2 xv[3], dpx[3], dpy[3];
            float2 l1 = splat(0.f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 dt = sub2(t.ne[c], t.nw[c]), db = sub2(t.se[c], t.sw[c]);
                const float2 top = fma2(g.tx, dt, t.nw[c]), bot = fma2(g.tx, db, t.sw[c]);
                const float2 dvert = sub2(bot, top);
                xv[c] = fma2(g.ty, dvert, top);
                if (GRAD) {
                    dpx[c] = mul2(g.mx, fma2(g.ty, sub2(db, dt), dt));
                    dpy[c] = mul2(g.my, dvert);
                }
                const float2 df = sub2(xv[c], splat(g.yv[c]));
                l1.x += fabsf(df.x);
                l1.y += fabsf(df.y);
            }
            const bool row_owned = (r >= y0) && (r < y1);
            if (emit_any && col_owned && row_owned) {   // trainer.py:480, :508 (on request)
                const int o = r * W + cx;
                if (ps.depth != nullptr) ps.depth[bp + o] = g.D;
                if (ps.warped != nullptr) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        ps.warped[b3p + c * plane + o] = xv[c].x;
                        if (S > 1) ps.warped[(size_t)p.B * 3 * plane + b3p + c * plane + o] = xv[c].y;
                    }
                }
            }
            // ---- queue A -> B: colours of row r (target, warped pair) and its L1 sum ----
            const int kab = r - r_begin, sab = kab % kPipeNAB;
            mbar_wait(emptyAB + sab, ((kab / kPipeNAB) & 1) ^ 1);
            float4* wab = qAB + (sab * kPipeQAB) * 32 + lane;
            wab[0 * 32] = make_float4(g.yv[0], g.yv[1], g.yv[2], 0.f);
            wab[1 * 32] = make_float4(l1.x, l1.y, xv[0].x, xv[0].y);
            wab[2 * 32] = make_float4(xv[1].x, xv[1].y, xv[2].x, xv[2].y);
            slot_publish(fullAB + sab, lane);
            // ---- queue A -> C: geometry of the owned rows, for the adjoint ----
            if (GRAD && row_owned) {
                const int kac = r - y0, sac = kac % kPipeNAC;
                mbar_wait(emptyAC + sac, ((kac / kPipeNAC) & 1) ^ 1);
                float4* wac = qAC + (sac * kPipeQAC) * 32 + lane;
                wac[0 * 32] = make_float4(g.yv[0], g.yv[1], g.yv[2], g.D);
                wac[1 * 32] = make_float4(xv[0].x, xv[0].y, xv[1].x, xv[1].y);
                wac[2 * 32] = make_float4(xv[2].x, xv[2].y, dpx[0].x, dpx[0].y);
                wac[3 * 32] = make_float4(dpx[1].x, dpx[1].y, dpx[2].x, dpx[2].y);
                wac[4 * 32] = make_float4(dpy[0].x, dpy[0].y, dpy[1].x, dpy[1].y);
                wac[5 * 32] = make_float4(dpy[2].x, dpy[2].y, g.invz.x, g.invz.y);
                wac[6 * 32] = make_float4(g.u.x, g.u.y, g.v.x, g.v.y);
                slot_publish(fullAC + sac, lane);
            }
        };
        Geo gA, gB;
        Taps tA, tB;
        issue(r_begin, gA, tA);
#pragma unroll 1
        for (int r = r_begin; r <= r_end; r += 2) {
            issue(r + 1, gB, tB);            // rows past r_end are clamped by reflect1 / clampi: harmless
            finish(r, gA, tA);
            if (r + 1 <= r_end) {
                issue(r + 2, gA, tA);
                finish(r + 1, gB, tB);
            }
        }
    } else if (warp == 1) {
        // ======================================= stage B ==========================================
        const float* __restrict__ id_g = p.identity;
        const float* __restrict__ nz_g = ps.noise;
        const int bip = b * n_id * plane;
        float hy1[3], hy2[3], hyy1[3], hyy2[3];
        float2 hx1[3], hx2[3], hxx1[3], hxx2[3], hxy1[3], hxy2[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            hy1[c] = hy2[c] = hyy1[c] = hyy2[c] = 0.f;
            hx1[c] = hx2[c] = hxx1[c] = hxx2[c] = hxy1[c] = hxy2[c] = splat(0.f);
        }
        float loss_acc = 0.f;
        float2 l1_prev = splat(0.f);
        const float kssim9 = SSIM ? (0.85f / 27.0f) : 0.f;      // 0.85 / 3 channels / 9 window taps
        // identity losses / noise of window row r-1, fetched one row step early
        float idA[2], nzA[2], idB[2], nzB[2];
        auto load_id = [&](int r, float (&idv)[2], float (&nzv)[2]) {
            idv[0] = idv[1] = nzv[0] = nzv[1] = 0.f;
            if (n_id > 0) {
                const int o = bip + clampi(r - 1, 0, H - 1) * W + rx;
                idv[0] = __ldg(at(id_g, o));
                if (n_id > 1) idv[1] = __ldg(at(id_g, o + plane));
                if (nz_g != nullptr) {
                    nzv[0] = __ldg(at(nz_g, o));
                    if (n_id > 1) nzv[1] = __ldg(at(nz_g, o + plane));
                }
            }
        };
        auto step = [&](const int r, const float (&idv)[2], const float (&nzv)[2], float (&idN)[2], float (&nzN)[2],
                        float (&hyA)[3], float (&hyB)[3], float (&hyyA)[3], float (&hyyB)[3],
                        float2 (&hxA)[3], float2 (&hxB)[3], float2 (&hxxA)[3], float2 (&hxxB)[3],
                        float2 (&hxyA)[3], float2 (&hxyB)[3]) {
            load_id(r + 1, idN, nzN);
            const int kab = r - r_begin, sab = kab % kPipeNAB;
            mbar_wait(fullAB + sab, (kab / kPipeNAB) & 1);
            const float4* rab = qAB + (sab * kPipeQAB) * 32 + lane;
            const float4 a0 = rab[0 * 32], a1 = rab[1 * 32], a2 = rab[2 * 32];
            slot_release(emptyAB + sab, lane);
            const float yv[3] = {a0.x, a0.y, a0.z};
            const float2 l1_cur = f2(a1.x, a1.y);
            const float2 xv[3] = {f2(a1.z, a1.w), f2(a2.x, a2.y), f2(a2.z, a2.w)};

            const int py = r - 1;
            const bool p_valid = (r >= y0) && (py >= 0) && (py < H) && col_in_image && lane_inner;
            float2 rp;
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
                        ssim_sum = add2(ssim_sum, ssim_pair<GRAD>(Sx, Sxx, Sxy, my_, myy + kSsimC1, sy_ + kSsimC2,
                                                                  pa[c], pb[c], pe[c]));
                    }
                }
                // trainer.py:527 (0.85 * SSIM.mean(1) + 0.15 * L1.mean(1)) or :523 (L1 only)
                rp = SSIM ? fma2(ssim_sum, splat(0.85f / 3.0f), mul2(l1_prev, splat(0.15f / 3.0f)))
                          : mul2(l1_prev, splat(1.0f / 3.0f));
            }
            l1_prev = l1_cur;

            // candidates in the reference's order: identity (+noise) first, then reprojection
            // (trainer.py:597); torch.min returns the first minimum.  Straight-line for all lanes,
            // masked by p_valid at the end.
            float best = 3.0e38f;
            int best_i = 0;
            const int pix = py * W + cx;
            if (n_id > 0) {
                float n0 = nzv[0], n1 = nzv[1];
                if (nz_g == nullptr)
                    philox2_normal2((uint32_t)p.seed ^ ((uint32_t)(p.seed >> 32) * 0x9E3779B9u),
                                    (uint32_t)(bp + pix), (uint32_t)pass_i, n0, n1);
                best = fmaf(n0, kTieNoise, idv[0]);
                if (n_id > 1) {
                    const float cand = fmaf(n1, kTieNoise, idv[1]);
                    if (cand < best) { best = cand; best_i = 1; }
                }
            }
            float2 wgt;
            if (avg) {
                const float m = (S > 1) ? (rp.x + rp.y) / 2.0f : rp.x;
                if (m < best) { best = m; best_i = n_id; }
                wgt = (best_i == n_id) ? ((S > 1) ? splat(0.5f) : f2(1.f, 0.f)) : splat(0.f);
            } else {
                if (rp.x < best) { best = rp.x; best_i = n_id; }
                if (S > 1 && rp.y < best) { best = rp.y; best_i = n_id + 1; }
                wgt = f2(best_i == n_id ? 1.f : 0.f, best_i == n_id + 1 ? 1.f : 0.f);
            }
            if (!p_valid) wgt = splat(0.f);
            if (p_valid && col_owned && py >= y0 && py < y1) {
                loss_acc += best;
                if (ps.argmin != nullptr) ps.argmin[bp + pix] = (uint8_t)best_i;
            }
            // ---- queue B -> C: adjoint coefficients of window row p = r-1 (p in [y0-1, y1]) ----
            if (GRAD && r >= y0) {
                const float2 wsc = mul2(wgt, splat(p.inv_n));
                const float2 base = mul2(wsc, splat(kssim9));
                const int kbc = r - y0, sbc = kbc % kPipeNBC;
                mbar_wait(emptyBC + sbc, ((kbc / kPipeNBC) & 1) ^ 1);
                float4* wbc = qBC + (sbc * kPipeQBC) * 32 + lane;
                float2 cf[9];
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    const int c = m % 3;
                    // d rp / d x_q = (0.85/27) * (pa + x_q pb + y_q pe) for q in the window of p
                    cf[m] = SSIM ? mul2(base, m < 3 ? pa[c] : (m < 6 ? pb[c] : pe[c])) : splat(0.f);
                }
                wbc[0 * 32] = make_float4(cf[0].x, cf[0].y, cf[1].x, cf[1].y);
                wbc[1 * 32] = make_float4(cf[2].x, cf[2].y, cf[3].x, cf[3].y);
                wbc[2 * 32] = make_float4(cf[4].x, cf[4].y, cf[5].x, cf[5].y);
                wbc[3 * 32] = make_float4(cf[6].x, cf[6].y, cf[7].x, cf[7].y);
                wbc[4 * 32] = make_float4(cf[8].x, cf[8].y, wsc.x, wsc.y);
                slot_publish(fullBC + sbc, lane);
            }
        };
        load_id(r_begin, idA, nzA);
#pragma unroll 1
        for (int r = r_begin; r <= r_end; r += 2) {
            step(r, idA, nzA, idB, nzB, hy1, hy2, hyy1, hyy2, hx1, hx2, hxx1, hxx2, hxy1, hxy2);
            if (r + 1 <= r_end)
                step(r + 1, idB, nzB, idA, nzA, hy2, hy1, hyy2, hyy1, hx2, hx1, hxx2, hxx1, hxy2, hxy1);
            else { idA[0] = idB[0]; idA[1] = idB[1]; nzA[0] = nzB[0]; nzA[1] = nzB[1]; }
        }
        const float v = warp_sum(loss_acc);
        if (lane == 0) out[0] = v;
    } else if (GRAD) {
        // ======================================= stage C ==========================================
        float2 hc1[9], hc2[9];
        float2 gP[12];
#pragma unroll
        for (int m = 0; m < 9; ++m) hc1[m] = hc2[m] = splat(0.f);
#pragma unroll
        for (int e = 0; e < 12; ++e) gP[e] = splat(0.f);
        float2 wq_prev = splat(0.f);
        float acc0 = 0.f, acc1 = 0.f;
        int cur = 0, jbase = 0;
        if (kk > 1) {
            cur = (int)fmaxf(fmaf(rscale, (float)y0 + 0.5f, -0.5f), 0.f);
            jbase = (int)fmaxf(fmaf(rscale, (float)x0 + 0.5f, -0.5f), 0.f);
        }
        const int low_cols = kSweepTW / kk + 3;
        const float wl = (cx == 1) ? 2.f : 1.f, wr = (cx == W - 2) ? 2.f : 1.f;   // reflection fold (columns)
        const float kl1 = SSIM ? (0.15f / 3.0f) : (1.0f / 3.0f);
        auto flush_row = [&](int irow, float h) {
            sweep_flush_row(sG, ps.grad_disp + (bdp + irow * wd), h, lane, jbase, low_cols, wd, kk, x0, x1, rscale);
        };
        // window rows p = y0-1 .. y1; pixel row q = p-1 is complete once p has arrived
        auto step = [&](const int pw, float2 (&hcA)[9], float2 (&hcB)[9]) {
            const int kbc = pw - (y0 - 1), sbc = kbc % kPipeNBC;
            mbar_wait(fullBC + sbc, (kbc / kPipeNBC) & 1);
            const float4* rbc = qBC + (sbc * kPipeQBC) * 32 + lane;
            const float4 b0 = rbc[0 * 32], b1 = rbc[1 * 32], b2 = rbc[2 * 32], b3 = rbc[3 * 32], b4 = rbc[4 * 32];
            slot_release(emptyBC + sbc, lane);
            const float2 cf[9] = {f2(b0.x, b0.y), f2(b0.z, b0.w), f2(b1.x, b1.y), f2(b1.z, b1.w), f2(b2.x, b2.y),
                                  f2(b2.z, b2.w), f2(b3.x, b3.y), f2(b3.z, b3.w), f2(b4.x, b4.y)};
            const float2 wsc = f2(b4.z, b4.w);
            const int qy = pw - 1;
            const bool do_q = (qy >= y0) && (qy < y1);
            const float wt = (qy == 1) ? 2.f : 1.f, wb = (qy == H - 2) ? 2.f : 1.f;   // reflection fold (rows)
            float2 V[9];
            if (SSIM) {
#pragma unroll
                for (int m = 0; m < 9; ++m) {
                    const float2 cl = shfl_up2(cf[m]), cr = shfl_down2(cf[m]);
                    const float2 hn = fma2(splat(wl), cl, fma2(splat(wr), cr, cf[m]));
                    V[m] = fma2(splat(wt), hcB[m], fma2(splat(wb), hn, hcA[m]));
                    hcB[m] = hn;
                }
            }
            float g_d = 0.f;
            if (do_q) {   // warp-uniform
                const int kac = qy - y0, sac = kac % kPipeNAC;
                mbar_wait(fullAC + sac, (kac / kPipeNAC) & 1);
                const float4* rc = qAC + (sac * kPipeQAC) * 32 + lane;
                const float4 q0r = rc[0 * 32], q1r = rc[1 * 32], q2r = rc[2 * 32], q3r = rc[3 * 32];
                const float4 q4r = rc[4 * 32], q5r = rc[5 * 32], q6r = rc[6 * 32];
                slot_release(emptyAC + sac, lane);
                const float yq[3] = {q0r.x, q0r.y, q0r.z};
                const float Dq = q0r.w;
                const float2 xq[3] = {f2(q1r.x, q1r.y), f2(q1r.z, q1r.w), f2(q2r.x, q2r.y)};
                const float2 dxq[3] = {f2(q2r.z, q2r.w), f2(q3r.x, q3r.y), f2(q3r.z, q3r.w)};
                const float2 dyq[3] = {f2(q4r.x, q4r.y), f2(q4r.z, q4r.w), f2(q5r.x, q5r.y)};
                const float2 invzq = f2(q5r.z, q5r.w), uq = f2(q6r.x, q6r.y), vq = f2(q6r.z, q6r.w);
                const float2 kw = mul2(wq_prev, splat(kl1));   // wq_prev: winner weight of row q
                float2 du = splat(0.f), dv = splat(0.f);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float2 df = sub2(xq[c], splat(yq[c]));
                    // d|x - y| / dx = sign(x - y); sign(0) taken as +1 (see pml_sweep.cuh)
                    float2 g = f2(copysignf(kw.x, df.x), copysignf(kw.y, df.y));
                    if (SSIM) g = add2(g, fma2(xq[c], V[3 + c], fma2(splat(yq[c]), V[6 + c], V[c])));
                    du = fma2(g, dxq[c], du);
                    dv = fma2(g, dyq[c], dv);
                }
                du = f2(col_owned ? du.x : 0.f, col_owned ? du.y : 0.f);
                dv = f2(col_owned ? dv.x : 0.f, col_owned ? dv.y : 0.f);
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
                g_d = col_owned ? -p.disp_range * Dq * Dq * (gD.x + gD.y) : 0.f;   // d(1/sigma)/d disp, both frames
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
            wq_prev = wsc;
        };
#pragma unroll 1
        for (int pw = y0 - 1; pw <= y1; pw += 2) {
            step(pw, hc1, hc2);
            if (pw + 1 <= y1) step(pw + 1, hc2, hc1);
        }
        if (kk > 1) {
            flush_row(cur, acc0);
            if (cur + 1 <= hd - 1) flush_row(cur + 1, acc1);
        }
#pragma unroll
        for (int e = 0; e < 12; ++e) {
            const float a = warp_sum(gP[e].x), c = warp_sum(gP[e].y);
            if (lane == 0) {
                out[1 + e] = a;
                if (S > 1) out[1 + 12 + e] = c;
            }
        }
    }
}

inline size_t pipe_smem_bytes() { return (size_t)kPipeFloats * sizeof(float) + 16; }

}  // namespace pml
