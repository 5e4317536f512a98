// Fused view-synthesis + reprojection-loss kernels (the hot path).
//
// Restates, per scale, trainer.py:469-515 (disp upsample, disp_to_depth, BackprojectDepth,
// Project3D, grid_sample border/bilinear) and trainer.py:546-610 (SSIM+L1 reprojection loss,
// identity automask with tie-break noise, per-pixel min, mean) in ONE sweep that also evaluates
// the adjoint (d loss / d disp_s, d loss / d T_f) while the rows are on chip.  cam_points,
// pix_coords and warped images never exist in HBM.
//
// Work decomposition ("rolling strips"): a CTA owns a vertical strip of TW = blockDim.x-4 image
// columns x TH rows of one (scale, image).  Thread t owns column x0-2+t and the CTA marches down
// the rows.  At row step r it (A) warps row r of every source frame, (B) closes the 3x3 SSIM
// windows centred on row r-1 (vertical sums roll in registers, horizontal neighbours come from
// a shared-memory row), takes the per-pixel min/argmin and emits the SSIM adjoint coefficients,
// (C) closes the transposed 3x3 sums centred on row r-2 and turns them into pose / disparity
// gradients.  Only 4 halo rows and 4 halo columns per strip are recomputed.
#pragma once
#include "pml_common.cuh"

namespace pml {

struct PassDev {
    const float* disp;
    const float* noise;
    uint8_t* argmin;
    float* depth;
    float* warped;
    float* grad_disp;
    const float* fw;   // [B,S,H,W] per-frame weights of the reprojection losses (predictive mask) or null
    float* gfw;        // [B,S,H,W] out: d loss / d fw (photometric part) or null
    int hd, wd, k;   // k = H / hd (1, 2, 4, 8, ...)
    float rscale;    // hd / H : ATen area_pixel_compute_scale for align_corners=False
    int low_cols, low_rows;  // extent of the per-CTA low-res accumulator (k > 1)
};

struct PhotoParams {
    int B, H, W, n_pass;
    unsigned flags;
    float min_disp, disp_range, eps;
    unsigned long long seed;
    const unsigned long long* seed_dev;   // optional: xor'ed into seed on the device (CUDA-graph replays with fresh noise)
    const float* target;
    const float* src[PML_MAX_SOURCES];
    const float* K;
    const float* invK;
    const float* T[PML_MAX_SOURCES];
    const float* identity;  // [B, n_id, H, W] identity reprojection losses (automask)
    PassDev pass[PML_MAX_PASSES];
    int TW, TH, n_strips, n_chunks, cta_per_pass;   // cta_per_pass: work items (CTAs or warps) per pass
    int S;            // number of source frames (run-time copy of the template parameter)
    int n_items;      // total work items (pml_sweep.cuh: one per warp)
    // pml_sweep.cuh with more than two source frames: the frames are swept pair by pair
    int mode;         // 0: selection in the sweep (S <= 2); 1: reprojection losses only -> rp;
                      // 2: selection given by pass.argmin (written by the mode-3 sweep);
                      // 3: last pair: selection over identity + rp of earlier pairs + own pair
    int f_base;       // first frame of the pair handled by this launch
    int pair_n;       // frames in the pair: 1 or 2
    float* rp;        // [n_pass][S][B][H][W] reprojection losses (written in mode 1, read in mode 3)
    uint8_t* presel;  // [n_pass][B][H][W] select_prepass_kernel: best candidate ahead of the last pair (its value: rp slot f_base)
    float* part;      // [n_cta][part_stride]: loss partial, then S x 12 dL/dP partials
    int part_stride;
    float inv_n;      // 1 / (B*H*W)
    // chunked batch (pml_segments), honoured by the warp-strip sweep only
    int n_seg, seg_size;
    ChunkPtrs target_c, K_c, invK_c, src_c[PML_MAX_SOURCES];
};

constexpr int kGeoFields = 13;  // per frame in the row ring: x[3] dpx[3] dpy[3] invz u v wgt

template <int S>
__host__ __device__ constexpr int ring_slot_floats() { return 4 + kGeoFields * S; }  // y[3], D, frames

// ---------------------------------------------------------------------------------------------
// identity reprojection loss (trainer.py:556-569): reprojection_loss(source_f, target), once per
// step, shared by every scale that uses the same images.
// ---------------------------------------------------------------------------------------------
template <bool SSIM>
__global__ void __launch_bounds__(256)
identity_kernel(const float* __restrict__ target,
                const float* s0, const float* s1, const float* s2, const float* s3,
                float* __restrict__ out, int B, int H, int W, int S, int avg) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    int b = blockIdx.z;
    if (x >= W) return;
    const size_t plane = (size_t)H * W;
    const float* srcs[4] = {s0, s1, s2, s3};
    int xs[3] = {reflect1(x - 1, W), x, reflect1(x + 1, W)};
    int ys[3] = {reflect1(y - 1, H), y, reflect1(y + 1, H)};
    float acc = 0.f;
    for (int f = 0; f < S; ++f) {
        float ssim_sum = 0.f, l1_sum = 0.f;
        for (int c = 0; c < 3; ++c) {
            const float* tp = target + ((size_t)b * 3 + c) * plane;
            const float* sp = srcs[f] + ((size_t)b * 3 + c) * plane;
            float yc = __ldg(tp + (size_t)y * W + x), xc = __ldg(sp + (size_t)y * W + x);
            l1_sum += fabsf(yc - xc);
            if (SSIM) {
                float Sx = 0, Sy = 0, Sxx = 0, Syy = 0, Sxy = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        float xv = __ldg(sp + (size_t)ys[j] * W + xs[i]);
                        float yv = __ldg(tp + (size_t)ys[j] * W + xs[i]);
                        Sx += xv; Sy += yv;
                        Sxx = fmaf(xv, xv, Sxx); Syy = fmaf(yv, yv, Syy); Sxy = fmaf(xv, yv, Sxy);
                    }
                float d0, d1, d2;
                ssim_sum += ssim_window<false>(Sx, Sy, Sxx, Syy, Sxy, d0, d1, d2);
            }
        }
        float rp = SSIM ? fmaf(0.85f, ssim_sum * (1.0f / 3.0f), 0.15f * (l1_sum * (1.0f / 3.0f)))
                        : l1_sum * (1.0f / 3.0f);
        if (avg) acc += rp;
        else out[(((size_t)b * S + f) * H + y) * W + x] = rp;
    }
    if (avg) out[((size_t)b * H + y) * W + x] = acc / (float)S;
}

// ---------------------------------------------------------------------------------------------
// main rolling-strip kernel
// ---------------------------------------------------------------------------------------------
template <int S, bool GRAD, bool SSIM>
__global__ void __launch_bounds__(128)
photometric_kernel(const PhotoParams p) {
    PML_DYN_SMEM(float, smem);
    const int NT = blockDim.x;
    const int tid = threadIdx.x;

    // ---- which strip -------------------------------------------------------------------
    int wi = blockIdx.x;
    const int pass_i = wi / p.cta_per_pass;
    wi -= pass_i * p.cta_per_pass;
    const int b = wi / (p.n_chunks * p.n_strips);
    wi -= b * (p.n_chunks * p.n_strips);
    const int chunk = wi / p.n_strips;
    const int strip = wi - chunk * p.n_strips;
    const PassDev& ps = p.pass[pass_i];

    const int H = p.H, W = p.W;
    const int x0 = strip * p.TW, x1 = min(x0 + p.TW, W);
    const int y0 = chunk * p.TH, y1 = min(y0 + p.TH, H);
    const int cx = x0 - 2 + tid;                         // column on the reflect-padded grid
    const int rx = reflect1(clampi(cx, -1, W), W);       // image column actually evaluated
    const bool col_in_image = (cx >= 0) && (cx < W);
    const bool col_owned = (cx >= x0) && (cx < x1);
    const int tl = max(tid - 1, 0), tr = min(tid + 1, NT - 1);

    const bool automask = !(p.flags & PML_FLAG_NO_AUTOMASK);
    const bool avg = (p.flags & PML_FLAG_AVG_REPROJ) != 0;
    const int n_id = automask ? (avg ? 1 : S) : 0;

    // ---- shared memory carve-up ----------------------------------------------------------
    constexpr int RS = ring_slot_floats<S>();
    float* sP = smem;                        // [S][12]
    float* sIK = sP + S * 12;                // [9] (+3 pad)
    float* sY = sIK + 12;                    // [3][NT]
    float* sX = sY + 3 * NT;                 // [S][3][NT]
    float* sC = sX + S * 3 * NT;             // [S][9][NT]   (GRAD)
    float* sRing = sC + (GRAD ? S * 9 * NT : 0);      // [3][RS][NT]
    float* sG = sRing + 3 * RS * NT;         // [NT]
    float* sLam = sG + NT;                   // [NT]
    int* sJ0 = reinterpret_cast<int*>(sLam + NT);      // [NT]
    float* sAcc = reinterpret_cast<float*>(sJ0 + NT);  // [low_rows][low_cols]  (GRAD, k>1)
    float* sRed = sAcc + ((GRAD && ps.k > 1) ? ps.low_rows * ps.low_cols : 0);  // [NT/32][1+12S]

    // P = (K @ T)[:3, :]  (layers.py:183) and the 3x3 of inv_K (layers.py:164)
    if (tid < S * 12) {
        int f = tid / 12, e = tid - f * 12, i = e >> 2, j = e & 3;
        const float* Kb = p.K + (size_t)b * 16;
        const float* Tb = p.T[f] + (size_t)b * 16;
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) a = fmaf(Kb[i * 4 + k], Tb[k * 4 + j], a);
        sP[tid] = a;
    }
    if (tid >= 64 - 9 && tid < 64) {  // NT >= 64 always
        int e = tid - (64 - 9);
        sIK[e] = p.invK[(size_t)b * 16 + (e / 3) * 4 + (e % 3)];
    }
    if (GRAD && ps.k > 1) {
        for (int i = tid; i < ps.low_rows * ps.low_cols; i += NT) sAcc[i] = 0.f;
    }
    __syncthreads();

    float P[S][12];
#pragma unroll
    for (int f = 0; f < S; ++f)
#pragma unroll
        for (int e = 0; e < 12; ++e) P[f][e] = sP[f * 12 + e];
    // back-projection ray r = inv_K[:3,:3] @ (x, y, 1): column part is fixed per thread
    const float fxc = (float)rx;
    const float rc0 = sIK[0] * fxc, rc1 = sIK[3] * fxc, rc2 = sIK[6] * fxc;

    // horizontal part of the disparity upsample (trainer.py:474): fixed per thread
    int j0 = rx, j1 = rx;
    float lam = 0.f;
    if (ps.k > 1) {
        float sx = fmaxf(fmaf(ps.rscale, (float)rx + 0.5f, -0.5f), 0.f);
        j0 = (int)sx;
        j1 = min(j0 + 1, ps.wd - 1);
        lam = sx - (float)j0;
    }
    // low-res accumulator geometry (transposed upsample)
    int jbase = 0, ibase = 0;
    if (GRAD && ps.k > 1) {
        float sxb = fmaxf(fmaf(ps.rscale, (float)x0 + 0.5f, -0.5f), 0.f);
        float syb = fmaxf(fmaf(ps.rscale, (float)y0 + 0.5f, -0.5f), 0.f);
        jbase = (int)sxb;
        ibase = (int)syb;
        sJ0[tid] = j0;
        sLam[tid] = lam;
        sG[tid] = 0.f;
    }

    const size_t plane = (size_t)H * W;
    const float* tgt_b = p.target + (size_t)b * 3 * plane;
    const float* disp_b = ps.disp + (size_t)b * ps.hd * ps.wd;
    const float wscale = (float)W / (float)(W - 1), hscale = (float)H / (float)(H - 1);
    const float wmax = (float)(W - 1), hmax = (float)(H - 1);

    // rolling state
    float hy1[3], hy2[3], hyy1[3], hyy2[3];
    float hx1[S][3], hx2[S][3], hxx1[S][3], hxx2[S][3], hxy1[S][3], hxy2[S][3];
    float hc1[GRAD ? S : 1][9], hc2[GRAD ? S : 1][9];
    float gP[GRAD ? S : 1][12];
    float loss_acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { hy1[c] = hy2[c] = hyy1[c] = hyy2[c] = 0.f; }
#pragma unroll
    for (int f = 0; f < S; ++f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { hx1[f][c] = hx2[f][c] = hxx1[f][c] = hxx2[f][c] = hxy1[f][c] = hxy2[f][c] = 0.f; }
    }
#pragma unroll
    for (int f = 0; f < (GRAD ? S : 1); ++f) {
#pragma unroll
        for (int m = 0; m < 9; ++m) { hc1[f][m] = hc2[f][m] = 0.f; }
#pragma unroll
        for (int e = 0; e < 12; ++e) gP[f][e] = 0.f;
    }
    const float wl = (cx == 1) ? 2.f : 1.f, wr = (cx == W - 2) ? 2.f : 1.f;  // reflection fold (columns)
    const float kssim = SSIM ? (0.85f / 3.0f) : 0.f;
    const float kl1 = SSIM ? (0.15f / 3.0f) : (1.0f / 3.0f);

    int slotA = 0;  // ring slot of row r; (slotA+2)%3 = row r-1; (slotA+1)%3 = row r-2
    const int r_end = GRAD ? (y1 + 1) : y1;
    for (int r = y0 - 2; r <= r_end; ++r) {
        // =============================== (A) warp row r ===================================
        const int ry = reflect1(clampi(r, -1, H), H);
        float* ringA = sRing + (size_t)slotA * RS * NT + tid;
        float yv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            yv[c] = __ldg(tgt_b + c * plane + (size_t)ry * W + rx);
            sY[c * NT + tid] = yv[c];
            ringA[c * NT] = yv[c];
        }
        // disparity at (rx, ry): bilinear upsample of disp_s, align_corners=False
        float d;
        if (ps.k > 1) {
            float sy = fmaxf(fmaf(ps.rscale, (float)ry + 0.5f, -0.5f), 0.f);
            int i0 = (int)sy, i1 = min(i0 + 1, ps.hd - 1);
            float mu = sy - (float)i0;
            float v00 = __ldg(disp_b + (size_t)i0 * ps.wd + j0), v01 = __ldg(disp_b + (size_t)i0 * ps.wd + j1);
            float v10 = __ldg(disp_b + (size_t)i1 * ps.wd + j0), v11 = __ldg(disp_b + (size_t)i1 * ps.wd + j1);
            float top = fmaf(lam, v01, (1.f - lam) * v00), bot = fmaf(lam, v11, (1.f - lam) * v10);
            d = fmaf(mu, bot, (1.f - mu) * top);
        } else {
            d = __ldg(disp_b + (size_t)ry * W + rx);
        }
        const float sigma = fmaf(p.disp_range, d, p.min_disp);   // layers.py:23
        const float D = __fdiv_rn(1.0f, sigma);                    // layers.py:24
        ringA[3 * NT] = D;
        const float fy = (float)ry;
        const float r0 = rc0 + fmaf(sIK[1], fy, sIK[2]);
        const float r1 = rc1 + fmaf(sIK[4], fy, sIK[5]);
        const float r2 = rc2 + fmaf(sIK[7], fy, sIK[8]);
        const float X0 = D * r0, X1 = D * r1, X2 = D * r2;        // layers.py:165
        const bool emit = col_owned && (r >= y0) && (r < y1);
        if (ps.depth != nullptr && emit) ps.depth[(size_t)b * plane + (size_t)r * W + cx] = D;

        float xv[S][3];
#pragma unroll
        for (int f = 0; f < S; ++f) {
            const float c0 = fmaf(P[f][0], X0, fmaf(P[f][1], X1, fmaf(P[f][2], X2, P[f][3])));
            const float c1 = fmaf(P[f][4], X0, fmaf(P[f][5], X1, fmaf(P[f][6], X2, P[f][7])));
            const float c2 = fmaf(P[f][8], X0, fmaf(P[f][9], X1, fmaf(P[f][10], X2, P[f][11])));
            const float invz = __frcp_rn(c2 + p.eps);             // layers.py:187
            const float u = c0 * invz, v = c1 * invz;
            // layers.py:190-192 + grid_sample unnormalise (align_corners=False): ix = u*W/(W-1) - 0.5
            const float ixr = fmaf(u, wscale, -0.5f), iyr = fmaf(v, hscale, -0.5f);
            const float ix = fminf(fmaxf(ixr, 0.f), wmax), iy = fminf(fmaxf(iyr, 0.f), hmax);
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const float tx = ix - fx0, ty = iy - fy0;
            const int xi0 = (int)fx0, yi0 = (int)fy0;
            const int xi1 = min(xi0 + 1, W - 1), yi1 = min(yi0 + 1, H - 1);
            const float mx = (ixr > 0.f && ixr < wmax) ? wscale : 0.f;   // clip backward * d ix/d u
            const float my = (iyr > 0.f && iyr < hmax) ? hscale : 0.f;
            const float* sb = p.src[f] + (size_t)b * 3 * plane;
            float* rf = ringA + (4 + kGeoFields * f) * NT;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* sp = sb + c * plane;
                const float nw = __ldg(sp + (size_t)yi0 * W + xi0), ne = __ldg(sp + (size_t)yi0 * W + xi1);
                const float sw = __ldg(sp + (size_t)yi1 * W + xi0), se = __ldg(sp + (size_t)yi1 * W + xi1);
                const float top = fmaf(tx, ne - nw, nw), bot = fmaf(tx, se - sw, sw);
                const float val = fmaf(ty, bot - top, top);
                xv[f][c] = val;
                sX[(f * 3 + c) * NT + tid] = val;
                rf[c * NT] = val;
                if (GRAD) {
                    rf[(3 + c) * NT] = mx * fmaf(ty, (se - sw) - (ne - nw), ne - nw);
                    rf[(6 + c) * NT] = my * fmaf(tx, (se - ne) - (sw - nw), sw - nw);
                }
                if (ps.warped != nullptr && emit)
                    ps.warped[(((size_t)f * p.B + b) * 3 + c) * plane + (size_t)r * W + cx] = val;
            }
            if (GRAD) { rf[9 * NT] = invz; rf[10 * NT] = u; rf[11 * NT] = v; }
        }
        __syncthreads();  // S1: sY / sX rows visible

        // ============= transposed upsample of the previous step's disparity gradients =======
        if (GRAD && ps.k > 1) {
            const int qy = r - 3;  // row whose g_d sits in sG
            if (qy >= y0 && qy < y1 && tid < ps.low_cols) {
                const int j = jbase + tid;
                // full-res columns touching low-res column j: [k*j - k/2, k*j + 3k/2 - 1]
                int xa = max(ps.k * j - ps.k / 2, x0), xb = min(ps.k * j + (3 * ps.k) / 2 - 1, x1 - 1);
                float h = 0.f;
                for (int x = xa; x <= xb; ++x) {
                    int t = x - x0 + 2;
                    int jj0 = sJ0[t];
                    float l = sLam[t];
                    int jj1 = min(jj0 + 1, ps.wd - 1);
                    float w = (jj0 == j ? 1.f - l : 0.f) + (jj1 == j ? l : 0.f);
                    h = fmaf(w, sG[t], h);
                }
                float sy = fmaxf(fmaf(ps.rscale, (float)qy + 0.5f, -0.5f), 0.f);
                int i0 = (int)sy, i1 = min(i0 + 1, ps.hd - 1);
                float mu = sy - (float)i0;
                sAcc[(i0 - ibase) * ps.low_cols + tid] += (1.f - mu) * h;
                sAcc[(i1 - ibase) * ps.low_cols + tid] += mu * h;
            }
        }

        // ====================== (B) close the windows centred on row r-1 ===================
        float hyn[3], hyyn[3];
        float yl[3], yr[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            yl[c] = sY[c * NT + tl];
            yr[c] = sY[c * NT + tr];
            hyn[c] = yl[c] + yv[c] + yr[c];
            hyyn[c] = fmaf(yl[c], yl[c], fmaf(yv[c], yv[c], yr[c] * yr[c]));
        }
        const int py = r - 1;
        const bool stats_ready = (r >= y0);
        const bool p_valid = stats_ready && (py >= 0) && (py < H) && col_in_image && (tid >= 1) && (tid <= NT - 2);
        const float* ringB = sRing + (size_t)((slotA + 2) % 3) * RS * NT + tid;   // row r-1
        float rp[S];
        float coef[GRAD ? S : 1][9];
#pragma unroll
        for (int f = 0; f < S; ++f) {
            float ssim_sum = 0.f, l1_sum = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float xl = sX[(f * 3 + c) * NT + tl], xr = sX[(f * 3 + c) * NT + tr], xc = xv[f][c];
                const float hxn = xl + xc + xr;
                const float hxxn = fmaf(xl, xl, fmaf(xc, xc, xr * xr));
                const float hxyn = fmaf(xl, yl[c], fmaf(xc, yv[c], xr * yr[c]));
                if (SSIM) {
                    const float Sx = hx2[f][c] + hx1[f][c] + hxn;
                    const float Sxx = hxx2[f][c] + hxx1[f][c] + hxxn;
                    const float Sxy = hxy2[f][c] + hxy1[f][c] + hxyn;
                    const float Sy = hy2[c] + hy1[c] + hyn[c];
                    const float Syy = hyy2[c] + hyy1[c] + hyyn[c];
                    float dSx = 0.f, dSxx = 0.f, dSxy = 0.f;
                    ssim_sum += ssim_window<GRAD, false>(Sx, Sy, Sxx, Syy, Sxy, dSx, dSxx, dSxy);
                    if (GRAD) {
                        // d rp / d x_q = kssim * (dSx + 2 x_q dSxx + y_q dSxy) for q in the window
                        coef[f][c] = kssim * dSx;
                        coef[f][3 + c] = kssim * 2.0f * dSxx;
                        coef[f][6 + c] = kssim * dSxy;
                    }
                }
                hx2[f][c] = hx1[f][c]; hx1[f][c] = hxn;
                hxx2[f][c] = hxx1[f][c]; hxx1[f][c] = hxxn;
                hxy2[f][c] = hxy1[f][c]; hxy1[f][c] = hxyn;
                // L1 at p = (r-1, cx): own-column values of the previous row
                const float xp = ringB[(4 + kGeoFields * f + c) * NT];
                const float yp = ringB[c * NT];
                l1_sum += fabsf(yp - xp);
            }
            rp[f] = SSIM ? fmaf(0.85f, ssim_sum * (1.0f / 3.0f), 0.15f * (l1_sum * (1.0f / 3.0f)))
                         : l1_sum * (1.0f / 3.0f);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { hy2[c] = hy1[c]; hy1[c] = hyn[c]; hyy2[c] = hyy1[c]; hyy1[c] = hyyn[c]; }

        float wgt[S];
#pragma unroll
        for (int f = 0; f < S; ++f) wgt[f] = 0.f;
        if (p_valid) {
            // candidates in the reference's order: identity (+noise) first, then reprojection
            // (trainer.py:597); torch.min returns the first minimum.
            float best = 3.0e38f;
            int best_i = 0;
            const size_t pix = (size_t)py * W + cx;
            if (n_id > 0) {
                float nz[4] = {0.f, 0.f, 0.f, 0.f};
                if (ps.noise == nullptr) {
                    size_t lin = (size_t)b * plane + pix;
                    philox_normal4(p.seed_dev ? (p.seed ^ __ldg(p.seed_dev)) : p.seed, (uint32_t)lin, (uint32_t)(lin >> 32), (uint32_t)pass_i, nz);
                }
                for (int i = 0; i < n_id; ++i) {
                    size_t off = ((size_t)b * n_id + i) * plane + pix;
                    float nv = (ps.noise != nullptr) ? __ldg(ps.noise + off) : nz[i];
                    float cand = fmaf(nv, kTieNoise, __ldg(p.identity + off));
                    if (cand < best) { best = cand; best_i = i; }
                }
            }
            if (avg) {
                float m = 0.f;
#pragma unroll
                for (int f = 0; f < S; ++f) m += rp[f];
                m = m / (float)S;
                if (m < best) { best = m; best_i = n_id; }
                if (best_i == n_id) {
#pragma unroll
                    for (int f = 0; f < S; ++f) wgt[f] = 1.0f / (float)S;
                }
            } else {
#pragma unroll
                for (int f = 0; f < S; ++f)
                    if (rp[f] < best) { best = rp[f]; best_i = n_id + f; }
#pragma unroll
                for (int f = 0; f < S; ++f) wgt[f] = (best_i == n_id + f) ? 1.f : 0.f;
            }
            if (col_owned && py >= y0 && py < y1) {
                loss_acc += best;
                if (ps.argmin != nullptr) ps.argmin[(size_t)b * plane + pix] = (uint8_t)best_i;
            }
        }
        if (GRAD) {
            float* ringBw = sRing + (size_t)((slotA + 2) % 3) * RS * NT + tid;
#pragma unroll
            for (int f = 0; f < S; ++f) {
                const float wsc = wgt[f] * p.inv_n;
                ringBw[(4 + kGeoFields * f + 12) * NT] = wsc;
                if (SSIM) {
#pragma unroll
                    for (int m = 0; m < 9; ++m) sC[(f * 9 + m) * NT + tid] = p_valid ? coef[f][m] * wsc : 0.f;
                }
            }
        }
        __syncthreads();  // S2: coefficient row visible (also fences sG/sAcc use above)

        // ====================== (C) adjoint for the pixels of row r-2 ======================
        if (GRAD) {
            const int qy = r - 2;
            const bool do_q = (r >= y0 + 2);   // => y0 <= qy < y1
            const float wt = (qy == 1) ? 2.f : 1.f, wb = (qy == H - 2) ? 2.f : 1.f;  // reflection fold (rows)
            const float* ringC = sRing + (size_t)((slotA + 1) % 3) * RS * NT + tid;   // row r-2
            float g_d = 0.f;
            float rq0 = 0.f, rq1 = 0.f, rq2 = 0.f, Dq = 0.f, yq[3] = {0.f, 0.f, 0.f};
            if (do_q) {
                const float fq = (float)qy;
                rq0 = rc0 + fmaf(sIK[1], fq, sIK[2]);
                rq1 = rc1 + fmaf(sIK[4], fq, sIK[5]);
                rq2 = rc2 + fmaf(sIK[7], fq, sIK[8]);
                Dq = ringC[3 * NT];
#pragma unroll
                for (int c = 0; c < 3; ++c) yq[c] = ringC[c * NT];
            }
#pragma unroll
            for (int f = 0; f < S; ++f) {
                float gx[3] = {0.f, 0.f, 0.f};
                if (SSIM) {
                    float V[9];
#pragma unroll
                    for (int m = 0; m < 9; ++m) {
                        const float* row = sC + (f * 9 + m) * NT;
                        const float hn = fmaf(wl, row[tl], fmaf(wr, row[tr], row[tid]));
                        V[m] = fmaf(wt, hc2[f][m], fmaf(wb, hn, hc1[f][m]));
                        hc2[f][m] = hc1[f][m];
                        hc1[f][m] = hn;
                    }
                    if (do_q) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float xq = ringC[(4 + kGeoFields * f + c) * NT];
                            gx[c] = fmaf(xq, V[3 + c], fmaf(yq[c], V[6 + c], V[c]));
                        }
                    }
                }
                if (do_q && col_owned) {
                    const float* rf = ringC + (4 + kGeoFields * f) * NT;
                    const float wq = rf[12 * NT];   // winner weight * 1/N at q
                    float du = 0.f, dv = 0.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float xq = rf[c * NT];
                        const float df = xq - yq[c];
                        const float sgn = (df > 0.f) ? 1.f : ((df < 0.f) ? -1.f : 0.f);
                        const float g = fmaf(kl1 * wq, sgn, gx[c]);
                        du = fmaf(g, rf[(3 + c) * NT], du);
                        dv = fmaf(g, rf[(6 + c) * NT], dv);
                    }
                    const float invz = rf[9 * NT], u = rf[10 * NT], v = rf[11 * NT];
                    const float dc0 = du * invz, dc1 = dv * invz;
                    const float dc2 = -(fmaf(u, du, v * dv)) * invz;
                    const float Xq0 = Dq * rq0, Xq1 = Dq * rq1, Xq2 = Dq * rq2;
                    gP[f][0] = fmaf(dc0, Xq0, gP[f][0]); gP[f][1] = fmaf(dc0, Xq1, gP[f][1]);
                    gP[f][2] = fmaf(dc0, Xq2, gP[f][2]); gP[f][3] += dc0;
                    gP[f][4] = fmaf(dc1, Xq0, gP[f][4]); gP[f][5] = fmaf(dc1, Xq1, gP[f][5]);
                    gP[f][6] = fmaf(dc1, Xq2, gP[f][6]); gP[f][7] += dc1;
                    gP[f][8] = fmaf(dc2, Xq0, gP[f][8]); gP[f][9] = fmaf(dc2, Xq1, gP[f][9]);
                    gP[f][10] = fmaf(dc2, Xq2, gP[f][10]); gP[f][11] += dc2;
                    const float gX0 = fmaf(P[f][0], dc0, fmaf(P[f][4], dc1, P[f][8] * dc2));
                    const float gX1 = fmaf(P[f][1], dc0, fmaf(P[f][5], dc1, P[f][9] * dc2));
                    const float gX2 = fmaf(P[f][2], dc0, fmaf(P[f][6], dc1, P[f][10] * dc2));
                    const float gD = fmaf(rq0, gX0, fmaf(rq1, gX1, rq2 * gX2));
                    g_d = fmaf(-p.disp_range * Dq * Dq, gD, g_d);   // d(1/sigma)/d disp
                }
            }
            if (ps.k > 1) {
                sG[tid] = (do_q && col_owned) ? g_d : 0.f;
            } else if (do_q && col_owned && ps.grad_disp != nullptr) {
                ps.grad_disp[(size_t)b * plane + (size_t)qy * W + cx] += g_d;
            }
        }
        slotA = (slotA + 1) % 3;
    }

    // ------------------------------- epilogue ------------------------------------------------
    if (GRAD && ps.k > 1) {
        __syncthreads();
        {   // last row's g_d (qy = y1-1) is still in sG
            const int qy = y1 - 1;
            if (tid < ps.low_cols) {
                const int j = jbase + tid;
                int xa = max(ps.k * j - ps.k / 2, x0), xb = min(ps.k * j + (3 * ps.k) / 2 - 1, x1 - 1);
                float h = 0.f;
                for (int x = xa; x <= xb; ++x) {
                    int t = x - x0 + 2;
                    int jj0 = sJ0[t];
                    float l = sLam[t];
                    int jj1 = min(jj0 + 1, ps.wd - 1);
                    float w = (jj0 == j ? 1.f - l : 0.f) + (jj1 == j ? l : 0.f);
                    h = fmaf(w, sG[t], h);
                }
                float sy = fmaxf(fmaf(ps.rscale, (float)qy + 0.5f, -0.5f), 0.f);
                int i0 = (int)sy, i1 = min(i0 + 1, ps.hd - 1);
                float mu = sy - (float)i0;
                sAcc[(i0 - ibase) * ps.low_cols + tid] += (1.f - mu) * h;
                sAcc[(i1 - ibase) * ps.low_cols + tid] += mu * h;
            }
        }
        __syncthreads();
        float* gd = ps.grad_disp + (size_t)b * ps.hd * ps.wd;
        for (int i = tid; i < ps.low_rows * ps.low_cols; i += NT) {
            int ii = ibase + i / ps.low_cols, jj = jbase + i % ps.low_cols;
            float v = sAcc[i];
            if (ii < ps.hd && jj < ps.wd && v != 0.f) atomicAdd(gd + (size_t)ii * ps.wd + jj, v);
        }
    }

    // block reduction: warp shuffles, then one partial row per CTA (reduced later in fixed order)
    const int nred = GRAD ? 1 + 12 * S : 1;
    const int warp = tid >> 5, lane = tid & 31, nwarp = NT >> 5;
    {
        float v = warp_sum(loss_acc);
        if (lane == 0) sRed[warp * nred] = v;
    }
    if (GRAD) {
#pragma unroll
        for (int f = 0; f < S; ++f)
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                float v = warp_sum(gP[f][e]);
                if (lane == 0) sRed[warp * nred + 1 + f * 12 + e] = v;
            }
    }
    __syncthreads();
    if (tid < nred) {
        float v = 0.f;
        for (int w = 0; w < nwarp; ++w) v += sRed[w * nred + tid];
        p.part[(size_t)blockIdx.x * p.part_stride + tid] = v;
    }
}

// shared-memory bytes of photometric_kernel<S,GRAD,*> for a given launch shape
template <int S>
inline size_t photometric_smem_bytes(int NT, bool grad, int low_cells) {
    size_t fl = (size_t)S * 12 + 12 + 3 * NT + (size_t)S * 3 * NT + (grad ? (size_t)S * 9 * NT : 0) +
                (size_t)3 * ring_slot_floats<S>() * NT + 3 * NT + (grad ? low_cells : 0) +
                (size_t)(NT / 32) * (1 + 12 * S);
    return fl * sizeof(float) + 16;
}

// ---------------------------------------------------------------------------------------------
// finalize: fixed-order reduction of the per-CTA partials into losses / grad_T, plus the
// smoothness partials (see pml_smooth.cuh).  Stage 1: one CTA per (image, scale); stage 2: one warp
// per scale combines the images.  The work is a few thousand floats.
// ---------------------------------------------------------------------------------------------
struct FinalizeParams {
    int B, S, n_pass, cta_per_pass, cta_per_image, part_stride, with_grad;
    float inv_n;               // 1 / (B*H*W): to_optimise.mean(), trainer.py:610
    const float* part;         // photometric partials
    const float* K;            // [B,4,4]
    const float* smooth_part;  // [n_pass][B][smooth_blocks][3]: sum_x, sum_y, sum g_n*disp
    int smooth_blocks[PML_MAX_PASSES];
    int smooth_off[PML_MAX_PASSES];   // offset (in partial rows) of pass i
    const float* disp_mean;    // [n_pass][B]
    int hd[PML_MAX_PASSES], wd[PML_MAX_PASSES];
    float smooth_weight[PML_MAX_PASSES];
    float* image_part;         // [n_pass][B][4]: photometric sum, smooth x sum, smooth y sum, -
    float* losses;             // [n_pass][4]
    float* loss_vector;        // [n_pass] or null: loss_s once more, contiguous
    float* loss_total;         // [1] or null: sum_s loss_s / total_div (trainer.py:618-621)
    float total_div;
    float* grad_T;             // [n_pass][S][B][16]
    float* grad_disp_const;    // [n_pass][B]
    int n_seg, seg_size;
    ChunkPtrs K_c;
};

__global__ void __launch_bounds__(256)
finalize_image_kernel(const FinalizeParams q) {
    __shared__ float s_col[1 + 12 * PML_MAX_SOURCES];
    __shared__ float s_seg[256];       // flat [segment][column]: 8 x 32 or 2 x 128
    __shared__ float s_w[8][3];
    const int tid = threadIdx.x, b = blockIdx.x, pi = blockIdx.y;
    const int ncol = q.with_grad ? 1 + 12 * q.S : 1;
    // photometric partial columns of this image's work items (up to 1 + 12 * 8 = 97): two interleaved
    // segments per column, each summed in item order, then combined (fixed order => deterministic)
    // (up to 32 columns, i.e. S <= 2: eight segments of 32 columns -- four times shorter dependent chains)
    const bool narrow = ncol <= 32;
    {
        const int col = narrow ? (tid & 31) : (tid & 127), seg = narrow ? (tid >> 5) : (tid >> 7);
        const int nseg = narrow ? 8 : 2;
        float v = 0.f;
        if (col < ncol) {
            const float* base = q.part + (size_t)(pi * q.cta_per_pass + b * q.cta_per_image) * q.part_stride + col;
#pragma unroll 4      // the loads are independent of the running sum: four in flight, added in the same order
            for (int c = seg; c < q.cta_per_image; c += nseg) v += base[(size_t)c * q.part_stride];
        }
        s_seg[tid] = v;
    }
    __syncthreads();
    if (tid < ncol) {
        float t = 0.f;
        if (narrow) { for (int g = 0; g < 8; ++g) t += s_seg[g * 32 + tid]; }
        else t = s_seg[tid] + s_seg[128 + tid];
        s_col[tid] = t;
    }
    // smoothness partials of this image
    const int nb = q.smooth_blocks[pi];
    const float* sp = q.smooth_part + ((size_t)q.smooth_off[pi] + (size_t)b * nb) * 3;
    float vx = 0.f, vy = 0.f, vg = 0.f;
    for (int i = tid; i < nb; i += blockDim.x) { vx += sp[i * 3]; vy += sp[i * 3 + 1]; vg += sp[i * 3 + 2]; }
    vx = warp_sum(vx); vy = warp_sum(vy); vg = warp_sum(vg);
    if ((tid & 31) == 0) { s_w[tid >> 5][0] = vx; s_w[tid >> 5][1] = vy; s_w[tid >> 5][2] = vg; }
    __syncthreads();
    if (tid == 0) {
        float sx = 0.f, sy = 0.f, sg = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sx += s_w[w][0]; sy += s_w[w][1]; sg += s_w[w][2]; }
        float* ip = q.image_part + (size_t)(pi * q.B + b) * 4;
        ip[0] = s_col[0]; ip[1] = sx; ip[2] = sy; ip[3] = 0.f;
        if (q.with_grad) {
            const float m = q.disp_mean[pi * q.B + b] + 1e-7f;
            // d/d mean of disp/(mean+eps): -sum(g_n*disp)/(mean+eps)^2, spread by 1/(h*w)
            q.grad_disp_const[pi * q.B + b] =
                -q.smooth_weight[pi] * sg / (m * m) / ((float)q.hd[pi] * (float)q.wd[pi]);
        }
    }
    // dL/dT = K[:3,:]^T dL/dP  (layers.py:183: P = (K T)[:3])
    if (q.with_grad && tid < q.S * 16) {
        const int f = tid >> 4, e = tid & 15, kk = e >> 2, j = e & 3;
        int bl;
        const float* Kc = chunk_of(q.K, q.K_c, q.n_seg, q.seg_size, b, bl);
        const float* Kb = Kc + (size_t)bl * 16;
        const float* gp = s_col + 1 + f * 12;
        q.grad_T[((size_t)(pi * q.S + f) * q.B + b) * 16 + e] =
            fmaf(Kb[kk], gp[j], fmaf(Kb[4 + kk], gp[4 + j], Kb[8 + kk] * gp[8 + j]));
    }
}

// one warp: lane pi reduces scale pi over the images (fixed order), lane 0 then adds the scales in
// order like `total_loss += loss; total_loss /= num_scales` (trainer.py:618-621)
__global__ void __launch_bounds__(32)
finalize_loss_kernel(const FinalizeParams q) {
    __shared__ float s_loss[PML_MAX_PASSES];
    const int pi = threadIdx.x;
    if (pi < q.n_pass) {
        double ph = 0.0, sx = 0.0, sy = 0.0;   // a handful of adds per launch: double costs nothing here
        for (int b = 0; b < q.B; ++b) {
            const float* ip = q.image_part + (size_t)(pi * q.B + b) * 4;
            ph += (double)ip[0]; sx += (double)ip[1]; sy += (double)ip[2];
        }
        const float h = (float)q.hd[pi], w = (float)q.wd[pi];
        const float photo = (float)(ph * (double)q.inv_n);
        const float sm = (float)sx / ((float)q.B * h * (w - 1.f)) + (float)sy / ((float)q.B * (h - 1.f) * w);   // layers.py:215
        const float loss = photo + q.smooth_weight[pi] * sm;   // trainer.py:610,616
        q.losses[pi * 4 + 0] = loss;
        q.losses[pi * 4 + 1] = photo;
        q.losses[pi * 4 + 2] = sm;
        q.losses[pi * 4 + 3] = 0.f;
        if (q.loss_vector != nullptr) q.loss_vector[pi] = loss;
        s_loss[pi] = loss;
    }
    __syncwarp();
    if (pi == 0 && q.loss_total != nullptr) {
        float t = 0.f;
        for (int k = 0; k < q.n_pass; ++k) t += s_loss[k];
        q.loss_total[0] = t / q.total_div;
    }
}

}  // namespace pml
