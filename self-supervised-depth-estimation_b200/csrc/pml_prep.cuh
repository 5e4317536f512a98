// Warp-strip versions of the two kernels that run before the fused sweep (same scheme as
// pml_sweep.cuh: one warp marches down a strip of columns, neighbours travel by shuffles, vertical
// 3-sums roll in registers, every input byte is read once):
//
//  * identity_sweep_kernel -- the identity reprojection losses of the automask, trainer.py:556-569:
//    reprojection_loss(source_f, target) per pixel, once per step, re-used by every scale;
//  * smooth_sweep_kernel   -- get_smooth_loss (layers.py:202-215) on disp / (mean + 1e-7)
//    (trainer.py:612-616), forward sums and the adjoint with respect to disp in one sweep; it
//    *initialises* grad_disp_s, the fused sweep adds the photometric part on top.
#pragma once
#include "pml_common.cuh"
#include "pml_sweep.cuh"
#include "pml_smooth.cuh"

namespace pml {

constexpr int kPrepTW = 30;   // owned columns per strip: 32 lanes minus one halo column each side

// One SSIM window for the frame pair with the reference's rounding structure (layers.py:238-248):
// products are rounded separately and B1 / B2 are summed x-part + y-part first, so that a source
// frame that equals the target bit for bit (static camera, mono_dataset.py:165-170) gives numerator
// == denominator and the quotient, refined by one Newton step, is exactly 1 -> dissimilarity 0.
__device__ __forceinline__ float2 ssim_pair_exact(float2 Sx, float2 Sxx, float2 Sxy, float Sy, float Syy) {
    const float k9 = 1.0f / 9.0f;
    const float2 mx = mul2(Sx, splat(k9));
    const float my = Sy * k9;
    const float2 mxx = mul2(mx, mx), mxy = mul2(mx, splat(my));
    const float myy = __fmul_rn(my, my);
    const float2 sx = fma2(Sxx, splat(k9), f2(-mxx.x, -mxx.y));
    const float2 sxy = fma2(Sxy, splat(k9), f2(-mxy.x, -mxy.y));
    const float sy = fmaf(Syy, k9, -myy);
    const float2 A1 = fma2(mxy, splat(2.0f), splat(kSsimC1));
    const float2 A2 = fma2(sxy, splat(2.0f), splat(kSsimC2));
    const float2 B1 = add2(add2(mxx, splat(myy)), splat(kSsimC1));
    const float2 B2 = add2(add2(sx, splat(sy)), splat(kSsimC2));
    const float2 num = mul2(A1, A2), den = mul2(B1, B2);
    const float2 q = f2(rcp_approx(den.x), rcp_approx(den.y));
    float2 ratio = mul2(num, q);
    ratio = fma2(fma2(f2(-den.x, -den.y), ratio, num), q, ratio);
    return f2(__saturatef(fmaf(-0.5f, ratio.x, 0.5f)), __saturatef(fmaf(-0.5f, ratio.y, 0.5f)));
}

struct IdentityParams {
    const float* target;
    const float* src0;
    const float* src1;     // == src0 when S == 1
    float* out;            // [B, n_out, H, W], n_out = avg ? 1 : total source frames
    int B, H, W, S, avg;   // S: frames of THIS launch (1 or 2); more frames are swept pair by pair
    int n_out, plane_off;  // output planes per image, first plane written by this launch
    int accumulate;        // avg: add to what an earlier pair wrote
    float inv_total;       // avg: 1 / total source frames
    int TH, n_strips, n_chunks;
    int n_seg, seg_size;   // chunked batch (pml_segments)
    ChunkPtrs target_c, src0_c, src1_c;
};

// one warp = one (chunk, strip) item `bx` of image `b`
template <bool SSIM>
__device__ __forceinline__ void identity_sweep_body(const IdentityParams& p, const int bx, const int b) {
    const int lane = threadIdx.x;
    const int chunk = bx / p.n_strips, strip = bx - chunk * p.n_strips;
    const int H = p.H, W = p.W, plane = H * W;
    const int x0 = strip * kPrepTW, x1 = min(x0 + kPrepTW, W);
    const int y0 = chunk * p.TH, y1 = min(y0 + p.TH, H);
    const int cx = x0 - 1 + lane;
    const int rx = reflect1(clampi(cx, -1, W), W);
    const bool owned = (cx >= x0) && (cx < x1);
    int bl;
    const float* __restrict__ tgt_g = chunk_of(p.target, p.target_c, p.n_seg, p.seg_size, b, bl);
    const float* __restrict__ s0_g = chunk_of(p.src0, p.src0_c, p.n_seg, p.seg_size, b, bl);
    const float* __restrict__ s1_g = chunk_of(p.src1, p.src1_c, p.n_seg, p.seg_size, b, bl);
    const int b3p = bl * 3 * plane;
    float* out_b = p.out + ((size_t)b * p.n_out + p.plane_off) * plane;

    float hy1[3], hy2[3], hyy1[3], hyy2[3];
    float2 hx1[3], hx2[3], hxx1[3], hxx2[3], hxy1[3], hxy2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        hy1[c] = hy2[c] = hyy1[c] = hyy2[c] = 0.f;
        hx1[c] = hx2[c] = hxx1[c] = hxx2[c] = hxy1[c] = hxy2[c] = splat(0.f);
    }
    float2 l1_prev = splat(0.f);

    // streaming inputs are fetched two rows (one loop trip) ahead of their use: the kernel is bound by
    // memory latency, not arithmetic, so every warp keeps 18 loads in flight while it computes
    struct Row { float y[3]; float2 x[3]; };
    auto load = [&](int r, Row& row) {
        const int ry = reflect1(clampi(r, -1, H), H);
        const int o = b3p + ry * W + rx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            row.y[c] = __ldg(at(tgt_g, o + c * plane));
            row.x[c] = f2(__ldg(at(s0_g, o + c * plane)), __ldg(at(s1_g, o + c * plane)));
        }
    };
    auto step = [&](const int r, const Row& row, float (&hyA)[3], float (&hyB)[3], float (&hyyA)[3], float (&hyyB)[3],
                    float2 (&hxA)[3], float2 (&hxB)[3], float2 (&hxxA)[3], float2 (&hxxB)[3],
                    float2 (&hxyA)[3], float2 (&hxyB)[3]) {
        float2 ssim_sum = splat(0.f), l1_cur = splat(0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float yv = row.y[c];
            const float2 xv = row.x[c];
            l1_cur.x += fabsf(yv - xv.x);
            l1_cur.y += fabsf(yv - xv.y);
            if (SSIM) {
                const float yl = __shfl_up_sync(0xffffffffu, yv, 1), yr = __shfl_down_sync(0xffffffffu, yv, 1);
                const float2 xl = shfl_up2(xv), xr = shfl_down2(xv);
                const float hyn = yl + yv + yr;
                const float hyyn = fmaf(yl, yl, fmaf(yv, yv, yr * yr));
                const float2 hxn = add2(add2(xl, xv), xr);
                const float2 hxxn = fma2(xl, xl, fma2(xv, xv, mul2(xr, xr)));
                const float2 hxyn = fma2(xl, splat(yl), fma2(xv, splat(yv), mul2(xr, splat(yr))));
                const float Sy = hyB[c] + hyA[c] + hyn, Syy = hyyB[c] + hyyA[c] + hyyn;
                const float2 Sx = add2(add2(hxB[c], hxA[c]), hxn);
                const float2 Sxx = add2(add2(hxxB[c], hxxA[c]), hxxn);
                const float2 Sxy = add2(add2(hxyB[c], hxyA[c]), hxyn);
                hyB[c] = hyn; hyyB[c] = hyyn; hxB[c] = hxn; hxxB[c] = hxxn; hxyB[c] = hxyn;
                ssim_sum = add2(ssim_sum, ssim_pair_exact(Sx, Sxx, Sxy, Sy, Syy));
            }
        }
        // window row r-1 (trainer.py:527 / :523)
        const int py = r - 1;
        if (owned && py >= y0 && py < y1) {
            float2 rp;
            if (SSIM) rp = f2(fmaf(0.85f, ssim_sum.x * (1.0f / 3.0f), 0.15f * (l1_prev.x * (1.0f / 3.0f))),
                              fmaf(0.85f, ssim_sum.y * (1.0f / 3.0f), 0.15f * (l1_prev.y * (1.0f / 3.0f))));
            else rp = f2(l1_prev.x * (1.0f / 3.0f), l1_prev.y * (1.0f / 3.0f));
            const int pix = py * W + cx;
            if (p.avg) {
                const float v = ((p.S > 1) ? rp.x + rp.y : rp.x) * p.inv_total;   // trainer.py:565-566
                out_b[pix] = p.accumulate ? out_b[pix] + v : v;
            } else {
                out_b[pix] = rp.x;
                if (p.S > 1) out_b[plane + pix] = rp.y;
            }
        }
        l1_prev = l1_cur;
    };
    Row c0, c1, n0, n1;
    load(y0 - 1, c0);
    load(y0, c1);
#pragma unroll 1
    for (int r = y0 - 1; r <= y1; r += 2) {
        load(r + 2, n0);
        load(r + 3, n1);
        step(r, c0, hy1, hy2, hyy1, hyy2, hx1, hx2, hxx1, hxx2, hxy1, hxy2);
        step(r + 1, c1, hy2, hy1, hyy2, hyy1, hx2, hx1, hxx2, hxx1, hxy2, hxy1);
        c0 = n0;
        c1 = n1;
    }
}

// grid = (n_chunks * n_strips, B), 32 threads
template <bool SSIM>
__global__ void __launch_bounds__(32)
identity_sweep_kernel(const IdentityParams p) {
    identity_sweep_body<SSIM>(p, blockIdx.x, blockIdx.y);
}

// ---------------------------------------------------------------------------------------------
// smoothness term: grid = (sum over passes of B * strips_i * chunks_i), 32 threads.  Lane 0 and 31
// are halo columns (they supply the left neighbour's right-edge term and the right neighbour's
// values); a warp marches down its rows holding the previous row in registers, so that the edge to
// the row above and its adjoint need no second read.  Partials per item: sum_x, sum_y, sum g_n * disp.
// ---------------------------------------------------------------------------------------------
constexpr int kSmoothTH = 32;   // rows per item

__device__ __forceinline__ float sgn1(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

template <bool GRAD>
__device__ __forceinline__ void smooth_sweep_body(const SmoothParams& q, const int item) {
    const int lane = threadIdx.x;
    int pi = 0;
#pragma unroll 1
    for (int i = 1; i < q.n_pass; ++i) if (item >= q.pass[i].block_off) pi = i;
    const SmoothPass& ps = q.pass[pi];
    int rel = item - ps.block_off;
    const int b = rel / ps.blocks;
    rel -= b * ps.blocks;
    const int h = ps.h, w = ps.w, n = h * w;
    const int n_strips = (w + kPrepTW - 1) / kPrepTW;
    const int chunk = rel / n_strips, strip = rel - chunk * n_strips;
    const int x0 = strip * kPrepTW, x1 = min(x0 + kPrepTW, w);
    const int y0 = chunk * kSmoothTH, y1 = min(y0 + kSmoothTH, h);
    const int cx = x0 - 1 + lane;
    const int xc = clampi(cx, 0, w - 1);
    const bool owned = (cx >= x0) && (cx < x1);
    const bool has_right = (cx >= 0) && (cx + 1 < w) && (lane < 31);   // edge (cx, cx+1) exists and is computable here

    const float mean = image_mean(q, pi, b);
    if (item == ps.block_off + b * ps.blocks && lane == 0) q.disp_mean[pi * q.B + b] = mean;
    const float inv = __fdiv_rn(1.0f, mean + 1e-7f);
    const float nx_ = 1.0f / ((float)q.B * (float)h * (float)(w - 1));
    const float ny_ = 1.0f / ((float)q.B * (float)(h - 1) * (float)w);
    const float* dg = ps.disp + (size_t)b * n;
    int bl;
    const float* cgc = chunk_of(ps.color, q.color_c[pi], q.n_seg, q.seg_size, b, bl);
    const float* cg = cgc + (size_t)bl * 3 * n;

    float ex = 0.f, ey = 0.f, gd = 0.f;
    // previous row (y-1): normalised disparity, raw disparity, colours, its horizontal adjoint, and
    // the vertical edge term between rows y-2 and y-1
    float pn = 0.f, pd = 0.f, pc0 = 0.f, pc1 = 0.f, pc2 = 0.f, pgh = 0.f, pgv = 0.f;
    // rows y0-1 .. y1: row y0-1 only primes the vertical edge, row y1 only closes it.  Inputs are
    // fetched four rows ahead of their use (latency-bound streaming kernel).
    struct Row { float d, c0, c1, c2; };
    auto load = [&](int y, Row& r) {
        const int o = clampi(y, 0, h - 1) * w + xc;
        r.d = __ldg(dg + o);
        r.c0 = __ldg(cg + o); r.c1 = __ldg(cg + n + o); r.c2 = __ldg(cg + 2 * n + o);
    };
    auto step = [&](const int y, const Row& row) {
        const bool row_in = (y >= 0) && (y < h);
        const float d = row.d, c0 = row.c0, c1 = row.c1, c2 = row.c2;
        const float nd = d * inv;
        // horizontal edge (cx, cx+1) of this row
        const float ndr = __shfl_down_sync(0xffffffffu, nd, 1);
        const float r0 = __shfl_down_sync(0xffffffffu, c0, 1), r1 = __shfl_down_sync(0xffffffffu, c1, 1);
        const float r2 = __shfl_down_sync(0xffffffffu, c2, 1);
        float gh_edge = 0.f;   // d loss / d nd(cx) through this edge ( = - d loss / d nd(cx+1) )
        if (has_right && row_in) {
            const float wgt = expf(-(fabsf(c0 - r0) + fabsf(c1 - r1) + fabsf(c2 - r2)) * (1.0f / 3.0f));   // layers.py:209-212
            const float df = nd - ndr;
            if (owned && y >= y0 && y < y1) ex += fabsf(df) * wgt;
            gh_edge = sgn1(df) * wgt * nx_;
        }
        const float gh_left = __shfl_up_sync(0xffffffffu, gh_edge, 1);   // edge (cx-1, cx)
        const float gh = gh_edge - ((lane > 0) ? gh_left : 0.f);
        // vertical edge between rows y-1 and y
        float gv_edge = 0.f;
        if (row_in && y >= 1 && y > y0 - 1) {
            const float wgt = expf(-(fabsf(pc0 - c0) + fabsf(pc1 - c1) + fabsf(pc2 - c2)) * (1.0f / 3.0f));
            const float df = pn - nd;
            if (owned && (y - 1) >= y0 && (y - 1) < y1) ey += fabsf(df) * wgt;
            gv_edge = sgn1(df) * wgt * ny_;
        }
        // pixel (y-1, cx) is complete: horizontal adjoint pgh, vertical edges pgv (above) and gv_edge (below)
        if (GRAD && owned && (y - 1) >= y0 && (y - 1) < y1) {
            const float g = pgh + gv_edge - pgv;
            gd = fmaf(g, pd, gd);
            if (ps.grad_disp != nullptr) ps.grad_disp[(size_t)b * n + (y - 1) * w + cx] = ps.weight * g * inv;
        }
        pn = nd; pd = d; pc0 = c0; pc1 = c1; pc2 = c2; pgh = gh; pgv = gv_edge;
    };
    Row cur[4], nxt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) load(y0 - 1 + i, cur[i]);
#pragma unroll 1
    for (int y = y0 - 1; y <= y1; y += 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) load(y + 4 + i, nxt[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (y + i <= y1) step(y + i, cur[i]);
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
    }
    ex = warp_sum(ex); ey = warp_sum(ey); gd = warp_sum(gd);
    if (lane == 0) {
        float* o = q.part + (size_t)item * 3;
        o[0] = ex; o[1] = ey; o[2] = gd;
    }
}

template <bool GRAD>
__global__ void __launch_bounds__(32)
smooth_sweep_kernel(const SmoothParams q) {
    smooth_sweep_body<GRAD>(q, (int)blockIdx.x);
}

// Both preparation sweeps in ONE launch: the two are independent (identity losses read the images,
// the smoothness term the disparities), each alone leaves the machine under-filled (0.7 waves of
// latency-bound warps), together they overlap.  Identity items come first (they run longer).
template <bool SSIM, bool GRAD>
__global__ void __launch_bounds__(32)
prep_kernel(const IdentityParams ip, const SmoothParams sp, const int n_identity_items) {
    const int bx = (int)blockIdx.x;
    if (bx < n_identity_items) {
        const int per_image = ip.n_chunks * ip.n_strips;
        const int b = bx / per_image;
        identity_sweep_body<SSIM>(ip, bx - b * per_image, b);
    } else {
        smooth_sweep_body<GRAD>(sp, bx - n_identity_items);
    }
}

}  // namespace pml
