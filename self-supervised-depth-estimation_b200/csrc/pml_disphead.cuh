// DepthDecoder disparity heads (second half): the producer of outputs[("disp", s)],
//
//     disp_s = sigmoid(Conv3x3(x_s))          networks/depth_decoder.py:46-47, :62-66
//     Conv3x3 = ReflectionPad2d(1) + Conv2d(C_in, 1, 3)   layers.py:121-136,  C_in = 16 / 32 / 64 / 128 for s = 0..3
//
// fused into one forward and one backward kernel: the stock path pads (a full copy of x), convolves and applies
// the sigmoid in three launches and keeps the padded copy and the pre-activation for autograd.  Both kernels are
// HBM-bound byte movers (x is read once forward; backward reads x and writes g_x once), organised like the other
// sweeps of this library: one warp marches down a strip of 30 owned columns (+1 halo lane each side), horizontal
// neighbours travel by shuffles, vertical neighbours roll in registers, reflection is an index map on the loads.
#pragma once
#include "pml_common.cuh"

namespace pml {

constexpr int kHeadTW = 30;   // owned columns per strip
constexpr int kHeadTH = 16;   // owned rows per work item

struct HeadParams {
    const float* x;       // [B,C,h,w]
    const float* weight;  // [1,C,3,3]
    const float* bias;    // [1]
    float* disp;          // [B,1,h,w]  (forward: out; backward: the saved output)
    const float* g_disp;  // [B,1,h,w]
    float* g_x;           // [B,C,h,w] or null
    float* part;          // [items][10]: 9 weight-gradient partials + bias partial
    int B, C, h, w, n_strips, n_chunks;
};

__device__ __forceinline__ float sigmoidf(float z) { return __fdiv_rn(1.0f, 1.0f + expf(-z)); }

// grid = B * ceil(h / kHeadFR) * n_strips work items of kHeadFR rows x 30 columns; G = blockDim.x / 32 warps per item, warp g
// convolves the channels g, g + G, g + 2G, ... for the WHOLE item (the low-resolution heads have 64 / 128 channels over few
// pixels: without the split a 24 x 80 map would occupy 72 warps of the whole GPU): the ten input rows of a channel are loaded
// at once, the kHeadFR pre-activations stay in registers, and the warps meet in shared memory once, at the end (the first
// version exchanged partial rows through shared memory with two block barriers per row and was slower than cuDNN on the
// 24 x 80 head).  Dynamic smem = C * 12 floats (weights) + G * kHeadFR * 32 floats.
constexpr int kHeadFR = 8;
__global__ void __launch_bounds__(256)
disp_head_fwd_kernel(const HeadParams p) {
    PML_DYN_SMEM(float, sw);
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5, G = blockDim.x >> 5;
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    float* sred = sw + C * 12;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[(i / 9) * 12 + (i % 9)] = p.weight[i];
    __syncthreads();
    const int n_rows = (h + kHeadFR - 1) / kHeadFR;
    int item = blockIdx.x;
    const int strip = item % p.n_strips; item /= p.n_strips;
    const int chunk = item % n_rows;
    const int b = item / n_rows;
    const int x0 = strip * kHeadTW, x1 = min(x0 + kHeadTW, w);
    const int y0 = chunk * kHeadFR;
    const int cx = x0 - 1 + lane;
    const int rx = reflect1(clampi(cx, -1, w), w);
    const bool owned = (cx >= x0) && (cx < x1);
    const float* xb = p.x + (size_t)b * C * plane + rx;
    int ro[kHeadFR + 2];
#pragma unroll
    for (int j = 0; j < kHeadFR + 2; ++j) ro[j] = reflect1(clampi(y0 - 1 + j, -1, h), h) * w;
    float a[kHeadFR];
#pragma unroll
    for (int i = 0; i < kHeadFR; ++i) a[i] = 0.f;
#pragma unroll 2
    for (int c = g; c < C; c += G) {
        const float* xc = xb + (size_t)c * plane;
        float v[kHeadFR + 2];
#pragma unroll
        for (int j = 0; j < kHeadFR + 2; ++j) v[j] = __ldg(xc + ro[j]);
        const float4* wk = reinterpret_cast<const float4*>(sw + c * 12);
        const float4 w0 = wk[0], w1 = wk[1], w2 = wk[2];   // W[0][0..2] W[1][0] | W[1][1..2] W[2][0..1] | W[2][2]
        const float W[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
#pragma unroll
        for (int j = 0; j < kHeadFR + 2; ++j) {
            const float l = __shfl_up_sync(0xffffffffu, v[j], 1), rr = __shfl_down_sync(0xffffffffu, v[j], 1);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {      // input row j is tap row dy of output row j - dy
                const int i = j - dy;
                if (i >= 0 && i < kHeadFR) a[i] = fmaf(W[dy * 3], l, fmaf(W[dy * 3 + 1], v[j], fmaf(W[dy * 3 + 2], rr, a[i])));
            }
        }
    }
    const float bias = __ldg(p.bias);
    if (G > 1) {
#pragma unroll
        for (int i = 0; i < kHeadFR; ++i) sred[(g * kHeadFR + i) * 32 + lane] = a[i];
        __syncthreads();
        for (int t = threadIdx.x; t < kHeadFR * 32; t += blockDim.x) {      // (row i, lane): warps summed in warp order
            const int i = t >> 5, ln = t & 31;
            float s = 0.f;
            for (int k = 0; k < G; ++k) s += sred[(k * kHeadFR + i) * 32 + ln];
            const int ox = x0 - 1 + ln, oy = y0 + i;
            if (ox >= x0 && ox < x1 && oy < h) p.disp[(size_t)b * plane + oy * w + ox] = sigmoidf(s + bias);
        }
    } else if (owned) {
#pragma unroll
        for (int i = 0; i < kHeadFR; ++i)
            if (y0 + i < h) p.disp[(size_t)b * plane + (y0 + i) * w + cx] = sigmoidf(a[i] + bias);
    }
}

// Backward, one work item per (image, channel, chunk, strip).  With gz = g_disp * disp * (1 - disp):
//   g_x[c](q)     = sum over windows p and taps d with reflect(p + d) = q of W[c][d] * gz(p)
//   g_W[c][d]     = sum_p gz(p) * x_pad[c](p + d),      g_bias = sum_p gz(p)
// The reflection adjoint folds the padded ring back: because the row and the column index maps are independent it
// can be applied to the gz neighbourhood once per row -- l' = l + [cx == w-2] r, r' = r + [cx == 1] l, and the same
// between the rows above / below for q == 1 and q == h-2 -- after which every channel is nine plain FMAs.
template <bool WITH_GX>
__global__ void __launch_bounds__(32)
disp_head_bwd_kernel(const HeadParams p) {
    const int lane = threadIdx.x;
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    int item = blockIdx.x;
    const int strip = item % p.n_strips; item /= p.n_strips;
    const int chunk = item % p.n_chunks; item /= p.n_chunks;
    const int c = item % C;
    const int b = item / C;
    const int x0 = strip * kHeadTW, x1 = min(x0 + kHeadTW, w);
    const int y0 = chunk * kHeadTH, y1 = min(y0 + kHeadTH, h);
    const int cx = x0 - 1 + lane;
    const int rx = reflect1(clampi(cx, -1, w), w);
    const bool in_img = (cx >= 0) && (cx < w);
    const bool owned = (cx >= x0) && (cx < x1);
    const float f1 = (cx == 1) ? 1.f : 0.f, f2 = (cx == w - 2) ? 1.f : 0.f;
    float W[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) W[k] = __ldg(p.weight + c * 9 + k);
    const float* xc = p.x + ((size_t)b * C + c) * plane + rx;
    const float* gd = p.g_disp + (size_t)b * plane;
    const float* dd = p.disp + (size_t)b * plane;
    float* gx = WITH_GX ? p.g_x + ((size_t)b * C + c) * plane : nullptr;

    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    float accb = 0.f;
    // all rows of the item are requested before the first is consumed (the march used to wait for three loads per row)
    float xs[kHeadTH + 2], gs[kHeadTH + 2];
#pragma unroll
    for (int j = 0; j < kHeadTH + 2; ++j) {
        const int r = y0 - 1 + j;
        xs[j] = 0.f; gs[j] = 0.f;
        if (r <= y1) {
            xs[j] = __ldg(xc + reflect1(clampi(r, -1, h), h) * w);
            if (r >= 0 && r < h && in_img) {
                const float d = __ldg(dd + r * w + cx);
                gs[j] = __ldg(gd + r * w + cx) * d * (1.f - d);      // sigmoid backward
            }
        }
    }
    // rolling rows r-2 (U), r-1 (M): folded gz neighbourhood (l', v, r') and x_pad neighbourhood (l, v, r)
    float gU[3] = {0.f, 0.f, 0.f}, gM[3] = {0.f, 0.f, 0.f}, xU[3] = {0.f, 0.f, 0.f}, xM[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < kHeadTH + 2; ++j) {
        const int r = y0 - 1 + j;
        if (r > y1) break;
        const float xv = xs[j], gz = gs[j];
        const float xl = __shfl_up_sync(0xffffffffu, xv, 1), xr = __shfl_down_sync(0xffffffffu, xv, 1);
        const float gl = __shfl_up_sync(0xffffffffu, gz, 1), gr = __shfl_down_sync(0xffffffffu, gz, 1);
        const float gD[3] = {gl + f2 * gr, gz, gr + f1 * gl};
        const float xD[3] = {xl, xv, xr};
        const int q = r - 1;                                     // output row / window row completed by row r
        if (q >= y0 && q < y1) {
            if (WITH_GX) {
                const float g1 = (q == 1) ? 1.f : 0.f, g2 = (q == h - 2) ? 1.f : 0.f;
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float u = gU[k] + g2 * gD[k], dn = gD[k] + g1 * gU[k];
                    // neighbour k (0: left, 1: centre, 2: right) pairs with tap column 2 - k; rows above / at / below
                    // with tap rows 2 / 1 / 0
                    s = fmaf(W[6 + (2 - k)], u, fmaf(W[3 + (2 - k)], gM[k], fmaf(W[0 + (2 - k)], dn, s)));
                }
                if (owned) gx[q * w + cx] = s;
            }
            const float gq = owned ? gM[1] : 0.f;                // gz(q, cx)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                acc[0 + k] = fmaf(gq, xU[k], acc[0 + k]);        // x_pad row q-1
                acc[3 + k] = fmaf(gq, xM[k], acc[3 + k]);        // row q
                acc[6 + k] = fmaf(gq, xD[k], acc[6 + k]);        // row q+1
            }
            accb += gq;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { gU[k] = gM[k]; gM[k] = gD[k]; xU[k] = xM[k]; xM[k] = xD[k]; }
    }
    float* out = p.part + (size_t)blockIdx.x * 10;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) out[k] = v;
    }
    const float vb = warp_sum(accb);
    if (lane == 0) out[9] = vb;
}

// ---------------------------------------------------------------------------------------------------------------
// Tile kernels for the large heads (scales 0 and 1: few channels, many pixels).  The warp march above keeps one row of
// loads in flight per warp, which leaves the two big heads latency-bound at < 1 TB/s; here every thread owns a column
// of four output rows, addresses its 6 x 3 neighbourhood directly (L1 serves the horizontal and vertical overlap), and
// has the 18 loads of a channel (two channels forward) in flight at once.  Block = 32 x 8 threads = a 32 x 32 pixel tile.
// (Measured, scale 0: 4 rows x 2 channels per thread 53 us; 6 rows x 1 channel at 3 blocks per SM 89 us.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHeadTileRows = 4;    // output rows per thread, forward (a block of 32 x 8 threads covers 32 x 32 pixels)
constexpr int kHeadGxRows = 4;      // ... input-gradient kernel (32 x 32 pixels)

constexpr int kHeadTileH = 8 * kHeadTileRows, kHeadGxTileH = 8 * kHeadGxRows;   // tile heights

struct HeadTileIdx {
    int ro[kHeadTileRows + 2];   // row offsets (reflected, clamped into the image) of rows q0-1 .. q0+4
    int co[3];                   // column offsets of cx-1, cx, cx+1 (reflected)
    int cx, q0;
    bool col_ok;
};
__device__ __forceinline__ HeadTileIdx head_tile_idx(int h, int w) {
    HeadTileIdx t;
    t.cx = blockIdx.x * 32 + (threadIdx.x & 31);
    t.q0 = blockIdx.y * (8 * kHeadTileRows) + (threadIdx.x >> 5) * kHeadTileRows;
    t.col_ok = t.cx < w;
    const int cxc = min(t.cx, w - 1);
#pragma unroll
    for (int k = 0; k < 3; ++k) t.co[k] = reflect1(cxc + k - 1, w);
#pragma unroll
    for (int j = 0; j < kHeadTileRows + 2; ++j) t.ro[j] = reflect1(min(t.q0 - 1 + j, h), h) * w;   // row h reflects to h-2
    return t;
}

// forward: grid = (ceil(w/32), ceil(h/32), B); dynamic smem = C * 12 floats
__global__ void __launch_bounds__(256)
disp_head_fwd_tile_kernel(const HeadParams p) {
    PML_DYN_SMEM(float, sw);
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[(i / 9) * 12 + (i % 9)] = p.weight[i];
    __syncthreads();
    const HeadTileIdx t = head_tile_idx(h, w);
    const int b = blockIdx.z;
    const float* xc = p.x + (size_t)b * C * plane;
    float a[kHeadTileRows];
#pragma unroll
    for (int i = 0; i < kHeadTileRows; ++i) a[i] = 0.f;
#pragma unroll 2
    for (int c = 0; c < C; ++c, xc += plane) {      // two channels = 36 loads in flight per thread
        float v[kHeadTileRows + 2][3];
#pragma unroll
        for (int j = 0; j < kHeadTileRows + 2; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) v[j][k] = __ldg(xc + t.ro[j] + t.co[k]);
        const float4* wk = reinterpret_cast<const float4*>(sw + c * 12);
        const float4 w0 = wk[0], w1 = wk[1], w2 = wk[2];
        const float W[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
#pragma unroll
        for (int i = 0; i < kHeadTileRows; ++i)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int k = 0; k < 3; ++k) a[i] = fmaf(W[dy * 3 + k], v[i + dy][k], a[i]);
    }
    const float bias = __ldg(p.bias);
    if (t.col_ok) {
#pragma unroll
        for (int i = 0; i < kHeadTileRows; ++i)
            if (t.q0 + i < h) p.disp[(size_t)b * plane + (t.q0 + i) * w + t.cx] = sigmoidf(a[i] + bias);
    }
}

// ---- forward, staged: a 128 x 8 pixel tile per block (long row segments: a 32-column tile reads 128-160 bytes per DRAM
// page it opens and stalls at 2 TB/s however much is in flight -- measured with this very kernel at 32 x 32: 47 us);
// the 10 x 136 float window of four channels at a time travels to shared memory with 16-byte cp.async copies (no
// registers held while in flight), double-buffered: one stage loads while the other is consumed, five blocks per SM keep
// > 100 KB in flight per SM.  Needs w % 4 == 0 and a 16-byte aligned x (the window starts at column x0 - 4 so that every
// copy is aligned); reflection is an index map on the shared-memory reads -- the reflected row / column of a border
// tile lies inside its own window.
constexpr int kHeadStW = 128, kHeadStH = 8;      // tile; 128 threads = 32 groups of 4 columns x 4 groups of 2 rows
constexpr int kHeadStCB = 4, kHeadStRows = kHeadStH + 2, kHeadStCols = kHeadStW + 8;
constexpr int kHeadStFloats = kHeadStCB * kHeadStRows * kHeadStCols;     // one stage

__device__ __forceinline__ void head_cp_async16(float* smem_dst, const float* gsrc) {
#ifdef PML_HOST_EMU
    for (int i = 0; i < 4; ++i) smem_dst[i] = gsrc[i];
#else
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void head_cp_async_commit() {
#ifndef PML_HOST_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void head_cp_async_wait() {
#ifndef PML_HOST_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// grid = (ceil(w/128), ceil(h/8), B), 128 threads; dynamic smem = (2 stages + C * 12 weights) floats.
// A thread owns 4 columns x 2 rows: per channel 4 input rows x (one 128-bit + two 32-bit shared-memory loads), 72 FMAs and
// two 128-bit stores at the end -- the first version (one column x four rows per thread, 18 scalar loads per channel, a
// div/mod per staged chunk) was issue-bound at 84 % issue / 47 us.
__global__ void __launch_bounds__(128)
disp_head_fwd_staged_kernel(const HeadParams p) {
    PML_DYN_SMEM(float, sm);
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    float* sw = sm + 2 * kHeadStFloats;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[(i / 9) * 12 + (i % 9)] = p.weight[i];
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * kHeadStW, y0 = blockIdx.y * kHeadStH;
    const float* xb = p.x + (size_t)b * C * plane;
    const int wx0 = x0 - 4, wy0 = y0 - 1;                 // window origin
    // the 10 x 34 sixteen-byte chunks of one channel's window: this thread copies chunks t, t + 128, t + 256 of every channel
    constexpr int kChunks = kHeadStRows * (kHeadStCols / 4);
    int s_off[3], g_off[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int t = threadIdx.x + 128 * i;
        const int rr = t / (kHeadStCols / 4), g = t - rr * (kHeadStCols / 4);
        const int y = wy0 + rr, x = wx0 + 4 * g;
        const bool ok = t < kChunks && y >= 0 && y < h && x >= 0 && x < w;    // w % 4 == 0: a chunk is inside or outside as a whole
        s_off[i] = ok ? rr * kHeadStCols + 4 * g : -1;
        g_off[i] = y * w + x;
    }
    auto issue = [&](int c0, float* stage) {              // channels c0 .. c0+3 (those < C)
#pragma unroll
        for (int k = 0; k < kHeadStCB; ++k) {
            if (c0 + k < C) {
                const float* gc = xb + (size_t)(c0 + k) * plane;
                float* sc = stage + k * kHeadStRows * kHeadStCols;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (s_off[i] >= 0) head_cp_async16(sc + s_off[i], gc + g_off[i]);
            }
        }
        head_cp_async_commit();
    };
    const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int cx = x0 + 4 * cg, q0 = y0 + 2 * rg;         // first column / row of this thread's 4 x 2 outputs
    const int iv = 4 * cg + 4;                            // window column of cx (16-byte aligned)
    const int il = reflect1(min(cx, w) - 1, w) - wx0, ir = max(reflect1(min(cx + 4, w), w) - wx0, 0);
    int ro[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ro[j] = min(max(reflect1(min(q0 - 1 + j, h), h) - wy0, 0), kHeadStRows - 1) * kHeadStCols;
    float a[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[r][j] = 0.f;
    const int n_st = (C + kHeadStCB - 1) / kHeadStCB;
    issue(0, sm);
    for (int st = 0; st < n_st; ++st) {
        float* cur = sm + (st & 1) * kHeadStFloats;
        if (st + 1 < n_st) { issue((st + 1) * kHeadStCB, sm + ((st + 1) & 1) * kHeadStFloats); head_cp_async_wait<1>(); }
        else head_cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kHeadStCB; ++k) {
            const int c = st * kHeadStCB + k;
            if (c < C) {
                const float* sc = cur + k * kHeadStRows * kHeadStCols;
                float v[4][6];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 m = *reinterpret_cast<const float4*>(sc + ro[j] + iv);
                    v[j][0] = sc[ro[j] + il]; v[j][1] = m.x; v[j][2] = m.y; v[j][3] = m.z; v[j][4] = m.w; v[j][5] = sc[ro[j] + ir];
                }
                const float4* wk = reinterpret_cast<const float4*>(sw + c * 12);
                const float4 w0 = wk[0], w1 = wk[1], w2 = wk[2];
                const float W[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int d = 0; d < 3; ++d) a[r][j] = fmaf(W[dy * 3 + d], v[r + dy][j + d], a[r][j]);
            }
        }
        __syncthreads();     // the stage is free for the copy issued in the next iteration
    }
    const float bias = __ldg(p.bias);
    if (cx < w) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (q0 + r < h)
                *reinterpret_cast<float4*>(p.disp + (size_t)b * plane + (q0 + r) * w + cx) =
                    make_float4(sigmoidf(a[r][0] + bias), sigmoidf(a[r][1] + bias), sigmoidf(a[r][2] + bias), sigmoidf(a[r][3] + bias));
    }
}

// backward, input gradient: same tiling.  The folded gz neighbourhood of the four output rows (see the comment of
// disp_head_bwd_kernel) is computed once per thread and serves every channel: nine FMAs and one store per channel.
__global__ void __launch_bounds__(256)
disp_head_gx_tile_kernel(const HeadParams p) {
    PML_DYN_SMEM(float, sw);
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    for (int i = threadIdx.x; i < C * 9; i += blockDim.x) sw[(i / 9) * 12 + (i % 9)] = p.weight[i];
    __syncthreads();
    const int cx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int q0 = blockIdx.y * (8 * kHeadGxRows) + (threadIdx.x >> 5) * kHeadGxRows;
    const int b = blockIdx.z;
    const float* gd = p.g_disp + (size_t)b * plane;
    const float* dd = p.disp + (size_t)b * plane;
    // gz rows q0-1 .. q0+4, columns cx-1 .. cx+1, zero outside the image, column fold applied
    float g[kHeadGxRows + 2][3];
    const float f1 = (cx == 1) ? 1.f : 0.f, f2 = (cx == w - 2) ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < kHeadGxRows + 2; ++j) {
        const int r = q0 - 1 + j;
        float z[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int c = cx + k - 1;
            z[k] = 0.f;
            if (r >= 0 && r < h && c >= 0 && c < w) {
                const float d = __ldg(dd + r * w + c);
                z[k] = __ldg(gd + r * w + c) * d * (1.f - d);
            }
        }
        g[j][0] = z[0] + f2 * z[2]; g[j][1] = z[1]; g[j][2] = z[2] + f1 * z[0];
    }
    if (cx >= w) return;
    float* gx = p.g_x + (size_t)b * C * plane + cx;
#pragma unroll 2
    for (int c = 0; c < C; ++c, gx += plane) {
        const float4* wk = reinterpret_cast<const float4*>(sw + c * 12);
        const float4 w0 = wk[0], w1 = wk[1], w2 = wk[2];
        const float W[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
#pragma unroll
        for (int i = 0; i < kHeadGxRows; ++i) {
            const int q = q0 + i;
            const float g1 = (q == 1) ? 1.f : 0.f, g2 = (q == h - 2) ? 1.f : 0.f;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float u = g[i][k] + g2 * g[i + 2][k], dn = g[i + 2][k] + g1 * g[i][k];
                s = fmaf(W[6 + (2 - k)], u, fmaf(W[3 + (2 - k)], g[i + 1][k], fmaf(W[0 + (2 - k)], dn, s)));
            }
            if (q < h) gx[q * w] = s;
        }
    }
}

// backward, weight / bias gradient: the warp march, eight channels per warp.  Row r of x_pad meets the gz of the
// windows centred on rows r+1 (tap row 0), r (tap row 1) and r-1 (tap row 2), so the per-channel state is just the
// nine accumulators; the eight loads of row r+2 are issued before row r is consumed (two rows in flight per warp).
// grid = B * ceil(C/8) * n_chunks * n_strips warps; part layout as in disp_head_bwd_kernel.
constexpr int kHeadCB = 8;
__global__ void __launch_bounds__(32, 12)
disp_head_gw_kernel(const HeadParams p) {
    const int lane = threadIdx.x;
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    const int n_cg = (C + kHeadCB - 1) / kHeadCB;
    int item = blockIdx.x;
    const int strip = item % p.n_strips; item /= p.n_strips;
    const int chunk = item % p.n_chunks; item /= p.n_chunks;
    const int cg = item % n_cg;
    const int b = item / n_cg;
    const int c0 = cg * kHeadCB;
    const int x0 = strip * kHeadTW, x1 = min(x0 + kHeadTW, w);
    const int y0 = chunk * kHeadTH, y1 = min(y0 + kHeadTH, h);
    const int cx = x0 - 1 + lane;
    const int rx = reflect1(clampi(cx, -1, w), w);
    const bool owned = (cx >= x0) && (cx < x1);
    const float* xb = p.x + ((size_t)b * C + c0) * plane + rx;
    const float* gd = p.g_disp + (size_t)b * plane + min(max(cx, 0), w - 1);
    const float* dd = p.disp + (size_t)b * plane + min(max(cx, 0), w - 1);
    auto gz_row = [&](int r) {     // gz(r, cx) of the windows this item owns, else 0
        float z = 0.f;
        if (owned && r >= y0 && r < y1) {
            const float d = __ldg(dd + r * w);
            z = __ldg(gd + r * w) * d * (1.f - d);
        }
        return z;
    };
    auto x_row = [&](int r, float (&v)[kHeadCB]) {
        const int ry = reflect1(clampi(r, -1, h), h);
#pragma unroll
        for (int k = 0; k < kHeadCB; ++k) v[k] = (c0 + k < C) ? __ldg(xb + (size_t)k * plane + ry * w) : 0.f;
    };
    float acc[kHeadCB][9];
#pragma unroll
    for (int k = 0; k < kHeadCB; ++k)
#pragma unroll
        for (int d = 0; d < 9; ++d) acc[k][d] = 0.f;
    float accb = 0.f;
    float gA = 0.f, gB = gz_row(y0 - 1), gC = gz_row(y0);   // gz of rows r-1, r, r+1 for r = y0 - 1
    float va[kHeadCB], vb[kHeadCB], vc[kHeadCB];     // rows r, r+1 (in flight), r+2 (issued by the step)
    x_row(y0 - 1, va);
    x_row(y0, vb);
    auto step = [&](int r, const float (&v)[kHeadCB], float (&vn)[kHeadCB]) {
        x_row(r + 2, vn);
        const float gN = gz_row(r + 2);
#pragma unroll
        for (int k = 0; k < kHeadCB; ++k) {
            const float l = __shfl_up_sync(0xffffffffu, v[k], 1), rr = __shfl_down_sync(0xffffffffu, v[k], 1);
            acc[k][0] = fmaf(gC, l, acc[k][0]); acc[k][1] = fmaf(gC, v[k], acc[k][1]); acc[k][2] = fmaf(gC, rr, acc[k][2]);
            acc[k][3] = fmaf(gB, l, acc[k][3]); acc[k][4] = fmaf(gB, v[k], acc[k][4]); acc[k][5] = fmaf(gB, rr, acc[k][5]);
            acc[k][6] = fmaf(gA, l, acc[k][6]); acc[k][7] = fmaf(gA, v[k], acc[k][7]); acc[k][8] = fmaf(gA, rr, acc[k][8]);
        }
        accb += gB;
        gA = gB; gB = gC; gC = gN;
    };
#pragma unroll 1
    for (int r = y0 - 1; r <= y1; r += 3) {     // steps past y1 only meet gz = 0
        step(r, va, vc);
        step(r + 1, vb, va);
        step(r + 2, vc, vb);
    }
    const int per = p.n_chunks * p.n_strips;
    const int j = chunk * p.n_strips + strip;
#pragma unroll
    for (int k = 0; k < kHeadCB; ++k) {
        if (c0 + k >= C) break;
        float* out = p.part + ((size_t)(b * C + c0 + k) * per + j) * 10;
#pragma unroll
        for (int d = 0; d < 9; ++d) {
            const float v = warp_sum(acc[k][d]);
            if (lane == 0) out[d] = v;
        }
        if (k == 0) {
            const float vb2 = warp_sum(accb);
            if (lane == 0) out[9] = vb2;
        }
    }
}

// ---- backward, weight / bias gradient, staged like the forward: a block owns four channels of a 128-column stripe over
// 32 rows (four 128 x 8 tiles, double-buffered through shared memory), a thread the same 4 x 2 window positions in every
// tile: gz from two 128-bit loads per row, per channel 4 x (one 128-bit + two 32-bit) shared-memory loads and 72 FMAs into
// the nine accumulators of that channel; one block reduction at the end.  part layout as above with
// per = stripes x row groups (the launcher passes them in n_strips / n_chunks for disp_head_reduce_kernel).
constexpr int kHeadGwTiles = 4;      // tiles (of kHeadStH rows) per block
__global__ void __launch_bounds__(128)
disp_head_gw_staged_kernel(const HeadParams p) {
    PML_DYN_SMEM(float, sm);
    __shared__ float s_red[4][37];
    const int C = p.C, h = p.h, w = p.w, plane = h * w;
    const int n_cg = (C + kHeadStCB - 1) / kHeadStCB;
    const int sx = blockIdx.x, gy = blockIdx.y;
    const int cgp = blockIdx.z % n_cg, b = blockIdx.z / n_cg;
    const int c0 = cgp * kHeadStCB;
    const int x0 = sx * kHeadStW, wx0 = x0 - 4;
    const float* xb = p.x + ((size_t)b * C + c0) * plane;
    constexpr int kChunks = kHeadStRows * (kHeadStCols / 4);
    int s_off[3], g_col[3], g_row[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int t = threadIdx.x + 128 * i;
        const int rr = t / (kHeadStCols / 4), g = t - rr * (kHeadStCols / 4);
        const int x = wx0 + 4 * g;
        s_off[i] = (t < kChunks && x >= 0 && x < w) ? rr * kHeadStCols + 4 * g : -1;
        g_col[i] = x; g_row[i] = rr;
    }
    auto issue = [&](int y0, float* stage) {              // window rows y0-1 .. y0+8 of the four channels
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int y = y0 - 1 + g_row[i];
            if (s_off[i] >= 0 && y >= 0 && y < h) {
#pragma unroll
                for (int k = 0; k < kHeadStCB; ++k)
                    if (c0 + k < C) head_cp_async16(stage + k * kHeadStRows * kHeadStCols + s_off[i], xb + (size_t)k * plane + y * w + g_col[i]);
            }
        }
        head_cp_async_commit();
    };
    const int cg = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int cx = x0 + 4 * cg;
    const int iv = 4 * cg + 4;
    const int il = reflect1(min(cx, w) - 1, w) - wx0, ir = max(reflect1(min(cx + 4, w), w) - wx0, 0);
    float acc[kHeadStCB][9];
#pragma unroll
    for (int k = 0; k < kHeadStCB; ++k)
#pragma unroll
        for (int d = 0; d < 9; ++d) acc[k][d] = 0.f;
    float accb = 0.f;
    const int yb = gy * (kHeadGwTiles * kHeadStH);
    const float* gd = p.g_disp + (size_t)b * plane;
    const float* dd = p.disp + (size_t)b * plane;
    issue(yb, sm);
    for (int st = 0; st < kHeadGwTiles; ++st) {
        const int y0 = yb + st * kHeadStH, wy0 = y0 - 1;
        float* cur = sm + (st & 1) * kHeadStFloats;
        if (st + 1 < kHeadGwTiles) { issue(y0 + kHeadStH, sm + ((st + 1) & 1) * kHeadStFloats); head_cp_async_wait<1>(); }
        else head_cp_async_wait<0>();
        // gz of this thread's 4 x 2 window positions (0 outside the image)
        const int q0 = y0 + 2 * rg;
        float gz[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = g4;
            if (cx < w && q0 + r < h) {
                g4 = __ldg(reinterpret_cast<const float4*>(gd + (q0 + r) * w + cx));
                d4 = __ldg(reinterpret_cast<const float4*>(dd + (q0 + r) * w + cx));
            }
            gz[r][0] = g4.x * d4.x * (1.f - d4.x); gz[r][1] = g4.y * d4.y * (1.f - d4.y);
            gz[r][2] = g4.z * d4.z * (1.f - d4.z); gz[r][3] = g4.w * d4.w * (1.f - d4.w);
            accb += (gz[r][0] + gz[r][1]) + (gz[r][2] + gz[r][3]);
        }
        int ro[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ro[j] = min(max(reflect1(min(q0 - 1 + j, h), h) - wy0, 0), kHeadStRows - 1) * kHeadStCols;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kHeadStCB; ++k) {
            const float* sc = cur + k * kHeadStRows * kHeadStCols;
            float v[4][6];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 m = *reinterpret_cast<const float4*>(sc + ro[j] + iv);
                v[j][0] = sc[ro[j] + il]; v[j][1] = m.x; v[j][2] = m.y; v[j][3] = m.z; v[j][4] = m.w; v[j][5] = sc[ro[j] + ir];
            }
            // positions outside the image are skipped, not multiplied by gz = 0: their taps may lie in window cells
            // no copy has written (0 * stale shared memory)
#pragma unroll
            for (int r = 0; r < 2; ++r)
                if (cx < w && q0 + r < h) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int d = 0; d < 3; ++d) acc[k][dy * 3 + d] = fmaf(gz[r][j], v[r + dy][j + d], acc[k][dy * 3 + d]);
                }
        }
        __syncthreads();
    }
    // block reduction in a fixed order: lanes by shuffles, then the four warps
#pragma unroll
    for (int k = 0; k < kHeadStCB; ++k)
#pragma unroll
        for (int d = 0; d < 9; ++d) {
            const float v = warp_sum(acc[k][d]);
            if (cg == 0) s_red[rg][k * 9 + d] = v;
        }
    {
        const float v = warp_sum(accb);
        if (cg == 0) s_red[rg][36] = v;
    }
    __syncthreads();
    const int per = gridDim.x * gridDim.y, j = gy * gridDim.x + sx;
    if (threadIdx.x < 37) {
        const float v = (s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + (s_red[2][threadIdx.x] + s_red[3][threadIdx.x]);
        if (threadIdx.x < 36) {
            const int k = threadIdx.x / 9, d = threadIdx.x - 9 * k;
            if (c0 + k < C) p.part[((size_t)(b * C + c0 + k) * per + j) * 10 + d] = v;
        } else if (cgp == 0) {
            p.part[((size_t)(b * C) * per + j) * 10 + 9] = v;
        }
    }
}

// fixed-order reduction of the partials: block c sums the 9 weight gradients of channel c over (image, chunk, strip),
// one warp per tap (lanes stride over the items, lane 0 adds the 32 partial sums in lane order: the order is fixed); block C sums the bias
// gradient (taken from the items of channel 0).  288 threads.
__global__ void __launch_bounds__(288)
disp_head_reduce_kernel(const HeadParams p, float* __restrict__ g_weight, float* __restrict__ g_bias) {
    __shared__ double s_acc[9][32];
    const int c = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = p.n_chunks * p.n_strips;
    const bool bias = (c == p.C);
    const int ch = bias ? 0 : c, col = bias ? 9 : k;
    double a = 0.0;
    if (!bias || k == 0) {
        const int n = p.B * per;
        for (int i0 = lane; i0 < n; i0 += 32 * 8) {      // eight independent loads in flight per lane, added in index order
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u;
                const int b = i / per, j = i - b * per;
                v[u] = (i < n) ? p.part[((size_t)(b * p.C + ch) * per + j) * 10 + col] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) a += (double)v[u];
        }
    }
    s_acc[k][lane] = a;
    __syncthreads();
    if (lane == 0 && (!bias || k == 0)) {
        double t = 0.0;
        for (int m = 0; m < 32; ++m) t += s_acc[k][m];
        if (bias) g_bias[0] = (float)t;
        else g_weight[c * 9 + k] = (float)t;
    }
}

}  // namespace pml
