// C-ABI entry points of libpml.so (declared in include/pml.h).  Host side only validates,
// plans the launch shape, carves the caller-provided workspace and enqueues kernels.
#include "pml_common.cuh"
#include "pml_photometric.cuh"
#include "pml_sweep.cuh"
#include "pml_prep.cuh"
#include "pml_smooth.cuh"
#include "pml_layers.cuh"
#include "pml_metrics.cuh"
#include "pml_mask.cuh"
#include "pml_pyramid.cuh"
#include "pml_disphead.cuh"

#include <math.h>
#include <stdlib.h>
#include <atomic>
#ifndef PML_HOST_EMU
#include <nvtx3/nvToolsExt.h>   // header-only: a no-op unless a profiler (nsys / ncu) injects its library
#endif

namespace {

using namespace pml;

constexpr int kNumSM = 148;   // B200

inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

struct Plan {
    bool sweep;   // warp-strip kernel (pml_sweep.cuh) instead of the CTA-strip kernel
    bool two_sweeps;   // S > 2 or predictive mask: forward sweep per frame pair, selection, adjoint sweep
    int NT, TW, TH, n_strips, n_chunks, cta_per_pass, n_cta, part_stride;
    int TH_fwd, n_chunks_fwd;   // chunk geometry of forward-only sweeps that write no partials
    int n_id;
    int smooth_blocks[PML_MAX_PASSES], smooth_off[PML_MAX_PASSES], smooth_total;
    int max_chunks;
    size_t off_identity, off_part, off_mean, off_meanpart, off_imagepart, off_smooth, off_rp, off_argmin, off_presel, total;
};

inline void pml_event_record(void* ev, cudaStream_t st) {
#ifndef PML_HOST_EMU
    cudaEventRecord(static_cast<cudaEvent_t>(ev), st);
#else
    (void)ev; (void)st;
#endif
}

// Developer knobs (A/B experiments on the GPU box).  Read ONCE per process, never on the per-call path.
struct Knobs {
    int th, id_th, smem_pad, pyramid_scalar;
    Knobs() {
        auto geti = [](const char* name, int dflt) { const char* s = getenv(name); return (s && *s) ? atoi(s) : dflt; };
        th = geti("PML_TH", 0); id_th = geti("PML_ID_TH", 16); smem_pad = geti("PML_SMEM_PAD", 0);
        pyramid_scalar = geti("PML_PYRAMID_SCALAR", 0);
    }
};
const Knobs& knobs() { static const Knobs k; return k; }

// NVTX range around a phase of an entry point (tracing); shows up in nsys / ncu timelines
struct NvtxRange {
#ifndef PML_HOST_EMU
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
#else
    explicit NvtxRange(const char*) {}
#endif
};

// RAII: make the device that owns `ptr` current for the duration of an entry point.  The reference trainers keep
// their tensors on cuda:1 / cuda:3 and never call set_device (trainer.py:44,67); the legacy default stream handle
// is the same for every device, so without this the kernels would be enqueued on the wrong GPU.
struct DeviceGuard {
    int prev = -1, dev = -1;
    bool switched = false;
    explicit DeviceGuard(const void* ptr) {
#ifndef PML_HOST_EMU
        cudaPointerAttributes a;
        if (!ptr || cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return; }
        if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return;   // host pointer: the launch will fault loudly
        dev = a.device;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
#else
        (void)ptr;
#endif
    }
    ~DeviceGuard() {
#ifndef PML_HOST_EMU
        if (switched) cudaSetDevice(prev);
#endif
    }
    // true if `other` lives on another device than the guarded pointer (unknown pointers pass)
    bool foreign(const void* other) const {
#ifndef PML_HOST_EMU
        cudaPointerAttributes a;
        if (dev < 0 || !other || cudaPointerGetAttributes(&a, other) != cudaSuccess) { cudaGetLastError(); return false; }
        return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) && a.device != dev;
#else
        (void)other; return false;
#endif
    }
};

int validate(const pml_problem* p, bool grad) {
    if (!p) return PML_ERR_INVALID;
    if (p->B < 1 || p->H < 4 || p->W < 4 || p->n_pass < 1) return PML_ERR_INVALID;
    if (p->S < 1 || p->S > PML_MAX_SOURCES || p->n_pass > PML_MAX_PASSES) return PML_ERR_UNSUPPORTED;
    // the kernels index every tensor with 32-bit element offsets; the largest is [B, max(3, S), H, W]
    if ((long long)p->B * (p->S > 3 ? p->S : 3) * p->H * p->W >= (1LL << 31)) return PML_ERR_UNSUPPORTED;
    if (!p->target || !p->K || !p->inv_K || !p->losses) return PML_ERR_INVALID;
    if (grad && (!p->grad_T || !p->grad_disp_const)) return PML_ERR_INVALID;
    for (int f = 0; f < p->S; ++f)
        if (!p->sources[f] || !p->T[f]) return PML_ERR_INVALID;
    if (const pml_segments* sg = p->segments) {   // chunked batch: every chunk of every tensor in use must be there
        if (sg->n_seg < 1 || sg->n_seg > PML_MAX_SEGMENTS || sg->seg_size < 1 || sg->n_seg * sg->seg_size != p->B) return PML_ERR_INVALID;
        if (p->flags & PML_FLAG_KERNEL_CTA) return PML_ERR_UNSUPPORTED;
        for (int j = 0; j < sg->n_seg; ++j) {
            if (!sg->target[j] || !sg->K[j] || !sg->inv_K[j]) return PML_ERR_INVALID;
            for (int f = 0; f < p->S; ++f) if (!sg->sources[f][j]) return PML_ERR_INVALID;
            for (int i = 0; i < p->n_pass; ++i) if (!sg->smooth_color[i][j]) return PML_ERR_INVALID;
        }
    }
    for (int i = 0; i < p->n_pass; ++i) {
        const pml_pass& ps = p->pass[i];
        if (!ps.disp || !ps.smooth_color || ps.hd < 2 || ps.wd < 2) return PML_ERR_INVALID;
        if (grad && !ps.grad_disp) return PML_ERR_INVALID;
        if ((ps.frame_weight != nullptr) != (p->pass[0].frame_weight != nullptr)) return PML_ERR_INVALID;
        if (ps.frame_weight && !(p->flags & PML_FLAG_NO_AUTOMASK)) return PML_ERR_UNSUPPORTED;   // trainer.py:556 / :571
        if (p->H % ps.hd != 0 || p->W % ps.wd != 0) return PML_ERR_UNSUPPORTED;
        int k = p->H / ps.hd;
        if (p->W / ps.wd != k || (k & (k - 1)) != 0 || k > 64) return PML_ERR_UNSUPPORTED;
    }
    return PML_OK;
}

// Which fused kernel serves this problem.  The warp-strip sweep (pml_sweep.cuh) packs two source
// frames of a pixel into fp32x2 values; more frames are swept pair by pair (forward sweeps of all
// pairs but the last, the last pair's sweep selects over everything and carries its adjoint, then
// the adjoint sweeps of the others).  It is ahead of the first-generation CTA-strip kernel
// (pml_photometric.cuh, templated on S <= 4) everywhere -- S=3 at B=8, 320x1024: 2.5 vs 2.8 ms;
// S=4 at the headline size: 1.4 vs 3.9 ms -- which is kept as an independent implementation for the
// cross-check test.  PML_FLAG_KERNEL_CTA selects it.
bool use_sweep(const pml_problem* p) {
    if (p->pass[0].frame_weight) return true;   // predictive mask: handled by the MODE 3 / MODE 2 sweeps
    if (p->S > 4) return true;                  // the CTA-strip kernel is instantiated for S <= 4
    return !(p->flags & PML_FLAG_KERNEL_CTA);
}

Plan make_plan(const pml_problem* p, bool grad) {
    Plan pl;
    pl.sweep = use_sweep(p);
    // strip width: the candidate whose strips waste the fewest columns
    int best_nt = 64;
    double best_cost = 1e30;
    const int cands[3] = {64, 96, 128};
    for (int c = 0; c < 3; ++c) {
        int nt = cands[c], tw = nt - 4;
        int ns = (p->W + tw - 1) / tw;
        double cost = (double)ns * nt / p->W;
        if (cost < best_cost - 1e-9) { best_cost = cost; best_nt = nt; }
    }
    pl.NT = best_nt;
    if (pl.NT < 64) pl.NT = 64;
    if (pl.NT > 128) pl.NT = 128;
    pl.NT = (pl.NT / 32) * 32;
    pl.TW = pl.NT - 4;
    if (pl.sweep) { pl.NT = pml::kSweepWarps * 32; pl.TW = pml::kSweepTW; }
    pl.n_strips = (p->W + pl.TW - 1) / pl.TW;
    // strip height: the tallest chunk that still yields ~6 CTAs per SM (4 halo rows per chunk);
    // warp strips: ~1.5-2 waves of the 8 warps an SM holds (5 halo row steps per chunk; measured on
    // B200 at the headline size: TH 96 beats 48 and 32)
    // The forward-only instantiations need fewer registers (12 resident warps per SM instead of 8) and
    // get their own, shorter chunks; the launches of one call that write no partials (sweep mode 1) use
    // the forward geometry even inside a forward+backward call.
    int per_chunk = p->n_pass * p->B * pl.n_strips;
    auto chunk_rows = [&](int target_items) {
        int want = (target_items + per_chunk - 1) / per_chunk;
        if (want < 1) want = 1;
        int th = (p->H + want - 1) / want;
        if (th < 16) th = 16;
        th = ((th + 7) / 8) * 8;
        if (knobs().th > 0) th = knobs().th;
        if (th > p->H) th = p->H;
        if (th < 4) th = 4;
        return th;
    };
    // ~1.5 work items per resident-warp slot: 8 warps/SM for the forward+backward sweeps (255 registers),
    // 12 for the generic forward-only one (168), 16 for the forward-only default configuration (128)
    bool common_fwd = p->S == 2 && !(p->flags & (PML_FLAG_NO_AUTOMASK | PML_FLAG_AVG_REPROJ));
    for (int i = 0; i < p->n_pass; ++i)
        common_fwd = common_fwd && !p->pass[i].noise && !p->pass[i].depth && !p->pass[i].warped && !p->pass[i].frame_weight;
    // Warp strips: one warp per work item and `slots` resident warps per SM (8 for the forward+backward sweeps at 255
    // registers, 12 / 16 for the forward-only ones).  All items cost the same, so the sweep takes
    // ceil(items / (SMs * slots)) waves of (TH + 5) row steps (5 halo steps per chunk): pick the chunk count that
    // minimises that; among equals the shorter chunks (more items: better balance, more parallelism for small
    // problems).  Measured at the headline size (profiles/r02_ab_timings.txt): 1 chunk 0.3604 ms, 2 chunks 0.3672, 3: 0.3705.
    // More than two frames: the sweeps with adjoint are the last pair's (mode 3: 8 slots, or 12 for a lone last frame
    // at ~0.7 of a pair's cost per row step) and one per earlier pair (mode 2); they share the chunk geometry because
    // they share the partials.  Measured at BASELINE config 3 (37 x 8 x 4 = 1184 items per chunk = exactly one full
    // wave): 1 chunk 1.80 ms, 2: 1.65, 3: 1.63, 5: 1.64 -- full waves hide no imbalance, which the extra quarter wave
    // in the multi-frame cost stands for.
    const int n_pairs = (p->S + 1) / 2;
    const bool lone_last = (p->S > 2) && (p->S & 1);
    auto sweep_rows = [&](int slots, bool multi = false) {
        if (knobs().th > 0) return knobs().th > p->H ? p->H : (knobs().th < 4 ? 4 : knobs().th);
        const long long cap = (long long)kNumSM * slots;
        double best_cost = -1;
        int best_th = p->H;
        for (int chunks = 1; chunks <= 64; ++chunks) {
            int th = (p->H + chunks - 1) / chunks;
            th = ((th + 7) / 8) * 8;
            if (th > p->H) th = p->H;
            if (th < (multi ? 96 : 16) && chunks > 1) break;   // multi-frame, 192x640, 5 frames: 96 rows 1.445 ms, 64 rows 1.529
            const int n_ch = (p->H + th - 1) / th;
            const long long items = (long long)per_chunk * n_ch;
            double waves = (double)((items + cap - 1) / cap);
            if (multi) {
                const long long cap3 = (long long)kNumSM * (lone_last ? kSweepLoneCtas : slots);
                waves = (n_pairs - 1) * (waves + 0.25) + (lone_last ? 0.7 : 1.0) * ((double)((items + cap3 - 1) / cap3) + 0.25);
            }
            const double cost = waves * (th + 5);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_th = th; }
        }
        return best_th;
    };
    pl.TH_fwd = pl.sweep ? sweep_rows(kSweepFwdCtas) : chunk_rows(kNumSM * 6);
    pl.TH = (pl.sweep && !grad) ? pl.TH_fwd : (pl.sweep ? sweep_rows(p->S == 1 ? kSweepLoneCtas : kSweepPairCtas, p->S > 2) : chunk_rows(kNumSM * 6));
    pl.n_chunks = (p->H + pl.TH - 1) / pl.TH;
    pl.n_chunks_fwd = (p->H + pl.TH_fwd - 1) / pl.TH_fwd;
    pl.cta_per_pass = p->B * pl.n_chunks * pl.n_strips;
    pl.n_cta = pl.cta_per_pass * p->n_pass;
    pl.part_stride = ((1 + 12 * p->S) + 3) & ~3;
    const bool automask = !(p->flags & PML_FLAG_NO_AUTOMASK);
    const bool avg = (p->flags & PML_FLAG_AVG_REPROJ) != 0;
    pl.n_id = automask ? (avg ? 1 : p->S) : 0;
    pl.smooth_total = 0;
    for (int i = 0; i < p->n_pass; ++i) {
        pl.smooth_blocks[i] = pl.sweep ? ((p->pass[i].wd + pml::kPrepTW - 1) / pml::kPrepTW) *
                                         ((p->pass[i].hd + pml::kSmoothTH - 1) / pml::kSmoothTH)
                                       : (p->pass[i].hd * p->pass[i].wd + 255) / 256;
        pl.smooth_off[i] = pl.smooth_total;
        pl.smooth_total += pl.smooth_blocks[i] * p->B;
    }
    size_t off = 0;
    pl.off_identity = off; off = align16(off + (size_t)p->B * pl.n_id * p->H * p->W * sizeof(float));
    pl.off_part = off;     off = align16(off + (size_t)pl.n_cta * pl.part_stride * sizeof(float));
    pl.off_mean = off;     off = align16(off + (size_t)p->n_pass * p->B * sizeof(float));
    pl.max_chunks = 1;
    for (int i = 0; i < p->n_pass; ++i) {
        int c = (p->pass[i].hd * p->pass[i].wd + pml::kMeanChunk - 1) / pml::kMeanChunk;
        if (c > pl.max_chunks) pl.max_chunks = c;
    }
    pl.off_meanpart = off; off = align16(off + (size_t)p->n_pass * p->B * pl.max_chunks * sizeof(float));
    pl.off_imagepart = off; off = align16(off + (size_t)p->n_pass * p->B * 4 * sizeof(float));
    pl.off_smooth = off;   off = align16(off + (size_t)pl.smooth_total * 3 * sizeof(float));
    pl.off_rp = pl.off_argmin = pl.off_presel = off;
    pl.two_sweeps = pl.sweep && (p->S > 2 || p->pass[0].frame_weight != nullptr);
    if (pl.two_sweeps) {   // pair-by-pair sweep: reprojection losses of all frames + a selection map
        off = align16(off + (size_t)p->n_pass * p->S * p->B * p->H * p->W * sizeof(float));
        pl.off_argmin = off;
        off = align16(off + (size_t)p->n_pass * p->B * p->H * p->W);
        pl.off_presel = off;
        off = align16(off + (size_t)p->n_pass * p->B * p->H * p->W);
    }
    pl.total = off;
    return pl;
}

template <int S, bool GRAD, bool SSIM>
int launch_photo(const PhotoParams& pp, int n_cta, int NT, int low_cells, cudaStream_t st) {
    size_t smem = photometric_smem_bytes<S>(NT, GRAD, low_cells);
    if (smem > 227 * 1024) return PML_ERR_UNSUPPORTED;
    // Opt in to > 48 KB dynamic shared memory.  The high-water mark is the only process-wide state
    // of the library: monotonic, so concurrent callers can at worst repeat an idempotent call, and
    // it keeps attribute changes out of CUDA-graph capture once the shape has been seen.
    static std::atomic<int> configured{0};
    if ((int)smem > configured.load(std::memory_order_relaxed)) {
        if (cudaFuncSetAttribute(photometric_kernel<S, GRAD, SSIM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem) != cudaSuccess)
            return PML_ERR_CUDA;
        configured.store((int)smem, std::memory_order_relaxed);
    }
    PML_LAUNCH((photometric_kernel<S, GRAD, SSIM>), dim3(n_cta), dim3(NT), smem, st, pp);
    return PML_OK;
}

// the instantiation without run-time flag tests: default flags, in-kernel noise, no by-product stores
bool sweep_common(const PhotoParams& pp) {
    bool emit = false;   // by-products requested by any pass (never in sweep mode 2)
    for (int i = 0; i < pp.n_pass; ++i) emit = emit || pp.pass[i].depth != nullptr || pp.pass[i].warped != nullptr;
    if (pp.mode == 2) emit = false;
    bool common = !emit && (pp.mode != 0 ? pp.pair_n >= 1 : (pp.pair_n == 2 && pp.S == 2)) &&
                  !(pp.flags & (PML_FLAG_NO_AUTOMASK | PML_FLAG_AVG_REPROJ));
    for (int i = 0; i < pp.n_pass; ++i) common = common && pp.pass[i].noise == nullptr && pp.pass[i].fw == nullptr;
    return common;
}

template <bool GRAD, bool SSIM>
int launch_sweep(const PhotoParams& pp, cudaStream_t st) {
    // a single frame (the last one of an odd count; stereo-only training: S = 1): scalar instantiation, shorter ring
    const bool lone = (pp.mode == 3 || pp.mode == 0) && pp.pair_n == 1;
    const size_t smem = sweep_smem_bytes(GRAD && pp.mode != 1, !lone) + (size_t)(GRAD ? knobs().smem_pad : 0);   // < 48 KB: no opt-in needed
    if (pp.B > 65535) return PML_ERR_UNSUPPORTED;
    const dim3 grid(pp.n_chunks * pp.n_strips, pp.B, pp.n_pass);
    bool emit = false;   // by-products requested by any pass (never in sweep mode 2)
    for (int i = 0; i < pp.n_pass; ++i) emit = emit || pp.pass[i].depth != nullptr || pp.pass[i].warped != nullptr;
    const dim3 blk(kSweepWarps * 32);
    if (pp.mode == 2) emit = false;   // the adjoint sweeps of earlier pairs never write by-products
    const bool common = sweep_common(pp);
    if (pp.mode == 0 && !lone) {
        if (common)    PML_LAUNCH((sweep_kernel<GRAD, SSIM, 0, false, true>), grid, blk, smem, st, pp);
        else if (emit) PML_LAUNCH((sweep_kernel<GRAD, SSIM, 0, true>), grid, blk, smem, st, pp);
        else           PML_LAUNCH((sweep_kernel<GRAD, SSIM, 0, false>), grid, blk, smem, st, pp);
    } else if (lone && pp.mode == 0) {
        if (emit)      PML_LAUNCH((sweep_kernel<GRAD, SSIM, 0, true, false, false>), grid, blk, smem, st, pp);
        else           PML_LAUNCH((sweep_kernel<GRAD, SSIM, 0, false, false, false>), grid, blk, smem, st, pp);
    } else if (lone) {
        if (common)    PML_LAUNCH((sweep_kernel<GRAD, SSIM, 3, false, true, false>), grid, blk, smem, st, pp);
        else if (emit) PML_LAUNCH((sweep_kernel<GRAD, SSIM, 3, true, false, false>), grid, blk, smem, st, pp);
        else           PML_LAUNCH((sweep_kernel<GRAD, SSIM, 3, false, false, false>), grid, blk, smem, st, pp);
    } else if (pp.mode == 3) {
        if (common)    PML_LAUNCH((sweep_kernel<GRAD, SSIM, 3, false, true>), grid, blk, smem, st, pp);
        else if (emit) PML_LAUNCH((sweep_kernel<GRAD, SSIM, 3, true>), grid, blk, smem, st, pp);
        else           PML_LAUNCH((sweep_kernel<GRAD, SSIM, 3, false>), grid, blk, smem, st, pp);
    } else if (pp.mode == 1) {
        if (common)    PML_LAUNCH((sweep_kernel<false, SSIM, 1, false, true>), grid, blk, smem, st, pp);
        else if (emit) PML_LAUNCH((sweep_kernel<false, SSIM, 1, true>), grid, blk, smem, st, pp);
        else           PML_LAUNCH((sweep_kernel<false, SSIM, 1, false>), grid, blk, smem, st, pp);
    } else {
        if (common)    PML_LAUNCH((sweep_kernel<true, SSIM, 2, false, true>), grid, blk, smem, st, pp);
        else           PML_LAUNCH((sweep_kernel<true, SSIM, 2, false>), grid, blk, smem, st, pp);
    }
    return PML_OK;
}

template <bool GRAD, bool SSIM>
int dispatch_S(int S, const PhotoParams& pp, int n_cta, int NT, int low_cells, cudaStream_t st) {
    switch (S) {
        case 1: return launch_photo<1, GRAD, SSIM>(pp, n_cta, NT, low_cells, st);
        case 2: return launch_photo<2, GRAD, SSIM>(pp, n_cta, NT, low_cells, st);
        case 3: return launch_photo<3, GRAD, SSIM>(pp, n_cta, NT, low_cells, st);
        case 4: return launch_photo<4, GRAD, SSIM>(pp, n_cta, NT, low_cells, st);
    }
    return PML_ERR_UNSUPPORTED;
}

int run_loss(const pml_problem* p, void* ws, size_t ws_bytes, cudaStream_t st, bool grad) {
    int rc = validate(p, grad);
    if (rc != PML_OK) return rc;
    NvtxRange nvtx_call(grad ? "pml_loss_forward_backward" : "pml_loss_forward");
    DeviceGuard guard(p->target);
    if (guard.foreign(p->pass[0].disp) || guard.foreign(p->losses) || guard.foreign(p->sources[0]) || guard.foreign(ws))
        return PML_ERR_INVALID;   // tensors of one call must live on one device
    Plan pl = make_plan(p, grad);
    if (!ws || ws_bytes < pl.total) return PML_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(ws) & 15) != 0) return PML_ERR_INVALID;
    char* base = static_cast<char*>(ws);
    float* identity = reinterpret_cast<float*>(base + pl.off_identity);
    float* part = reinterpret_cast<float*>(base + pl.off_part);
    float* mean = reinterpret_cast<float*>(base + pl.off_mean);
    float* spart = reinterpret_cast<float*>(base + pl.off_smooth);
    float* meanpart = reinterpret_cast<float*>(base + pl.off_meanpart);
    float* imagepart = reinterpret_cast<float*>(base + pl.off_imagepart);
    const bool ssim = !(p->flags & PML_FLAG_NO_SSIM);

    // 1-2. smoothness term (writes grad_disp first; the photometric kernel adds onto it)
    SmoothParams sp;
    const pml_segments* sg = p->segments;
    sp.n_seg = sg ? sg->n_seg : 0; sp.seg_size = sg ? sg->seg_size : 0;
    if (sg)
        for (int i = 0; i < p->n_pass; ++i)
            for (int j = 0; j < sg->n_seg; ++j) sp.color_c[i].p[j] = sg->smooth_color[i][j];
    sp.B = p->B; sp.n_pass = p->n_pass; sp.disp_mean = mean; sp.part = spart;
    sp.mean_part = meanpart; sp.max_chunks = pl.max_chunks;
    for (int i = 0; i < p->n_pass; ++i) {
        const pml_pass& ps = p->pass[i];
        sp.pass[i].disp = ps.disp; sp.pass[i].color = ps.smooth_color;
        sp.pass[i].grad_disp = grad ? ps.grad_disp : nullptr;
        sp.pass[i].h = ps.hd; sp.pass[i].w = ps.wd;
        sp.pass[i].blocks = pl.smooth_blocks[i]; sp.pass[i].block_off = pl.smooth_off[i];
        sp.pass[i].weight = ps.smooth_weight;
    }
    PML_LAUNCH(disp_sum_kernel, dim3(pl.max_chunks, p->B, p->n_pass), dim3(256), 0, st, sp);
    // warp-strip path with automask: the smoothness sweep shares one launch with the identity sweep of
    // the first frame pair (prep_kernel); otherwise it is launched on its own
    const bool merged_prep = pl.sweep && pl.n_id > 0;
    if (pl.sweep && !merged_prep) {
        if (grad) PML_LAUNCH(smooth_sweep_kernel<true>, dim3(pl.smooth_total), dim3(32), 0, st, sp);
        else      PML_LAUNCH(smooth_sweep_kernel<false>, dim3(pl.smooth_total), dim3(32), 0, st, sp);
    } else if (!pl.sweep) {
        if (grad) PML_LAUNCH(smooth_kernel<true>, dim3(pl.smooth_total), dim3(256), 0, st, sp);
        else      PML_LAUNCH(smooth_kernel<false>, dim3(pl.smooth_total), dim3(256), 0, st, sp);
    }

    // 3. identity reprojection losses (automask), once for all passes
    if (pl.n_id > 0 && pl.sweep) {
        for (int fa = 0; fa < p->S; fa += 2) {   // one launch per pair of source frames
            IdentityParams ip;
            const int pair_n = (fa + 1 < p->S) ? 2 : 1;
            ip.target = p->target; ip.src0 = p->sources[fa]; ip.src1 = p->sources[fa + pair_n - 1];
            ip.n_seg = sp.n_seg; ip.seg_size = sp.seg_size;
            if (sg)
                for (int j = 0; j < sg->n_seg; ++j) {
                    ip.target_c.p[j] = sg->target[j]; ip.src0_c.p[j] = sg->sources[fa][j];
                    ip.src1_c.p[j] = sg->sources[fa + pair_n - 1][j];
                }
            ip.out = identity; ip.B = p->B; ip.H = p->H; ip.W = p->W; ip.S = pair_n;
            ip.avg = (p->flags & PML_FLAG_AVG_REPROJ) ? 1 : 0;
            ip.n_out = pl.n_id; ip.plane_off = ip.avg ? 0 : fa; ip.accumulate = (ip.avg && fa > 0) ? 1 : 0;
            ip.inv_total = 1.0f / (float)p->S;
            ip.TH = knobs().id_th;
            ip.n_strips = (p->W + kPrepTW - 1) / kPrepTW;
            ip.n_chunks = (p->H + ip.TH - 1) / ip.TH;
            const dim3 g(ip.n_chunks * ip.n_strips, p->B);
            if (fa == 0) {
                const int n_items = ip.n_chunks * ip.n_strips * p->B;
                const dim3 gm(n_items + pl.smooth_total);
                if (ssim) { if (grad) PML_LAUNCH((prep_kernel<true, true>), gm, dim3(32), 0, st, ip, sp, n_items);
                            else      PML_LAUNCH((prep_kernel<true, false>), gm, dim3(32), 0, st, ip, sp, n_items); }
                else      { if (grad) PML_LAUNCH((prep_kernel<false, true>), gm, dim3(32), 0, st, ip, sp, n_items);
                            else      PML_LAUNCH((prep_kernel<false, false>), gm, dim3(32), 0, st, ip, sp, n_items); }
            } else if (ssim) PML_LAUNCH(identity_sweep_kernel<true>, g, dim3(32), 0, st, ip);
            else             PML_LAUNCH(identity_sweep_kernel<false>, g, dim3(32), 0, st, ip);
        }
    } else if (pl.n_id > 0) {
        dim3 g((p->W + 255) / 256, p->H, p->B);
        const int avg = (p->flags & PML_FLAG_AVG_REPROJ) ? 1 : 0;
        if (ssim) PML_LAUNCH(identity_kernel<true>, g, dim3(256), 0, st, p->target, p->sources[0], p->sources[1],
                             p->sources[2], p->sources[3], identity, p->B, p->H, p->W, p->S, avg);
        else      PML_LAUNCH(identity_kernel<false>, g, dim3(256), 0, st, p->target, p->sources[0], p->sources[1],
                             p->sources[2], p->sources[3], identity, p->B, p->H, p->W, p->S, avg);
    }

    // 4. fused warp + loss (+ adjoint)
    PhotoParams pp;
    pp.B = p->B; pp.H = p->H; pp.W = p->W; pp.n_pass = p->n_pass; pp.flags = p->flags;
    pp.min_disp = (float)(1.0 / (double)p->max_depth);                                   // layers.py:21
    pp.disp_range = (float)(1.0 / (double)p->min_depth - 1.0 / (double)p->max_depth);     // layers.py:23
    pp.eps = p->eps; pp.seed = p->seed; pp.seed_dev = reinterpret_cast<const unsigned long long*>(p->seed_device);
    pp.target = p->target; pp.K = p->K; pp.invK = p->inv_K; pp.identity = identity;
    for (int f = 0; f < PML_MAX_SOURCES; ++f) { pp.src[f] = p->sources[f]; pp.T[f] = p->T[f]; }
    int low_cells = 0;
    for (int i = 0; i < p->n_pass; ++i) {
        const pml_pass& ps = p->pass[i];
        PassDev& d = pp.pass[i];
        d.disp = ps.disp; d.noise = ps.noise; d.argmin = ps.argmin; d.depth = ps.depth; d.warped = ps.warped;
        d.grad_disp = grad ? ps.grad_disp : nullptr;
        d.fw = ps.frame_weight; d.gfw = grad ? ps.grad_frame_weight : nullptr;
        d.hd = ps.hd; d.wd = ps.wd; d.k = p->H / ps.hd;
        d.rscale = (float)ps.hd / (float)p->H;
        d.low_cols = pl.TW / d.k + 3; d.low_rows = pl.TH / d.k + 3;
        if (d.k > 1 && d.low_cols * d.low_rows > low_cells) low_cells = d.low_cols * d.low_rows;
    }
    pp.TW = pl.TW; pp.TH = pl.TH; pp.n_strips = pl.n_strips; pp.n_chunks = pl.n_chunks;
    pp.n_items = pl.n_cta; pp.S = p->S;
    pp.mode = 0; pp.f_base = 0; pp.pair_n = p->S > 1 ? 2 : 1; pp.rp = nullptr; pp.presel = nullptr;
    pp.cta_per_pass = pl.cta_per_pass; pp.part = part; pp.part_stride = pl.part_stride;
    pp.inv_n = (float)(1.0 / ((double)p->B * p->H * p->W));
    pp.n_seg = sp.n_seg; pp.seg_size = sp.seg_size;
    if (sg)
        for (int j = 0; j < sg->n_seg; ++j) {
            pp.target_c.p[j] = sg->target[j]; pp.K_c.p[j] = sg->K[j]; pp.invK_c.p[j] = sg->inv_K[j];
            for (int f = 0; f < p->S; ++f) pp.src_c[f].p[j] = sg->sources[f][j];
        }
    if (p->prof_start) pml_event_record(p->prof_start, st);
    NvtxRange nvtx_sweep("pml: fused sweep");
    if (pl.two_sweeps) {
        // More than two source frames (or per-frame weights): (1) forward sweep per frame pair -> reprojection losses,
        // (2) selection over all candidates, (3) forward + adjoint sweep per pair with that selection.
        pp.rp = reinterpret_cast<float*>(base + pl.off_rp);
        for (int i = 0; i < p->n_pass; ++i)
            if (!pp.pass[i].argmin)
                pp.pass[i].argmin = reinterpret_cast<uint8_t*>(base + pl.off_argmin) + (size_t)i * p->B * p->H * p->W;
        // forward sweeps of all pairs but the last (reprojection losses -> rp); the last pair's sweep
        // selects over everything and carries its own adjoint; then the adjoint sweeps of the others
        const int last_fa = ((p->S - 1) / 2) * 2;
        pp.mode = 1; pp.TH = pl.TH_fwd; pp.n_chunks = pl.n_chunks_fwd;
        for (int fa = 0; fa < last_fa && rc == PML_OK; fa += 2) {
            pp.f_base = fa; pp.pair_n = 2;
            rc = ssim ? launch_sweep<false, true>(pp, st) : launch_sweep<false, false>(pp, st);
        }
        pp.TH = pl.TH; pp.n_chunks = pl.n_chunks;
        if (rc == PML_OK) {
            pp.mode = 3; pp.f_base = last_fa; pp.pair_n = (last_fa + 1 < p->S) ? 2 : 1;
            pp.presel = reinterpret_cast<uint8_t*>(base + pl.off_presel);
            if (sweep_common(pp)) {   // candidates ahead of the last pair -> one value + one index per pixel
                const dim3 g((p->H * p->W + 255) / 256, p->B);
                if (last_fa == 2)      PML_LAUNCH(select_prepass_kernel<2>, g, dim3(256), 0, st, pp);
                else if (last_fa == 4) PML_LAUNCH(select_prepass_kernel<4>, g, dim3(256), 0, st, pp);
                else                   PML_LAUNCH(select_prepass_kernel<6>, g, dim3(256), 0, st, pp);
            }
            if (grad) rc = ssim ? launch_sweep<true, true>(pp, st) : launch_sweep<true, false>(pp, st);
            else      rc = ssim ? launch_sweep<false, true>(pp, st) : launch_sweep<false, false>(pp, st);
        }
        pp.mode = 2;
        for (int fa = 0; grad && fa < last_fa && rc == PML_OK; fa += 2) {
            pp.f_base = fa; pp.pair_n = 2;
            rc = ssim ? launch_sweep<true, true>(pp, st) : launch_sweep<true, false>(pp, st);
        }
    } else if (pl.sweep) {
        if (grad) rc = ssim ? launch_sweep<true, true>(pp, st) : launch_sweep<true, false>(pp, st);
        else      rc = ssim ? launch_sweep<false, true>(pp, st) : launch_sweep<false, false>(pp, st);
    } else if (grad) rc = ssim ? dispatch_S<true, true>(p->S, pp, pl.n_cta, pl.NT, low_cells, st)
                               : dispatch_S<true, false>(p->S, pp, pl.n_cta, pl.NT, low_cells, st);
    else             rc = ssim ? dispatch_S<false, true>(p->S, pp, pl.n_cta, pl.NT, low_cells, st)
                               : dispatch_S<false, false>(p->S, pp, pl.n_cta, pl.NT, low_cells, st);
    if (rc != PML_OK) return rc;
    if (p->prof_stop) pml_event_record(p->prof_stop, st);

    // 5. fixed-order reduction of all partials
    FinalizeParams fq;
    fq.B = p->B; fq.S = p->S; fq.n_pass = p->n_pass; fq.cta_per_pass = pl.cta_per_pass;
    fq.cta_per_image = pl.n_chunks * pl.n_strips; fq.part_stride = pl.part_stride; fq.with_grad = grad ? 1 : 0;
    fq.inv_n = pp.inv_n;
    fq.n_seg = sp.n_seg; fq.seg_size = sp.seg_size;
    if (sg) for (int j = 0; j < sg->n_seg; ++j) fq.K_c.p[j] = sg->K[j];
    fq.part = part; fq.K = p->K; fq.smooth_part = spart; fq.disp_mean = mean;
    for (int i = 0; i < p->n_pass; ++i) {
        fq.smooth_blocks[i] = pl.smooth_blocks[i]; fq.smooth_off[i] = pl.smooth_off[i];
        fq.hd[i] = p->pass[i].hd; fq.wd[i] = p->pass[i].wd; fq.smooth_weight[i] = p->pass[i].smooth_weight;
    }
    fq.losses = p->losses; fq.loss_vector = p->loss_vector; fq.grad_T = p->grad_T; fq.grad_disp_const = p->grad_disp_const;
    fq.loss_total = p->loss_total; fq.total_div = p->loss_total_div > 0.f ? p->loss_total_div : (float)p->n_pass;
    fq.image_part = imagepart;
    PML_LAUNCH(finalize_image_kernel, dim3(p->B, p->n_pass), dim3(256), 0, st, fq);
    PML_LAUNCH(finalize_loss_kernel, dim3(1), dim3(32), 0, st, fq);
    return cudaGetLastError() == cudaSuccess ? PML_OK : PML_ERR_CUDA;
}

}  // namespace

extern "C" {

int pml_abi_version(void) { return PML_ABI_VERSION; }

const char* pml_strerror(int status) {
    switch (status) {
        case PML_OK: return "ok";
        case PML_ERR_INVALID: return "invalid argument (size, null or misaligned pointer)";
        case PML_ERR_UNSUPPORTED: return "unsupported configuration";
        case PML_ERR_WORKSPACE: return "workspace too small";
        case PML_ERR_CUDA: return "CUDA runtime error";
    }
    return "unknown status";
}

size_t pml_workspace_bytes(const pml_problem* p) {
    if (validate(p, false) != PML_OK) return 0;
    const size_t a = make_plan(p, true).total, b = make_plan(p, false).total;   // either call may follow
    return a > b ? a : b;
}

int pml_loss_forward(const pml_problem* p, void* ws, size_t ws_bytes, pml_stream_t stream) {
    return run_loss(p, ws, ws_bytes, reinterpret_cast<cudaStream_t>(stream), false);
}

int pml_loss_forward_backward(const pml_problem* p, void* ws, size_t ws_bytes, pml_stream_t stream) {
    return run_loss(p, ws, ws_bytes, reinterpret_cast<cudaStream_t>(stream), true);
}

int pml_scale_grads(int32_t n_pass, int32_t B, int32_t S, const int32_t* hd, const int32_t* wd,
                    float* const* grad_disp, const float* grad_disp_const, const float* grad_T,
                    const float* upstream, const float* upstream_total, float total_div,
                    float* grad_T_out, pml_stream_t stream) {
    if (n_pass < 1 || n_pass > PML_MAX_PASSES || B < 1 || S < 1 || S > PML_MAX_SOURCES) return PML_ERR_INVALID;
    if (!hd || !wd || !grad_disp || !grad_disp_const || !grad_T || (!upstream && !upstream_total) || !grad_T_out) return PML_ERR_INVALID;
    NvtxRange nvtx_call("pml_scale_grads");
    DeviceGuard guard(grad_T);
    pml::ScaleParams sp;
    sp.n_pass = n_pass; sp.B = B; sp.S = S;
    long long total = 0;
    for (int i = 0; i < n_pass; ++i) {
        if (!grad_disp[i]) return PML_ERR_INVALID;
        sp.g[i] = grad_disp[i]; sp.per_image[i] = hd[i] * wd[i];
        sp.off[i] = total;
        total += (long long)B * hd[i] * wd[i];
    }
    sp.off[n_pass] = total;
    sp.gconst = grad_disp_const; sp.gT = grad_T; sp.up = upstream; sp.up_total = upstream_total;
    sp.inv_div = 1.0f / (total_div > 0.f ? total_div : (float)n_pass); sp.gT_out = grad_T_out;
    int max_n = 1;
    for (int i = 0; i < n_pass; ++i) if (sp.per_image[i] > max_n) max_n = sp.per_image[i];
    int chunks = (max_n + pml::kScaleChunk - 1) / pml::kScaleChunk;
    const int pose_blocks = (S * B * 16 + 255) / 256;            // spread over the (x, y) extent of plane n_pass
    while (chunks * B < pose_blocks) ++chunks;
    if (B > 65535) return PML_ERR_UNSUPPORTED;
    PML_LAUNCH(pml::scale_grads_kernel, dim3(chunks, B, n_pass + 1), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), sp);
    return cudaGetLastError() == cudaSuccess ? PML_OK : PML_ERR_CUDA;
}

int pml_selection_masks(int32_t n_pass, int64_t n_pix, const uint8_t* const* argmin, int32_t n_id,
                        float* const* out, pml_stream_t stream) {
    if (n_pass < 1 || n_pass > PML_MAX_PASSES || n_pix < 1 || n_id < 1 || !argmin || !out) return PML_ERR_INVALID;
    pml::SelectionParams q;
    for (int i = 0; i < n_pass; ++i) {
        if (!argmin[i] || !out[i]) return PML_ERR_INVALID;
        q.argmin[i] = argmin[i]; q.out[i] = out[i];
    }
    q.n_pix = n_pix; q.n_id = n_id;
    DeviceGuard guard(out[0]);
    const long long per_block = 256LL * 16;
    PML_LAUNCH(pml::selection_mask_kernel, dim3((unsigned)((n_pix + per_block - 1) / per_block), n_pass), dim3(256), 0,
               reinterpret_cast<cudaStream_t>(stream), q);
    return cudaGetLastError() == cudaSuccess ? PML_OK : PML_ERR_CUDA;
}

}  // extern "C"

#include "pml_layers_api.inc"
