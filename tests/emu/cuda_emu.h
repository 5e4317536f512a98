// TEST INFRASTRUCTURE ONLY -- never linked into libpml.so, never imported by the package.
//
// A tiny SPMD emulator that lets g++ compile the *same* kernel sources as nvcc
// (self-supervised-depth-estimation_b200/csrc/*.cuh) and run them on the host: one user-level fiber
// per CUDA thread of a block, blocks spread over a pool of OS threads.  The build container has no
// GPU, so this is how kernel logic is debugged before GPU minutes are spent; the parity that
// counts is still measured on the B200 (tests marked `gpu`).  __syncthreads / __syncwarp are fiber
// yields until the group has arrived, warp shuffles a per-warp exchange buffer, atomics std::atomic_ref.
#pragma once
#include <atomic>
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <algorithm>
#include <vector>

#define PML_HOST_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static thread_local   /* one block at a time per OS thread */
#define __align__(n)
#define __restrict__ __restrict

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emu {
// Execution model: every CUDA thread of a block is a user-level fiber (ucontext) of ONE OS thread,
// scheduled round-robin; a barrier / shuffle is a fiber yield (~100 ns) instead of an OS-level
// rendezvous.  Blocks are independent and run in parallel on a small pool of OS threads.
struct Group { unsigned size = 0, count = 0, gen = 0; };
struct Lane {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    uint3_ tid;
    bool done = false;
};
struct Block {
    unsigned nthreads = 0;
    uint3_ bid, bdim, gdim;
    std::vector<Lane> lanes;
    Group all;
    std::vector<Group> warps;
    std::vector<uint32_t> warp_xchg;   // [nwarps][32]
    std::vector<unsigned char> smem;
    ucontext_t sched;
    unsigned cur = 0;
    const std::function<void()>* body = nullptr;
};
inline thread_local Block* t_block = nullptr;
inline thread_local uint3_ t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;

inline void yield_to_scheduler() {
    Block* b = t_block;
    swapcontext(&b->lanes[b->cur].ctx, &b->sched);
}
inline void group_wait(Group& g) {
    const unsigned my = g.gen;
    if (++g.count == g.size) { g.count = 0; ++g.gen; return; }
    while (g.gen == my) yield_to_scheduler();
}
inline void lane_entry() {
    Block* b = t_block;
    (*b->body)();
    b->lanes[b->cur].done = true;
    swapcontext(&b->lanes[b->cur].ctx, &b->sched);
}
inline void run_block(Block& blk) {
    t_block = &blk;
    t_blockIdx = blk.bid; t_blockDim = blk.bdim; t_gridDim = blk.gdim;
    unsigned live = blk.nthreads;
    for (unsigned t = 0; t < blk.nthreads; ++t) {
        Lane& l = blk.lanes[t];
        l.done = false;
        getcontext(&l.ctx);
        l.ctx.uc_stack.ss_sp = l.stack.data();
        l.ctx.uc_stack.ss_size = l.stack.size();
        l.ctx.uc_link = &blk.sched;
        makecontext(&l.ctx, (void (*)())lane_entry, 0);
    }
    while (live) {
        for (unsigned t = 0; t < blk.nthreads; ++t) {
            Lane& l = blk.lanes[t];
            if (l.done) continue;
            blk.cur = t;
            t_threadIdx = l.tid;
            swapcontext(&blk.sched, &l.ctx);
            if (l.done) --live;
        }
    }
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body_in) {
    const std::function<void()> body = body_in;
    const unsigned nt = block.x * block.y * block.z;
    const unsigned nwarps = (nt + 31) / 32;
    const unsigned long long nblocks = (unsigned long long)grid.x * grid.y * grid.z;
    unsigned nworkers = std::thread::hardware_concurrency();
    if (nworkers == 0) nworkers = 4;
    if (nworkers > 16) nworkers = 16;
    if (nworkers > nblocks) nworkers = (unsigned)nblocks;
    std::atomic<unsigned long long> next{0};
    auto worker = [&]() {
        Block blk;
        blk.nthreads = nt;
        blk.bdim = {block.x, block.y, block.z};
        blk.gdim = {grid.x, grid.y, grid.z};
        blk.lanes.resize(nt);
        for (unsigned t = 0; t < nt; ++t) {
            blk.lanes[t].stack.resize(256 * 1024);
            blk.lanes[t].tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        }
        blk.warps.resize(nwarps);
        blk.warp_xchg.resize(nwarps * 32);
        blk.smem.resize(smem_bytes + 64);
        blk.body = &body;
        for (;;) {
            const unsigned long long i = next.fetch_add(1);
            if (i >= nblocks) break;
            blk.bid = {(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((unsigned long long)grid.x * grid.y))};
            blk.all = Group{nt, 0, 0};
            for (unsigned w = 0; w < nwarps; ++w) blk.warps[w] = Group{std::min(32u, nt - w * 32), 0, 0};
            // dynamic shared memory starts out as whatever the previous CTA on that SM left behind: poison it with
            // the NaN bit pattern so that a kernel relying on its initial contents fails on the CPU too
            std::fill(blk.smem.begin(), blk.smem.end(), (unsigned char)0xFF);
            run_block(blk);
        }
    };
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < nworkers; ++w) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}
inline unsigned linear_tid() {
    return t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

static inline void __syncthreads() { emu::group_wait(emu::t_block->all); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::group_wait(emu::t_block->warps[emu::linear_tid() / 32]); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T emu_shfl(T v, unsigned src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    unsigned tid = emu::linear_tid(), w = tid / 32, lane = tid % 32;
    uint32_t* buf = emu::t_block->warp_xchg.data() + w * 32;
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    buf[lane] = bits;
    __syncwarp();
    uint32_t r = buf[src_lane % 32];
    __syncwarp();
    T out;
    std::memcpy(&out, &r, 4);
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, (emu::linear_tid() % 32) ^ (unsigned)m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    unsigned lane = emu::linear_tid() % 32;
    T r = emu_shfl(v, lane + d);
    return (lane + d < 32) ? r : v;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    unsigned lane = emu::linear_tid() % 32;
    T r = emu_shfl(v, lane >= (unsigned)d ? lane - d : lane);
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int l) { return emu_shfl(v, (unsigned)l); }
static inline int __any_sync(unsigned, int pred) {
    int r = pred ? 1 : 0;
    for (int m = 16; m > 0; m >>= 1) r |= __shfl_xor_sync(0xffffffffu, r, m);
    return r;
}

static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
static inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v); }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline void sincospif(float a, float* s, float* c) { *s = std::sin(3.14159265358979323846 * a); *c = std::cos(3.14159265358979323846 * a); }
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
static inline float __int_as_float(int a) { float f; std::memcpy(&f, &a, 4); return f; }
static inline int __float_as_int(float a) { int f; std::memcpy(&f, &a, 4); return f; }
static inline unsigned __float_as_uint(float a) { unsigned f; std::memcpy(&f, &a, 4); return f; }
static inline float __uint_as_float(unsigned a) { float f; std::memcpy(&f, &a, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __float2int_rd(float a) { return (int)std::floor(a); }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
// glibc's <math.h> already provides fmaxf/fminf/floorf/fabsf/sqrtf/expf/logf/cosf/sinf/fmaf and
// *declares* (but does not export) __expf/__logf/__cosf/__sinf, so map the CUDA fast-math
// intrinsics onto the accurate functions by macro.
#define __expf(x) expf(x)
#define __logf(x) logf(x)
#define __cosf(x) cosf(x)
#define __sinf(x) sinf(x)
static inline void emu_sincosf(float a, float* s, float* c) { *s = std::sin(a); *c = std::cos(a); }
#define __sincosf(a, s, c) emu_sincosf(a, s, c)
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }

// host-side runtime shims used by the C-ABI file
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

#define PML_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define PML_DYN_SMEM(type, name) \
    type* name = reinterpret_cast<type*>((reinterpret_cast<uintptr_t>(emu::t_block->smem.data()) + 15) & ~uintptr_t(15))
