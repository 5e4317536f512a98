// TEST INFRASTRUCTURE ONLY -- never linked into libpml.so, never imported by the package.
//
// A tiny SPMD emulator that lets g++ compile the *same* kernel sources as nvcc
// (self-supervised-depth-estimation_b200/csrc/*.cuh) and run them on host threads: one OS thread
// per CUDA thread of a block, blocks executed one after another.  The build container has no
// GPU, so this is how kernel logic is debugged before GPU minutes are spent; the parity that
// counts is still measured on the B200 (tests marked `gpu`).  __syncthreads maps to a
// std::barrier, warp shuffles to a per-warp exchange buffer, atomics to std::atomic_ref.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define PML_HOST_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n)
#define __restrict__ __restrict

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };

namespace emu {
struct Block {
    unsigned nthreads;
    std::barrier<>* bar;
    std::vector<std::unique_ptr<std::barrier<>>>* warp_bars;
    std::vector<uint32_t>* warp_xchg;   // [nwarps][32]
    std::vector<unsigned char>* smem;
};
inline thread_local uint3_ t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local Block* t_block = nullptr;

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& body) {
    unsigned nt = block.x * block.y * block.z;
    unsigned nwarps = (nt + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nt);
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        for (unsigned w = 0; w < nwarps; ++w) {
            unsigned lanes = std::min(32u, nt - w * 32);
            wb.emplace_back(new std::barrier<>(lanes));
        }
        std::vector<uint32_t> xchg(nwarps * 32);
        std::vector<unsigned char> smem(smem_bytes + 64);
        Block blk{nt, &bar, &wb, &xchg, &smem};
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t) {
            th.emplace_back([&, t]() {
                t_threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                t_blockIdx = {bx, by, bz};
                t_blockDim = {block.x, block.y, block.z};
                t_gridDim = {grid.x, grid.y, grid.z};
                t_block = &blk;
                body();
            });
        }
        for (auto& x : th) x.join();
    }
}
inline unsigned linear_tid() {
    return t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

static inline void __syncthreads() { emu::t_block->bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    (*emu::t_block->warp_bars)[emu::linear_tid() / 32]->arrive_and_wait();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
static inline T emu_shfl(T v, unsigned src_lane) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    unsigned tid = emu::linear_tid(), w = tid / 32, lane = tid % 32;
    uint32_t* buf = emu::t_block->warp_xchg->data() + w * 32;
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
    type* name = reinterpret_cast<type*>((reinterpret_cast<uintptr_t>(emu::t_block->smem->data()) + 15) & ~uintptr_t(15))
