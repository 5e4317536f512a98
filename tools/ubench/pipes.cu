// Micro-benchmarks of the sm_100a issue/LSU rates the fused sweep is budgeted against
// (tools only; not part of libpml.so).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, const float* in, int n) {
    __shared__ float sm[4096];
    int tid = threadIdx.x;
    float a0 = in[tid], a1 = in[tid + 1], a2 = in[tid + 2], a3 = in[tid + 3];
    float a4 = in[tid + 4], a5 = in[tid + 5], a6 = in[tid + 6], a7 = in[tid + 7];
    float b = in[tid + 8], c = in[tid + 9];
    for (int i = tid; i < 4096; i += blockDim.x) sm[i] = in[i % n];
    __syncthreads();
    if (MODE == 0) {          // scalar FFMA, 8 independent chains
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    } else if (MODE == 1) {   // packed FFMA2, 8 independent chains of pairs (16 floats)
        float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
        float2 p4 = make_float2(a1, a0), p5 = make_float2(a3, a2), p6 = make_float2(a5, a4), p7 = make_float2(a7, a6);
        float2 bb = make_float2(b, c), cc = make_float2(c, b);
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            p0 = __ffma2_rn(p0, bb, cc); p1 = __ffma2_rn(p1, bb, cc); p2 = __ffma2_rn(p2, bb, cc); p3 = __ffma2_rn(p3, bb, cc);
            p4 = __ffma2_rn(p4, bb, cc); p5 = __ffma2_rn(p5, bb, cc); p6 = __ffma2_rn(p6, bb, cc); p7 = __ffma2_rn(p7, bb, cc);
        }
        a0 = p0.x + p0.y; a1 = p1.x + p1.y; a2 = p2.x + p2.y; a3 = p3.x + p3.y;
        a4 = p4.x + p4.y; a5 = p5.x + p5.y; a6 = p6.x + p6.y; a7 = p7.x + p7.y;
    } else if (MODE == 2) {   // SHFL
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            a0 = __shfl_up_sync(0xffffffffu, a0, 1); a1 = __shfl_down_sync(0xffffffffu, a1, 1);
            a2 = __shfl_up_sync(0xffffffffu, a2, 1); a3 = __shfl_down_sync(0xffffffffu, a3, 1);
            a4 = __shfl_up_sync(0xffffffffu, a4, 1); a5 = __shfl_down_sync(0xffffffffu, a5, 1);
            a6 = __shfl_up_sync(0xffffffffu, a6, 1); a7 = __shfl_down_sync(0xffffffffu, a7, 1);
        }
    } else if (MODE == 3) {   // LDS.32 conflict-free, address chain-free
        int o = tid;
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            a0 += sm[(o) & 4095]; a1 += sm[(o + 256) & 4095]; a2 += sm[(o + 512) & 4095]; a3 += sm[(o + 768) & 4095];
            a4 += sm[(o + 1024) & 4095]; a5 += sm[(o + 1280) & 4095]; a6 += sm[(o + 1536) & 4095]; a7 += sm[(o + 1792) & 4095];
            o += 32;
        }
    } else if (MODE == 4) {   // LDS.64
        const float2* s2 = reinterpret_cast<const float2*>(sm);
        int o = tid;
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            float2 v0 = s2[(o) & 2047], v1 = s2[(o + 256) & 2047], v2 = s2[(o + 512) & 2047], v3 = s2[(o + 768) & 2047];
            a0 += v0.x; a1 += v0.y; a2 += v1.x; a3 += v1.y; a4 += v2.x; a5 += v2.y; a6 += v3.x; a7 += v3.y;
            o += 32;
        }
    } else if (MODE == 5) {   // mixed: 1 FFMA2 + 1 IADD-ish (alu) interleaved: dual pipe issue check
        float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
        float2 bb = make_float2(b, c), cc = make_float2(c, b);
        int i0 = tid, i1 = tid * 3, i2 = tid * 5, i3 = tid * 7;
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            p0 = __ffma2_rn(p0, bb, cc); i0 = (i0 ^ i) + i1;
            p1 = __ffma2_rn(p1, bb, cc); i1 = (i1 ^ i) + i2;
            p2 = __ffma2_rn(p2, bb, cc); i2 = (i2 ^ i) + i3;
            p3 = __ffma2_rn(p3, bb, cc); i3 = (i3 ^ i) + i0;
        }
        a0 = p0.x + p0.y + __int_as_float(i0 & 0xff); a1 = p1.x + p1.y + __int_as_float(i1 & 0xff);
        a2 = p2.x + p2.y + __int_as_float(i2 & 0xff); a3 = p3.x + p3.y + __int_as_float(i3 & 0xff);
    } else if (MODE == 6) {   // LDG.32 L1-resident gather (stride-1 across lanes, 8 KB window)
        int o = tid;
#pragma unroll 4
        for (int i = 0; i < ITERS; ++i) {
            a0 += __ldg(in + ((o) & 2047)); a1 += __ldg(in + ((o + 256) & 2047)); a2 += __ldg(in + ((o + 512) & 2047));
            a3 += __ldg(in + ((o + 768) & 2047)); a4 += __ldg(in + ((o + 1024) & 2047)); a5 += __ldg(in + ((o + 1280) & 2047));
            a6 += __ldg(in + ((o + 1536) & 2047)); a7 += __ldg(in + ((o + 1792) & 2047));
            o += 33;
        }
    }
    out[blockIdx.x * blockDim.x + tid] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int MODE>
void run(const char* name, int per_iter, float* out, const float* in, int n) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = 148 * 8;
    kern<MODE><<<grid, 256>>>(out, in, n);
    cudaEventRecord(e0);
    kern<MODE><<<grid, 256>>>(out, in, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = (double)grid * 8 /*warps*/ * ITERS * per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %.3f warp-inst/clk/SM  (%.3f per SMSP) [clock attr %d kHz]\n", name, ms,
           winst / cyc / 148.0, winst / cyc / 592.0, clk);
}

int main() {
    float *in, *out;
    int n = 1 << 16;
    cudaMalloc(&in, n * 4 + 64); cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMemset(in, 0, n * 4 + 64);
    run<0>("FFMA x8", 8, out, in, n);
    run<1>("FFMA2 x8", 8, out, in, n);
    run<2>("SHFL x8", 8, out, in, n);
    run<3>("LDS.32 x8 (+8 FADD)", 8, out, in, n);
    run<4>("LDS.64 x4 (+8 FADD)", 4, out, in, n);
    run<5>("FFMA2 + 2 int x4", 4, out, in, n);
    run<6>("LDG.32 L1-hit x8 (+8 FADD)", 8, out, in, n);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
