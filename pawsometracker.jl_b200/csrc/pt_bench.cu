// pt_bench.cu — measurement helpers exported for bench.py: the FP32 FMA peak
// of the device (the denominator of the FP32-pipe roofline; SURVEY §8d asks
// for a measured value instead of the nominal 148×128×2×clock) and an L2 flush.
#include "../../include/pawsome.h"
#include <cuda_runtime.h>
#include <cstdio>
#include "pt_kernels.cuh"

namespace {

// 16 independent accumulators per thread, dependent chains of FFMAs with
// register operands only: the plain scalar FFMA issue-rate ceiling.
__global__ void __launch_bounds__(256) fma_peak_scalar(float *out, int iters, float a, float b)
{
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same work issued as packed fma.rn.f32x2 (Blackwell FFMA2): two FMAs per
// issued instruction.
__global__ void __launch_bounds__(256) fma_peak_packed(float *out, int iters, float a, float b)
{
    unsigned long long acc[8];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x = (float)(threadIdx.x + i);
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(x), "f"(x + 0.5f));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Issue-port probe: NF packed FFMA2 per NA independent integer ops (LOP3/IADD chains on other registers).
// If FFMA2 held the dispatch port for both of its pipe cycles, adding ALU ops would lengthen the loop;
// if they issue in its shadow, the time stays at the FFMA2-only value.
template <int NA>
__global__ void __launch_bounds__(256) ffma2_alu_mix(float *out, int iters, float a, float b)
{
    unsigned long long acc[8];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
    unsigned int z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x = (float)(threadIdx.x + i);
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(x), "f"(x + 0.5f));
        z[i] = threadIdx.x * 2654435761u + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
                if (i < NA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(z[(i + 1) & 7]), "r"((unsigned)u));
            }
        }
    }
    float s = 0.f;
    unsigned int zz = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi; zz ^= z[i];
    }
    if (s == 123.456f || zz == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)zz;
}

__global__ void flush_kernel(float4 *p, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

} // namespace

extern "C" {

// Measures FP32 FMA throughput (TFLOP/s, 2 flops per FMA). packed=0: scalar
// FFMA; packed=1: fma.rn.f32x2.  Best of `reps` timed launches.
PT_API int pt_measure_fp32_peak(int device, int packed, int reps, double *tflops)
{
    if (!tflops) return PT_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PT_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PT_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float *out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * blocks * threads) != cudaSuccess) return PT_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(e0);
        if (packed) fma_peak_packed<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        else fma_peak_scalar<<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return PT_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 16.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (r >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return PT_OK;
}

// Issue-port probe (see ffma2_alu_mix): returns ms per launch for na ∈ {0, 4, 8} ALU ops per 8 FFMA2.
PT_API int pt_probe_ffma2_issue(int device, int na, double *ms_out)
{
    if (!ms_out) return PT_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PT_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PT_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    float *out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * blocks * threads) != cudaSuccess) return PT_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        if (na == 0) ffma2_alu_mix<0><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        else if (na == 4) ffma2_alu_mix<4><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        else ffma2_alu_mix<8><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return PT_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 1 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *ms_out = best;
    return PT_OK;
}

// Profiling aid: phase timestamps of dog_window45_argmax go to dev_buf ([n][T][6] int64), NULL = off.
PT_API int pt_debug_window45_timing(void *dev_buf)
{
    pt::window45_set_debug((long long *)dev_buf);
    return PT_OK;
}

// Overwrites `bytes` of scratch on `stream` so nothing useful stays in L2.
PT_API int pt_flush_l2(void *scratch, size_t bytes, void *stream)
{
    if (!scratch || bytes < 16) return PT_ERR_ARG;
    flush_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>((float4 *)scratch, bytes / 16);
    return cudaGetLastError() == cudaSuccess ? PT_OK : PT_ERR_CUDA;
}

} // extern "C"
