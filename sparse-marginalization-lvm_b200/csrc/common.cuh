// Shared device helpers for the sm_100a kernels of the sparse-marginalization path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "smlvm.h"

namespace smlvm {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNumSMsB200 = 148;

#define SMLVM_LAUNCH_CHECK()                                  \
    do {                                                      \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess) return static_cast<int>(e__); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// SM count of the current device (148 on B200); cached per process.
int device_sm_count();

// Streaming 128-bit accesses: inputs/outputs are touched once, keep them out of L1.
__device__ __forceinline__ float4 ldg_stream4(const float* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// A 4-byte gather (one loss entry per non-zero of P): ask L2 for the smallest fill it offers, keep the line out of L1.
__device__ __forceinline__ float ldg_gather(const float* p)
{
    float r;
#if defined(SMLVM_GATHER_PLAIN)
    r = __ldg(p);
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(r) : "l"(p));
#endif
    return r;
}
// The same from PINNED HOST memory (the loss matrix left in place, hotpath.HostPipelinedMargStep(zero_copy_losses=True)):
// every gather is a PCIe read, so no L2 prefetch hint -- the smallest request the memory system makes.
__device__ __forceinline__ float ldg_gather_host(const float* p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void stg_stream(float* p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// t / d for a grid-stride index: 32-bit when it fits (a 64-bit division by a run-time divisor is a ~70-instruction call)
__device__ __forceinline__ int64_t div_idx(int64_t t, int d)
{
    return ((unsigned long long)t >> 31) == 0 ? (int64_t)((unsigned)t / (unsigned)d) : t / d;
}

// Butterfly reductions over the G (power of two) lanes of a sub-warp group.
template <int G, typename T, typename Op>
__device__ __forceinline__ T group_reduce(T v, Op op)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(kFull, v, o, G));
    return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v)
{
    return group_reduce<G>(v, [](float a, float b) { return fmaxf(a, b); });
}
template <int G>
__device__ __forceinline__ float group_sum(float v)
{
    return group_reduce<G>(v, [](float a, float b) { return a + b; });
}
template <int G>
__device__ __forceinline__ int group_sum(int v)
{
    return group_reduce<G>(v, [](int a, int b) { return a + b; });
}
template <int G>
__device__ __forceinline__ double group_sum(double v)
{
    return group_reduce<G>(v, [](double a, double b) { return a + b; });
}

// Block-wide sum of doubles (blockDim.x multiple of 32, <= 1024). `smem` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem)
{
    v = group_sum<32>(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double r = (threadIdx.x < nw) ? smem[threadIdx.x] : 0.0;
    if (w == 0) r = group_sum<32>(r);
    if (threadIdx.x == 0) smem[0] = r;
    __syncthreads();
    r = smem[0];
    return r;
}

}  // namespace smlvm
