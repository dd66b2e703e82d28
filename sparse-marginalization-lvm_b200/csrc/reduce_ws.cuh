// Deterministic fp64 totals across CTAs (shared by the weighted-loss kernels and the fused loss + backward kernel).
#pragma once

#include "common.cuh"

namespace smlvm {

constexpr int kMaxCtas = 2048;

// workspace layout: [0] uint32 arrival counter (self-resetting), [64..] double partials[3][kMaxCtas]
struct ReduceWs {
    unsigned int* counter;
    double* part;  // [3 * kMaxCtas]
};
__host__ __device__ inline ReduceWs ws_view(void* ws)
{
    ReduceWs r;
    r.counter = reinterpret_cast<unsigned int*>(ws);
    r.part = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + 64);
    return r;
}

// Every CTA deposits up to 3 partial sums; the last CTA to arrive adds them up in CTA order.
__device__ __forceinline__ void finish_totals(ReduceWs ws, double a, double b, float scale,
                                              float* out_a, float* out_b, double* smem)
{
    a = block_sum(a, smem);
    b = block_sum(b, smem);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        ws.part[blockIdx.x] = a;
        ws.part[kMaxCtas + blockIdx.x] = b;
        __threadfence();
        unsigned int t = atomicAdd(ws.counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double sa = 0.0, sb = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
            sa += __ldcg(ws.part + i);
            sb += __ldcg(ws.part + kMaxCtas + i);
        }
        sa = block_sum(sa, smem);
        sb = block_sum(sb, smem);
        if (threadIdx.x == 0) {
            if (out_a != nullptr) *out_a = (float)((double)scale * sa);
            if (out_b != nullptr) *out_b = (float)((double)scale * sb);
            *ws.counter = 0u;  // leave the workspace ready for the next call
        }
    }
}

}  // namespace smlvm
