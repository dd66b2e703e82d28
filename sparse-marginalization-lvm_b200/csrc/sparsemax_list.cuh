// The reference's sparsemax threshold evaluated on a short per-thread list (shared by the tile forward kernel and the
// k-sparsemax epilogue of the top-k path).
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace smlvm {

// Descending bitonic network over a register array (compile-time indices, no divergence).
template <int N>
__device__ __forceinline__ void sort_desc_regs(float (&v)[N])
{
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const bool desc = (i & k) == 0;
                    const float hi = fmaxf(v[i], v[l]), lo = fminf(v[i], v[l]);
                    v[i] = desc ? hi : lo;
                    v[l] = desc ? lo : hi;
                }
            }
        }
    }
}

// (k, tau) of the reference from the candidate list of one row (thread-private column `lv` of
// the [slot][lane] shared-memory list, n <= N entries): the candidates are loaded into
// registers, sorted descending by a bitonic network (compile-time indices, no divergence) and
// the reference's own support test is evaluated position by position -- fp64 cumulative sum in
// sorted order, C_r = fl32(fl32(cumsum_r) - 1), support where fl32(r * S_r) > C_r, k = number
// of such r, tau = C_k / k.  Padding is -inf: it sorts last and fails the test.
template <int N>
__device__ __forceinline__ void list_threshold(const float* lv, int n, int& k_out, float& tau_out)
{
    float v[N];
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = (j < n) ? lv[j * 32] : -CUDART_INF_F;
    sort_desc_regs<N>(v);
    double cs = 0.0;
    float c_last = -1.f;
    int kk = 0, last_true = -1;
#pragma unroll
    for (int r = 0; r < N; ++r) {
        cs += (double)v[r];
        const float C = __fsub_rn((float)cs, 1.0f);
        if (__fmul_rn((float)(r + 1), v[r]) > C) {
            ++kk;
            c_last = C;
            last_true = r;
        }
    }
    if (__any_sync(kFull, last_true + 1 != kk && kk > 0)) {
        // trues not contiguous (rounding): the reference takes C at index k, not at the last true
        if (last_true + 1 != kk && kk > 0) {
            cs = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r)
                if (r < kk) cs += (double)v[r];
            c_last = __fsub_rn((float)cs, 1.0f);
        }
    }
    kk = max(kk, 1);
    k_out = kk;
    tau_out = __fdiv_rn(c_last, (float)kk);
}

}  // namespace smlvm
