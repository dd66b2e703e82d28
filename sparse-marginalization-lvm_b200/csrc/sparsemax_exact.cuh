// Pieces of the batched sparsemax shared by the register kernels (sparsemax.cu) and the
// TMA-staged tile kernel (sparsemax_tile.cu): the column <-> (slot, lane) map, the sub-warp
// bitonic sort and the sort-based evaluation of the reference's (k, tau).
#pragma once

#include <math_constants.h>

#include "common.cuh"

namespace smlvm {

// Column held by slot e of lane l.  VEC: 128-bit chunks interleaved across the
// group (coalesced float4 accesses); scalar: plain striping.
template <int E, int G, bool VEC>
__device__ __forceinline__ int col_of(int e, int l)
{
    if (VEC) return 4 * ((e >> 2) * G + l) + (e & 3);
    return e * G + l;
}

// A row of K fp32 in the registers of a G-lane group (E values per lane, E*G >= K; columns >= K read `fill`).
template <int E, int G, bool VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ base, int K, int l, float fill,
                                         float (&v)[E])
{
    if (VEC) {
#pragma unroll
        for (int c = 0; c < E / 4; ++c) {
            int col = 4 * (c * G + l);
            float4 q = make_float4(fill, fill, fill, fill);
            if (col < K) q = ldg_stream4(base + col);
            v[4 * c + 0] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int col = e * G + l;
            v[e] = (col < K) ? ldg_stream(base + col) : fill;
        }
    }
}

template <int E, int G, bool VEC>
__device__ __forceinline__ void store_row(float* __restrict__ base, int K, int l,
                                          const float (&v)[E])
{
    if (VEC) {
#pragma unroll
        for (int c = 0; c < E / 4; ++c) {
            int col = 4 * (c * G + l);
            if (col < K)
                stg_stream4(base + col, make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int col = e * G + l;
            if (col < K) stg_stream(base + col, v[e]);
        }
    }
}

// Descending bitonic sort of the E*G values of a group; sorted position of
// slot e in lane l is l*E + e.
template <int E, int G>
__device__ __forceinline__ void group_sort_desc(float (&s)[E], int l)
{
    constexpr int N = E * G;
#pragma unroll
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= E) {
                const int lj = j / E;  // lane stride
                const bool lower = (l & lj) == 0;
                // every slot of a lane shares (pos & k) when k >= E... k > j >= E
                const bool desc = ((l * E) & k) == 0;
                const bool keep_max = (lower == desc);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    float o = __shfl_xor_sync(kFull, s[e], lj, G);
                    s[e] = keep_max ? fmaxf(s[e], o) : fminf(s[e], o);
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if ((e & j) == 0) {
                        const int f = e | j;
                        // direction of the block containing pos = l*E + e
                        const bool desc = (((l * E + e) & k) == 0);
                        float hi = fmaxf(s[e], s[f]), lo = fminf(s[e], s[f]);
                        s[e] = desc ? hi : lo;
                        s[f] = desc ? lo : hi;
                    }
                }
            }
        }
    }
}


// Full-warp groups reduce in one REDUX/CREDUX instruction (sm_100a has the fp32 max form);
// sub-warp groups fall back to shuffles (a partial-mask redux serialises per mask).
template <int G>
__device__ __forceinline__ float gmax_fast(float v)
{
    if (G == 32) {
        float m;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
        return m;
    }
    return group_max<G>(v);
}
template <int G>
__device__ __forceinline__ int gsum_fast(int v)
{
    if (G == 32) return __reduce_add_sync(kFull, v);
    return group_sum<G>(v);
}
template <int G>
__device__ __forceinline__ int gmin_fast(int v)
{
    if (G == 32) return __reduce_min_sync(kFull, v);
    return group_reduce<G>(v, [](int a, int b) { return min(a, b); });
}
constexpr int ilog2_c(int n) { return n <= 1 ? 0 : 1 + ilog2_c(n >> 1); }


// The reference's own definition of (k, tau), by sorting (entmax
// _sparsemax_threshold_and_support evaluated with torch CPU ops): v[] holds the shifted row
// X' = X - max(X) of a G-lane group (padding = -inf), Kr = number of real columns.
// Returns k = #{r : fl32(r * S_r) > C_r} (clamped to >= 1) and tau = C_k / k to every lane
// of the group.  Must be called by all 32 lanes of the warp.
template <int E, int G>
__device__ __forceinline__ void exact_threshold_sorted(const float (&v)[E], int l, int Kr, int& k_out,
                                                       float& tau_out)
{
    float s[E];
#pragma unroll
    for (int e = 0; e < E; ++e) s[e] = v[e];
    group_sort_desc<E, G>(s, l);

    // fp64 inclusive prefix sums in sorted order
    double cs[E];
    double run = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        run += (double)s[e];
        cs[e] = run;
    }
    double incl = run;  // lane totals -> exclusive offset of this lane
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        double t = __shfl_up_sync(kFull, incl, o, G);
        if (l >= o) incl += t;
    }
    // (not incl - run: a lane holding -inf padding would turn that into NaN)
    double excl = __shfl_up_sync(kFull, incl, 1, G);
    if (l == 0) excl = 0.0;

    float Cc[E];
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int pos = l * E + e;
        Cc[e] = __fsub_rn((float)(excl + cs[e]), 1.0f);
        const float lhs = __fmul_rn((float)(pos + 1), s[e]);
        cnt += (pos < Kr && lhs > Cc[e]) ? 1 : 0;
    }
    int ks = group_sum<G>(cnt);
    ks = max(ks, 1);

    // tau = C[k-1] / k : fetch slot (k-1) % E of lane (k-1) / E
    float ck = Cc[0];
#pragma unroll
    for (int e = 1; e < E; ++e) ck = (((ks - 1) % E) == e) ? Cc[e] : ck;
    ck = __shfl_sync(kFull, ck, (ks - 1) / E, G);
    k_out = ks;
    tau_out = __fdiv_rn(ck, (float)ks);
}

// d entropy / d p for entropy() of marg.py:6-28: -(log p~ + 1) where the clamp
// passes gradient (eps <= p <= 1-eps) and p > 0; 0 elsewhere.
__device__ __forceinline__ float dent_dp(float p)
{
    const float eps = 1.1920928955078125e-07f;
    return (p >= eps && p <= 1.f - eps) ? -(__logf(p) + 1.f) : 0.f;  // MUFU.LG2: |err| < 2e-6
}


}  // namespace smlvm
