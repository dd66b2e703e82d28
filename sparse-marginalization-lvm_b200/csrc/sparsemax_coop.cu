// Batched sparsemax forward for rows with WIDE supports (flat score rows) and for K = 256 / 512, sm_100a.
//
// Same semantics as sparsemax.cu / sparsemax_tile.cu (entmax.sparsemax evaluated with torch CPU ops,
// lvmhelpers/marg.py:51; the signal game's K = 256 scores at temperature 10, experiments/signal-game/archs.py:42).
//
// Why a third forward kernel.  The tile kernel (thread per row, candidates sorted in registers) does work
// proportional to the support: 50 warp instructions per row at a support of 3.5, 162 at 20 and > 180 once the
// candidates no longer fit its 32-entry list (ncu r02: profiles/r02_ncu_summary.md), at 10 resident warps per SM --
// latency-bound long before HBM.  The register kernel (warp per row) pays a warp-wide reduction for every step of
// the threshold search: 427 / 805 instructions per row at K = 128 / 256.  Here
//
//   * G = K / 32 lanes own a row, 32 values per lane in registers (a warp handles 32 / G rows per step, loaded as
//     128-bit accesses that cover G * 16 contiguous bytes per row and instruction); no shared memory, ~24 resident
//     warps per SM;
//   * the threshold search is Michelot's fixed point in fp32 from a lower bound given by the 4G class maxima of the
//     row (a per-lane maximum per float4 component costs nothing extra in the max pass): 2-3 passes of 3
//     instructions per element, reductions over G lanes only;
//   * the reference's tau is then evaluated ONCE on the proposed support (fp64 sum, fl32(fl32(sum) - 1) / k) and the
//     output pass checks, with the rounding margins of DESIGN.md 4.1, that the reference's sort-based support test
//     yields exactly this support: no entry within [tau - margin_out, tau + margin_in].  A row that fails (an entry
//     within ~1e-6 of the threshold, non-finite scores) is re-done by the whole warp with the reference's own
//     definition (bitonic sort + fp64 cumulative sums, sparsemax_exact.cuh), straight from global memory.
//
// The cost no longer depends on the support size: ~140 instructions per row at K = 128 whatever the score scale.
#include <math_constants.h>

#include <stdlib.h>
#include "common.cuh"
#include "sparsemax_exact.cuh"

namespace smlvm {

constexpr int kCoopThreads = 256;
constexpr int kCoopE = 32;   // values per lane

// Predicated accumulations, one SASS instruction per line (the compiler's select + add forms cost 5 instead of 3
// instructions per element and pass: ncu r02, 2050 -> 1250 warp instructions per 8 rows at K = 128).
__device__ __forceinline__ void acc_if_gt(float v, float t, float& s, int& c)
{
    asm("{\n .reg .pred q;\n setp.gt.f32 q, %2, %3;\n @q add.f32 %0, %0, %2;\n @q add.s32 %1, %1, 1;\n}"
        : "+f"(s), "+r"(c) : "f"(v), "f"(t));
}
__device__ __forceinline__ void acc64_if_gt(float v, float t, double& s)
{
    asm("{\n .reg .pred q;\n .reg .f64 d;\n setp.gt.f32 q, %1, %2;\n cvt.f64.f32 d, %1;\n @q add.f64 %0, %0, d;\n}"
        : "+d"(s) : "f"(v), "f"(t));
}
// One entry of the output pass: p = max(x' - tau, 0); entries at or below thr_in feed the running maximum `below`
// (NaN-propagating: a NaN score fails the check); entries above it add p log2 p to h.
template <bool ENT>
__device__ __forceinline__ float emit_entry(float xe, float neg_tau, float thr_in, float one_m_eps, float& below, float& h)
{
    float pe;
    if (ENT) {
        asm("{\n .reg .pred hi;\n .reg .f32 ps, lg;\n setp.gt.f32 hi, %3, %5;\n @!hi max.NaN.f32 %1, %1, %3;\n"
            " add.f32 %0, %3, %4;\n max.f32 %0, %0, 0f00000000;\n min.f32 ps, %0, %6;\n lg2.approx.ftz.f32 lg, ps;\n"
            " @hi fma.rn.f32 %2, ps, lg, %2;\n}"
            : "=f"(pe), "+f"(below), "+f"(h) : "f"(xe), "f"(neg_tau), "f"(thr_in), "f"(one_m_eps));
    } else {
        asm("{\n .reg .pred hi;\n setp.gt.f32 hi, %2, %4;\n @!hi max.NaN.f32 %1, %1, %2;\n"
            " add.f32 %0, %2, %3;\n max.f32 %0, %0, 0f00000000;\n}"
            : "=f"(pe), "+f"(below) : "f"(xe), "f"(neg_tau), "f"(thr_in));
    }
    return pe;
}
__device__ __forceinline__ void mark_if_zero(float xe, unsigned bit, unsigned& eq)
{
    asm("{\n .reg .pred q;\n setp.eq.f32 q, %1, 0f00000000;\n @q or.b32 %0, %0, %2;\n}" : "+r"(eq) : "f"(xe), "r"(bit));
}

// One row, whole warp, the reference's definition by sorting; reads X and writes P in global memory.
// E2 = columns per lane (32 * E2 >= K).
template <int E2>
__device__ __noinline__ void coop_row_exact(const float* __restrict__ xr, float* __restrict__ pr, int K, int lane,
                                            bool want_ent, int& k_out, int& nz_out, float& h_out, int& first_out)
{
    float v[E2];
#pragma unroll
    for (int e = 0; e < E2; ++e) {
        const int col = e * 32 + lane;
        v[e] = (col < K) ? __ldg(xr + col) : -CUDART_INF_F;
    }
    float m = v[0];
#pragma unroll
    for (int e = 1; e < E2; ++e) m = fmaxf(m, v[e]);
    m = gmax_fast<32>(m);
    int first = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < E2; ++e) {
        const int col = e * 32 + lane;
        if (col < K && v[e] == m) first = min(first, col);
        v[e] = __fsub_rn(v[e], m);
    }
    first_out = gmin_fast<32>(first);
    int k;
    float tau;
    exact_threshold_sorted<E2, 32>(v, lane, K, k, tau);
    const float eps = 1.1920928955078125e-07f;
    int nz = 0;
    float h = 0.f;
#pragma unroll
    for (int e = 0; e < E2; ++e) {
        const int col = e * 32 + lane;
        const float pe = fmaxf(__fsub_rn(v[e], tau), 0.f);
        if (col < K) pr[col] = pe;
        if (pe > 0.f) {
            ++nz;
            if (want_ent) {
                const float ps = fminf(fmaxf(pe, eps), 1.f - eps);
                h = fmaf(ps, __logf(ps), h);
            }
        }
    }
    k_out = k;
    nz_out = gsum_fast<32>(nz);
    h_out = group_sum<32>(h);
}

template <int G, bool VEC, bool AMAX, bool ENT>
__global__ void __launch_bounds__(kCoopThreads, 3)
sparsemax_fwd_coop(const float* __restrict__ x, float* __restrict__ p, int32_t* __restrict__ supp,
                   int32_t* __restrict__ nnz, float* __restrict__ ent, int64_t* __restrict__ amax, int64_t rows, int K)
{
    constexpr int E = kCoopE;
    constexpr int RPW = kWarp / G;
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kCoopThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kCoopThreads / kWarp) + (threadIdx.x >> 5);
    const float kf_margin = (3.06f * (float)K + 1.f) * 1.2e-7f;

    for (; w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int Kr = live ? K : 0;   // dead groups see an empty row
        const float* xr = x + (live ? row : 0) * (int64_t)K;

        float v[E];
        load_row<E, G, VEC>(xr, Kr, l, -CUDART_INF_F, v);

        // ---- row max; per-lane class maxima (one per float4 component) for the starting bound -----
        float cm[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int e = 0; e < E; ++e) cm[e & 3] = fmaxf(cm[e & 3], v[e]);
        float m = group_max<G>(fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
        if (!live) m = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = __fsub_rn(v[e], m);   // X' (padding stays -inf)

        // ---- lower bound on tau: Michelot on the 4G class maxima (tau of a subset never exceeds tau of the row,
        //      and every iterate from below stays below it) ------------------------------------------
        float t = -1.0f;
        {
            float c4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) c4[j] = __fsub_rn(cm[j], m);
#pragma unroll
            for (int it = 0; it < 3; ++it) {
                float s = 0.f;
                int c = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool in = c4[j] > t;
                    s += in ? c4[j] : 0.f;
                    c += in ? 1 : 0;
                }
                s = group_sum<G>(s);
                c = group_sum<G>(c);
                t = fmaxf(t, __fdividef(s - 1.f, (float)max(c, 1)));
            }
        }

        // ---- Michelot's fixed point over the row (fp32; the check below does not trust it) ----------
        int C = 0, cprev = -1;
        bool conv = false;
        for (int iter = 0; iter < 64; ++iter) {
            float s0 = 0.f, s1 = 0.f;
            int c0 = 0, c1 = 0;
#pragma unroll
            for (int e = 0; e < E; e += 2) {
                acc_if_gt(v[e], t, s0, c0);
                acc_if_gt(v[e + 1], t, s1, c1);
            }
            const float S = group_sum<G>(s0 + s1);
            C = group_sum<G>(c0 + c1);
            if (__all_sync(kFull, C == cprev)) { conv = true; break; }   // supports stopped shrinking in every row of the warp
            cprev = C;
            t = fmaxf(t, __fdividef(S - 1.f, (float)max(C, 1)));
        }

        // ---- the reference's tau on the proposed support ---------------------------------------------
        int k = max(C, 1);
        float tau;
        {
            double d0 = 0.0, d1 = 0.0;
#pragma unroll
            for (int e = 0; e < E; e += 2) {
                acc64_if_gt(v[e], t, d0);
                acc64_if_gt(v[e + 1], t, d1);
            }
            const double sd = group_sum<G>(d0 + d1);
            tau = __fdiv_rn(__fsub_rn((float)sd, 1.0f), (float)k);
        }
        // margins: inside needs x' - tau > 1e-6, outside (tau - x') k > 2^-24 (3.06 K + 1)
        const float thr_in = tau + 1e-6f;
        const float thr_out = tau - __fdividef(kf_margin, (float)k) - 3e-7f;

        // ---- P, entropy, argmax and the check in one pass ----------------------------------------------
        // The check: every entry is inside (x' > thr_in) or outside (x' < thr_out), i.e. the largest entry that is
        // not inside lies below thr_out; with thr_out <= t <= thr_in the inside entries are exactly the proposed
        // support.  NaN poisons the running maximum and sends the row to the sort.
        float below0 = -CUDART_INF_F, below1 = -CUDART_INF_F;
        float h0 = 0.f, h1 = 0.f;
        unsigned eq = 0u;
        const float one_m_eps = 1.f - 1.1920928955078125e-07f;
        const float neg_tau = -tau;
#pragma unroll
        for (int e = 0; e < E; e += 2) {
            if (AMAX) {
                mark_if_zero(v[e], 1u << e, eq);
                mark_if_zero(v[e + 1], 2u << e, eq);
            }
            v[e] = emit_entry<ENT>(v[e], neg_tau, thr_in, one_m_eps, below0, h0);
            v[e + 1] = emit_entry<ENT>(v[e + 1], neg_tau, thr_in, one_m_eps, below1, h1);
        }
        float below;
        asm("max.NaN.f32 %0, %1, %2;" : "=f"(below) : "f"(below0), "f"(below1));
        const int nbad = (conv && t >= thr_out && t <= thr_in && below < thr_out) ? 0 : 1;
        const unsigned gm = (G == 32) ? kFull : (((1u << G) - 1u) << (g * G));
        const unsigned bal = __ballot_sync(kFull, nbad != 0);   // (hoisted: every lane votes)
        const bool redo = live && ((bal & gm) != 0u);
        if (live && !redo) store_row<E, G, VEC>(p + row * (int64_t)K, K, l, v);
        float h = group_sum<G>(h0 + h1) * 0.69314718055994531f;
        int first = 0x7fffffff;
        if (AMAX) {
            if (eq != 0u) first = col_of<E, G, VEC>(__ffs(eq) - 1, l);
            first = group_reduce<G>(first, [](int a, int b) { return min(a, b); });
        }
        int nz = k;

        // ---- rows that failed the check: the whole warp, one row at a time ---------------------------
        unsigned todo = __ballot_sync(kFull, redo && l == 0);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r2 = w * RPW + src / G;
            int k2, nz2, f2;
            float h2;
            coop_row_exact<G>(x + r2 * (int64_t)K, p + r2 * (int64_t)K, K, lane, ENT, k2, nz2, h2, f2);
            if (lane / G == src / G) { k = k2; nz = nz2; h = h2; first = f2; }
        }

        if (live && l == 0) {
            if (supp != nullptr) supp[row] = k;
            if (nnz != nullptr) nnz[row] = nz;
            if (ENT && ent != nullptr) ent[row] = -h;
            if (AMAX && amax != nullptr) amax[row] = first;
        }
    }
}

template <int G>
static int launch_coop(const float* x, float* p, int32_t* supp, int32_t* nnz, float* ent, int64_t* amax, int64_t rows,
                       int K, bool vec, cudaStream_t st)
{
    const int rows_per_cta = (kCoopThreads / kWarp) * (kWarp / G);
    const int64_t need = (rows + rows_per_cta - 1) / rows_per_cta;
    static int per_sm = -1;
    if (per_sm < 0) { const char* ev = getenv("SMLVM_COOP_CTAS_PER_SM"); per_sm = ev ? atoi(ev) : 8; if (per_sm < 1) per_sm = 8; }
    // 8 CTAs per SM in the grid (3-4 resident, grid-stride beyond): measured at 1M x 128 against exactly one resident wave
    // of 3 per SM -- 0.2252 -> 0.2140 ms (support 8.5), 0.2353 -> 0.2265 (20), 0.2518 -> 0.2417 (49); 12 and 16 the same,
    // 24+ slower again.  The hardware scheduler evens out what a fixed share per CTA leaves as a tail.
    const int64_t cap = (int64_t)device_sm_count() * per_sm;
    const int grid = (int)(need < cap ? need : cap);
#define SMLVM_COOP_LAUNCH(V, A, H) \
    sparsemax_fwd_coop<G, V, A, H><<<grid, kCoopThreads, 0, st>>>(x, p, supp, nnz, ent, amax, rows, K)
    // (unaligned rows, K % 4 != 0: scalar accesses, always with argmax and entropy)
    if (!vec) SMLVM_COOP_LAUNCH(false, true, true);
    else if (amax != nullptr && ent != nullptr) SMLVM_COOP_LAUNCH(true, true, true);
    else if (amax != nullptr) SMLVM_COOP_LAUNCH(true, true, false);
    else if (ent != nullptr) SMLVM_COOP_LAUNCH(true, false, true);
    else SMLVM_COOP_LAUNCH(true, false, false);
#undef SMLVM_COOP_LAUNCH
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

bool sparsemax_coop_supported(int cols) { return cols > 32 && cols <= 512; }

int sparsemax_fwd_coop_launch(const float* x, float* p, int32_t* supp, int32_t* nnz, float* ent, int64_t* amax,
                              int64_t rows, int cols, bool vec, cudaStream_t st)
{
    if (cols <= 64) return launch_coop<2>(x, p, supp, nnz, ent, amax, rows, cols, vec, st);
    if (cols <= 128) return launch_coop<4>(x, p, supp, nnz, ent, amax, rows, cols, vec, st);
    if (cols <= 256) return launch_coop<8>(x, p, supp, nnz, ent, amax, rows, cols, vec, st);
    if (cols <= 512) return launch_coop<16>(x, p, supp, nnz, ent, amax, rows, cols, vec, st);
    return SMLVM_ERR_UNSUPPORTED;
}

}  // namespace smlvm
