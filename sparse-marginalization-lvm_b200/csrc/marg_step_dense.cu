// The categorical step behind the forward for rows with WIDE supports, in ONE launch that reads P and L once (sm_100a).
//
//   weighted loss        row_loss[r] = sum_c P[r,c] L[r,c],  total = scale * sum_r row_loss[r]     (marg.py:125-126)
//   its backward         dX = sparsemax Jacobian of (scale * L + dent[r] * d entropy / d P),  dL = scale * P
//   support compaction   (value, row, col) of P > 0 at offsets[r] + rank, column order          (structmarg.py:116-126)
//
// The slot-driven fused kernel (sparsemax_bwd_tile.cu) does this from 64 bytes of support slots per row when supports
// hold a few entries.  With flat score rows (the signal game at temperature 10, the sweep at score scale 0.1: supports
// of 20-60 entries) the slots do not apply and round 1 ran three dense-read kernels -- compaction (reads P), weighted
// loss (P, L), backward (P, L; writes dX, dL): 4.0 GB at 1M x 128, 0.92 ms.  Here G = K / 32 lanes own a row with 32
// values of P and of L per lane in registers; everything above comes out of that one read: 2.4 GB, HBM-bound.
//
// Compaction inside a lane group: a lane's non-zero mask gives 8 nibble counts (one per 128-bit chunk); a packed
// shuffle scan over the G lanes ranks the chunks of a row in column order (column = 4 (chunk * G + lane) + q).
#include <math_constants.h>

#include "common.cuh"
#include "reduce_ws.cuh"
#include "sparsemax_exact.cuh"

namespace smlvm {

constexpr int kDenseThreads = 256;

// GRAD = false: the compaction alone (smlvm_compact_fill without support slots on large batches).
template <int G, bool LISTS, bool GRAD>
__global__ void __launch_bounds__(kDenseThreads, 2)
marg_step_dense_kernel(const float* __restrict__ p, const int32_t* __restrict__ supp, const float* __restrict__ loss,
                       float scale, const float* __restrict__ dent, const int64_t* __restrict__ offsets,
                       float* __restrict__ vals, int32_t* __restrict__ row_idx, int32_t* __restrict__ col_idx,
                       int64_t capacity, float* __restrict__ row_loss, float* __restrict__ total, ReduceWs ws,
                       float* __restrict__ dx, float* __restrict__ dloss, int64_t rows, int K)
{
    constexpr int E = 32;
    constexpr int RPW = kWarp / G;
    __shared__ double red[32];
    extern __shared__ __align__(16) float stage[];   // LISTS: per warp 1024 values + 1024 (row, column) tags
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kDenseThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kDenseThreads / kWarp) + (threadIdx.x >> 5);
    double acc_total = 0.0;

    for (; w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int Kr = live ? K : 0;
        const int64_t off = (live ? row : 0) * (int64_t)K;
        float pv[E], gv[E];
        load_row<E, G, true>(p + off, Kr, l, 0.f, pv);
        int64_t obase = 0;
        if (LISTS && live) obase = __ldg(offsets + row);
        unsigned mask = 0u;
        if (GRAD) {
        load_row<E, G, true>(loss + off, Kr, l, 0.f, gv);
        const int kk = live ? __ldg(supp + row) : 1;
        const float de = (dent != nullptr && live) ? __ldg(dent + row) : 0.f;

        // ---- weighted loss of the row; dL = scale * P ----------------------------------------------
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int e = 0; e < E; e += 2) {
            s0 = fmaf(pv[e], gv[e], s0);
            s1 = fmaf(pv[e + 1], gv[e + 1], s1);
        }
        const float rl = group_sum<G>(s0 + s1);
        if (live && l == 0) {
            if (row_loss != nullptr) row_loss[row] = rl;
            acc_total += (double)rl;
        }
        if (dloss != nullptr) {
            float dl[E];
#pragma unroll
            for (int e = 0; e < E; ++e) dl[e] = scale * pv[e];
            if (live) store_row<E, G, true>(dloss + off, K, l, dl);
        }

        // ---- gradient w.r.t. the scores (the operation order of sparsemax_bwd_reg) -----------------
        float g0 = 0.f, g1 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            float ge = scale * gv[e];
            if (dent != nullptr) ge = fmaf(de, dent_dp(pv[e]), ge);
            const bool on = pv[e] != 0.f;
            ge = on ? ge : 0.f;
            gv[e] = ge;
            if (e & 1) g1 += ge; else g0 += ge;
            if (LISTS) mask |= (pv[e] > 0.f) ? (1u << e) : 0u;
        }
        const float vhat = __fdiv_rn(group_sum<G>(g0 + g1), (float)kk);
#pragma unroll
        for (int e = 0; e < E; ++e) gv[e] = (pv[e] != 0.f) ? __fsub_rn(gv[e], vhat) : 0.f;
        if (live) store_row<E, G, true>(dx + off, K, l, gv);
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) mask |= (pv[e] > 0.f) ? (1u << e) : 0u;
        }

        // ---- support compaction -----------------------------------------------------------------------
        // The entries of the warp's 32 / G consecutive rows are one contiguous run of the lists.  Each lane ranks
        // its entries inside the run and parks them in the warp's staging area; the run then leaves with coalesced
        // stores (a 4-byte store per entry straight to global memory costs a 32-byte sector operation each: the
        // first version of this kernel spent 0.7 of its 1.14 ms there at supports of 49).
        if (LISTS) {
            // nibble counts of this lane, four per word (chunk c of the lane = bits 4c..4c+3 of the mask)
            unsigned m = mask;
            m = (m & 0x55555555u) + ((m >> 1) & 0x55555555u);
            m = (m & 0x33333333u) + ((m >> 2) & 0x33333333u);   // popcount of every nibble, in place
            unsigned lo = (m & 0xFu) | ((m & 0xF0u) << 4) | ((m & 0xF00u) << 8) | ((m & 0xF000u) << 12);
            unsigned hi = ((m >> 16) & 0xFu) | (((m >> 16) & 0xF0u) << 4) | (((m >> 16) & 0xF00u) << 8) |
                          (((m >> 16) & 0xF000u) << 12);
            // inclusive scan over the lanes of the group, byte-wise (a chunk holds at most 4 G <= 64 entries)
            unsigned ilo = lo, ihi = hi;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                const unsigned tl = __shfl_up_sync(kFull, ilo, o, G), th = __shfl_up_sync(kFull, ihi, o, G);
                if (l >= o) { ilo += tl; ihi += th; }
            }
            const unsigned tlo = __shfl_sync(kFull, ilo, G - 1, G), thi = __shfl_sync(kFull, ihi, G - 1, G);
            const unsigned xlo = ilo - lo, xhi = ihi - hi;   // entries of this chunk in the lanes before this one
            const int64_t wbase = __shfl_sync(kFull, obase, 0);          // first entry of the warp's run
            int cbase = (int)(obase - wbase);                // entries of the run before this chunk of this row
            float* sv = stage + (threadIdx.x >> 5) * (2 * kWarp * E);
            unsigned* sc = reinterpret_cast<unsigned*>(sv) + kWarp * E;
            const unsigned tag = ((unsigned)g << 16) | (unsigned)(4 * l);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const unsigned tot = ((c < 4 ? tlo : thi) >> (8 * (c & 3))) & 0xFFu;
                const unsigned exc = ((c < 4 ? xlo : xhi) >> (8 * (c & 3))) & 0xFFu;
                int o = cbase + (int)exc;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (((mask >> (4 * c + q)) & 1u) && (unsigned)o < (unsigned)(kWarp * E)) {   // (guard: offsets that
                        sv[o] = pv[4 * c + q];                                                    //  are not the scan of nnz)
                        sc[o] = tag + (unsigned)(4 * c * G + q);   // (row of the warp) << 16 | column
                        ++o;
                    }
                }
                cbase += (int)tot;
            }
            // run length = end of the last live row of the warp (cbase now = end of this row inside the run)
            const int nlive = (int)min((int64_t)RPW, rows - w * RPW);
            const int T = min(__shfl_sync(kFull, cbase, (nlive - 1) * G), kWarp * E);
            __syncwarp();
            const int64_t row0 = w * RPW;
            for (int j = lane; j < T; j += 32) {
                const int64_t o = wbase + j;
                if (o < capacity) {
                    const unsigned cc = sc[j];
                    if (vals != nullptr) vals[o] = sv[j];
                    if (row_idx != nullptr) row_idx[o] = (int32_t)(row0 + (cc >> 16));
                    if (col_idx != nullptr) col_idx[o] = (int32_t)(cc & 0xFFFFu);
                }
            }
            __syncwarp();
        }
    }
    if (GRAD && total != nullptr) finish_totals(ws, acc_total, 0.0, scale, total, nullptr, red);
}

bool step_dense_supported(int cols) { return cols > 32 && cols <= 512 && cols % 4 == 0; }

template <int G>
static int launch_dense(const float* p, const int32_t* supp, const float* loss, float scale, const float* dent,
                        const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx, int64_t capacity,
                        float* row_loss, float* total, void* workspace, float* dx, float* dloss, int64_t rows, int K,
                        cudaStream_t st)
{
    const int rows_per_cta = (kDenseThreads / kWarp) * (kWarp / G);
    const int64_t need = (rows + rows_per_cta - 1) / rows_per_cta;
    static int per_sm = -1;
    if (per_sm < 0) { const char* ev = getenv("SMLVM_DENSE_CTAS_PER_SM"); per_sm = ev ? atoi(ev) : 13; if (per_sm < 1) per_sm = 13; }
    // 13 CTAs per SM in the grid (2 resident; grid-stride beyond): measured at 1M x 128 against one resident wave of 2 per
    // SM -- 0.4219 -> 0.3932 ms (support 8.5), 0.4506 -> 0.4219 (20), 0.5263 -> 0.4834 (49); 4 / 8 per SM 0.412 / 0.401, 16 the
    // same as 13, 24+ slower again.  The per-CTA totals still fit the reduction workspace (kMaxCtas).
    int64_t cap = (int64_t)device_sm_count() * per_sm;
    if (cap > kMaxCtas) cap = kMaxCtas;
    const int grid = (int)(need < cap ? need : cap);
    const size_t smem = (size_t)(kDenseThreads / kWarp) * 2 * kWarp * 32 * sizeof(float);   // 64 KB
    if (offsets != nullptr) {
        static bool configured_dev[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured_dev[dev & 63]) {
            cudaError_t ea = cudaFuncSetAttribute(marg_step_dense_kernel<G, true, true>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (ea == cudaSuccess)
                ea = cudaFuncSetAttribute(marg_step_dense_kernel<G, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem);
            if (ea != cudaSuccess) return (int)ea;
            configured_dev[dev & 63] = true;
        }
    }
    if (loss == nullptr)
        marg_step_dense_kernel<G, true, false><<<grid, kDenseThreads, smem, st>>>(p, supp, loss, scale, dent, offsets, vals,
                                                                               row_idx, col_idx, capacity, row_loss, total,
                                                                               ws_view(workspace), dx, dloss, rows, K);
    else if (offsets != nullptr)
        marg_step_dense_kernel<G, true, true><<<grid, kDenseThreads, smem, st>>>(p, supp, loss, scale, dent, offsets, vals, row_idx,
                                                                       col_idx, capacity, row_loss, total,
                                                                       ws_view(workspace), dx, dloss, rows, K);
    else
        marg_step_dense_kernel<G, false, true><<<grid, kDenseThreads, 0, st>>>(p, supp, loss, scale, dent, offsets, vals, row_idx,
                                                                        col_idx, capacity, row_loss, total,
                                                                        ws_view(workspace), dx, dloss, rows, K);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

// the compaction alone: (value, row, col) of p > 0 from a dense read of p, coalesced stores
int compact_fill_coop_launch(const float* p, const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx,
                             int64_t rows, int K, int64_t capacity, cudaStream_t st);

int step_dense_launch(const float* p, const int32_t* supp, const float* loss, float scale, const float* dent,
                      const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx, int64_t capacity,
                      float* row_loss, float* total, void* workspace, float* dx, float* dloss, int64_t rows, int K,
                      cudaStream_t st)
{
#define SMLVM_DENSE_ARGS p, supp, loss, scale, dent, offsets, vals, row_idx, col_idx, capacity, row_loss, total, workspace, dx, dloss, rows, K, st
    if (K <= 64) return launch_dense<2>(SMLVM_DENSE_ARGS);
    if (K <= 128) return launch_dense<4>(SMLVM_DENSE_ARGS);
    if (K <= 256) return launch_dense<8>(SMLVM_DENSE_ARGS);
    return launch_dense<16>(SMLVM_DENSE_ARGS);
#undef SMLVM_DENSE_ARGS
}

int compact_fill_coop_launch(const float* p, const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx,
                             int64_t rows, int K, int64_t capacity, cudaStream_t st)
{
    return step_dense_launch(p, nullptr, nullptr, 0.f, nullptr, offsets, vals, row_idx, col_idx, capacity, nullptr, nullptr,
                             nullptr, nullptr, nullptr, rows, K, st);
}

}  // namespace smlvm
