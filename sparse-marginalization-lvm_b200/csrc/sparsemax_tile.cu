// Batched sparsemax forward, TMA-staged tile kernel for sm_100a (K in {32,64,128}).
//
// Same semantics as sparsemax.cu (entmax.sparsemax evaluated with torch CPU ops,
// lvmhelpers/marg.py:51, lvmhelpers/structmarg.py:48) but organised around the observation
// that the register kernel is ISSUE-bound (ncu r01: 427 warp instructions per row, 79 % issue
// utilisation, 36 % of the HBM roofline): with one warp per row every element of the row pays
// for every step of the threshold search.  Here
//
//   * a warp owns a tile of 32 consecutive rows (32*K*4 contiguous bytes), fetched by ONE
//     `cp.async.bulk` (TMA) into shared memory and written back by one bulk store; each warp
//     runs its own 2-stage mbarrier ring, so there is no CTA-wide synchronisation;
//   * a THREAD owns a row.  It reads the row from shared memory in 128-bit chunks, rotated
//     by the lane id so that the 32 rows (stride K words) never collide on a bank;
//   * pass A finds the row max; pass B builds a bitmask of the candidates X' > thr (X' = X -
//     max, thr = -1 first: nothing at or below max-1 can be in the support) and their sum,
//     from which Michelot's bound tau >= (sum - 1) / count tightens thr until at most 32
//     candidates are left (one pass is enough for peaked rows);
//   * the candidates -- a superset of the support with a rounding margin -- are insertion-
//     sorted and the REFERENCE'S OWN definition (fp64 cumulative sums in sorted order,
//     fl32(r*S_r) > fl32(fl32(cumsum_r) - 1), tau = C_k / k) is evaluated on them: the sorted
//     candidates are exactly the head of the reference's full sort, and every position behind
//     them fails the support test by more than the rounding error, so (k, tau)
//     and hence P are those of the reference;
//   * rows whose support does not fit 32 candidates (near-uniform rows) are handed to the
//     whole warp, which evaluates them with the bitonic-sort path of sparsemax_exact.cuh.
//
// Work per row drops from ~430 to ~50 warp instructions; HBM traffic stays at the algorithmic
// 8K bytes per row.  Measured against the register kernel on 2^27 elements (tools/fwd_sweep.py,
// scores N(0, s^2)): 2.4x faster at K = 128, s = 1 (support 3.5), 2.1x at s = 0.3 (8.5), 1.4x at
// s = 0.1 (20), 0.9x at s = 0.03 (49 of 128 in the support).  K = 256 is left to the register
// kernel: with 32 KB tiles only 3 warp pipelines fit an SM and flat rows run 1.5-2.4x slower.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"
#include "sparsemax_exact.cuh"
#include "sparsemax_list.cuh"
#include "tile_ptx.cuh"

namespace smlvm {

constexpr int kCap = 32;        // candidate-list entries per row
constexpr int kSupportSlots = SMLVM_SUPPORT_SLOTS;
constexpr int kMaxPasses = 3;   // bound-tightening passes before a row is treated as wide

// per-warp shared memory of the tile kernel
template <int K, int S>
struct TileSmem {
    static constexpr int kTileFloats = kTileRows * K;
    static constexpr size_t kWarpBytes = (size_t)S * kTileFloats * 4 + kCap * 32 * 4 + kCap * 32 * 1;
    static constexpr size_t total(int W) { return W * kWarpBytes + (size_t)W * S * 8; }
};

// Cooperative (whole-warp) evaluation of one row held in shared memory: the sort-based
// reference definition.  Overwrites the row with P; returns the per-row scalars to every lane.
template <int K>
__device__ __forceinline__ void warp_row_exact(float* srow, int lane, int& k_out, int& nz_out, float& h_out,
                                               int& first_out)
{
    constexpr int E = K / 32;
    constexpr bool VEC = (E % 4 == 0);
    float v[E];
    if (VEC) {
#pragma unroll
        for (int c = 0; c < E / 4; ++c) {
            const float4 q = *reinterpret_cast<const float4*>(srow + 4 * (c * 32 + lane));
            v[4 * c + 0] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = srow[e * 32 + lane];
    }
    float m = v[0];
#pragma unroll
    for (int e = 1; e < E; ++e) m = fmaxf(m, v[e]);
    m = gmax_fast<32>(m);
    int first = 0x7fffffff;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if (v[e] == m) first = min(first, col_of<E, 32, VEC>(e, lane));
        v[e] = __fsub_rn(v[e], m);
    }
    first_out = gmin_fast<32>(first);
    int k;
    float tau;
    exact_threshold_sorted<E, 32>(v, lane, K, k, tau);
    const float eps = 1.1920928955078125e-07f;
    int nz = 0;
    float h = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const float pe = fmaxf(__fsub_rn(v[e], tau), 0.f);
        v[e] = pe;
        if (pe > 0.f) {
            ++nz;
            const float ps = fminf(fmaxf(pe, eps), 1.f - eps);
            h = fmaf(ps, __logf(ps), h);
        }
    }
    if (VEC) {
#pragma unroll
        for (int c = 0; c < E / 4; ++c)
            *reinterpret_cast<float4*>(srow + 4 * (c * 32 + lane)) =
                make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) srow[e * 32 + lane] = v[e];
    }
    k_out = k;
    nz_out = gsum_fast<32>(nz);
    h_out = group_sum<32>(h);
}


template <int K, int S, int W>
__global__ void __launch_bounds__(W * 32, 1)
sparsemax_fwd_tile(const float* __restrict__ x, float* __restrict__ p, int32_t* __restrict__ supp,
                   int32_t* __restrict__ nnz, float* __restrict__ ent, int64_t* __restrict__ amax,
                   float* __restrict__ slots, int64_t rows)
{
    static_assert(S == 1 || S == 2, "1- or 2-stage ring");
    constexpr int NC = K / 4;   // 128-bit chunks per row
    constexpr int MW = K / 32;  // candidate-mask words per row
    constexpr int TILE_F = kTileRows * K;
    // superset margin: candidates are X' > bound - kMargin; covers the fp32 error of the
    // Michelot bound (<= 2K ulp) and the rounding slack of the reference's support test behind
    // the candidates (<= 6K ulp), DESIGN.md §4.1
    constexpr float kMargin = 32.f * K * 5.9604645e-08f;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tiles = reinterpret_cast<float*>(smem_raw) + (size_t)warp * S * TILE_F;
    float* lval = reinterpret_cast<float*>(smem_raw) + (size_t)W * S * TILE_F + warp * (kCap * 32) + lane;
    uint8_t* lrho = reinterpret_cast<uint8_t*>(reinterpret_cast<float*>(smem_raw) + (size_t)W * S * TILE_F +
                                               W * (kCap * 32)) + warp * (kCap * 32) + lane;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)W * TileSmem<K, S>::kWarpBytes) + warp * S;

    if (threadIdx.x == 0) {
        for (int i = 0; i < W * S; ++i)
            mbar_init(smem_u32(reinterpret_cast<uint64_t*>(smem_raw + (size_t)W * TileSmem<K, S>::kWarpBytes) + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t ntiles = (rows + kTileRows - 1) / kTileRows;
    const int64_t stride = (int64_t)gridDim.x * W;
    int64_t t = (int64_t)warp * gridDim.x + blockIdx.x;

    auto issue_load = [&](int stage, int64_t tile) {
        const int64_t r0 = tile * kTileRows;
        const uint32_t bytes = (uint32_t)min((int64_t)kTileRows, rows - r0) * K * 4u;
        const uint32_t bar = smem_u32(bars + stage);
        mbar_expect_tx(bar, bytes);
        bulk_load(smem_u32(tiles + stage * TILE_F), x + r0 * K, bytes, bar);
    };

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (t + s * stride < ntiles) issue_load(s, t + s * stride);
    }

    uint32_t parity = 0;  // bit s: phase to wait for on stage s
    int stage = 0;
    for (int64_t it = 0; t < ntiles; t += stride, ++it) {
        mbar_wait(smem_u32(bars + stage), (parity >> stage) & 1u);
        parity ^= 1u << stage;

        const int64_t row0 = t * kTileRows;
        const int trows = (int)min((int64_t)kTileRows, rows - row0);
        const bool live = lane < trows;
        float* tile = tiles + stage * TILE_F;
        // Row `lane` of the tile, read in 128-bit chunks.  Lane l visits chunk i ^ (l mod NC) at
        // step i, so the 32 rows (stride K words) never collide on a bank; the row base is
        // aligned to the row size, hence chunk address = swz ^ (i << 4): one LOP3 per access.
        const uint32_t rowa = smem_u32(tile + (live ? lane : 0) * K);
        const uint32_t lsw = (uint32_t)lane & (uint32_t)((NC - 1) & 31);
        const uint32_t swz = rowa ^ (lsw << 4);

        // ---- pass A: class maxima -> row max and a first lower bound on tau -------------------
        // For ANY subset T of the row, tau >= (sum_T X' - 1) / |T|  (Michelot).  The maxima of
        // NCLS disjoint classes (float4 component x chunk index mod NCLS/4) are NCLS distinct
        // entries; the best prefix of their sorted order gives the starting bound (>= -1, the
        // bound of {max}).  16 classes keep rows with supports of 20-30 entries (K = 256 at the
        // signal-game score scale) under the 32-candidate list without a tightening pass.
        constexpr int NCLS = (NC >= 16) ? 16 : 8;
        float cm[NCLS];
#pragma unroll
        for (int j = 0; j < NCLS; ++j) cm[j] = -CUDART_INF_F;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const float4 q = lds128(swz ^ (i << 4));
            const int o = (i & (NCLS / 4 - 1)) * 4;
            cm[o + 0] = fmaxf(cm[o + 0], q.x);
            cm[o + 1] = fmaxf(cm[o + 1], q.y);
            cm[o + 2] = fmaxf(cm[o + 2], q.z);
            cm[o + 3] = fmaxf(cm[o + 3], q.w);
        }
        sort_desc_regs<NCLS>(cm);
        const float m = cm[0];
        float thr = -1.0f;
        {
            float run = 0.f;
#pragma unroll
            for (int j = 2; j <= NCLS; ++j) {
                run += __fsub_rn(cm[j - 1], m);
                thr = fmaxf(thr, (run - 1.f) * (1.f / (float)j));
            }
            thr -= kMargin;
        }

        // refill the other stage: its tile was stored one iteration ago, the store has long
        // finished reading it (S == 2); the new tile has the rest of this iteration to land
        if (S == 2 && it > 0 && lane == 0) {
            bulk_wait_read0();
            if (t + stride < ntiles) issue_load(stage ^ 1, t + stride);
        }

        // ---- pass B: candidate mask {X > max + thr}, 2 instructions per element ---------------
        // (compared unshifted against the rounded-DOWN sum, a superset of {fl(X - max) > thr};
        //  the extraction below applies the exact shifted test)
        uint32_t mask[MW];
        {
            const float thr_abs = __fadd_rd(m, thr);
#pragma unroll
            for (int w = 0; w < MW; ++w) mask[w] = 0u;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const float4 q = lds128(swz ^ (i << 4));
                const float xq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (xq[e] > thr_abs) mask[(4 * i + e) >> 5] |= 1u << ((4 * i + e) & 31);
            }
        }
        int cnt = 0;
#pragma unroll
        for (int w = 0; w < MW; ++w) cnt += __popc(mask[w]);
        bool need = cnt > kCap, slow = false;
        // wide rows only: tighten the bound with Michelot passes over the row (mask + sum)
        if (__any_sync(kFull, need)) {
            for (int pass = 0; pass < kMaxPasses; ++pass) {
                if (need) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < MW; ++w) mask[w] = 0u;
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const float4 q = lds128(swz ^ (i << 4));
                        const float xv[4] = {__fsub_rn(q.x, m), __fsub_rn(q.y, m), __fsub_rn(q.z, m),
                                             __fsub_rn(q.w, m)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            if (xv[e] > thr) {
                                mask[(4 * i + e) >> 5] |= 1u << ((4 * i + e) & 31);
                                s += xv[e];
                            }
                        }
                    }
                    cnt = 0;
#pragma unroll
                    for (int w = 0; w < MW; ++w) cnt += __popc(mask[w]);
                    if (cnt <= kCap) {
                        need = false;
                    } else {
                        const float tn = __fdividef(s - 1.f, (float)cnt) - kMargin;
                        if (tn > thr) thr = tn;
                        else { slow = true; need = false; }  // > kCap entries at the bound: wide support
                    }
                }
                if (!__any_sync(kFull, need)) break;
            }
            if (need) slow = true;
        }

        // ---- candidates -> shared-memory list (value, swizzled index), warp-uniform loops -----
        // (a per-thread `while (mask)` would serialise the 32 rows of the warp: ncu r01 showed
        //  3.4 active threads per instruction with nested data-dependent loops)
        int n = 0;
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            uint32_t mw = slow ? 0u : mask[w];
            const int steps = __reduce_max_sync(kFull, __popc(mw));
            for (int jj = 0; jj < steps; ++jj) {
                if (mw) {
                    const int rho = w * 32 + __ffs(mw) - 1;
                    mw &= mw - 1;
                    const float xv = __fsub_rn(lds32(rowa + (((uint32_t)rho ^ (lsw << 2)) << 2)), m);
                    if (xv > thr) {
                        lval[n * 32] = xv;
                        lrho[n * 32] = (uint8_t)rho;
                        ++n;
                    }
                }
            }
        }
        const int nmax = __reduce_max_sync(kFull, n);

        int k = 1, nz = 0, first = 0x7fffffff;
        float h = 0.f;
        float tau = 0.f;
        if (nmax <= 8) list_threshold<8>(lval, n, k, tau);
        else if (nmax <= 16) list_threshold<16>(lval, n, k, tau);
        else list_threshold<32>(lval, n, k, tau);

        // ---- zero-fill the row, then P on the support + entropy + argmax -----------------------
        if (live && !slow) {
#pragma unroll
            for (int i = 0; i < NC; ++i) sts128(swz ^ (i << 4), make_float4(0.f, 0.f, 0.f, 0.f));
        }
        {
            const float eps = 1.1920928955078125e-07f;
            for (int j = 0; j < nmax; ++j) {
                if (j < n) {
                    const float sv = lval[j * 32];
                    const float pe = fmaxf(__fsub_rn(sv, tau), 0.f);
                    if (pe > 0.f) {
                        const uint8_t rho = lrho[j * 32];
                        const uint32_t col = (uint32_t)rho ^ (lsw << 2);
                        if (live) sts32(rowa + (col << 2), pe);
                        lval[nz * 32] = pe;   // compact the support to the front of the list (nz <= j):
                        lrho[nz * 32] = rho;  // read back by the slot emission below
                        ++nz;
                        const float ps = fminf(fmaxf(pe, eps), 1.f - eps);
                        h = fmaf(ps, __logf(ps), h);
                        if (sv == 0.f) first = min(first, (int)col);
                    }
                }
            }
        }

        // ---- wide-support rows (more than kCap entries at the bound), still one thread per row ----
        // Michelot's iteration to a fixed point over the whole row (no list), the reference's tau
        // from an fp64 sum over the support, then a check WITH ROUNDING MARGINS that the reference's
        // sort-based support test yields exactly this support (same margins as the register kernel,
        // sparsemax.cu): inside needs x' - tau > 1e-6, outside (tau - x') k > 2^-24 (3.06 K + 1).
        // A row that fails the check (an entry within ~1e-6 of the threshold) goes to the whole-warp
        // sort below.  ~45 instructions per element instead of a 430-instruction cooperative sort.
        bool coop = false;
        if (__any_sync(kFull, slow)) {
            if (slow) {
                // whole-row passes: 8 chunks unrolled per step (rolled outer loop keeps the code and
                // register footprint of this rarely-taken path small), four independent accumulator
                // chains per pass (one per float4 component) for instruction-level parallelism
                float tq = thr;
                int c = 0, c_prev = -1;
                bool converged = false;
                for (int iter = 0; iter < 64; ++iter) {
                    float s4[4] = {0.f, 0.f, 0.f, 0.f};
                    int c4[4] = {0, 0, 0, 0};
#pragma unroll 1
                    for (int i0 = 0; i0 < NC; i0 += 8) {
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) {
                            const float4 q = lds128(swz ^ ((i0 + ii) << 4));
                            const float xv[4] = {__fsub_rn(q.x, m), __fsub_rn(q.y, m), __fsub_rn(q.z, m),
                                                 __fsub_rn(q.w, m)};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (xv[e] > tq) { s4[e] += xv[e]; ++c4[e]; }
                        }
                    }
                    c = (c4[0] + c4[1]) + (c4[2] + c4[3]);
                    if (c == c_prev) { converged = true; break; }
                    c_prev = c;
                    tq = fmaxf(tq, __fdividef((s4[0] + s4[1]) + (s4[2] + s4[3]) - 1.f, (float)max(c, 1)));
                }
                const int kk = max(c, 1);
                double d4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
                for (int i0 = 0; i0 < NC; i0 += 8) {
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        const float4 q = lds128(swz ^ ((i0 + ii) << 4));
                        const float xv[4] = {__fsub_rn(q.x, m), __fsub_rn(q.y, m), __fsub_rn(q.z, m),
                                             __fsub_rn(q.w, m)};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (xv[e] > tq) d4[e] += (double)xv[e];
                    }
                }
                const double sd = (d4[0] + d4[1]) + (d4[2] + d4[3]);
                const float kf = (float)kk;
                tau = __fdiv_rn(__fsub_rn((float)sd, 1.0f), kf);
                const float thr_in = tau + 1e-6f;
                const float thr_out = tau - __fdividef((3.06f * (float)K + 1.f) * 1.2e-7f, kf) - 3e-7f;
                bool b4[4] = {!converged, false, false, false};
#pragma unroll 1
                for (int i0 = 0; i0 < NC; i0 += 8) {
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii) {
                        const float4 q = lds128(swz ^ ((i0 + ii) << 4));
                        const float xv[4] = {__fsub_rn(q.x, m), __fsub_rn(q.y, m), __fsub_rn(q.z, m),
                                             __fsub_rn(q.w, m)};
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            b4[e] |= (xv[e] > tq) ? !(xv[e] > thr_in) : !(xv[e] < thr_out);
                    }
                }
                if (b4[0] | b4[1] | b4[2] | b4[3]) {
                    coop = true;
                } else if (live) {
                    const float eps = 1.1920928955078125e-07f;
                    k = kk;
                    float h4[4] = {0.f, 0.f, 0.f, 0.f};
                    int n4[4] = {0, 0, 0, 0};
#pragma unroll 1
                    for (int i0 = 0; i0 < NC; i0 += 8) {
#pragma unroll
                        for (int ii = 0; ii < 8; ++ii) {
                            const int i = i0 + ii;
                            const float4 q = lds128(swz ^ (i << 4));
                            float pe[4] = {__fsub_rn(q.x, m), __fsub_rn(q.y, m), __fsub_rn(q.z, m),
                                           __fsub_rn(q.w, m)};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (pe[e] == 0.f) first = min(first, (int)((uint32_t)(4 * i + e) ^ (lsw << 2)));
                                pe[e] = fmaxf(__fsub_rn(pe[e], tau), 0.f);
                                if (pe[e] > 0.f) {
                                    ++n4[e];
                                    const float ps = fminf(fmaxf(pe[e], eps), 1.f - eps);
                                    h4[e] = fmaf(ps, __logf(ps), h4[e]);
                                }
                            }
                            sts128(swz ^ (i << 4), make_float4(pe[0], pe[1], pe[2], pe[3]));
                        }
                    }
                    nz = (n4[0] + n4[1]) + (n4[2] + n4[3]);
                    h = (h4[0] + h4[1]) + (h4[2] + h4[3]);
                }
            }
        }

        // ---- rows with an entry at the threshold: the whole warp, one row at a time (exact sort) ----
        unsigned sm = __ballot_sync(kFull, coop && live);
        while (sm) {
            const int r = __ffs(sm) - 1;
            sm &= sm - 1;
            int k2, nz2, f2;
            float h2;
            warp_row_exact<K>(tile + r * K, lane, k2, nz2, h2, f2);
            if (lane == r) { k = k2; nz = nz2; h = h2; first = f2; }
        }

        // ---- per-row support slots for the compaction (saves it the re-read of P) ----------------
        // up to 8 (value, column) pairs per row, 64 B; a row that does not fit (or did not come
        // through the candidate list) is marked "scan P" and the fill falls back for that row
        if (slots != nullptr && live) {
            float sv[kSupportSlots];
            int sc[kSupportSlots];
            const bool fits = !slow && nz <= kSupportSlots;
#pragma unroll
            for (int j = 0; j < kSupportSlots; ++j) {
                const bool on = fits && j < nz;
                sv[j] = on ? lval[j * 32] : 0.f;     // the list holds P on its first nz entries (sorted by value)
                sc[j] = on ? (int)((uint32_t)lrho[j * 32] ^ (lsw << 2)) : -1;
            }
            if (!fits) sc[0] = -2;
            float4* dst = reinterpret_cast<float4*>(slots + (row0 + lane) * (2 * kSupportSlots));
#pragma unroll
            for (int j = 0; j < kSupportSlots; j += 2)
                dst[j / 2] = make_float4(sv[j], __int_as_float(sc[j]), sv[j + 1], __int_as_float(sc[j + 1]));
        }

        if (live) {
            const int64_t row = row0 + lane;
            if (supp != nullptr) supp[row] = k;
            if (nnz != nullptr) nnz[row] = nz;
            if (ent != nullptr) ent[row] = -h;
            if (amax != nullptr) amax[row] = first;
        }

        // ---- tile of P back to HBM -----------------------------------------------------------
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store(p + row0 * K, smem_u32(tile), (uint32_t)trows * K * 4u);
            bulk_commit();
            if (S == 1) {
                bulk_wait_read0();
                if (t + stride < ntiles) issue_load(0, t + stride);
            }
        }
        __syncwarp();
        if (S == 2) stage ^= 1;
    }
    if (lane == 0) bulk_wait_read0();
}

template <int K, int S, int W>
static int launch_tile(const float* x, float* p, int32_t* supp, int32_t* nnz, float* ent, int64_t* amax,
                       float* slots, int64_t rows, cudaStream_t st)
{
    auto kern = sparsemax_fwd_tile<K, S, W>;
    const size_t smem = TileSmem<K, S>::total(W);
    static bool configured_dev[64] = {false};
    static int ctas_per_sm = 1;
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, W * 32, smem) != cudaSuccess ||
            ctas_per_sm < 1)
            ctas_per_sm = 1;
        configured = true;
    }
    const int64_t ntiles = (rows + kTileRows - 1) / kTileRows;
    const int64_t need = (ntiles + W - 1) / W;
    static int gmul = -1;
    if (gmul < 0) { const char* ev = getenv("SMLVM_TILE_GRID"); gmul = ev ? atoi(ev) : 1; if (gmul < 1) gmul = 1; }
    const int64_t cap = (int64_t)device_sm_count() * ctas_per_sm * gmul;
    const int grid = (int)(need < cap ? need : cap);
    kern<<<grid, W * 32, smem, st>>>(x, p, supp, nnz, ent, amax, slots, rows);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

bool sparsemax_tile_supported(int cols) { return cols == 32 || cols == 64 || cols == 128; }

int sparsemax_fwd_tile_launch(const float* x, float* p, int32_t* supp, int32_t* nnz, float* ent, int64_t* amax,
                              float* slots, int64_t rows, int cols, cudaStream_t st)
{
    // (stages, warps) per K: as many independent warp pipelines as 227 KB of shared memory hold
    static int variant = -1;
    if (variant < 0) {
        const char* e = getenv("SMLVM_TILE_VARIANT");
        variant = e ? atoi(e) : 0;
    }
    switch (cols) {
        case 32: return launch_tile<32, 2, 8>(x, p, supp, nnz, ent, amax, slots, rows, st);
        case 64:
            if (variant == 1) return launch_tile<64, 1, 12>(x, p, supp, nnz, ent, amax, slots, rows, st);
            return launch_tile<64, 2, 8>(x, p, supp, nnz, ent, amax, slots, rows, st);
        case 128:
            if (variant == 1) return launch_tile<128, 2, 6>(x, p, supp, nnz, ent, amax, slots, rows, st);
            // (fewer warp pipelines, measured at 1M x 128, N(0,1): 9 warps 0.2066, 8 warps 0.2089, 7 warps 0.2240 ms
            // against 0.2054 ms for 10 -- the forward wants them all)
            return launch_tile<128, 1, 10>(x, p, supp, nnz, ent, amax, slots, rows, st);
        default: return SMLVM_ERR_UNSUPPORTED;
    }
}

}  // namespace smlvm
