// Sparsemax backward over the support only, fused clear + scatter kernel for sm_100a (K % 4 == 0).
//
// dx is zero off the support.  A warp owns 32 consecutive rows: it clears their K*128 bytes with
// coalesced 128-bit streaming stores, stages the rows' support entries (contiguous in the ragged
// arrays) with coalesced loads, and scatters g_j - mean(g) into the rows it has just cleared -- the
// lines are still in L2, so DRAM sees every byte of dx written once.  No shared-memory tile: at
// 4 KB of staging per warp 48 warps per SM hide the load latency (a TMA-tile variant with 36 KB per
// warp ran at 5 warps per SM and 9 % issue utilisation: 0.20 ms, the same as memset + scatter).
#include <stdlib.h>

#include "common.cuh"
#include "reduce_ws.cuh"
#include "sparsemax_exact.cuh"
#include "tile_ptx.cuh"

namespace smlvm {

constexpr int kStage = 256;       // support entries of a 32-row group staged in shared memory
constexpr int kBwdWarps = 8;

__global__ void __launch_bounds__(kBwdWarps * 32)
bwd_ragged_tile_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ cols,
                       const float* __restrict__ vals, const int32_t* __restrict__ supp,
                       const float* __restrict__ dp, const float* __restrict__ loss,
                       const float* __restrict__ loss_scale_dev, float loss_scale_host,
                       const float* __restrict__ dent, float* __restrict__ dx, float* __restrict__ dloss,
                       int64_t rows, int K, bool clear, int64_t n_items)
{
    __shared__ float s_all[kBwdWarps][4][kStage];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_val = s_all[warp][0];
    float* s_g = s_all[warp][1];
    float* s_loss = s_all[warp][2];
    int* s_col = reinterpret_cast<int*>(s_all[warp][3]);
    float ls = loss_scale_host;
    if (loss_scale_dev != nullptr) ls *= __ldg(loss_scale_dev);

    const int64_t ngroups = (rows + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * kBwdWarps;
    for (int64_t t = (int64_t)blockIdx.x * kBwdWarps + warp; t < ngroups; t += stride) {
        const int64_t row0 = t * 32;
        const int trows = (int)min((int64_t)32, rows - row0);
        const bool live = lane < trows;
        const int64_t row = row0 + lane;
        // (entries beyond the caller's lists -- a compaction that overflowed its capacity -- are ignored)
        const int64_t b = live ? min(__ldg(offsets + row), n_items) : 0, e = live ? min(__ldg(offsets + row + 1), n_items) : 0;
        const float kf = live ? (float)__ldg(supp + row) : 1.f;
        const float de = (dent != nullptr && live) ? __ldg(dent + row) : 0.f;

        if (clear) {  // clear the group's rows (contiguous trows*K floats)
            float* base = dx + row0 * K;
            const int n4 = trows * K / 4;
            for (int i = lane; i < n4; i += 32) stg_stream4(base + 4 * i, make_float4(0.f, 0.f, 0.f, 0.f));
        }

        const int64_t tb = __shfl_sync(kFull, b, 0);
        const int64_t te = __shfl_sync(kFull, e, trows - 1);
        const int tcnt = (int)(te - tb);
        __syncwarp();  // orders the clears before the scatters below (and protects the staging arrays)
        if (tcnt <= kStage) {
            for (int i = lane; i < tcnt; i += 32) {
                const float pj = __ldg(vals + tb + i);
                s_val[i] = pj;
                s_col[i] = __ldg(cols + tb + i);
                s_g[i] = dp != nullptr ? __ldg(dp + tb + i) : 0.f;
                s_loss[i] = loss != nullptr ? __ldg(loss + tb + i) : 0.f;
                if (dloss != nullptr) dloss[tb + i] = ls * pj;
            }
            __syncwarp();
            const int jb = (int)(b - tb), je = (int)(e - tb);
            float sum = 0.f;
            for (int j = jb; j < je; ++j) {
                float g = fmaf(ls, s_loss[j], s_g[j]);
                if (dent != nullptr) g = fmaf(de, dent_dp(s_val[j]), g);
                s_g[j] = g;  // own segment
                sum += g;
            }
            const float vhat = __fdiv_rn(sum, kf);
            float* out = dx + row * K;
            for (int j = jb; j < je; ++j) out[s_col[j]] = __fsub_rn(s_g[j], vhat);
        } else {
            float sum = 0.f;
            for (int64_t j = b; j < e; ++j) {
                const float pj = __ldg(vals + j);
                float g = dp != nullptr ? __ldg(dp + j) : 0.f;
                if (loss != nullptr) g = fmaf(ls, __ldg(loss + j), g);
                if (dent != nullptr) g = fmaf(de, dent_dp(pj), g);
                sum += g;
                if (dloss != nullptr) dloss[j] = ls * pj;
            }
            const float vhat = __fdiv_rn(sum, kf);
            float* out = dx + row * K;
            for (int64_t j = b; j < e; ++j) {
                const float pj = __ldg(vals + j);
                float g = dp != nullptr ? __ldg(dp + j) : 0.f;
                if (loss != nullptr) g = fmaf(ls, __ldg(loss + j), g);
                if (dent != nullptr) g = fmaf(de, dent_dp(pj), g);
                out[__ldg(cols + j)] = __fsub_rn(g, vhat);
            }
        }
        __syncwarp();
    }
}

// Variant that builds each 32-row tile of dx in shared memory (zero fill, scatter) and writes it with
// one bulk (TMA) store: DRAM sees dx written once, in full lines, and no scattered 4-byte stores.
__global__ void __launch_bounds__(512, 1)
bwd_ragged_tma_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ cols,
                      const float* __restrict__ vals, const int32_t* __restrict__ supp,
                      const float* __restrict__ dp, const float* __restrict__ loss,
                      const float* __restrict__ loss_scale_dev, float loss_scale_host,
                      const float* __restrict__ dent, float* __restrict__ dx, float* __restrict__ dloss,
                      int64_t rows, int K, int64_t n_items)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tile_f = kTileRows * K;
    float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * (tile_f + 4 * kStage);
    float* s_val = tile + tile_f;
    float* s_g = s_val + kStage;
    float* s_loss = s_g + kStage;
    int* s_col = reinterpret_cast<int*>(s_loss + kStage);
    const uint32_t ta = smem_u32(tile);
    float ls = loss_scale_host;
    if (loss_scale_dev != nullptr) ls *= __ldg(loss_scale_dev);

    const int64_t ntiles = (rows + kTileRows - 1) / kTileRows;
    for (int64_t t = (int64_t)warp * gridDim.x + blockIdx.x; t < ntiles; t += (int64_t)gridDim.x * nwarps) {
        const int64_t row0 = t * kTileRows;
        const int trows = (int)min((int64_t)kTileRows, rows - row0);
        const bool live = lane < trows;
        const int64_t row = row0 + lane;
        // (entries beyond the caller's lists -- a compaction that overflowed its capacity -- are ignored)
        const int64_t b = live ? min(__ldg(offsets + row), n_items) : 0, e = live ? min(__ldg(offsets + row + 1), n_items) : 0;
        const float kf = live ? (float)__ldg(supp + row) : 1.f;
        const float de = (dent != nullptr && live) ? __ldg(dent + row) : 0.f;
        const int64_t tb = __shfl_sync(kFull, b, 0);
        const int64_t te = __shfl_sync(kFull, e, trows - 1);
        const int tcnt = (int)(te - tb);
        const bool staged = tcnt <= kStage;
        // issue the staged loads before waiting on the previous store: their latency overlaps it
        float r_val[kStage / 32], r_g[kStage / 32], r_loss[kStage / 32];
        int r_col[kStage / 32];
        if (staged) {
#pragma unroll
            for (int u = 0; u < kStage / 32; ++u) {
                const int i = lane + 32 * u;
                const bool in = i < tcnt;
                r_val[u] = in ? __ldg(vals + tb + i) : 0.f;
                r_col[u] = in ? __ldg(cols + tb + i) : 0;
                r_g[u] = (in && dp != nullptr) ? __ldg(dp + tb + i) : 0.f;
                r_loss[u] = (in && loss != nullptr) ? __ldg(loss + tb + i) : 0.f;
            }
        }
        // the previous store of this warp must have finished READING the tile
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        for (int i = lane; i < tile_f / 4; i += 32) sts128(ta + (i << 4), make_float4(0.f, 0.f, 0.f, 0.f));
        if (staged) {
#pragma unroll
            for (int u = 0; u < kStage / 32; ++u) {
                const int i = lane + 32 * u;
                if (i < tcnt) {
                    s_val[i] = r_val[u];
                    s_col[i] = r_col[u];
                    s_g[i] = r_g[u];
                    s_loss[i] = r_loss[u];
                    if (dloss != nullptr) dloss[tb + i] = ls * r_val[u];
                }
            }
        }
        __syncwarp();
        if (staged) {
            const int jb = (int)(b - tb), je = (int)(e - tb);
            float sum = 0.f;
            for (int j = jb; j < je; ++j) {
                float g = fmaf(ls, s_loss[j], s_g[j]);
                if (dent != nullptr) g = fmaf(de, dent_dp(s_val[j]), g);
                s_g[j] = g;
                sum += g;
            }
            const float vhat = __fdiv_rn(sum, kf);
            for (int j = jb; j < je; ++j) tile[lane * K + s_col[j]] = __fsub_rn(s_g[j], vhat);
        } else {
            float sum = 0.f;
            for (int64_t j = b; j < e; ++j) {
                const float pj = __ldg(vals + j);
                float g = dp != nullptr ? __ldg(dp + j) : 0.f;
                if (loss != nullptr) g = fmaf(ls, __ldg(loss + j), g);
                if (dent != nullptr) g = fmaf(de, dent_dp(pj), g);
                sum += g;
                if (dloss != nullptr) dloss[j] = ls * pj;
            }
            const float vhat = __fdiv_rn(sum, kf);
            for (int64_t j = b; j < e; ++j) {
                const float pj = __ldg(vals + j);
                float g = dp != nullptr ? __ldg(dp + j) : 0.f;
                if (loss != nullptr) g = fmaf(ls, __ldg(loss + j), g);
                if (dent != nullptr) g = fmaf(de, dent_dp(pj), g);
                tile[lane * K + __ldg(cols + j)] = __fsub_rn(g, vhat);
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store(dx + row0 * K, ta, (uint32_t)trows * K * 4u);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait_read0();
}

// Dense backward driven by the forward's support slots (the gradient sources dp / loss are DENSE
// matrices, as in marg.py:125, but only their entries on the support matter): per 32-row tile the warp
// builds dX, then dL = ls * P, in ONE shared-memory tile (zero fill, scatter of <= 8 values per row)
// and writes each with one bulk (TMA) store.  Reads 64 B of slots + one sector of loss (and dp) per
// non-zero instead of the 8K..12K bytes of the rows of P, loss, dp; rows marked "scan" are processed
// densely by the whole warp straight into the tile.  One 16 KB tile per warp keeps 12 warps per SM
// (two tiles per warp: 6 warps, 0.28 ms -- the kernel is bound by the latency of the two dependent
// load rounds, slots then gathers, which more warps hide).
// FUSED: the same launch also produces what the step needs from the same slots -- the weighted loss itself
// (row_loss[r] = sum_j val_j L[r, col_j], total = total_scale * sum_r row_loss, fp64, fixed order: marg.py:125-126) and
// the support compaction (value, row, col) at offsets[r] in column order (structmarg.py:116-126), both clamped to the
// caller's capacity.  One read of the slots and ONE random gather of L per non-zero serve loss, gradient and lists
// (separately: wloss 0.088 ms + fill 0.040 ms + backward 0.214 ms at 1M x 128; the gather cost ~2x its bytes each time).
struct FusedExtras {
    float* row_loss;          // [rows] or NULL
    float* total;             // device scalar or NULL
    float total_scale;
    ReduceWs ws;
    const int64_t* offsets;   // [rows + 1] or NULL: no compaction
    float* vals;
    int32_t* row_idx;
    int32_t* col_idx;
    int64_t capacity;         // entries the three lists can hold (writes beyond are dropped)
};

template <bool FUSED>
__global__ void __launch_bounds__(512, 1)
bwd_slots_tma_kernel(const float* __restrict__ p, const int32_t* __restrict__ supp, const float* __restrict__ dp,
                     const float* __restrict__ loss, const float* __restrict__ loss_scale_dev,
                     float loss_scale_host, const float* __restrict__ dent, const float* __restrict__ slots,
                     const float* __restrict__ loss_slots, float* __restrict__ dx, float* __restrict__ dloss,
                     int64_t rows, int K, FusedExtras fx, bool loss_host)
{
    __shared__ double red[20];
    double acc_total = 0.0;
    constexpr int NS = SMLVM_SUPPORT_SLOTS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tile_f = kTileRows * K;
    float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * tile_f;
    const uint32_t ta = smem_u32(tile);
    float ls = loss_scale_host;
    if (loss_scale_dev != nullptr) ls *= __ldg(loss_scale_dev);

    const int64_t ntiles = (rows + kTileRows - 1) / kTileRows;
    for (int64_t t = (int64_t)warp * gridDim.x + blockIdx.x; t < ntiles; t += (int64_t)gridDim.x * nwarps) {
        const int64_t row0 = t * kTileRows;
        const int trows = (int)min((int64_t)kTileRows, rows - row0);
        const bool live = lane < trows;
        const int64_t row = row0 + lane;
        // ---- all loads first: slots, per-row scalars, then the gathers ---------------------------
        float v[NS], lv[NS], gv[NS];
        int c[NS];
        const float4* src = reinterpret_cast<const float4*>(slots + (live ? row : 0) * (2 * NS));
#pragma unroll
        for (int j = 0; j < NS; j += 2) {
            const float4 q = __ldg(src + j / 2);
            v[j] = q.x; c[j] = __float_as_int(q.y); v[j + 1] = q.z; c[j + 1] = __float_as_int(q.w);
        }
        const float kf = live ? (float)__ldg(supp + row) : 1.f;
        const float de = (dent != nullptr && live) ? __ldg(dent + row) : 0.f;
        const bool scan = live && c[0] == -2;
        if (loss_slots != nullptr && loss != nullptr) {
            // the weighted-loss forward already gathered L on the support: stream it
            const float4* ls4 = reinterpret_cast<const float4*>(loss_slots + (live ? row : 0) * NS);
            const float4 a4 = __ldg(ls4), b4 = __ldg(ls4 + 1);
            lv[0] = a4.x; lv[1] = a4.y; lv[2] = a4.z; lv[3] = a4.w;
            lv[4] = b4.x; lv[5] = b4.y; lv[6] = b4.z; lv[7] = b4.w;
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const bool on = live && !scan && c[j] >= 0;
            if (loss_slots == nullptr || loss == nullptr)
                lv[j] = (on && loss != nullptr) ? (loss_host ? ldg_gather_host(loss + row * K + c[j]) : ldg_gather(loss + row * K + c[j])) : 0.f;
            gv[j] = (on && dp != nullptr) ? ldg_gather(dp + row * K + c[j]) : 0.f;
        }
        float vhat = 0.f;
        if (live && !scan) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                float g = fmaf(ls, lv[j], gv[j]);
                if (dent != nullptr) g = fmaf(de, dent_dp(v[j]), g);
                g = (c[j] >= 0) ? g : 0.f;
                gv[j] = g;
                sum += g;
            }
            vhat = __fdiv_rn(sum, kf);
        }
        const unsigned scan_rows = __ballot_sync(kFull, scan);

        if (FUSED) {
            // ---- weighted loss of the row, in slot order (the arithmetic of wloss_slots_kernel) -------------------
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NS; ++j) s = fmaf(v[j], (live && !scan && c[j] >= 0) ? lv[j] : 0.f, s);
            if (scan) s = 0.f;
            unsigned todo = scan_rows;
            while (todo) {
                const int rl = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t rr = row0 + rl;
                float part = 0.f;
                for (int col = lane; col < K; col += 32) part = fmaf(__ldg(p + rr * K + col), __ldg(loss + rr * K + col), part);
                part = group_sum<32>(part);
                if (lane == rl) s = part;
            }
            if (live) {
                if (fx.row_loss != nullptr) fx.row_loss[row] = s;
                acc_total += (double)s;
            }
            // ---- support compaction: the row's pairs sorted by column at offsets[row] --------------------------
            if (fx.offsets != nullptr) {
                const int64_t ob = live ? __ldg(fx.offsets + row) : 0;
                if (live && !scan) {
                    float sv[NS];
                    int sc[NS];
#pragma unroll
                    for (int j = 0; j < NS; ++j) { sv[j] = v[j]; sc[j] = (c[j] < 0) ? 0x7fffffff : c[j]; }
#pragma unroll
                    for (int pass = 0; pass < NS; ++pass) {
#pragma unroll
                        for (int j = pass & 1; j + 1 < NS; j += 2) {
                            const bool sw = sc[j] > sc[j + 1];
                            const int ca = sw ? sc[j + 1] : sc[j], cb = sw ? sc[j] : sc[j + 1];
                            const float va = sw ? sv[j + 1] : sv[j], vb = sw ? sv[j] : sv[j + 1];
                            sc[j] = ca; sc[j + 1] = cb; sv[j] = va; sv[j + 1] = vb;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        if (sc[j] != 0x7fffffff && ob + j < fx.capacity) {
                            if (fx.vals != nullptr) fx.vals[ob + j] = sv[j];
                            if (fx.row_idx != nullptr) fx.row_idx[ob + j] = (int32_t)row;
                            if (fx.col_idx != nullptr) fx.col_idx[ob + j] = sc[j];
                        }
                    }
                }
                unsigned todo2 = scan_rows;
                while (todo2) {   // rows that did not fit the slots: the warp scans P, 32 columns at a time
                    const int rl = __ffs(todo2) - 1;
                    todo2 &= todo2 - 1;
                    const int64_t rr = row0 + rl;
                    int64_t out = __shfl_sync(kFull, ob, rl);
                    for (int c0 = 0; c0 < K; c0 += 32) {
                        const int col = c0 + lane;
                        const float pv = col < K ? __ldg(p + rr * K + col) : 0.f;
                        const unsigned bal = __ballot_sync(kFull, pv > 0.f);
                        const int64_t o = out + __popc(bal & ((1u << lane) - 1u));
                        if (pv > 0.f && o < fx.capacity) {
                            if (fx.vals != nullptr) fx.vals[o] = pv;
                            if (fx.row_idx != nullptr) fx.row_idx[o] = (int32_t)rr;
                            if (fx.col_idx != nullptr) fx.col_idx[o] = col;
                        }
                        out += __popc(bal);
                    }
                }
            }
        }

        for (int phase = 0; phase < (dloss != nullptr ? 2 : 1); ++phase) {
            // the previous store of this warp must have finished READING the tile
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
            for (int i = lane; i < tile_f / 4; i += 32) sts128(ta + (i << 4), make_float4(0.f, 0.f, 0.f, 0.f));
            __syncwarp();
            if (live && !scan) {
#pragma unroll
                for (int j = 0; j < NS; ++j)
                    if (c[j] >= 0) tile[lane * K + c[j]] = (phase == 0) ? __fsub_rn(gv[j], vhat) : ls * v[j];
            }
            // rows whose support does not fit the slots: dense, by the whole warp
            unsigned todo = scan_rows;
            while (todo) {
                const int rl = __ffs(todo) - 1;
                todo &= todo - 1;
                const int64_t rr = row0 + rl;
                if (phase == 1) {
                    for (int col = lane; col < K; col += 32) tile[rl * K + col] = ls * __ldg(p + rr * K + col);
                    continue;
                }
                const float de_r = __shfl_sync(kFull, de, rl), kf_r = __shfl_sync(kFull, kf, rl);
                float sum = 0.f;
                for (int col = lane; col < K; col += 32) {
                    const float pv = __ldg(p + rr * K + col);
                    float g = dp != nullptr ? __ldg(dp + rr * K + col) : 0.f;
                    if (loss != nullptr) g = fmaf(ls, __ldg(loss + rr * K + col), g);
                    if (dent != nullptr) g = fmaf(de_r, dent_dp(pv), g);
                    g = (pv != 0.f) ? g : 0.f;
                    tile[rl * K + col] = g;  // parked; centred below
                    sum += g;
                }
                sum = group_sum<32>(sum);
                const float vh = __fdiv_rn(sum, kf_r);
                for (int col = lane; col < K; col += 32) {
                    const float pv = __ldg(p + rr * K + col);
                    tile[rl * K + col] = (pv != 0.f) ? __fsub_rn(tile[rl * K + col], vh) : 0.f;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store((phase == 0 ? dx : dloss) + row0 * K, ta, (uint32_t)trows * K * 4u);
                bulk_commit();
            }
        }
    }
    if (lane == 0) bulk_wait_read0();
    if (FUSED && fx.total != nullptr) finish_totals(fx.ws, acc_total, 0.0, fx.total_scale, fx.total, nullptr, red);
}

bool bwd_slots_supported(int cols) { return cols % 4 == 0 && cols >= 4 && cols <= 512; }

template <bool FUSED>
static int bwd_slots_launch_t(const float* p, const int32_t* supp, const float* dp, const float* loss, const float* lsd,
                              float lsh, const float* dent, const float* slots, const float* loss_slots, float* dx,
                              float* dloss, int64_t rows, int K, const FusedExtras& fx, cudaStream_t st)
{
    const size_t per_warp = (size_t)kTileRows * K * 4;
    // Tiles of up to 16 KB (K <= 128): as many warps as fit the 164 KB shared-memory configuration, i.e. 10 at K = 128 --
    // measured against 12 (192 KB, the 196 KB configuration, 60 instead of 92 KB of L1): 0.271 -> 0.260 ms at 1M x 128 with
    // 3.5 entries per row, 0.258 -> 0.238 ms with 1.9; 11 and 9 warps 0.264, 8 warps 0.277 (SMLVM_BWD_SLOTS_WARPS sweeps).
    int warps = (int)(((per_warp <= 16 * 1024 ? 162 : 200) * 1024) / per_warp);
    warps = warps > 16 ? 16 : (warps < 1 ? 1 : warps);
    {
        static int wenv = -1;
        if (wenv < 0) { const char* ev = getenv("SMLVM_BWD_SLOTS_WARPS"); wenv = ev ? atoi(ev) : 0; }
        if (wenv > 0 && wenv <= 16 && per_warp * wenv <= (size_t)200 * 1024) warps = wenv;
    }
    const size_t smem = per_warp * warps;
    static bool configured_dev[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured_dev[dev & 63]) {
        cudaError_t ea = cudaFuncSetAttribute(bwd_slots_tma_kernel<FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              200 * 1024);
        if (ea != cudaSuccess) return (int)ea;
        configured_dev[dev & 63] = true;
    }
    static int gmul = -1;
    if (gmul < 0) { const char* ev = getenv("SMLVM_BWD_SLOTS_GRID"); gmul = ev ? atoi(ev) : 1; if (gmul < 1) gmul = 1; }
    const int64_t nt = (rows + kTileRows - 1) / kTileRows;
    const int64_t nd = (nt + warps - 1) / warps;
    const int64_t cp = (int64_t)device_sm_count() * gmul;
    // the loss matrix may be pinned host memory read in place (include/smlvm.h, smlvm_marg_step_fused): plain gathers then
    bool loss_host = false;
    if (loss != nullptr) {
        static int hint = -1;
        if (hint < 0) {
            const char* ev = getenv("SMLVM_GATHER_HOST_HINT");     // development: 1 keeps the 64-byte L2 hint on host memory
            hint = (ev && atoi(ev)) ? 1 : 0;
        }
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, loss) == cudaSuccess) loss_host = at.type == cudaMemoryTypeHost && hint == 0;
        else cudaGetLastError();
    }
    bwd_slots_tma_kernel<FUSED><<<(int)(nd < cp ? nd : cp), warps * 32, smem, st>>>(p, supp, dp, loss, lsd, lsh, dent, slots,
                                                                                     loss_slots, dx, dloss, rows, K, fx, loss_host);
    cudaError_t e2 = cudaGetLastError();
    return e2 == cudaSuccess ? SMLVM_OK : (int)e2;
}

int bwd_slots_launch(const float* p, const int32_t* supp, const float* dp, const float* loss, const float* lsd,
                     float lsh, const float* dent, const float* slots, const float* loss_slots, float* dx,
                     float* dloss, int64_t rows, int K, cudaStream_t st)
{
    FusedExtras none = {};
    return bwd_slots_launch_t<false>(p, supp, dp, loss, lsd, lsh, dent, slots, loss_slots, dx, dloss, rows, K, none, st);
}

// loss forward + backward (+ compaction) of the categorical step in one launch; see FusedExtras
int step_fused_launch(const float* p, const int32_t* supp, const float* loss, float scale, const float* dent,
                      const float* slots, const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx,
                      int64_t capacity, float* row_loss, float* total, void* workspace, float* dx, float* dloss,
                      int64_t rows, int K, cudaStream_t st)
{
    FusedExtras fx;
    fx.row_loss = row_loss;
    fx.total = total;
    fx.total_scale = scale;
    fx.ws = ws_view(workspace);
    fx.offsets = offsets;
    fx.vals = vals;
    fx.row_idx = row_idx;
    fx.col_idx = col_idx;
    fx.capacity = capacity;
    return bwd_slots_launch_t<true>(p, supp, nullptr, loss, nullptr, scale, dent, slots, nullptr, dx, dloss, rows, K, fx, st);
}

bool bwd_ragged_tile_supported(int cols) { return cols % 4 == 0 && cols >= 4; }

int bwd_ragged_tile_launch(const int64_t* offsets, const int32_t* cols_idx, const float* vals, const int32_t* supp,
                           const float* dp, const float* loss, const float* lsd, float lsh, const float* dent,
                           float* dx, float* dloss, int64_t rows, int K, int64_t n_items, cudaStream_t st)
{
    const int64_t ngroups = (rows + 31) / 32;
    const int64_t need = (ngroups + kBwdWarps - 1) / kBwdWarps;
    const int64_t cap = (int64_t)device_sm_count() * 6;
    const int grid = (int)(need < cap ? need : cap);
    static int tma = -1;
    if (tma < 0) {
        const char* ev = getenv("SMLVM_BWD_RAGGED_TMA");
        tma = (ev && atoi(ev) == 0) ? 0 : 1;
    }
    if (tma && K <= 512 && rows >= 4096) {
        const size_t per_warp = ((size_t)kTileRows * K + 4 * kStage) * 4;
        int warps = (int)((200 * 1024) / per_warp);
        warps = warps > 16 ? 16 : (warps < 1 ? 1 : warps);
        {
            static int wenv = -1;
            if (wenv < 0) { const char* ev = getenv("SMLVM_BWD_RAGGED_WARPS"); wenv = ev ? atoi(ev) : 0; }
            if (wenv > 0 && wenv < warps) warps = wenv;
        }
        const size_t smem = per_warp * warps;
        static bool configured_dev[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured_dev[dev & 63]) {
            cudaError_t ea = cudaFuncSetAttribute(bwd_ragged_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  200 * 1024);
            if (ea != cudaSuccess) return (int)ea;
            configured_dev[dev & 63] = true;
        }
        const int64_t nt = (rows + kTileRows - 1) / kTileRows;
        const int64_t nd = (nt + warps - 1) / warps;
        const int64_t cp = (int64_t)device_sm_count();
        bwd_ragged_tma_kernel<<<(int)(nd < cp ? nd : cp), warps * 32, smem, st>>>(offsets, cols_idx, vals, supp, dp, loss,
                                                                                   lsd, lsh, dent, dx, dloss, rows, K, n_items);
        cudaError_t e2 = cudaGetLastError();
        return e2 == cudaSuccess ? SMLVM_OK : (int)e2;
    }
    // clear dx with a memset node (7.2 TB/s), then scatter: fusing the clear into the kernel measured
    // no better (0.21 ms vs 0.20), the scattered stores queue behind the streaming clears in the LSU
    static int fused = -1;
    if (fused < 0) {
        const char* ev = getenv("SMLVM_BWD_RAGGED_FUSED_CLEAR");
        fused = (ev && atoi(ev)) ? 1 : 0;
    }
    if (!fused) {
        cudaError_t em = cudaMemsetAsync(dx, 0, (size_t)rows * K * sizeof(float), st);
        if (em != cudaSuccess) return (int)em;
    }
    bwd_ragged_tile_kernel<<<grid, kBwdWarps * 32, 0, st>>>(offsets, cols_idx, vals, supp, dp, loss, lsd, lsh, dent, dx,
                                                           dloss, rows, K, fused != 0, n_items);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

}  // namespace smlvm
