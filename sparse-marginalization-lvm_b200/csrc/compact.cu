// Support compaction for sm_100a: row-major stream compaction of the non-zero
// entries of a sparse distribution matrix into a ragged batch.
//
// replaces `mask = p.view(-1) > 0; t[mask]` (lvmhelpers/structmarg.py:116-126);
// output order == torch.nonzero(p > 0).
//
//   counts  (per-row number of p > 0; normally already written by the sparsemax
//            forward kernel, smlvm_compact_count exists for foreign inputs)
//   scan    single-pass exclusive prefix sum with decoupled look-back
//           (one ticketed tile per CTA, 64-bit {flag,value} descriptors)
//   fill    G-lane group per row, ballot + popc ranks, coalesced ragged stores
//
// Algorithmic bytes per row: 4K (read P) + 4 + 8 (count, offset) + 12 s_b.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace smlvm {

// ---- scan -----------------------------------------------------------------------

// 8192 rows per tile: 128 tiles at 1M rows, i.e. 4 look-back hops of 32 descriptors (2048-row tiles: 16 hops, 17.5 us)
constexpr int kScanThreads = 512;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// descriptor: bits 63..62 = state (0 empty, 1 tile aggregate, 2 inclusive prefix), low 62 = value
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagInc = 2ull << 62;
constexpr unsigned long long kValMask = (1ull << 62) - 1;

__global__ void __launch_bounds__(kScanThreads)
scan_counts_kernel(const int32_t* __restrict__ counts, int64_t* __restrict__ offsets, int64_t rows,
                   unsigned int* __restrict__ ticket, unsigned long long* __restrict__ desc)
{
    __shared__ long long s_warp[kScanThreads / 32];
    __shared__ long long s_prefix;
    __shared__ unsigned int s_tile;
    const int tid = threadIdx.x;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);  // tiles start in ticket order => no deadlock
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)tid * kScanItems;

    int v[kScanItems];
    if (base + kScanItems <= rows) {
        const int4* src = reinterpret_cast<const int4*>(counts + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; ++q) {
            const int4 a = __ldg(src + q);
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < rows) ? __ldg(counts + base + i) : 0;
    }
    long long local = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) local += v[i];

    // block exclusive scan of `local`
    long long incl = local;
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    long long warp_off = 0, tile_total = 0;
#pragma unroll
    for (int i = 0; i < kScanThreads / 32; ++i) {
        long long t = s_warp[i];
        if (i < w) warp_off += t;
        tile_total += t;
    }
    long long excl = warp_off + incl - local;

    // publish aggregate, look back for the exclusive prefix of this tile: warp 0 inspects 32
    // predecessor descriptors per step (a single thread walking them one by one made this kernel
    // a 29 us latency chain at 512 tiles)
    if (w == 0) {
        long long prefix = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch(&desc[0], kFlagInc | (unsigned long long)tile_total);
        } else {
            if (lane == 0) atomicExch(&desc[tile], kFlagAgg | (unsigned long long)tile_total);
            long long top = (long long)tile - 1;
            while (true) {
                const long long idx = top - lane;
                // before tile 0 there is a virtual inclusive prefix of 0
                const unsigned long long d = (idx >= 0) ? atomicAdd(&desc[idx], 0ull) : kFlagInc;
                const unsigned long long fl = d & ~kValMask;
                const unsigned inc = __ballot_sync(kFull, fl == kFlagInc);
                const unsigned empty = __ballot_sync(kFull, fl == 0ull);
                const int first_inc = inc ? (__ffs(inc) - 1) : 32;
                const unsigned needed = (first_inc >= 31) ? kFull : ((2u << first_inc) - 1u);
                if (empty & needed) { __nanosleep(20); continue; }  // a needed predecessor is not published yet
                long long val = (lane <= first_inc) ? (long long)(d & kValMask) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(kFull, val, o);
                prefix += val;
                if (first_inc < 32) break;
                top -= 32;
            }
            if (lane == 0) atomicExch(&desc[tile], kFlagInc | (unsigned long long)(prefix + tile_total));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    long long run = s_prefix + excl;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < rows) offsets[base + i] = run;
        run += v[i];
    }
    // the thread owning the last row also writes the grand total
    if (base < rows && base + kScanItems >= rows) offsets[rows] = run;
}

// ---- count / fill ---------------------------------------------------------------

constexpr int kThreads = 256;

template <int E, int G, bool VEC>
__device__ __forceinline__ void load_p(const float* __restrict__ base, int K, int l, float (&v)[E])
{
    if (VEC) {
#pragma unroll
        for (int c = 0; c < E / 4; ++c) {
            int col = 4 * (c * G + l);
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < K) q = ldg_stream4(base + col);
            v[4 * c + 0] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
            int col = e * G + l;
            v[e] = (col < K) ? ldg_stream(base + col) : 0.f;
        }
    }
}

// FILL = false: counts only.  One G-lane group per row, E slots per lane per pass;
// rows wider than E*G are walked in passes of E*G columns.
template <int E, int G, bool VEC, bool FILL>
__global__ void __launch_bounds__(kThreads)
compact_kernel(const float* __restrict__ p, const int64_t* __restrict__ offsets,
               int32_t* __restrict__ counts, float* __restrict__ vals, int32_t* __restrict__ row_idx,
               int32_t* __restrict__ col_idx, int64_t rows, int K, int64_t capacity)
{
    constexpr int RPW = kWarp / G;
    constexpr int SPAN = E * G;
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    const unsigned gmask = (G == 32) ? kFull : (((1u << G) - 1u) << (g * G));
    const unsigned lt = gmask & ((1u << lane) - 1u);
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kThreads / kWarp) + (threadIdx.x >> 5);

    for (; w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const float* pr = p + (live ? row : 0) * (int64_t)K;
        int64_t out = (FILL && live) ? __ldg(offsets + row) : 0;
        int total = 0;
        for (int c0 = 0; c0 < K; c0 += SPAN) {  // uniform trip count across the warp
            const int Kp = live ? min(K - c0, SPAN) : 0;
            float v[E];
            load_p<E, G, VEC>(pr + c0, Kp, l, v);
            if (VEC) {
#pragma unroll
                for (int c = 0; c < E / 4; ++c) {
                    unsigned f = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) f |= (v[4 * c + q] > 0.f) ? (1u << q) : 0u;
                    unsigned b0 = __ballot_sync(kFull, f & 1u), b1 = __ballot_sync(kFull, f & 2u);
                    unsigned b2 = __ballot_sync(kFull, f & 4u), b3 = __ballot_sync(kFull, f & 8u);
                    if (FILL) {
                        int rank = __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (f & (1u << q)) {
                                const int64_t o = out + rank;
                                if (o < capacity) {
                                    if (vals != nullptr) vals[o] = v[4 * c + q];
                                    if (row_idx != nullptr) row_idx[o] = (int32_t)row;
                                    if (col_idx != nullptr) col_idx[o] = c0 + 4 * (c * G + l) + q;
                                }
                                ++rank;
                            }
                        }
                    }
                    const int chunk = __popc(b0 & gmask) + __popc(b1 & gmask) + __popc(b2 & gmask) +
                                      __popc(b3 & gmask);
                    out += chunk;
                    total += chunk;
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const bool nz = v[e] > 0.f;
                    unsigned b = __ballot_sync(kFull, nz);
                    if (FILL && nz && out + __popc(b & lt) < capacity) {
                        const int64_t o = out + __popc(b & lt);
                        if (vals != nullptr) vals[o] = v[e];
                        if (row_idx != nullptr) row_idx[o] = (int32_t)row;
                        if (col_idx != nullptr) col_idx[o] = c0 + e * G + l;
                    }
                    const int chunk = __popc(b & gmask);
                    out += chunk;
                    total += chunk;
                }
            }
        }
        if (!FILL && live && l == 0) counts[row] = total;
    }
}

static int group_lanes(int K)
{
    int g = 1;
    while (g < 32 && g * 4 < K) g <<= 1;
    return g;
}

template <bool FILL>
static int launch_compact(const float* p, const int64_t* offsets, int32_t* counts, float* vals,
                          int32_t* row_idx, int32_t* col_idx, int64_t rows, int K, int64_t capacity, cudaStream_t st)
{
    const int G = group_lanes(K);
    const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
    const int rows_per_cta = (kThreads / kWarp) * (kWarp / G);
    int64_t need = (rows + rows_per_cta - 1) / rows_per_cta;
    int64_t cap = (int64_t)device_sm_count() * 16;
    const int grid = (int)(need < cap ? (need > 0 ? need : 1) : cap);
#define SMLVM_COMPACT_CASE(GG)                                                                     \
    case GG:                                                                                       \
        if (vec)                                                                                   \
            compact_kernel<4, GG, true, FILL><<<grid, kThreads, 0, st>>>(p, offsets, counts, vals, \
                                                                         row_idx, col_idx, rows, K, capacity); \
        else                                                                                       \
            compact_kernel<4, GG, false, FILL><<<grid, kThreads, 0, st>>>(p, offsets, counts, vals, \
                                                                          row_idx, col_idx, rows, K, capacity); \
        break;
    switch (G) {
        SMLVM_COMPACT_CASE(1)
        SMLVM_COMPACT_CASE(2)
        SMLVM_COMPACT_CASE(4)
        SMLVM_COMPACT_CASE(8)
        SMLVM_COMPACT_CASE(16)
        SMLVM_COMPACT_CASE(32)
    }
#undef SMLVM_COMPACT_CASE
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

// compact_tile.cu
int compact_fill_slots_launch(const float* p, const float* slots, const int64_t* offsets, float* vals,
                              int32_t* row_idx, int32_t* col_idx, int64_t rows, int cols, int64_t capacity, cudaStream_t st);
bool compact_tile_supported(int cols);
int compact_fill_tile_launch(const float* p, const int64_t* offsets, float* vals, int32_t* row_idx,
                             int32_t* col_idx, int64_t rows, int cols, int64_t capacity, cudaStream_t st);

// SMLVM_COMPACT=reg|tile overrides the path choice (tests run both).
// marg_step_dense.cu
bool step_dense_supported(int cols);
int compact_fill_coop_launch(const float* p, const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx,
                             int64_t rows, int K, int64_t capacity, cudaStream_t st);

static int compact_path_override()
{
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("SMLVM_COMPACT");
        cached = (e == nullptr) ? 0 : (strcmp(e, "reg") == 0 ? 1 : (strcmp(e, "tile") == 0 ? 2 : 0));
    }
    return cached;
}
constexpr int64_t kCompactTileMinRows = 4096;

}  // namespace smlvm

using namespace smlvm;

extern "C" size_t smlvm_scan_workspace_bytes(int64_t rows)
{
    if (rows < 0) rows = 0;
    int64_t tiles = (rows + kScanTile - 1) / kScanTile + 1;
    return 16 + (size_t)tiles * sizeof(unsigned long long);
}

extern "C" int smlvm_scan_counts(const int32_t* counts, int64_t* offsets, int64_t rows,
                                 void* workspace, size_t workspace_bytes, void* stream)
{
    if (rows < 0 || offsets == nullptr) return SMLVM_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    if (rows == 0) {
        cudaError_t e = cudaMemsetAsync(offsets, 0, sizeof(int64_t), st);
        return e == cudaSuccess ? SMLVM_OK : (int)e;
    }
    if (counts == nullptr || workspace == nullptr) return SMLVM_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(counts) & 15u) != 0) return SMLVM_ERR_ARG;
    const size_t need = smlvm_scan_workspace_bytes(rows);
    if (workspace_bytes < need) return SMLVM_ERR_WORKSPACE;
    cudaError_t e = cudaMemsetAsync(workspace, 0, need, st);
    if (e != cudaSuccess) return (int)e;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    unsigned long long* desc =
        reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(workspace) + 16);
    const int64_t tiles = (rows + kScanTile - 1) / kScanTile;
    scan_counts_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(counts, offsets, rows, ticket, desc);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_compact_count(const float* p, int32_t* counts, int64_t rows, int32_t cols,
                                   void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (p == nullptr || counts == nullptr) return SMLVM_ERR_ARG;
    return launch_compact<false>(p, nullptr, counts, nullptr, nullptr, nullptr, rows, cols, 0,
                                 as_stream(stream));
}

extern "C" int smlvm_compact_fill(const float* p, const int64_t* offsets, const float* support_slots,
                                  float* vals, int32_t* row_idx, int32_t* col_idx, int64_t rows,
                                  int32_t cols, int64_t capacity, void* stream)
{
    if (rows < 0 || cols < 1 || capacity < 0) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (p == nullptr || offsets == nullptr) return SMLVM_ERR_ARG;
    if (support_slots != nullptr)
        return compact_fill_slots_launch(p, support_slots, offsets, vals, row_idx, col_idx, rows, cols, capacity,
                                         as_stream(stream));
    const int ovr = compact_path_override();
    // wide rows of a large batch: G lanes per row, the run of 32 / G rows staged and stored coalesced (marg_step_dense.cu)
    if (ovr == 0 && step_dense_supported(cols) && rows >= kCompactTileMinRows && (reinterpret_cast<uintptr_t>(p) & 15u) == 0)
        return compact_fill_coop_launch(p, offsets, vals, row_idx, col_idx, rows, cols, capacity, as_stream(stream));
    if (compact_tile_supported(cols) && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && ovr != 1 &&
        (rows >= kCompactTileMinRows || ovr == 2))
        return compact_fill_tile_launch(p, offsets, vals, row_idx, col_idx, rows, cols, capacity, as_stream(stream));
    return launch_compact<true>(p, offsets, nullptr, vals, row_idx, col_idx, rows, cols, capacity,
                                as_stream(stream));
}
