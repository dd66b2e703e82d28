// Support compaction (fill pass), TMA-staged tile kernel for sm_100a (K in {32,64,128,256}).
//
// Same output as compact_kernel<FILL> in compact.cu -- (value, row, col) of every p > 0 at
// offsets[row] + rank, i.e. torch.nonzero(p > 0) order, lvmhelpers/structmarg.py:116-126 --
// but the group-per-row version is ISSUE-bound (ncu r01: 129 warp instructions per row, 71 %
// issue utilisation, 52 % of the HBM roofline at 1M x 128): four ballots and their popcounts
// per 128-bit chunk are paid by every row although only ~3 of 128 entries are non-zero.
//
// Here a warp owns a tile of 32 rows (one cp.async.bulk into shared memory, 2-stage mbarrier
// ring per warp, no CTA-wide synchronisation) and a THREAD owns a row: it scans the row in
// XOR-swizzled 128-bit chunks (conflict-free across the 32 rows) into a K-bit mask, un-swizzles
// the mask by a nibble-index butterfly so that set bits are in column order, and emits its
// entries in a warp-uniform loop.  ~17 warp instructions per row; traffic unchanged (read P
// once, write 12 B per non-zero).
#include "common.cuh"
#include "tile_ptx.cuh"

namespace smlvm {

template <int K, int W>
__global__ void __launch_bounds__(W * 32, 1)
compact_fill_tile(const float* __restrict__ p, const int64_t* __restrict__ offsets, float* __restrict__ vals,
                  int32_t* __restrict__ row_idx, int32_t* __restrict__ col_idx, int64_t rows, int64_t capacity)
{
    constexpr int S = 2;
    constexpr int NC = K / 4;
    constexpr int MW = K / 32;
    constexpr int TILE_F = kTileRows * K;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tiles = reinterpret_cast<float*>(smem_raw) + (size_t)warp * S * TILE_F;
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(smem_raw + (size_t)W * S * TILE_F * 4);
    uint64_t* bars = bars_all + warp * S;
    if (threadIdx.x == 0) {
        for (int i = 0; i < W * S; ++i) mbar_init(smem_u32(bars_all + i), 1);
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
        bulk_load(smem_u32(tiles + stage * TILE_F), p + r0 * K, bytes, bar);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (t + s * stride < ntiles) issue_load(s, t + s * stride);
    }

    const uint32_t lsw = (uint32_t)lane & (uint32_t)((NC - 1) & 31);
    uint32_t parity = 0;
    int stage = 0;
    for (; t < ntiles; t += stride) {
        const int64_t row0 = t * kTileRows;
        const int trows = (int)min((int64_t)kTileRows, rows - row0);
        const bool live = lane < trows;
        const int64_t row = row0 + lane;
        int64_t out = live ? __ldg(offsets + row) : 0;  // in flight while the tile lands

        mbar_wait(smem_u32(bars + stage), (parity >> stage) & 1u);
        parity ^= 1u << stage;
        const uint32_t rowa = smem_u32(tiles + stage * TILE_F + (live ? lane : 0) * K);
        const uint32_t swz = rowa ^ (lsw << 4);

        uint32_t mask[MW];
#pragma unroll
        for (int w = 0; w < MW; ++w) mask[w] = 0u;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const float4 q = lds128(swz ^ (i << 4));
            const float xq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (xq[e] > 0.f) mask[(4 * i + e) >> 5] |= 1u << ((4 * i + e) & 31);
        }
        unswizzle_mask<MW>(mask, lsw);

        // entries in column order; warp-uniform trip counts (word by word)
#pragma unroll
        for (int w = 0; w < MW; ++w) {
            uint32_t mw = live ? mask[w] : 0u;
            const int steps = __reduce_max_sync(kFull, __popc(mw));
            for (int jj = 0; jj < steps; ++jj) {
                if (mw) {
                    const int col = w * 32 + __ffs(mw) - 1;
                    mw &= mw - 1;
                    if (out < capacity) {
                        if (vals != nullptr) vals[out] = lds32(rowa + (col << 2));
                        if (row_idx != nullptr) row_idx[out] = (int32_t)row;
                        if (col_idx != nullptr) col_idx[out] = col;
                    }
                    ++out;
                }
            }
        }

        // this stage is free again: every lane is done reading it
        __syncwarp();
        if (lane == 0 && t + S * stride < ntiles) issue_load(stage, t + S * stride);
        stage ^= 1;
    }
}

template <int K, int W>
static int launch_compact_tile(const float* p, const int64_t* offsets, float* vals, int32_t* row_idx,
                               int32_t* col_idx, int64_t rows, int64_t capacity, cudaStream_t st)
{
    auto kern = compact_fill_tile<K, W>;
    const size_t smem = (size_t)W * 2 * kTileRows * K * 4 + (size_t)W * 2 * 8;
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
    const int64_t cap = (int64_t)device_sm_count() * ctas_per_sm;
    const int grid = (int)(need < cap ? need : cap);
    kern<<<grid, W * 32, smem, st>>>(p, offsets, vals, row_idx, col_idx, rows, capacity);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

// Fill from the per-row support slots the forward kernel can emit (slots[rows][8] of (value, column),
// 64 B per row, unordered; column -1 = empty, slot 0 column -2 = "this row does not fit: scan P").
// A thread owns a row: it sorts its <= 8 pairs by column (register network) and writes them at
// offsets[row]; rows that do not fit are compacted by the whole warp from P.  Streams 64 + 8 B per
// row instead of the 4K-byte row.
__global__ void __launch_bounds__(256)
compact_fill_slots_kernel(const float* __restrict__ p, const float* __restrict__ slots,
                          const int64_t* __restrict__ offsets, float* __restrict__ vals,
                          int32_t* __restrict__ row_idx, int32_t* __restrict__ col_idx, int64_t rows, int K,
                          int64_t capacity)
{
    constexpr int NS = SMLVM_SUPPORT_SLOTS;
    const int lane = threadIdx.x & 31;
    const int64_t ngroups = (rows + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * 8;
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < ngroups; t += stride) {
        const int64_t r = t * 32 + lane;
        const bool live = r < rows;
        const int64_t b = live ? __ldg(offsets + r) : 0;
        const int cnt = live ? (int)(__ldg(offsets + r + 1) - b) : 0;
        float v[NS];
        int c[NS];
        const float4* src = reinterpret_cast<const float4*>(slots + (live ? r : 0) * (2 * NS));
#pragma unroll
        for (int j = 0; j < NS; j += 2) {
            const float4 q = __ldg(src + j / 2);
            v[j] = q.x; c[j] = __float_as_int(q.y); v[j + 1] = q.z; c[j + 1] = __float_as_int(q.w);
        }
        const bool scan = live && (cnt > NS || c[0] == -2);
        if (live && !scan) {
            // sort by column ascending, empty slots (-1) last: odd-even transposition network
#pragma unroll
            for (int j = 0; j < NS; ++j) c[j] = (c[j] < 0) ? 0x7fffffff : c[j];
#pragma unroll
            for (int pass = 0; pass < NS; ++pass) {
#pragma unroll
                for (int j = pass & 1; j + 1 < NS; j += 2) {
                    const bool sw = c[j] > c[j + 1];
                    const int ca = sw ? c[j + 1] : c[j], cb = sw ? c[j] : c[j + 1];
                    const float va = sw ? v[j + 1] : v[j], vb = sw ? v[j] : v[j + 1];
                    c[j] = ca; c[j + 1] = cb; v[j] = va; v[j + 1] = vb;
                }
            }
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                if (j < cnt && b + j < capacity) {
                    if (vals != nullptr) vals[b + j] = v[j];
                    if (row_idx != nullptr) row_idx[b + j] = (int32_t)r;
                    if (col_idx != nullptr) col_idx[b + j] = c[j];
                }
            }
        }
        // rows that did not fit the slots: the warp scans P, 32 columns at a time
        unsigned todo = __ballot_sync(kFull, scan);
        while (todo) {
            const int src_lane = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t rr = t * 32 + src_lane;
            int64_t out = __shfl_sync(kFull, b, src_lane);
            const float* pr = p + rr * K;
            for (int c0 = 0; c0 < K; c0 += 32) {
                const int col = c0 + lane;
                const float pv = col < K ? __ldg(pr + col) : 0.f;
                const unsigned bal = __ballot_sync(kFull, pv > 0.f);
                if (pv > 0.f && out + __popc(bal & ((1u << lane) - 1u)) < capacity) {
                    const int64_t o = out + __popc(bal & ((1u << lane) - 1u));
                    if (vals != nullptr) vals[o] = pv;
                    if (row_idx != nullptr) row_idx[o] = (int32_t)rr;
                    if (col_idx != nullptr) col_idx[o] = col;
                }
                out += __popc(bal);
            }
        }
    }
}

int compact_fill_slots_launch(const float* p, const float* slots, const int64_t* offsets, float* vals,
                              int32_t* row_idx, int32_t* col_idx, int64_t rows, int cols, int64_t capacity, cudaStream_t st)
{
    int64_t need = (rows + 255) / 256;
    int64_t cap = (int64_t)device_sm_count() * 16;
    compact_fill_slots_kernel<<<(int)(need < cap ? need : cap), 256, 0, st>>>(p, slots, offsets, vals, row_idx,
                                                                            col_idx, rows, cols, capacity);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

// slots from P for callers whose forward path did not emit them (small batches, other widths): a
// warp per row gathers the first 8 non-zeros; rows with more are marked "scan P".
__global__ void __launch_bounds__(256)
support_slots_kernel(const float* __restrict__ p, float* __restrict__ slots, int64_t rows, int K)
{
    constexpr int NS = SMLVM_SUPPORT_SLOTS;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += warps_total) {
        float* dst = slots + r * (2 * NS);
        if (lane < NS) {
            dst[2 * lane] = 0.f;
            dst[2 * lane + 1] = __int_as_float(-1);
        }
        __syncwarp();
        int out = 0;
        for (int c0 = 0; c0 < K; c0 += 32) {
            const int col = c0 + lane;
            const float pv = col < K ? __ldg(p + r * K + col) : 0.f;
            const unsigned bal = __ballot_sync(kFull, pv > 0.f);
            const int o = out + __popc(bal & ((1u << lane) - 1u));
            if (pv > 0.f && o < NS) {
                dst[2 * o] = pv;
                dst[2 * o + 1] = __int_as_float(col);
            }
            out += __popc(bal);
        }
        __syncwarp();
        if (out > NS && lane == 0) dst[1] = __int_as_float(-2);
    }
}

int support_slots_launch(const float* p, float* slots, int64_t rows, int cols, cudaStream_t st)
{
    int64_t need = (rows + 7) / 8;
    int64_t cap = (int64_t)device_sm_count() * 16;
    support_slots_kernel<<<(int)(need < cap ? need : cap), 256, 0, st>>>(p, slots, rows, cols);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

bool compact_tile_supported(int cols) { return cols == 32 || cols == 64 || cols == 128 || cols == 256; }

int compact_fill_tile_launch(const float* p, const int64_t* offsets, float* vals, int32_t* row_idx,
                             int32_t* col_idx, int64_t rows, int cols, int64_t capacity, cudaStream_t st)
{
    switch (cols) {
        case 32: return launch_compact_tile<32, 8>(p, offsets, vals, row_idx, col_idx, rows, capacity, st);
        case 64: return launch_compact_tile<64, 8>(p, offsets, vals, row_idx, col_idx, rows, capacity, st);
        case 128: return launch_compact_tile<128, 7>(p, offsets, vals, row_idx, col_idx, rows, capacity, st);
        case 256: return launch_compact_tile<256, 3>(p, offsets, vals, row_idx, col_idx, rows, capacity, st);
        default: return SMLVM_ERR_UNSUPPORTED;
    }
}

}  // namespace smlvm
