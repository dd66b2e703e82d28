// k-best bit-vectors of n independent bits for sm_100a.
//
// replaces lvmhelpers.pbinary_topk.batched_topk (lvmhelpers/pbinary_topk.pyx:23-34) ->
// topk()/merge_branch() (lvmhelpers/binary_topk.h:57-129).  Results are bit-exact with the
// reference: same list initialisation (x0 >= 0, :110), same fp32 score accumulation in bit
// order (:27-29), same merge tie rule ("bit on" only if STRICTLY greater, :72), same
// truncation to k after every bit.
//
// Mapping: one G-lane group per row (G = 8/16/32 >= k), lane r owns list slot r: its fp32
// score and its configuration as W 32-bit words in registers.  The sequential list merge of
// the reference becomes a merge-path search: output slot r finds by binary search (shuffles
// inside the group) how many "bit off" items precede it, then gathers score and words from
// the source lane.  n steps of O(log k) shuffles per row; the scores row is staged in
// shared memory; the fp32 {0,1} output (4*k*n B/row, the API's format) is written coalesced.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace smlvm {

constexpr int kThreads = 256;

template <int G, int W>
__global__ void __launch_bounds__(kThreads)
topk_bits_kernel(const float* __restrict__ x, float* __restrict__ configs,
                 uint32_t* __restrict__ bits, float* __restrict__ dp_scores, int64_t rows, int n,
                 int k, int n_pad)
{
    extern __shared__ float xs_all[];
    constexpr int GPW = kWarp / G;
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    float* xs = xs_all + ((threadIdx.x >> 5) * GPW + g) * n_pad;
    const int words_n = (n + 31) >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kThreads / kWarp) + (threadIdx.x >> 5);

    for (; w * GPW < rows; w += warps_total) {
        const int64_t row = w * GPW + g;
        const bool live = row < rows;
        const float* xr = x + (live ? row : 0) * (int64_t)n;
        __syncwarp();
        for (int c = l; c < n; c += G) xs[c] = live ? ldg_stream(xr + c) : 0.f;
        __syncwarp();

        // list = [(x0,{0}), (0,{})] ordered by x0 >= 0        (binary_topk.h:104-116)
        const float x0 = xs[0];
        const bool on_first = x0 >= 0.f;
        float sc = 0.f;
        uint32_t cfg[W];
#pragma unroll
        for (int i = 0; i < W; ++i) cfg[i] = 0u;
        if (l == 0) { sc = on_first ? x0 : 0.f; cfg[0] = on_first ? 1u : 0u; }
        if (l == 1) { sc = on_first ? 0.f : x0; cfg[0] = on_first ? 0u : 1u; }
        int len = 2;

        for (int i = 1; i < n; ++i) {
            const float v = xs[i];
            const int outlen = min(k, 2 * len);
            const int r = min(l, 2 * len - 1);
            int lo = max(0, r - len), hi = min(r, len);
            // merge path: a = number of "off" items among the first r outputs
            for (int s = 0; (1 << s) <= len; ++s) {
                const int mid = (lo + hi) >> 1;
                const float am = __shfl_sync(kFull, sc, min(mid, len - 1), G);
                const float bm = __fadd_rn(__shfl_sync(kFull, sc, max(r - mid - 1, 0), G), v);
                if (lo < hi) {
                    if (am >= bm) lo = mid + 1; else hi = mid;
                }
            }
            const int a = lo, b = r - lo;
            const float Aa = __shfl_sync(kFull, sc, min(a, len - 1), G);
            const float Bb = __fadd_rn(__shfl_sync(kFull, sc, min(b, len - 1), G), v);
            const bool takeA = (a < len) && (b >= len || Aa >= Bb);
            const int src = takeA ? a : b;
            sc = takeA ? Aa : Bb;
            const int wi = i >> 5;
            const uint32_t bit = takeA ? 0u : (1u << (i & 31));
#pragma unroll
            for (int q = 0; q < W; ++q) {
                if (q <= wi) {  // words above the current bit are still all zero
                    uint32_t t = __shfl_sync(kFull, cfg[q], src, G);
                    cfg[q] = t | (q == wi ? bit : 0u);
                }
            }
            len = outlen;
        }

        if (live) {
            if (dp_scores != nullptr && l < k) dp_scores[row * k + l] = sc;
            if (bits != nullptr && l < k) {
#pragma unroll
                for (int q = 0; q < W; ++q)
                    if (q < words_n) bits[(row * k + l) * words_n + q] = cfg[q];
            }
        }
        if (configs != nullptr) {
            for (int j = 0; j < k; ++j) {
                float* out = configs + ((live ? row : 0) * k + j) * (int64_t)n;
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const uint32_t word = __shfl_sync(kFull, cfg[q], j, G);
                    if (live && q < words_n) {
                        const int hi_c = min(n, q * 32 + 32);
                        for (int c = q * 32 + l; c < hi_c; c += G)
                            stg_stream(out + c, (float)((word >> (c & 31)) & 1u));
                    }
                }
            }
        }
    }
}

// scores_k[r,j] = sum_i bit(r,j,i) * x[r,i]; one thread per (row, j) walking the set bits of its
// configuration; fp64 accumulate (order-insensitive to ~1e-16, rounded once to fp32).
// (Warp-per-row variants measured slower at 1M x 128, k = 10: butterfly sums per configuration 0.92 ms, a
// reduce-scatter of the 16 fp64 partials over the lane bits 1.34 ms, against 0.67 ms here.)
__global__ void __launch_bounds__(kThreads)
bits_score_fwd_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ x,
                      float* __restrict__ scores_k, int64_t rows, int n, int k)
{
    const int words_n = (n + 31) >> 5;
    const int64_t total = rows * k;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * kThreads) {
        const int64_t r = div_idx(t, k);
        const uint32_t* b = bits + t * words_n;
        const float* xr = x + r * (int64_t)n;
        double acc = 0.0;
        for (int q = 0; q < words_n; ++q) {
            uint32_t word = __ldg(b + q);
            while (word) {
                const int i = __ffs(word) - 1;
                word &= word - 1;
                acc += (double)__ldg(xr + q * 32 + i);
            }
        }
        scores_k[t] = (float)acc;
    }
}

// Same thread mapping for n % 4 == 0: instead of chasing the set bits (ffs -> address -> load -> add, a
// ~60-cycle dependent step per bit) the thread streams its row of x as float4s (L1 hits: the k threads of a
// row read the same 4n bytes) and adds under the bit predicates into four fp64 accumulators.
__global__ void __launch_bounds__(kThreads)
bits_score_fwd_vec_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ x,
                          float* __restrict__ scores_k, int64_t rows, int n, int k)
{
    const int words_n = (n + 31) >> 5;
    const int64_t total = rows * k;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * kThreads) {
        const int64_t r = div_idx(t, k);
        const uint32_t* b = bits + t * words_n;
        const float4* xr = reinterpret_cast<const float4*>(x + r * (int64_t)n);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int q = 0; q < words_n; ++q) {
            const uint32_t word = __ldg(b + q);
            const int cmax = min(8, (n - q * 32) >> 2);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < cmax) {
                    const float4 f = __ldg(xr + q * 8 + c);
                    const uint32_t nib = word >> (4 * c);
                    if (nib & 1u) a0 += (double)f.x;
                    if (nib & 2u) a1 += (double)f.y;
                    if (nib & 4u) a2 += (double)f.z;
                    if (nib & 8u) a3 += (double)f.w;
                }
            }
        }
        scores_k[t] = (float)((a0 + a1) + (a2 + a3));
    }
}

// Re-scoring for k-best lists (n % 4 == 0, n <= 128).  The vec kernel above converts every score k times
// (k*n F2F.F64.F32 per row on the quarter-rate conversion pipe: what bounds it) although the k
// configurations of a k-best list differ from the best one in a handful of bits.  Here a warp takes 8 rows
// at a time: (1) 4 lanes per row stream the row once (float4s, kept in shared memory) and sum it under the
// bits of configuration 0 in fp64 (lane partials + butterfly); (2) a lane per (row, j) walks the set bits
// of cfg_j ^ cfg_0 (one loop over the 128-bit difference: the warp runs max-over-lanes iterations) and adds
// +-x to that base -- or, when the configuration has fewer set bits than its difference (arbitrary
// patterns), or the base is not finite, the set bits of cfg_j from zero.
// Any order of the fp64 sum is within 1e-16 of any other, rounded once to fp32.
constexpr int kScoreRows = 8;
constexpr int kScorePad = 16;  // floats between staged rows: the two rows of a 128-bit store phase hit disjoint banks

__global__ void __launch_bounds__(kThreads)
bits_score_fwd_delta_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ x,
                            float* __restrict__ scores_k, int64_t rows, int n, int k, uint32_t k_magic)
{
    extern __shared__ __align__(16) float xs_all[];
    constexpr int kWarps = kThreads / kWarp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int words_n = (n + 31) >> 5, n4 = n >> 2, ns = n + kScorePad;
    float* xs = xs_all + warp * kScoreRows * ns;
    double* base_s = reinterpret_cast<double*>(xs_all + kWarps * kScoreRows * ns) + warp * kScoreRows;
    const int g = lane & 3, rr1 = lane >> 2;
    const int64_t warps_total = (int64_t)gridDim.x * kWarps;
    for (int64_t blk = (int64_t)blockIdx.x * kWarps + warp; blk * kScoreRows < rows; blk += warps_total) {
        const int64_t r0 = blk * kScoreRows;
        const int cnt = (int)((rows - r0) < (int64_t)kScoreRows ? (rows - r0) : (int64_t)kScoreRows);
        __syncwarp();
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (rr1 < cnt) {
            const float4* xr = reinterpret_cast<const float4*>(x + (r0 + rr1) * (int64_t)n);
            const uint32_t* b0 = bits + (r0 + rr1) * (int64_t)k * words_n;
            float4* xo = reinterpret_cast<float4*>(xs + rr1 * ns);
            auto take = [&](int c) {
                const float4 f = __ldg(xr + c);
                xo[c] = f;
                const uint32_t nib = __ldg(b0 + (c >> 3)) >> (4 * (c & 7));
                if (nib & 1u) a0 += (double)f.x;
                if (nib & 2u) a1 += (double)f.y;
                if (nib & 4u) a2 += (double)f.z;
                if (nib & 8u) a3 += (double)f.w;
            };
            if (n4 == 32) {  // n = 128: eight float4s per lane, no loop control
#pragma unroll
                for (int it = 0; it < 8; ++it) take(g + 4 * it);
            } else {
#pragma unroll 4
                for (int c = g; c < n4; c += 4) take(c);
            }
        }
        double a = (a0 + a1) + (a2 + a3);
        a += __shfl_xor_sync(kFull, a, 2);
        a += __shfl_xor_sync(kFull, a, 1);
        if (g == 0 && rr1 < cnt) base_s[rr1] = a;
        __syncwarp();
        const int tasks = cnt * k;
        for (int t = lane; t < tasks; t += 32) {
            const int rr = k == 1 ? t : (int)__umulhi((uint32_t)t, k_magic);  // t / k, k_magic = ceil(2^32 / k)
            const int j = t - rr * k;
            const uint32_t* b0 = bits + (r0 + rr) * (int64_t)k * words_n;
            const uint32_t* bj = b0 + j * words_n;
            uint32_t cw[4], dw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                cw[q] = q < words_n ? __ldg(bj + q) : 0u;
                dw[q] = q < words_n ? (cw[q] ^ __ldg(b0 + q)) : 0u;
            }
            const int pc = __popc(cw[0]) + __popc(cw[1]) + __popc(cw[2]) + __popc(cw[3]);
            const int pd = __popc(dw[0]) + __popc(dw[1]) + __popc(dw[2]) + __popc(dw[3]);
            const double base = base_s[rr];
            const bool use_diff = pd < pc && fabs(base) <= 1.7976931348623157e308;
            const uint64_t c_lo = cw[0] | ((uint64_t)cw[1] << 32), c_hi = cw[2] | ((uint64_t)cw[3] << 32);
            uint64_t m_lo = use_diff ? (dw[0] | ((uint64_t)dw[1] << 32)) : c_lo;
            uint64_t m_hi = use_diff ? (dw[2] | ((uint64_t)dw[3] << 32)) : c_hi;
            double acc = use_diff ? base : 0.0;
            const float* xrow = xs + rr * ns;
            while ((m_lo | m_hi) != 0ull) {
                const bool lo = m_lo != 0ull;
                const uint64_t cur = lo ? m_lo : m_hi;
                const int i = __ffsll((long long)cur) - 1;
                const uint64_t rest = cur & (cur - 1ull);
                m_lo = lo ? rest : m_lo;
                m_hi = lo ? m_hi : rest;
                const double v = (double)xrow[i + (lo ? 0 : 64)];
                acc += (((lo ? c_lo : c_hi) >> i) & 1ull) ? v : -v;
            }
            scores_k[r0 * k + t] = (float)acc;
        }
    }
}

// dx[r,i] = sum_j ds[r,j] * bit(r,j,i).  One warp per row: ds and the k*W words are staged in
// shared memory, lane l owns columns l, l+32, ...; every shared read is a broadcast.
__global__ void __launch_bounds__(kThreads)
bits_score_bwd_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ ds,
                      float* __restrict__ dx, int64_t rows, int n, int k)
{
    extern __shared__ __align__(16) float stage_all[];
    const int words_n = (n + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kpad = (k + 3) & ~3;                      // keeps the word array 16-byte aligned
    const int per_warp = kpad + ((k * words_n + 3) & ~3);
    float* dss = stage_all + warp * per_warp;
    uint32_t* ws = reinterpret_cast<uint32_t*>(dss + kpad);
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / kWarp);
    for (int64_t r = (int64_t)blockIdx.x * (kThreads / kWarp) + warp; r < rows; r += warps_total) {
        __syncwarp();
        for (int j = lane; j < k; j += 32) dss[j] = __ldg(ds + r * k + j);
        for (int e = lane; e < k * words_n; e += 32) ws[e] = __ldg(bits + r * k * words_n + e);
        __syncwarp();
        if (words_n == 4) {
            // n in (96, 128]: the four words of a configuration in one 128-bit shared load
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int j = 0; j < k; ++j) {
                const uint4 w4 = *reinterpret_cast<const uint4*>(ws + 4 * j);
                const float dj = dss[j];
                a0 += ((w4.x >> lane) & 1u) ? dj : 0.f;
                a1 += ((w4.y >> lane) & 1u) ? dj : 0.f;
                a2 += ((w4.z >> lane) & 1u) ? dj : 0.f;
                a3 += ((w4.w >> lane) & 1u) ? dj : 0.f;
            }
            float* out = dx + r * n;
            stg_stream(out + lane, a0);
            stg_stream(out + 32 + lane, a1);
            stg_stream(out + 64 + lane, a2);
            if (96 + lane < n) stg_stream(out + 96 + lane, a3);
        } else {
            for (int q = 0; q < words_n; ++q) {
                float acc = 0.f;
                for (int j = 0; j < k; ++j)
                    acc += ((ws[j * words_n + q] >> lane) & 1u) ? dss[j] : 0.f;
                const int i = q * 32 + lane;
                if (i < n) stg_stream(dx + r * n + i, acc);
            }
        }
    }
}

// Same sums for n % CPT == 0, straight from global memory: a thread per CPT (8 / 16 / 32) columns of one row
// of dx, its k configuration words and k upstream gradients are independent loads (L1 hits: a row's k*W
// words and k floats are shared by the n/CPT threads of the row), no staging and no warp barrier.  The
// loads bound this kernel (4 columns per thread: 0.38 ms, 8: 0.26 ms at 1M x 128, k = 10).
template <int CPT>
__global__ void __launch_bounds__(256)
bits_score_bwd_vec_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ ds, float* __restrict__ dx,
                          int64_t rows, int n, int k)
{
    const int words_n = (n + 31) >> 5, qn = n / CPT;
    const int64_t total = rows * qn, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t r = div_idx(t, qn);
        const int c = (int)(t - r * qn) * CPT;
        const uint32_t* b = bits + r * (int64_t)k * words_n + (c >> 5);
        const float* d = ds + r * k;
        const int sh = c & 31;
        float a[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) a[e] = 0.f;
#pragma unroll 2
        for (int j = 0; j < k; ++j) {
            const uint32_t w = __ldg(b + j * words_n) >> sh;
            const float dj = __ldg(d + j);
#pragma unroll
            for (int e = 0; e < CPT; ++e) a[e] += ((w >> e) & 1u) ? dj : 0.f;
        }
#pragma unroll
        for (int e = 0; e < CPT; e += 4)
            stg_stream4(dx + t * CPT + e, make_float4(a[e], a[e + 1], a[e + 2], a[e + 3]));
    }
}

template <int G>
static void launch_topk_w(int W, int grid, size_t smem, cudaStream_t st, const float* x,
                          float* configs, uint32_t* bits, float* dp_scores, int64_t rows, int n,
                          int k, int n_pad)
{
#define SMLVM_TOPK_CASE(WW)                                                                       \
    case WW:                                                                                      \
        cudaFuncSetAttribute(topk_bits_kernel<G, WW>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)smem);                                                          \
        topk_bits_kernel<G, WW><<<grid, kThreads, smem, st>>>(x, configs, bits, dp_scores, rows, n, \
                                                              k, n_pad);                          \
        break;
    switch (W) {
        SMLVM_TOPK_CASE(1)
        SMLVM_TOPK_CASE(2)
        SMLVM_TOPK_CASE(4)
        SMLVM_TOPK_CASE(8)
        SMLVM_TOPK_CASE(16)
    }
#undef SMLVM_TOPK_CASE
}

// topk_tile.cu
bool topk_tile_supported(int n, int k);
int topk_tile_launch(const float* x, float* configs, uint32_t* bits, float* dp_scores, int64_t rows, int n,
                     int k, cudaStream_t st);

// topk_blk.cu
bool topk_blk_supported(int n, int k);
int topk_blk_launch(const float* x, float* configs, uint32_t* bits, float* dp_scores, int64_t rows, int n,
                    int k, cudaStream_t st);

// The thread-per-row kernels win on throughput (topk_blk.cu: 64-bit list entries, configurations materialised every
// 28 bits; topk_tile.cu: configurations riding along), the group-per-row kernel on latency (a row is a
// chain of n dependent merges): small batches stay on the latter.  SMLVM_TOPK=reg|tile|blk overrides.
// (Round 2 measured a third variant -- configurations parked in a pool of k slots per row, list entries (score, slot
// id), only the min(#A, #B) entries that survive on both sides of a merge copied: 1/8 of the shared-memory bytes, but
// 394 -> 250 (unrolled) warp instructions per bit against 220, the copies run at the warp's worst row: 0.527 vs 0.490
// ms at 262144 x 128, k = 10, 1.15 vs 0.94 ms at k = 16.  Dropped)
static int topk_path_override()
{
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("SMLVM_TOPK");
        cached = (e == nullptr) ? 0 : (strcmp(e, "reg") == 0 ? 1 : (strcmp(e, "tile") == 0 ? 2 : (strcmp(e, "blk") == 0 ? 3 : 0)));
    }
    return cached;
}
constexpr int64_t kTopkTileMinRows = 16384;

// out[e, :] = the n bits of packed row index[e] (index == NULL: row e) as fp32 {0,1}: the
// boolean-mask indexing `bit_vector_z.view(-1, n)[mask]` of structmarg.py:116-126 from packed bits,
// so that the dense [B,k,n] fp32 tensor is never materialised.  A thread per float4 of the output
// (four in flight), consecutive threads on consecutive float4s.
__global__ void __launch_bounds__(256)
bits_expand_vec_kernel(const uint32_t* __restrict__ bits, const int64_t* __restrict__ index, float* __restrict__ out,
                       int64_t n_out, int n)
{
    // a warp per output row, four rows in flight; lane l writes the float4s l, l + 32, ... of the row (no 64-bit
    // division per 16 bytes as in the first version: 324 -> see DESIGN.md 4.4); {0,1} as bit * 0x3F800000
    const int W = (n + 31) >> 5, q4 = n >> 2;  // float4s per row (n % 4 == 0)
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e0 = warp * 4; e0 < n_out; e0 += nwarps * 4) {
        int64_t src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) src[u] = (e0 + u < n_out) ? (index != nullptr ? __ldg(index + e0 + u) : e0 + u) : -1;
        for (int q = lane; q < q4; q += 32) {
            uint32_t word[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) word[u] = src[u] >= 0 ? __ldg(bits + src[u] * W + (q >> 3)) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (src[u] >= 0) {
                    const uint32_t nib = word[u] >> ((q & 7) * 4);
                    stg_stream4(out + (e0 + u) * n + q * 4,
                                make_float4(__uint_as_float((nib & 1u) * 0x3F800000u), __uint_as_float(((nib >> 1) & 1u) * 0x3F800000u),
                                            __uint_as_float(((nib >> 2) & 1u) * 0x3F800000u),
                                            __uint_as_float(((nib >> 3) & 1u) * 0x3F800000u)));
                }
            }
        }
    }
}
// any n / alignment: one warp per output row
__global__ void __launch_bounds__(256)
bits_expand_kernel(const uint32_t* __restrict__ bits, const int64_t* __restrict__ index, float* __restrict__ out,
                   int64_t n_out, int n)
{
    const int W = (n + 31) >> 5, lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n_out; e += warps_total) {
        const int64_t src = index != nullptr ? __ldg(index + e) : e;
        const uint32_t* bw = bits + src * W;
        float* o = out + e * n;
        for (int c = lane; c < n; c += 32) stg_stream(o + c, (float)((__ldg(bw + (c >> 5)) >> (c & 31)) & 1u));
    }
}

// out[e] = src[idx[e]] for rows of `row_chunks` chunks of sizeof(T) bytes: `encoder_input[idxs]`,
// `x.unsqueeze(1).repeat(1,k,1).view(B*k,-1)[mask]` (structmarg.py:95-124, 236-240).  A thread per
// chunk, four in flight, consecutive threads on consecutive chunks of the output.
template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_small_kernel(const T* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ out, int64_t n_out,
                         int64_t row_chunks)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_out; e += stride) {
        const int64_t from = (int64_t)__ldg(idx + e) * row_chunks;
        for (int64_t c = 0; c < row_chunks; ++c) out[e * row_chunks + c] = __ldg(src + from + c);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const T* __restrict__ src, const int32_t* __restrict__ idx, T* __restrict__ out, int64_t n_out,
                   int64_t row_chunks)
{
    // a warp per output row, four rows in flight; lane l copies the chunks l, l + 32, ... of the row
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e0 = warp * 4; e0 < n_out; e0 += nwarps * 4) {
        int64_t from[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) from[u] = (e0 + u < n_out) ? (int64_t)__ldg(idx + e0 + u) * row_chunks : -1;
        for (int64_t c = lane; c < row_chunks; c += 32) {
            T v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (from[u] >= 0) v[u] = __ldg(src + from[u] + c);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (from[u] >= 0) out[(e0 + u) * row_chunks + c] = v[u];
        }
    }
}

}  // namespace smlvm

using namespace smlvm;

extern "C" int smlvm_topk_bits(const float* x, float* configs, uint32_t* bits, float* dp_scores,
                               int64_t rows, int32_t n, int32_t k, void* stream)
{
    if (rows < 0 || n < 1 || k < 2) return SMLVM_ERR_ARG;  // k > 1: binary_topk.h:102
    if (n < 31 && (1 << n) < k) return SMLVM_ERR_ARG;      // fewer than k bit-vectors exist
    if (n > SMLVM_TOPK_MAX_BITS || k > SMLVM_TOPK_MAX_K) return SMLVM_ERR_UNSUPPORTED;
    if (rows == 0) return SMLVM_OK;
    if (x == nullptr) return SMLVM_ERR_ARG;
    {
        const int ovr = topk_path_override();
        // block-local configurations pay when a configuration is wide (3-4 words) and the lists are long
        const bool blk_wins = n > 64 && k >= 8;
        if (topk_blk_supported(n, k) && ovr != 1 && ovr != 2 && ((rows >= kTopkTileMinRows && blk_wins) || ovr == 3))
            return topk_blk_launch(x, configs, bits, dp_scores, rows, n, k, as_stream(stream));
        if (topk_tile_supported(n, k) && ovr != 1 && (rows >= kTopkTileMinRows || ovr == 2))
            return topk_tile_launch(x, configs, bits, dp_scores, rows, n, k, as_stream(stream));
    }
    const int G = k <= 8 ? 8 : (k <= 16 ? 16 : 32);
    const int words_n = (n + 31) >> 5;
    int W = 1;
    while (W < words_n) W <<= 1;
    const int n_pad = (n + 3) & ~3;
    const int groups_per_cta = (kThreads / kWarp) * (kWarp / G);
    const size_t smem = (size_t)groups_per_cta * n_pad * sizeof(float);
    int64_t need = (rows + groups_per_cta - 1) / groups_per_cta;
    int64_t cap = (int64_t)device_sm_count() * 8;
    const int grid = (int)(need < cap ? need : cap);
    cudaStream_t st = as_stream(stream);
    if (G == 8) launch_topk_w<8>(W, grid, smem, st, x, configs, bits, dp_scores, rows, n, k, n_pad);
    else if (G == 16) launch_topk_w<16>(W, grid, smem, st, x, configs, bits, dp_scores, rows, n, k, n_pad);
    else launch_topk_w<32>(W, grid, smem, st, x, configs, bits, dp_scores, rows, n, k, n_pad);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_bits_score_fwd(const uint32_t* bits, const float* x, float* scores_k,
                                    int64_t rows, int32_t n, int32_t k, void* stream)
{
    if (rows < 0 || n < 1 || k < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (bits == nullptr || x == nullptr || scores_k == nullptr) return SMLVM_ERR_ARG;
    int64_t need = (rows * k + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)device_sm_count() * 16;
    static int fwd_env = -1;  // SMLVM_BITS_FWD=vec keeps the per-configuration streaming kernel (A/B runs)
    if (fwd_env < 0) {
        const char* e = getenv("SMLVM_BITS_FWD");
        fwd_env = (e && strcmp(e, "vec") == 0) ? 1 : 0;
    }
    if ((n & 3) == 0 && n <= 128 && fwd_env == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
        constexpr int kWarps = kThreads / kWarp;
        const int64_t need_d = (rows + kWarps * kScoreRows - 1) / (kWarps * kScoreRows), cap_d = (int64_t)device_sm_count() * 8;
        const size_t smem = (size_t)kWarps * kScoreRows * ((n + kScorePad) * sizeof(float) + sizeof(double));
        const uint32_t k_magic = (uint32_t)((0x100000000ull + (uint64_t)k - 1) / (uint64_t)k);
        bits_score_fwd_delta_kernel<<<(int)(need_d < cap_d ? need_d : cap_d), kThreads, smem, as_stream(stream)>>>(
            bits, x, scores_k, rows, n, k, k_magic);
    } else if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0)
        bits_score_fwd_vec_kernel<<<(int)(need < cap ? need : cap), kThreads, 0, as_stream(stream)>>>(
            bits, x, scores_k, rows, n, k);
    else
        bits_score_fwd_kernel<<<(int)(need < cap ? need : cap), kThreads, 0, as_stream(stream)>>>(
            bits, x, scores_k, rows, n, k);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_bits_score_bwd(const uint32_t* bits, const float* dscores_k, float* dx,
                                    int64_t rows, int32_t n, int32_t k, void* stream)
{
    if (rows < 0 || n < 1 || k < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (bits == nullptr || dscores_k == nullptr || dx == nullptr) return SMLVM_ERR_ARG;
    int64_t need = (rows + 7) / 8;
    int64_t cap = (int64_t)device_sm_count() * 8;
    const size_t smem = (size_t)(kThreads / kWarp) * (((k + 3) & ~3) + ((k * ((n + 31) / 32) + 3) & ~3)) * sizeof(float);
    if ((n & 7) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15u) == 0) {
        static int cpt_env = -1;
        if (cpt_env < 0) {
            const char* e = getenv("SMLVM_BITS_BWD_CPT");
            cpt_env = e ? atoi(e) : 0;
        }
        int cpt = (n & 31) == 0 ? 32 : ((n & 15) == 0 ? 16 : 8);
        if (cpt_env == 8 || (cpt_env == 16 && (n & 15) == 0) || (cpt_env == 32 && (n & 31) == 0)) cpt = cpt_env;
        const int64_t need_v = (rows * (n / cpt) + 255) / 256, cap_v = (int64_t)device_sm_count() * 16;
        const int grid_v = (int)(need_v < cap_v ? need_v : cap_v);
        cudaStream_t st = as_stream(stream);
        if (cpt == 32) bits_score_bwd_vec_kernel<32><<<grid_v, 256, 0, st>>>(bits, dscores_k, dx, rows, n, k);
        else if (cpt == 16) bits_score_bwd_vec_kernel<16><<<grid_v, 256, 0, st>>>(bits, dscores_k, dx, rows, n, k);
        else bits_score_bwd_vec_kernel<8><<<grid_v, 256, 0, st>>>(bits, dscores_k, dx, rows, n, k);
        SMLVM_LAUNCH_CHECK();
        return SMLVM_OK;
    }
    bits_score_bwd_kernel<<<(int)(need < cap ? need : cap), kThreads, smem, as_stream(stream)>>>(
        bits, dscores_k, dx, rows, n, k);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_bits_expand(const uint32_t* bits, const int64_t* index, float* out, int64_t n_out,
                                 int32_t n, void* stream)
{
    if (n_out < 0 || n < 1) return SMLVM_ERR_ARG;
    if (n_out == 0) return SMLVM_OK;
    if (bits == nullptr || out == nullptr) return SMLVM_ERR_ARG;
    const int64_t cap = (int64_t)device_sm_count() * 16;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
        const int64_t need = (n_out + 31) / 32;   // 8 warps x 4 rows per CTA and pass
        bits_expand_vec_kernel<<<(int)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(bits, index, out, n_out, n);
    } else {
        const int64_t need = (n_out + 7) / 8;
        bits_expand_kernel<<<(int)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(bits, index, out, n_out, n);
    }
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_gather_rows(const void* src, const int32_t* idx, void* out, int64_t n_out, int64_t row_bytes,
                                 void* stream)
{
    if (n_out < 0 || row_bytes < 1) return SMLVM_ERR_ARG;
    if (n_out == 0) return SMLVM_OK;
    if (src == nullptr || idx == nullptr || out == nullptr) return SMLVM_ERR_ARG;
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | (uintptr_t)row_bytes;
    const int chunk = (al & 15u) == 0 ? 16 : ((al & 7u) == 0 ? 8 : ((al & 3u) == 0 ? 4 : 0));
    if (chunk == 0) return SMLVM_ERR_UNSUPPORTED;
    const int64_t row_chunks = row_bytes / chunk;
    const int64_t cap = (int64_t)device_sm_count() * 16;
    // rows of at least 8 chunks: a warp per row; shorter rows (labels, scalars): a thread per row
    const bool small = row_chunks < 8;
    const int64_t need = small ? (n_out + 255) / 256 : (n_out + 31) / 32;
    const int grid = (int)(need < cap ? need : cap);
    cudaStream_t st = as_stream(stream);
    if (small) {
        if (chunk == 16) gather_rows_small_kernel<uint4><<<grid, 256, 0, st>>>((const uint4*)src, idx, (uint4*)out, n_out, row_chunks);
        else if (chunk == 8) gather_rows_small_kernel<uint2><<<grid, 256, 0, st>>>((const uint2*)src, idx, (uint2*)out, n_out, row_chunks);
        else gather_rows_small_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)src, idx, (uint32_t*)out, n_out, row_chunks);
        SMLVM_LAUNCH_CHECK();
        return SMLVM_OK;
    }
    if (chunk == 16)
        gather_rows_kernel<uint4><<<grid, 256, 0, st>>>((const uint4*)src, idx, (uint4*)out, n_out, row_chunks);
    else if (chunk == 8)
        gather_rows_kernel<uint2><<<grid, 256, 0, st>>>((const uint2*)src, idx, (uint2*)out, n_out, row_chunks);
    else
        gather_rows_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)src, idx, (uint32_t*)out, n_out, row_chunks);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}
