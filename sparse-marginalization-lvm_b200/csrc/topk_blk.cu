// k-best bit-vectors, thread-per-row kernel with BLOCK-LOCAL configurations for sm_100a (k <= 16, n <= 128).
//
// Same results as topk_tile.cu / topk_bits.cu (bit-exact with lvmhelpers/binary_topk.h:57-129: list seeded by x0 >= 0
// (:110), fp32 scores accumulated in bit order (:27-29), "bit on" taken only if STRICTLY greater (:72), truncation to
// k after every bit).  topk_tile.cu lets the 128-bit configurations ride along with the scores through every merge
// step: 10 shared-memory wavefronts per step, 8 of them configuration bytes (ncu r01: l1tex data pipe at 85 %, 48 %
// issue utilisation).  Here a list entry is 64 bits: (score, tag), tag = 4 bits of ANCESTOR id + the 28 bits the entry
// has decided since the last materialisation.  A merge step of bit i only ever keeps an entry as it is (the A side) or
// keeps it with bit i set (the B side), so inside a block of 28 bits
//     B entry = (score + x_i, tag | 1 << (i - block start))
// and nothing else moves: one 64-bit load + one 64-bit store per step (4 wavefronts), no data-dependent loop, the 32
// rows of a warp stay in lockstep.  Every 28 bits (5 times for n = 128) the k configurations are materialised once:
// pool_next[r] = pool[ancestor of entry r] | (its 28 bits << block start), and entry r becomes its own ancestor r --
// so after the last block the pool holds the configurations in list order.  (The other way to stop moving
// configurations -- a slot pool with per-bit copies of the entries that survive on both sides -- was measured in round
// 2 and dropped: its copy loop runs at the worst row of the warp)
#include "common.cuh"

namespace smlvm {

template <int W> struct CfgB;
template <> struct CfgB<1> { uint32_t w[1]; };
template <> struct __align__(8) CfgB<2> { uint32_t w[2]; };
template <> struct __align__(16) CfgB<4> { uint32_t w[4]; };

constexpr int kBlkBits = 28;
constexpr uint32_t kBlkMask = (1u << kBlkBits) - 1u;

__device__ __forceinline__ uint32_t blk_sh_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint2 blk_lds64(uint32_t a)
{
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ void blk_sts64(uint32_t a, uint32_t x, uint32_t y)
{
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

// One bit: merge of the list with itself + v.  SL = [2][k][32] (+ one slot of slack) of (score, tag).
// KT = k at compile time (0: run-time k): the merge of a full list is unrolled -- list positions are immediates, the two
// read pointers are shared-memory addresses advanced under the predicate.
template <int KT>
__device__ __forceinline__ void blk_merge_bit(uint2* SL, float v, uint32_t bit, int lane, int k, int& cur, int& len)
{
    const uint2* Sc = SL + (cur * k) * 32 + lane;
    uint2* Sn = SL + ((cur ^ 1) * k) * 32 + lane;
    const int L = len;                       // (data-independent: the same in every row)
    const int outlen = min(k, 2 * L);
    uint2 h = Sc[0];
    float sa = __uint_as_float(h.x), sb = __fadd_rn(sa, v);
    uint32_t ta = h.y, tb = h.y | bit;
    if (KT > 0 && L == KT) {
        // full list: a + b = r < k, neither side can run out (the read behind the last step lands in the slack)
        // Both possible next heads are loaded at the top of the step, before the compare that picks one of them: the
        // shared-memory latency leaves the chain compare -> select -> add -> compare (ncu: 29 % of the stall samples
        // sat on the short scoreboard of the one dependent load).
        uint32_t pa = blk_sh_addr(Sc) + 256, pb = pa;       // addresses of the NEXT entry of each side
        const uint32_t sn = blk_sh_addr(Sn);
#pragma unroll
        for (int r = 0; r < KT; ++r) {
            const uint2 na = blk_lds64(pa), nb = blk_lds64(pb);
            const bool takeB = sb > sa;                                  // strictly greater (binary_topk.h:72)
            blk_sts64(sn + r * 256, takeB ? __float_as_uint(sb) : __float_as_uint(sa), takeB ? tb : ta);
            if (takeB) pb += 256; else pa += 256;
            const float t = __fadd_rn(__uint_as_float(nb.x), v);
            sb = takeB ? t : sb;
            tb = takeB ? (nb.y | bit) : tb;
            sa = takeB ? sa : __uint_as_float(na.x);
            ta = takeB ? ta : na.y;
        }
    } else if (L == k) {
        int a = 0, b = 0;
        for (int r = 0; r < k; ++r) {
            const bool takeB = sb > sa;
            Sn[r * 32] = takeB ? make_uint2(__float_as_uint(sb), tb) : make_uint2(__float_as_uint(sa), ta);
            if (takeB) ++b; else ++a;
            h = Sc[(takeB ? b : a) * 32];
            if (takeB) { sb = __fadd_rn(__uint_as_float(h.x), v); tb = h.y | bit; }
            else { sa = __uint_as_float(h.x); ta = h.y; }
        }
    } else {
        int a = 0, b = 0;
        for (int r = 0; r < outlen; ++r) {
            const bool takeB = (b < L) && (a >= L || sb > sa);          // merge_branch, binary_topk.h:57-98
            Sn[r * 32] = takeB ? make_uint2(__float_as_uint(sb), tb) : make_uint2(__float_as_uint(sa), ta);
            if (takeB) ++b; else ++a;
            h = Sc[min(takeB ? b : a, L - 1) * 32];
            if (takeB) { sb = __fadd_rn(__uint_as_float(h.x), v); tb = h.y | bit; }
            else { sa = __uint_as_float(h.x); ta = h.y; }
        }
    }
    len = outlen;
    cur ^= 1;
}

// End of a block of bits [start, start + 28): pool[r] = pool[ancestor(r)] | bits(r) << start; entry r -> ancestor r.
// In place: the (at most 16) new configurations are gathered into registers first (several entries may share an
// ancestor, and ancestors are overwritten), then stored in list order -- one pool per row, 17 instead of 12 resident
// warps per SM at k = 10, n = 128.
template <int W>
__device__ __forceinline__ void blk_materialise(uint2* SLcur, CfgB<W>* pool, int lane, int len, int start)
{
    constexpr int kMaxK = 16;
    const int q0 = start >> 5, off = start & 31;
    CfgB<W> c[kMaxK];
#pragma unroll
    for (int r = 0; r < kMaxK; ++r) {
        if (r < len) {
            const uint32_t tag = SLcur[r * 32 + lane].y;
            c[r] = pool[(tag >> kBlkBits) * 32 + lane];
            const uint32_t bits = tag & kBlkMask;
            const uint32_t lo = bits << off, hi = off > 32 - kBlkBits ? bits >> (32 - off) : 0u;
#pragma unroll
            for (int q = 0; q < W; ++q) {
                if (q == q0) c[r].w[q] |= lo;
                if (q == q0 + 1) c[r].w[q] |= hi;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kMaxK; ++r) {
        if (r < len) {
            pool[r * 32 + lane] = c[r];
            SLcur[r * 32 + lane].y = (uint32_t)r << kBlkBits;
        }
    }
}

template <int W, int STAGE, int KT>
__global__ void __launch_bounds__(640, 1)
topk_blk_kernel(const float* __restrict__ x, float* __restrict__ configs, uint32_t* __restrict__ bits,
                float* __restrict__ dp_scores, int64_t rows, int n, int k)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int words_n = (n + 31) >> 5;
    const size_t per_warp = ((size_t)k * W * 32 + (2 * k + 1) * 64 + STAGE * 33) * 4;
    CfgB<W>* P0 = reinterpret_cast<CfgB<W>*>(smem_raw + warp * per_warp);     // [k][32] pool of ancestors
    uint2* SL = reinterpret_cast<uint2*>(P0 + k * 32);                        // [2][k][32] (+ one slot of slack)
    float* xs = reinterpret_cast<float*>(SL + (2 * k + 1) * 32);              // [STAGE][33]

    const int64_t ntiles = (rows + 31) / 32;
    for (int64_t t = (int64_t)blockIdx.x * nwarps + warp; t < ntiles; t += (int64_t)gridDim.x * nwarps) {
        const int64_t row0 = t * 32;
        const int trows = (int)min((int64_t)32, rows - row0);
        int cur = 0, len = 2, start = 0;   // start: first bit of the current block
        // the scores of the NEXT stage are loaded into registers before the merges of the current one (the global-load
        // latency of a stage sat on the long scoreboard: 17 % of the stall samples), and parked in xs afterwards
        constexpr int kRowsPerLoad = 32 / STAGE, kLoads = 32 / kRowsPerLoad;
        const int scol = lane & (STAGE - 1), srb = lane / STAGE;
        float nxt[kLoads];
        auto fetch = [&](int i0f) {
            const int c = i0f + scol;
#pragma unroll
            for (int j = 0; j < kLoads; ++j) {
                const int r = srb + kRowsPerLoad * j;
                nxt[j] = (r < trows && c < n) ? __ldcs(x + (row0 + r) * n + c) : 0.f;
            }
        };
        fetch(0);
        for (int i0 = 0; i0 < n; i0 += STAGE) {
            // ---- stage bits i0 .. i0+STAGE-1 of the 32 score rows, transposed (as topk_tile.cu) ----------
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kLoads; ++j) xs[scol * 33 + srb + kRowsPerLoad * j] = nxt[j];
            if (i0 + STAGE < n) fetch(i0 + STAGE);
            __syncwarp();
            int ib = 0;
            if (i0 == 0) {
                // list = [(x0,{0}), (0,{})] ordered by x0 >= 0        (binary_topk.h:104-116): both entries descend
                // from the empty configuration in slot 0, bit 0 is the first bit of block 0
                const float x0 = xs[lane];
                const bool on_first = x0 >= 0.f;
                SL[0 * 32 + lane] = make_uint2(__float_as_uint(on_first ? x0 : 0.f), on_first ? 1u : 0u);
                SL[1 * 32 + lane] = make_uint2(__float_as_uint(on_first ? 0.f : x0), on_first ? 0u : 1u);
                CfgB<W> z;
#pragma unroll
                for (int q = 0; q < W; ++q) z.w[q] = 0u;
                P0[lane] = z;
                ib = 1;
            }
            const int iend = min(STAGE, n - i0);
            for (; ib < iend; ++ib) {
                const int i = i0 + ib;
                if (i - start == kBlkBits) {    // the block is full: materialise before bit i opens the next one
                    blk_materialise<W>(SL + (cur * k) * 32, P0, lane, len, start);
                    start += kBlkBits;
                }
                blk_merge_bit<KT>(SL, xs[ib * 33 + lane], 1u << (i - start), lane, k, cur, len);
            }
        }
        blk_materialise<W>(SL + (cur * k) * 32, P0, lane, len, start);
        __syncwarp();

        // ---- outputs, one row at a time by the whole warp (coalesced): the pool is in list order ------------
        const uint2* Sf = SL + (cur * k) * 32;
        const uint32_t* Cf = reinterpret_cast<const uint32_t*>(P0);  // word q of entry j, row r: [(j*32 + r)*W + q]
        if (dp_scores != nullptr) {
            for (int idx = lane; idx < trows * k; idx += 32) {
                const int r = idx / k, j = idx - r * k;
                dp_scores[row0 * k + idx] = __uint_as_float(Sf[j * 32 + r].x);
            }
        }
        if (bits != nullptr) {
            const int per_row = k * words_n;
            if (words_n == W) {   // (power of two: no run-time divisions -- they were 8 % of the kernel's instructions)
                for (int r = 0; r < trows; ++r)
                    for (int e = lane; e < per_row; e += 32)
                        bits[(row0 + r) * per_row + e] = Cf[((e / W) * 32 + r) * W + (e % W)];
            } else {
                for (int idx = lane; idx < trows * per_row; idx += 32) {
                    const int r = idx / per_row, rem = idx - r * per_row;
                    const int j = rem / words_n, q = rem - j * words_n;
                    bits[row0 * per_row + idx] = Cf[(j * 32 + r) * W + q];
                }
            }
        }
        if (configs != nullptr) {
            const int per_row = k * n;
            const bool al16 = (reinterpret_cast<uintptr_t>(configs) & 15u) == 0;
            if (n == 128 && al16) {
                const int sh = (lane & 7) * 4;
                for (int r = 0; r < trows; ++r) {
                    float* out = configs + (row0 + r) * per_row + lane * 4;
                    for (int j = 0; j < k; ++j) {
                        const uint32_t nib = Cf[(j * 32 + r) * W + (lane >> 3)] >> sh;
                        stg_stream4(out + j * 128, make_float4(__uint_as_float((nib & 1u) * 0x3F800000u),
                                                               __uint_as_float(((nib >> 1) & 1u) * 0x3F800000u),
                                                               __uint_as_float(((nib >> 2) & 1u) * 0x3F800000u),
                                                               __uint_as_float(((nib >> 3) & 1u) * 0x3F800000u)));
                    }
                }
            } else if ((n & 3) == 0 && al16) {
                for (int r = 0; r < trows; ++r) {
                    float* out = configs + (row0 + r) * per_row;
                    int j = 0, c = lane * 4;
                    while (c >= n) { c -= n; ++j; }
                    for (int e = lane * 4; e < per_row; e += 128) {
                        const uint32_t nib = Cf[(j * 32 + r) * W + (c >> 5)] >> (c & 31);
                        stg_stream4(out + e, make_float4(__uint_as_float((nib & 1u) * 0x3F800000u),
                                                         __uint_as_float(((nib >> 1) & 1u) * 0x3F800000u),
                                                         __uint_as_float(((nib >> 2) & 1u) * 0x3F800000u),
                                                         __uint_as_float(((nib >> 3) & 1u) * 0x3F800000u)));
                        c += 128;
                        while (c >= n) { c -= n; ++j; }
                    }
                }
            } else {
                for (int r = 0; r < trows; ++r) {
                    float* out = configs + (row0 + r) * per_row;
                    for (int j = 0; j < k; ++j)
                        for (int c = lane; c < n; c += 32)
                            stg_stream(out + j * n + c, (float)((Cf[(j * 32 + r) * W + (c >> 5)] >> (c & 31)) & 1u));
                }
            }
        }
        __syncwarp();
    }
}

bool topk_blk_supported(int n, int k) { return k <= 16 && n <= 128; }

int topk_blk_launch(const float* x, float* configs, uint32_t* bits, float* dp_scores, int64_t rows, int n,
                    int k, cudaStream_t st)
{
    const int words_n = (n + 31) >> 5;
    const int W = words_n <= 1 ? 1 : (words_n <= 2 ? 2 : 4);
    auto per_warp_bytes = [&](int stage) { return ((size_t)k * W * 32 + (2 * k + 1) * 64 + stage * 33) * 4; };
    // one CTA per SM with as many warps as 224 KB of lists hold, at most 20 (640 threads x <= 102 registers): the kernel
    // is bound by the latency of its merge chain, and the number of resident warps also sets how many rounds of tiles a
    // batch needs (262144 rows = 8192 tiles: 3 rounds with 19 warps per SM, 4 with 16)
    auto warps_per_sm = [&](int stage) {
        const int w = (int)((224 * 1024) / per_warp_bytes(stage));
        return w > 20 ? 20 : (w < 1 ? 1 : w);
    };
    const int stage = 8;   // 8 staged bits: more resident warps
    const size_t per_warp = per_warp_bytes(stage);
    int warps = warps_per_sm(stage);
    const int64_t ntiles = (rows + 31) / 32;
    const int64_t slots = (int64_t)device_sm_count();
    if (ntiles < (int64_t)warps * slots) {  // small batches: spread the tiles over the SMs
        const int w2 = (int)((ntiles + slots - 1) / slots);
        warps = w2 < 1 ? 1 : (w2 < warps ? w2 : warps);
    }
    const size_t smem = per_warp * warps;
    const int64_t need = (ntiles + warps - 1) / warps;
    const int grid = (int)(need < slots ? need : slots);
#define SMLVM_TOPK_BLK_LAUNCH(WW, SS, KK)                                                                     \
    do {                                                                                                      \
        cudaError_t e = cudaFuncSetAttribute(topk_blk_kernel<WW, SS, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             224 * 1024);                                                     \
        if (e != cudaSuccess) return (int)e;                                                                  \
        topk_blk_kernel<WW, SS, KK><<<grid, warps * 32, smem, st>>>(x, configs, bits, dp_scores, rows, n, k); \
    } while (0)
    // the experiments' k = 10 (scripts/bit_vector/*.sh) and the largest k get the unrolled merge
#define SMLVM_TOPK_BLK_K(WW, SS)                                                                              \
    do {                                                                                                      \
        if (k == 10) SMLVM_TOPK_BLK_LAUNCH(WW, SS, 10);                                                       \
        else if (k == 16) SMLVM_TOPK_BLK_LAUNCH(WW, SS, 16);                                                  \
        else SMLVM_TOPK_BLK_LAUNCH(WW, SS, 0);                                                                \
    } while (0)
#define SMLVM_TOPK_BLK_CASE(WW)                                                                               \
    case WW:                                                                                                  \
        SMLVM_TOPK_BLK_K(WW, 8);                                                                              \
        break;
    switch (W) {
        SMLVM_TOPK_BLK_CASE(1)
        SMLVM_TOPK_BLK_CASE(2)
        SMLVM_TOPK_BLK_CASE(4)
    }
#undef SMLVM_TOPK_BLK_CASE
#undef SMLVM_TOPK_BLK_K
#undef SMLVM_TOPK_BLK_LAUNCH
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

}  // namespace smlvm
