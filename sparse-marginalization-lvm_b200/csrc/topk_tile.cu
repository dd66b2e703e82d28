// k-best bit-vectors, thread-per-row kernel for sm_100a (k <= 16, n <= 128).
//
// Same results as topk_bits_kernel in topk_bits.cu (bit-exact with
// lvmhelpers/binary_topk.h:57-129: list seeded by x0 >= 0 (:110), fp32 scores accumulated in
// bit order (:27-29), "bit on" taken only if STRICTLY greater (:72), truncation to k after
// every bit) but organised for instruction count: the group-per-row kernel spends ~3200 warp
// instructions per row on merge-path binary searches and configuration shuffles and reaches 8 %
// of the HBM roofline.  Here a THREAD owns a row and runs the reference's own two-pointer list
// merge; the control flow (list length, loop bounds) is data-independent, so the 32 rows of a
// warp run in lockstep without divergence, and the lists live in shared memory as
// [slot][lane] so that every access is conflict-free whatever slot a lane points at.
//
//   shared memory per warp: STAGE bits x 32 rows of scores transposed as xs[STAGE][33] (sector-sized
//   global reads, conflict-free column reads; restaged every STAGE bits), two score lists
//   S[2][k][32] and two configuration lists C[2][k][32] of W-word vectors (ping-pong; one
//   64/128-bit access moves a configuration, the new bit is OR-ed in with one LOP3 per word).
//   The fp32 {0,1} output (4*k*n B/row, the API's format) and the packed bits are written by the
//   whole warp, one row at a time, fully coalesced.
#include "common.cuh"

namespace smlvm {

// a configuration = W 32-bit words moved as ONE shared-memory access ([slot][lane] of Cfg<W>)
// STAGE = bits of the scores staged per round (transposed, [STAGE][33] floats per warp): 32, or 8 when the
// shorter stage buys >= 25 % more resident warps (k = 10, n = 128: 14 KB per warp, 14 warps per SM instead of
// 11 -- throughput follows the resident warps almost linearly; otherwise the extra rounds cost more)

template <int W> struct Cfg;
template <> struct Cfg<1> { uint32_t w[1]; };
template <> struct __align__(8) Cfg<2> { uint32_t w[2]; };
template <> struct __align__(16) Cfg<4> { uint32_t w[4]; };

// Moves words 0 .. WQ of a configuration (the higher words are still zero in every slot while the
// bits of block WQ are being merged), OR-ing `bit` into word WQ when `on`.
template <int W, int WQ>
__device__ __forceinline__ void move_cfg(const char* src, Cfg<W>* dst, bool on, uint32_t bit)
{
    if (WQ == 0) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(src);
        if (on) w |= bit;
        dst->w[0] = w;
    } else if (WQ == 1) {
        uint2 w = *reinterpret_cast<const uint2*>(src);
        if (on) w.y |= bit;
        *reinterpret_cast<uint2*>(dst) = w;
    } else {
        Cfg<W> cw = *reinterpret_cast<const Cfg<W>*>(src);
        if (on) cw.w[WQ] |= bit;
        *dst = cw;
    }
}

// Bits ib0 .. iend-1 of one 32-bit block (word WQ of the configurations) for the row of this lane.
// While the list is still growing (len < k) the merge checks both list ends (merge_branch,
// binary_topk.h:57-98); once it is full, a + b = r < k = len for every output, so neither list can
// run out: the loop is compare / select / move, with the new bit OR-ed into word WQ only.
template <int W, int WQ>
__device__ __forceinline__ void topk_merge_block(Cfg<W>* C, float* S, const float* xs, int lane, int k, int ib0,
                                                 int iend, int shift0, int& cur, int& len)
{
    for (int ib = ib0; ib < iend; ++ib) {
        const float v = xs[ib * 33 + lane];
        const uint32_t bit = 1u << (shift0 + ib);
        const char* ScB = reinterpret_cast<const char*>(S + (cur * k) * 32 + lane);
        float* Sn = S + ((cur ^ 1) * k) * 32 + lane;
        const char* CcB = reinterpret_cast<const char*>(C + (cur * k) * 32 + lane);
        Cfg<W>* Cn = C + ((cur ^ 1) * k) * 32 + lane;
        const float s0 = *reinterpret_cast<const float*>(ScB);
        if (len == k) {
            int b = 0;
            float sa = s0, sb = __fadd_rn(s0, v);
            for (int r = 0; r < k; ++r) {
                const bool takeB = sb > sa;                    // strictly greater (binary_topk.h:72)
                const int src = takeB ? b : r - b;             // a = r - b
                Sn[r * 32] = takeB ? sb : sa;
                move_cfg<W, WQ>(CcB + src * (32 * (int)sizeof(Cfg<W>)), Cn + r * 32, takeB, bit);
                const float head = *reinterpret_cast<const float*>(ScB + (src + 1) * 128);
                if (takeB) { ++b; sb = __fadd_rn(head, v); } else { sa = head; }
            }
        } else {
            const int outlen = min(k, 2 * len);
            int a = 0, b = 0;
            float sa = s0, sb = __fadd_rn(s0, v);
            for (int r = 0; r < outlen; ++r) {
                const bool takeB = (b < len) && (a >= len || sb > sa);
                const int src = takeB ? b : a;
                Sn[r * 32] = takeB ? sb : sa;
                move_cfg<W, WQ>(CcB + src * (32 * (int)sizeof(Cfg<W>)), Cn + r * 32, takeB, bit);
                // advance the list that was consumed and fetch its next head
                const int nxt = src + 1;
                const float head = *reinterpret_cast<const float*>(ScB + min(nxt, len - 1) * 128);
                if (takeB) { b = nxt; sb = __fadd_rn(head, v); } else { a = nxt; sa = head; }
            }
            len = outlen;
        }
        cur ^= 1;
    }
}

template <int W, int STAGE>
__global__ void __launch_bounds__(512)
topk_tile_kernel(const float* __restrict__ x, float* __restrict__ configs, uint32_t* __restrict__ bits,
                 float* __restrict__ dp_scores, int64_t rows, int n, int k)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int words_n = (n + 31) >> 5;
    const size_t per_warp = ((size_t)2 * k * W * 32 + STAGE * 33 + (2 * k + 1) * 32) * 4;
    Cfg<W>* C = reinterpret_cast<Cfg<W>*>(smem_raw + warp * per_warp);        // [2][k][32]
    float* xs = reinterpret_cast<float*>(C + 2 * k * 32);                     // [STAGE][33]: STAGE bits x 32 rows
    float* S = xs + STAGE * 33;                                                // [2][k][32] (+ one slot of slack)

    const int64_t ntiles = (rows + 31) / 32;
    for (int64_t t = (int64_t)blockIdx.x * nwarps + warp; t < ntiles; t += (int64_t)gridDim.x * nwarps) {
        const int64_t row0 = t * 32;
        const int trows = (int)min((int64_t)32, rows - row0);
        int cur = 0, len = 2;
        for (int i0 = 0; i0 < n; i0 += STAGE) {
            // ---- stage bits i0..i0+STAGE-1 of the 32 score rows, transposed: lane -> (column l & 7, rows
            //      (l >> 3) + 4j), one 32-byte sector per row and load, 8 loads in flight per lane ----------
            __syncwarp();
            {
                const int col = lane & (STAGE - 1), rb = lane / STAGE;
                const int c = i0 + col;
                constexpr int kRowsPerLoad = 32 / STAGE, kLoads = 32 / kRowsPerLoad;
                float tmp[kLoads];
#pragma unroll
                for (int j = 0; j < kLoads; ++j) {
                    const int r = rb + kRowsPerLoad * j;
                    tmp[j] = (r < trows && c < n) ? __ldcs(x + (row0 + r) * n + c) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < kLoads; ++j) xs[col * 33 + rb + kRowsPerLoad * j] = tmp[j];
            }
            __syncwarp();
            int ib = 0;
            if (i0 == 0) {
                // list = [(x0,{0}), (0,{})] ordered by x0 >= 0        (binary_topk.h:104-116)
                const float x0 = xs[lane];
                const bool on_first = x0 >= 0.f;
                S[0 * 32 + lane] = on_first ? x0 : 0.f;
                S[1 * 32 + lane] = on_first ? 0.f : x0;
                Cfg<W> c0, c1;
#pragma unroll
                for (int q = 0; q < W; ++q) { c0.w[q] = 0u; c1.w[q] = 0u; }
                // every slot of both sides starts from zero: a block only moves words 0 .. its own
                for (int slot = 2; slot < 2 * k; ++slot) C[slot * 32 + lane] = c0;
                c0.w[0] = on_first ? 1u : 0u;
                c1.w[0] = on_first ? 0u : 1u;
                C[0 * 32 + lane] = c0;
                C[1 * 32 + lane] = c1;
                ib = 1;
            }
            const int iend = min(STAGE, n - i0);
            const int sh0 = i0 & 31;
            if (W == 1) {
                topk_merge_block<W, 0>(C, S, xs, lane, k, ib, iend, sh0, cur, len);
            } else if (W == 2) {
                if (i0 < 32) topk_merge_block<W, 0>(C, S, xs, lane, k, ib, iend, sh0, cur, len);
                else topk_merge_block<W, (W > 1 ? 1 : 0)>(C, S, xs, lane, k, ib, iend, sh0, cur, len);
            } else {
                switch (i0 >> 5) {
                    case 0: topk_merge_block<W, 0>(C, S, xs, lane, k, ib, iend, sh0, cur, len); break;
                    case 1: topk_merge_block<W, (W > 1 ? 1 : 0)>(C, S, xs, lane, k, ib, iend, sh0, cur, len); break;
                    case 2: topk_merge_block<W, (W > 2 ? 2 : 0)>(C, S, xs, lane, k, ib, iend, sh0, cur, len); break;
                    default: topk_merge_block<W, (W > 3 ? 3 : 0)>(C, S, xs, lane, k, ib, iend, sh0, cur, len); break;
                }
            }
        }
        __syncwarp();

        // ---- outputs, one row at a time by the whole warp (coalesced) ------------------------
        const float* Sf = S + (cur * k) * 32;
        const uint32_t* Cf = reinterpret_cast<const uint32_t*>(C + (cur * k) * 32);  // word q of slot j, row r: [(j*32 + r)*W + q]
        if (dp_scores != nullptr) {
            for (int idx = lane; idx < trows * k; idx += 32) {
                const int r = idx / k, j = idx - r * k;
                dp_scores[row0 * k + idx] = Sf[j * 32 + r];
            }
        }
        if (bits != nullptr) {
            const int per_row = k * words_n;
            for (int idx = lane; idx < trows * per_row; idx += 32) {
                const int r = idx / per_row, rem = idx - r * per_row;
                const int j = rem / words_n, q = rem - j * words_n;
                bits[row0 * per_row + idx] = Cf[(j * 32 + r) * W + q];
            }
        }
        if (configs != nullptr) {
            const int per_row = k * n;
            // {0,1} as fp32 without the conversion pipe: bit * 0x3F800000 (the first version spent 40 % of the
            // kernel's instructions here, four I2F per 128-bit store)
            const bool al16 = (reinterpret_cast<uintptr_t>(configs) & 15u) == 0;
            if (n == 128 && al16) {
                // a configuration is exactly one 512-byte line: lane l stores bits 4l .. 4l+3 of entry j
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
                    // flat float4 index over [k][n] with (j, c) carried incrementally: all 32
                    // lanes store whatever n is
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
    }
}

bool topk_tile_supported(int n, int k) { return k <= 16 && n <= 128; }

int topk_tile_launch(const float* x, float* configs, uint32_t* bits, float* dp_scores, int64_t rows, int n,
                     int k, cudaStream_t st)
{
    const int words_n = (n + 31) >> 5;
    const int W = words_n <= 1 ? 1 : (words_n <= 2 ? 2 : 4);
    auto per_warp_bytes = [&](int stage) { return ((size_t)2 * k * W * 32 + stage * 33 + (2 * k + 1) * 32) * 4; };
    // resident warps per SM: 200 KB of lists, at most 32 (two CTAs of <= 16 warps)
    auto warps_per_sm = [&](int stage) {
        const int w = (int)((200 * 1024) / per_warp_bytes(stage));
        return w > 32 ? 32 : (w < 1 ? 1 : w);
    };
    const int stage = (4 * warps_per_sm(8) >= 5 * warps_per_sm(32)) ? 8 : 32;
    const size_t per_warp = per_warp_bytes(stage);
    const int wsm = warps_per_sm(stage);
    const int ctas_per_sm = wsm > 16 ? 2 : 1;
    int warps = wsm / ctas_per_sm;
    const int64_t ntiles = (rows + 31) / 32;
    const int64_t slots = (int64_t)device_sm_count() * ctas_per_sm;
    if (ntiles < (int64_t)warps * slots) {  // small batches: spread the tiles over the SMs
        const int w2 = (int)((ntiles + slots - 1) / slots);
        warps = w2 < 1 ? 1 : (w2 < warps ? w2 : warps);
    }
    const size_t smem = per_warp * warps;
    const int64_t need = (ntiles + warps - 1) / warps;
    const int grid = (int)(need < slots ? need : slots);
#define SMLVM_TOPK_TILE_LAUNCH(WW, SS)                                                                        \
    do {                                                                                                      \
        cudaError_t e = cudaFuncSetAttribute(topk_tile_kernel<WW, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                             200 * 1024);                                                     \
        if (e != cudaSuccess) return (int)e;                                                                  \
        topk_tile_kernel<WW, SS><<<grid, warps * 32, smem, st>>>(x, configs, bits, dp_scores, rows, n, k);    \
    } while (0)
#define SMLVM_TOPK_TILE_CASE(WW)                                                                              \
    case WW:                                                                                                  \
        if (stage == 8) SMLVM_TOPK_TILE_LAUNCH(WW, 8);                                                        \
        else SMLVM_TOPK_TILE_LAUNCH(WW, 32);                                                                  \
        break;
    switch (W) {
        SMLVM_TOPK_TILE_CASE(1)
        SMLVM_TOPK_TILE_CASE(2)
        SMLVM_TOPK_TILE_CASE(4)
    }
#undef SMLVM_TOPK_TILE_CASE
#undef SMLVM_TOPK_TILE_LAUNCH
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SMLVM_OK : (int)e;
}

}  // namespace smlvm
