// Probability-weighted loss for sm_100a: fused (segmented) reduce.
//
// replaces
//   dense : encoder_probs.unsqueeze(1).bmm(losses.unsqueeze(-1)) + .mean()  lvmhelpers/marg.py:125-126
//   ragged: (p @ loss_components) / batch_size                              lvmhelpers/structmarg.py:140,242
//           -p . log p / batch_size  (structured entropy)                   lvmhelpers/structmarg.py:51-58,200-202
//
// Totals are accumulated in fp64 and finished by the last CTA in a fixed order, so the
// scalar is run-to-run deterministic.  Streaming kernels: 8 B per (p, loss) pair forward,
// 8 B read + 8 B written backward.
#include <math_constants.h>

#include "common.cuh"
#include "reduce_ws.cuh"

namespace smlvm {

constexpr int kThreads = 256;

template <int G>
__global__ void __launch_bounds__(kThreads)
wloss_dense_kernel(const float* __restrict__ p, const float* __restrict__ loss,
                   float* __restrict__ row_loss, float* __restrict__ total, float scale,
                   int64_t rows, int K, bool vec, ReduceWs ws)
{
    __shared__ double smem[kThreads / 32 + 1];
    constexpr int RPW = kWarp / G;
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kThreads / kWarp) + (threadIdx.x >> 5);
    double acc = 0.0;
    for (; w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int64_t off = (live ? row : 0) * (int64_t)K;
        float s = 0.f;
        if (live) {
            if (vec) {
                for (int c = 4 * l; c < K; c += 4 * G) {
                    float4 a = ldg_stream4(p + off + c), b = ldg_stream4(loss + off + c);
                    s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s);
                    s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
                }
            } else {
                for (int c = l; c < K; c += G) s = fmaf(ldg_stream(p + off + c), ldg_stream(loss + off + c), s);
            }
        }
        s = group_sum<G>(s);
        if (live && l == 0) {
            if (row_loss != nullptr) row_loss[row] = s;
            acc += (double)s;
        }
    }
    if (total != nullptr) finish_totals(ws, acc, 0.0, scale, total, nullptr, smem);
}

__global__ void __launch_bounds__(kThreads)
wloss_ragged_kernel(const float* __restrict__ p, const float* __restrict__ loss,
                    const int64_t* __restrict__ offsets, float* __restrict__ row_loss,
                    float* __restrict__ total, float* __restrict__ neg_plogp, float scale,
                    int64_t rows, int64_t n_items, ReduceWs ws)
{
    __shared__ double smem[kThreads / 32 + 1];
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kThreads;
    double acc = 0.0, ent = 0.0;
    if (total != nullptr || neg_plogp != nullptr) {
        for (int64_t i = tid; i < n_items; i += nthreads) {
            const float pi = __ldg(p + i);
            if (loss != nullptr) acc += (double)(pi * __ldg(loss + i));
            if (neg_plogp != nullptr && pi > 0.f) ent -= (double)(pi * logf(pi));   // (p = 0: padding entries)
        }
    }
    if (row_loss != nullptr) {
        // one 8-lane group per segment (supports are short: tens of items)
        constexpr int G = 8;
        constexpr int RPW = kWarp / G;
        const int l = threadIdx.x % G;
        const int g = (threadIdx.x & 31) / G;
        // warp-uniform outer loop: the group reduction needs all 32 lanes present
        for (int64_t r0 = (tid / kWarp) * RPW; r0 < rows; r0 += (nthreads / kWarp) * RPW) {
            const int64_t r = r0 + g;
            const bool live = r < rows;
            const int64_t b = live ? __ldg(offsets + r) : 0, e = live ? __ldg(offsets + r + 1) : 0;
            float s = 0.f;
            for (int64_t i = b + l; i < e; i += G) s = fmaf(__ldg(p + i), __ldg(loss + i), s);
            s = group_sum<G>(s);
            if (live && l == 0) row_loss[r] = s;
        }
    }
    if (total != nullptr || neg_plogp != nullptr)
        finish_totals(ws, acc, ent, scale, total, neg_plogp, smem);
}

__global__ void __launch_bounds__(kThreads)
wloss_ragged_bwd_kernel(const float* __restrict__ p, const float* __restrict__ loss,
                        const float* __restrict__ g, const float* __restrict__ gent, float scale,
                        float* __restrict__ dp, float* __restrict__ dloss, int64_t n_items)
{
    const float gs = (g != nullptr ? __ldg(g) : 0.f) * scale;
    const float ge = (gent != nullptr ? __ldg(gent) : 0.f) * scale;
    const int64_t nthreads = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_items; i += nthreads) {
        const float pi = __ldg(p + i);
        if (dp != nullptr) {
            float d = (loss != nullptr) ? gs * __ldg(loss + i) : 0.f;
            if (gent != nullptr && pi > 0.f) d -= ge * (logf(pi) + 1.f);
            dp[i] = d;
        }
        if (dloss != nullptr) dloss[i] = gs * pi;
    }
}

static int grid_1d(int64_t work_items, int per_cta)
{
    int64_t need = (work_items + per_cta - 1) / per_cta;
    int64_t cap = (int64_t)device_sm_count() * 8;
    if (cap > kMaxCtas) cap = kMaxCtas;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// Dense weighted loss driven by the forward's support slots: row_loss[r] = sum_j val_j * L[r, col_j]
// -- 64 B of slots + one sector of L per non-zero instead of the 8K-byte rows of P and L.  A thread
// owns a row; rows marked "scan" (support does not fit the slots) are dotted densely by the warp.
__global__ void __launch_bounds__(kThreads)
wloss_slots_kernel(const float* __restrict__ p, const float* __restrict__ loss, const float* __restrict__ slots,
                   float* __restrict__ loss_slots, float* __restrict__ row_loss, float* __restrict__ total,
                   float scale, int64_t rows, int K, ReduceWs ws)
{
    constexpr int NS = SMLVM_SUPPORT_SLOTS;
    __shared__ double smem[kThreads / 32 + 1];
    const int lane = threadIdx.x & 31;
    const int64_t ngroups = (rows + 31) / 32;
    const int64_t stride = (int64_t)gridDim.x * (kThreads / kWarp);
    double acc_total = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * (kThreads / kWarp) + (threadIdx.x >> 5); t < ngroups; t += stride) {
        const int64_t r = t * 32 + lane;
        const bool live = r < rows;
        float v[NS];
        int c[NS];
        const float4* src = reinterpret_cast<const float4*>(slots + (live ? r : 0) * (2 * NS));
#pragma unroll
        for (int j = 0; j < NS; j += 2) {
            const float4 q = __ldg(src + j / 2);
            v[j] = q.x; c[j] = __float_as_int(q.y); v[j + 1] = q.z; c[j + 1] = __float_as_int(q.w);
        }
        const bool scan = live && c[0] == -2;
        float lv[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) lv[j] = (live && !scan && c[j] >= 0) ? ldg_gather(loss + r * K + c[j]) : 0.f;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NS; ++j) s = fmaf(v[j], lv[j], s);  // empty slots carry v = 0
        if (scan) s = 0.f;
        if (loss_slots != nullptr && live) {  // the gathered losses, for the backward to stream
            float4* dst = reinterpret_cast<float4*>(loss_slots + r * NS);
            dst[0] = make_float4(lv[0], lv[1], lv[2], lv[3]);
            dst[1] = make_float4(lv[4], lv[5], lv[6], lv[7]);
        }
        unsigned todo = __ballot_sync(kFull, scan);
        while (todo) {
            const int src_lane = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t rr = t * 32 + src_lane;
            float part = 0.f;
            for (int col = lane; col < K; col += 32) part = fmaf(__ldg(p + rr * K + col), __ldg(loss + rr * K + col), part);
            part = group_sum<32>(part);
            if (lane == src_lane) s = part;
        }
        if (live) {
            if (row_loss != nullptr) row_loss[r] = s;
            acc_total += (double)s;
        }
    }
    if (total != nullptr) finish_totals(ws, acc_total, 0.0, scale, total, nullptr, smem);
}

// out[r] = log sum_{j in segment r} exp(sign * v_j + shift); one 8-lane group per segment
// (supports are short), max-shifted, -inf for an empty segment.
__global__ void __launch_bounds__(kThreads)
segment_logsumexp_kernel(const float* __restrict__ v, const int64_t* __restrict__ offsets, float sign,
                         float shift, float* __restrict__ out, int64_t rows)
{
    constexpr int G = 8;
    constexpr int RPW = kWarp / G;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kThreads;
    const int l = threadIdx.x % G;
    const int g = (threadIdx.x & 31) / G;
    for (int64_t r0 = (tid / kWarp) * RPW; r0 < rows; r0 += (nthreads / kWarp) * RPW) {
        const int64_t r = r0 + g;
        const bool live = r < rows;
        const int64_t b = live ? __ldg(offsets + r) : 0, e = live ? __ldg(offsets + r + 1) : 0;
        float mx = -CUDART_INF_F;
        for (int64_t i = b + l; i < e; i += G) mx = fmaxf(mx, fmaf(sign, __ldg(v + i), shift));
        mx = group_max<G>(mx);
        float acc = 0.f;
        if (mx > -CUDART_INF_F)
            for (int64_t i = b + l; i < e; i += G) acc += expf(fmaf(sign, __ldg(v + i), shift) - mx);
        acc = group_sum<G>(acc);
        if (live && l == 0) out[r] = (mx > -CUDART_INF_F) ? mx + logf(acc) : -CUDART_INF_F;
    }
}

}  // namespace smlvm

using namespace smlvm;

extern "C" int smlvm_segment_logsumexp(const float* v, const int64_t* offsets, float sign, float shift,
                                       float* out, int64_t rows, void* stream)
{
    if (rows < 0) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (v == nullptr || offsets == nullptr || out == nullptr) return SMLVM_ERR_ARG;
    int64_t need = (rows + 31) / 32;
    int64_t cap = (int64_t)device_sm_count() * 16;
    segment_logsumexp_kernel<<<(int)(need < cap ? need : cap), kThreads, 0, as_stream(stream)>>>(
        v, offsets, sign, shift, out, rows);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" size_t smlvm_reduce_workspace_bytes(void) { return 64 + 3 * sizeof(double) * kMaxCtas; }

extern "C" int smlvm_wloss_dense_fwd(const float* p, const float* loss, const float* support_slots,
                                     float* loss_slots, float* row_loss, float* total, float scale,
                                     int64_t rows, int32_t cols, void* workspace, size_t workspace_bytes,
                                     void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (p == nullptr || loss == nullptr) return SMLVM_ERR_ARG;
    if (total != nullptr && (workspace == nullptr || workspace_bytes < smlvm_reduce_workspace_bytes()))
        return SMLVM_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    if (support_slots != nullptr && rows > 0) {
        const int grid = grid_1d(rows, kThreads);
        wloss_slots_kernel<<<grid, kThreads, 0, st>>>(p, loss, support_slots, loss_slots, row_loss, total, scale,
                                                      rows, cols, ws_view(workspace));
        SMLVM_LAUNCH_CHECK();
        return SMLVM_OK;
    }
    int G = 1;
    while (G < 32 && G * 4 < cols) G <<= 1;
    const bool vec = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(loss) & 15u) == 0);
    const int rows_per_cta = (kThreads / kWarp) * (kWarp / G);
    const int grid = grid_1d(rows, rows_per_cta);
    ReduceWs ws = ws_view(workspace);
    switch (G) {
        case 1: wloss_dense_kernel<1><<<grid, kThreads, 0, st>>>(p, loss, row_loss, total, scale, rows, cols, vec, ws); break;
        case 2: wloss_dense_kernel<2><<<grid, kThreads, 0, st>>>(p, loss, row_loss, total, scale, rows, cols, vec, ws); break;
        case 4: wloss_dense_kernel<4><<<grid, kThreads, 0, st>>>(p, loss, row_loss, total, scale, rows, cols, vec, ws); break;
        case 8: wloss_dense_kernel<8><<<grid, kThreads, 0, st>>>(p, loss, row_loss, total, scale, rows, cols, vec, ws); break;
        case 16: wloss_dense_kernel<16><<<grid, kThreads, 0, st>>>(p, loss, row_loss, total, scale, rows, cols, vec, ws); break;
        default: wloss_dense_kernel<32><<<grid, kThreads, 0, st>>>(p, loss, row_loss, total, scale, rows, cols, vec, ws); break;
    }
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_wloss_ragged_fwd(const float* p, const float* loss, const int64_t* offsets,
                                      float* row_loss, float* total, float* neg_plogp, float scale,
                                      int64_t rows, int64_t n_items, void* workspace,
                                      size_t workspace_bytes, void* stream)
{
    if (rows < 0 || n_items < 0 || p == nullptr) return SMLVM_ERR_ARG;
    if ((total != nullptr || row_loss != nullptr) && loss == nullptr) return SMLVM_ERR_ARG;
    if (row_loss != nullptr && offsets == nullptr) return SMLVM_ERR_ARG;
    if ((total != nullptr || neg_plogp != nullptr) &&
        (workspace == nullptr || workspace_bytes < smlvm_reduce_workspace_bytes()))
        return SMLVM_ERR_WORKSPACE;
    int64_t work = n_items > rows * 8 ? n_items : rows * 8;
    const int grid = grid_1d(work, kThreads);
    wloss_ragged_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(
        p, loss, offsets, row_loss, total, neg_plogp, scale, rows, n_items, ws_view(workspace));
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_wloss_ragged_bwd(const float* p, const float* loss, const float* g,
                                      const float* gent, float scale, float* dp, float* dloss,
                                      int64_t n_items, void* stream)
{
    if (n_items < 0 || p == nullptr) return SMLVM_ERR_ARG;
    if (n_items == 0) return SMLVM_OK;
    const int grid = grid_1d(n_items, kThreads);
    wloss_ragged_bwd_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(p, loss, g, gent, scale, dp,
                                                                      dloss, n_items);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}
