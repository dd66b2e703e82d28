// 1.5-entmax forward / backward and the Fenchel-Young losses of the SSVAE experiment, sm_100a.
//
// replaces: entmax.entmax15 (the reference's default normalizer argument, lvmhelpers/marg.py:39,46) and
//           entmax.SparsemaxLoss / Entmax15Loss (experiments/semi_supervised-vae/train.py:9,312-325) --
//           SURVEY.md 8(f) rank 4.  `entmax` is a third-party package absent from the reference tree; the
//           definitions followed here are its published ones (Peters, Niculae & Martins 2019):
//   entmax15:  X' = (X - max X) / 2,  P = [(X' - tau)_+]^2  with  sum P = 1
//              backward  g = sqrt(P),  dX = dP g - (sum dP g / sum g) g
//   FY loss:   L = (1 - sum_j p_j^alpha) / (alpha (alpha - 1)) + <p - e_y, X>,   dL/dX = p - e_y
//
// The package finds tau by sorting (tau_rho = mean_rho - sqrt((1 - ss_rho) / rho) over the rho largest entries, largest
// rho with tau_rho <= X'_(rho)).  Here a row lives in the registers of a G-lane group and tau is found without a sort:
// f(tau) = sum (X' - tau)_+^2 - 1 is convex and decreasing, so Newton steps from tau = -1 (f(-1) >= 0: the maximum is
// 0) increase monotonically and never pass the root; once the set {X' > tau} stops changing, the root of ITS quadratic
// (the package's closed form on that set, fp64) is taken and accepted when no entry of the set falls below it -- then
// it is the root of f itself.  Tolerance against the sort-based definition: 1e-6 on P (tests/test_entmax15_gpu.py).
#include <math_constants.h>

#include "common.cuh"
#include "sparsemax_exact.cuh"

namespace smlvm {

constexpr int kE15Threads = 256;

template <int E, int G, bool VEC>
__global__ void __launch_bounds__(kE15Threads)
entmax15_fwd_kernel(const float* __restrict__ x, float* __restrict__ p, float* __restrict__ ent,
                    int64_t* __restrict__ amax, int64_t rows, int K)
{
    constexpr int RPW = kWarp / G;
    const int lane = threadIdx.x & 31, l = lane % G, g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kE15Threads / kWarp);
    for (int64_t w = (int64_t)blockIdx.x * (kE15Threads / kWarp) + (threadIdx.x >> 5); w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int Kr = live ? K : 0;
        float v[E];
        load_row<E, G, VEC>(x + (live ? row : 0) * (int64_t)K, Kr, l, -CUDART_INF_F, v);
        float m = v[0];
#pragma unroll
        for (int e = 1; e < E; ++e) m = fmaxf(m, v[e]);
        m = group_max<G>(m);
        if (!live) m = 0.f;
        int first = 0x7fffffff;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (v[e] == m) first = min(first, col_of<E, G, VEC>(e, l));
            v[e] = __fsub_rn(v[e], m) * 0.5f;
        }
        // Newton from below on f(t) = sum (v - t)_+^2 - 1, sums of the shifted entries d = v - t in fp64
        double t = -1.0;
        int cprev = -1;
        double tau = -1.0;
        for (int iter = 0; iter < 64; ++iter) {
            double s1 = 0.0, s2 = 0.0;
            int c = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double d = (double)v[e] - t;
                if (d > 0.0) { s1 += d; s2 = fma(d, d, s2); ++c; }
            }
            s1 = group_sum<G>(s1);
            s2 = group_sum<G>(s2);
            c = group_sum<G>(c);
            bool done = !live || c == 0;
            tau = t;
            // (every lane of the warp takes the same path through the collectives below: the branch is on a vote)
            const bool stable = !done && c == cprev;
            if (__any_sync(kFull, stable)) {
                // root of the quadratic of the current set: c delta^2 - 2 s1 delta + (s2 - 1) = 0 (smaller root)
                const double disc = fmax(s1 * s1 - (double)c * (s2 - 1.0), 0.0);
                const double tc = t + (s1 - sqrt(disc)) / (double)max(c, 1);
                int below = 0;
#pragma unroll
                for (int e = 0; e < E; ++e) below += ((double)v[e] > t && (double)v[e] < tc) ? 1 : 0;
                below = group_sum<G>(below);
                if (stable && below == 0) { done = true; tau = tc; }
            }
            if (__all_sync(kFull, done)) break;
            cprev = c;
            if (!done) t = fmax(t, t + (s2 - 1.0) / (2.0 * s1));   // f >= 0 left of the root: the step is >= 0
        }
        const float tf = (float)tau;
        const float eps = 1.1920928955078125e-07f;
        float h = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float d = fmaxf(__fsub_rn(v[e], tf), 0.f);
            const float pe = d * d;
            v[e] = pe;
            if (ent != nullptr && pe > 0.f) {
                const float ps = fminf(fmaxf(pe, eps), 1.f - eps);
                h = fmaf(ps, __logf(ps), h);
            }
        }
        if (live) store_row<E, G, VEC>(p + row * (int64_t)K, K, l, v);
        if (ent != nullptr) {
            h = group_sum<G>(h);
            if (live && l == 0) ent[row] = -h;
        }
        if (amax != nullptr) {
            first = group_reduce<G>(first, [](int a, int b) { return min(a, b); });
            if (live && l == 0) amax[row] = first;
        }
    }
}

template <int E, int G, bool VEC>
__global__ void __launch_bounds__(kE15Threads)
entmax15_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ dx, int64_t rows, int K)
{
    constexpr int RPW = kWarp / G;
    const int lane = threadIdx.x & 31, l = lane % G, g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kE15Threads / kWarp);
    for (int64_t w = (int64_t)blockIdx.x * (kE15Threads / kWarp) + (threadIdx.x >> 5); w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int Kr = live ? K : 0;
        const int64_t off = (live ? row : 0) * (int64_t)K;
        float pv[E], gv[E];
        load_row<E, G, VEC>(p + off, Kr, l, 0.f, pv);
        load_row<E, G, VEC>(dp + off, Kr, l, 0.f, gv);
        float sg = 0.f, sd = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            pv[e] = sqrtf(pv[e]);           // gppr
            gv[e] = gv[e] * pv[e];          // dX = dY * gppr
            sg += pv[e];
            sd += gv[e];
        }
        sg = group_sum<G>(sg);
        sd = group_sum<G>(sd);
        const float q = sg > 0.f ? __fdiv_rn(sd, sg) : 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) gv[e] = fmaf(-q, pv[e], gv[e]);
        if (live) store_row<E, G, VEC>(dx + off, K, l, gv);
    }
}

// loss[r] = (1 - sum_j p^alpha) / (alpha (alpha - 1)) + <p - e_y, x>;  pm = p - e_y (the gradient w.r.t. x)
template <int E, int G, bool VEC>
__global__ void __launch_bounds__(kE15Threads)
fy_loss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ p, const int64_t* __restrict__ target, int alpha15,
                   float* __restrict__ loss, float* __restrict__ pm, int64_t rows, int K)
{
    constexpr int RPW = kWarp / G;
    const int lane = threadIdx.x & 31, l = lane % G, g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kE15Threads / kWarp);
    for (int64_t w = (int64_t)blockIdx.x * (kE15Threads / kWarp) + (threadIdx.x >> 5); w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int Kr = live ? K : 0;
        const int64_t off = (live ? row : 0) * (int64_t)K;
        float pv[E], xv[E];
        load_row<E, G, VEC>(p + off, Kr, l, 0.f, pv);
        load_row<E, G, VEC>(x + off, Kr, l, 0.f, xv);
        const int y = live ? (int)__ldg(target + row) : -1;
        float spa = 0.f, dot = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            spa += alpha15 ? pv[e] * sqrtf(pv[e]) : pv[e] * pv[e];
            if (col_of<E, G, VEC>(e, l) == y) pv[e] -= 1.f;
            dot = fmaf(pv[e], xv[e], dot);
        }
        spa = group_sum<G>(spa);
        dot = group_sum<G>(dot);
        if (live) {
            if (pm != nullptr) store_row<E, G, VEC>(pm + off, K, l, pv);
            if (l == 0) loss[row] = (1.f - spa) * (alpha15 ? (1.f / 0.75f) : 0.5f) + dot;
        }
    }
}

// out[r, :] = scale[r] * a[r, :]   (backward of the Fenchel-Young loss: grad_output[r] * (p - e_y))
__global__ void __launch_bounds__(256)
row_scale_kernel(const float* __restrict__ a, const float* __restrict__ scale, float* __restrict__ out, int64_t rows, int K)
{
    const int64_t total = rows * K, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        out[i] = __ldg(scale + i / K) * __ldg(a + i);
}

struct E15Shape { int E, G; };
// Large batches of rows wider than 32: 32 values per lane, K / 32 lanes per row (the layout of sparsemax_coop.cu) --
// the Newton passes run in fp64 and their group reductions are 64-bit shuffles: 74 warp instructions per row and pass
// with a warp per row (ncu r02: 1.70 ms at 1M x 128), 24 with 4 lanes per row.
static E15Shape e15_shape(int K, int64_t rows)
{
    if (rows >= 4096 && K > 32 && K <= 512) return {32, K <= 64 ? 2 : (K <= 128 ? 4 : (K <= 256 ? 8 : 16))};
    if (K <= 4) return {4, 1};
    if (K <= 8) return {4, 2};
    if (K <= 16) return {4, 4};
    if (K <= 32) return {4, 8};
    if (K <= 64) return {4, 16};
    if (K <= 128) return {4, 32};
    if (K <= 256) return {8, 32};
    if (K <= 512) return {16, 32};
    return {32, 32};
}
static int e15_grid(int64_t rows, int G)
{
    const int rows_per_cta = (kE15Threads / kWarp) * (kWarp / G);
    const int64_t need = (rows + rows_per_cta - 1) / rows_per_cta;
    const int64_t cap = (int64_t)device_sm_count() * 8;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

#define SMLVM_E15_DISPATCH(KERNEL, vec, ...)                                                         \
    do {                                                                                             \
        const E15Shape sh = e15_shape(cols, rows);                                                         \
        const int grid = e15_grid(rows, sh.G);                                                       \
        switch (sh.E * 100 + sh.G) {                                                                 \
            case 401: if (vec) KERNEL<4, 1, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<4, 1, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 402: if (vec) KERNEL<4, 2, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<4, 2, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 404: if (vec) KERNEL<4, 4, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<4, 4, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 408: if (vec) KERNEL<4, 8, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<4, 8, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 416: if (vec) KERNEL<4, 16, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<4, 16, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 432: if (vec) KERNEL<4, 32, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<4, 32, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 3202: if (vec) KERNEL<32, 2, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<32, 2, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 3204: if (vec) KERNEL<32, 4, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<32, 4, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 3208: if (vec) KERNEL<32, 8, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<32, 8, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 3216: if (vec) KERNEL<32, 16, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<32, 16, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 832: if (vec) KERNEL<8, 32, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<8, 32, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            case 1632: if (vec) KERNEL<16, 32, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<16, 32, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
            default: if (vec) KERNEL<32, 32, true><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); else KERNEL<32, 32, false><<<grid, kE15Threads, 0, st>>>(__VA_ARGS__); break; \
        }                                                                                            \
    } while (0)

static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

}  // namespace smlvm

using namespace smlvm;

extern "C" int smlvm_entmax15_fwd(const float* x, float* p, float* entropy, int64_t* argmax, int64_t rows, int32_t cols,
                                  void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (x == nullptr || p == nullptr) return SMLVM_ERR_ARG;
    if (cols > 1024) return SMLVM_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const bool vec = (cols % 4 == 0) && al16(x) && al16(p);
    SMLVM_E15_DISPATCH(entmax15_fwd_kernel, vec, x, p, entropy, argmax, rows, cols);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_entmax15_bwd(const float* p, const float* dp, float* dx, int64_t rows, int32_t cols, void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (p == nullptr || dp == nullptr || dx == nullptr) return SMLVM_ERR_ARG;
    if (cols > 1024) return SMLVM_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const bool vec = (cols % 4 == 0) && al16(p) && al16(dp) && al16(dx);
    SMLVM_E15_DISPATCH(entmax15_bwd_kernel, vec, p, dp, dx, rows, cols);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_fy_loss_fwd(const float* x, const float* p, const int64_t* target, int32_t alpha15, float* loss,
                                 float* p_minus_y, int64_t rows, int32_t cols, void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (x == nullptr || p == nullptr || target == nullptr || loss == nullptr) return SMLVM_ERR_ARG;
    if (cols > 1024) return SMLVM_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const bool vec = (cols % 4 == 0) && al16(x) && al16(p) && (p_minus_y == nullptr || al16(p_minus_y));
    const int a15 = alpha15 ? 1 : 0;
    SMLVM_E15_DISPATCH(fy_loss_fwd_kernel, vec, x, p, target, a15, loss, p_minus_y, rows, cols);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_row_scale(const float* a, const float* scale, float* out, int64_t rows, int32_t cols, void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (a == nullptr || scale == nullptr || out == nullptr) return SMLVM_ERR_ARG;
    const int64_t need = (rows * cols + 255) / 256, cap = (int64_t)device_sm_count() * 16;
    row_scale_kernel<<<(int)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(a, scale, out, rows, cols);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}
