// Batched sparsemax forward / backward for sm_100a.
//
// Semantics follow entmax.sparsemax as the reference calls it
// (lvmhelpers/marg.py:51, lvmhelpers/structmarg.py:48) evaluated with torch CPU
// ops, because support sets must be bit-exact against that path:
//   X' = X - max(X)                      (fp32)
//   S  = sort(X', descending)
//   C_r = fl32(fl32(sum_{j<=r} S_j in fp64) - 1)      (torch CPU cumsum accumulates fp64)
//   k  = #{ r : fl32(r * S_r) > C_r }    (strict)
//   tau = C_k / k                        (fp32 divide)
//   P  = max(X' - tau, 0)
// Backward (SparsemaxFunction.backward):
//   g <- dP with g = 0 where P == 0;  dX = where(P != 0, g - sum(g)/k, 0)
//
// Layout: a row of K fp32 lives in the registers of a G-lane sub-warp group
// (E values per lane, E*G >= K), so a warp handles 32/G rows at once; the sort
// is a bitonic network whose strides < E are register-local and whose strides
// >= E are shuffles.  Rows wider than 1024 go through a shared-memory path
// (one CTA per row).  Every row is read once and written once (HBM-bound
// design point: 8K B/row forward, 12K B/row backward).
#include <math_constants.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "sparsemax_exact.cuh"

namespace smlvm {

constexpr int kThreads = 256;

// ---- register path ---------------------------------------------------------

template <int E, int G, bool VEC>
__global__ void __launch_bounds__(kThreads)
sparsemax_fwd_reg(const float* __restrict__ x, float* __restrict__ p, int32_t* __restrict__ supp,
                  int32_t* __restrict__ nnz, float* __restrict__ ent, int64_t* __restrict__ amax,
                  int64_t rows, int K)
{
    constexpr int RPW = kWarp / G;  // rows per warp
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kThreads / kWarp) + (threadIdx.x >> 5);

    for (; w * RPW < rows; w += warps_total) {
        const int64_t row = w * RPW + g;
        const bool live = row < rows;
        const int Kr = live ? K : 0;  // dead groups see an empty row (all -inf)
        const float* xr = x + (live ? row : 0) * (int64_t)K;

        float v[E];
        load_row<E, G, VEC>(xr, Kr, l, -CUDART_INF_F, v);

        // row max and its first index
        float m = v[0];
#pragma unroll
        for (int e = 1; e < E; ++e) m = fmaxf(m, v[e]);
        m = gmax_fast<G>(m);

        if (amax != nullptr) {
            int first = 0x7fffffff;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                int col = col_of<E, G, VEC>(e, l);
                if (col < Kr && v[e] == m) first = min(first, col);
            }
            first = gmin_fast<G>(first);
            if (live && l == 0) amax[row] = first;
        }
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = __fsub_rn(v[e], m);  // X' (padding stays -inf)

        // ---- fast path: threshold search (Michelot / Newton on sum max(x - tau, 0) = 1) -----
        // in fixed point with single-instruction warp reductions, then one exact evaluation of the reference's tau on the proposed
        // support and a check, with rounding margins, that the reference's sort-based support
        // test yields exactly this support.  Rows that fail the check (an entry within ~1e-6
        // of the threshold, or a proposal spoilt by the 2^-SHIFT grid) take the sort path below.
        constexpr int LOGN = ilog2_c(E * G);
        constexpr int SHIFT = (30 - LOGN) < 24 ? (30 - LOGN) : 24;  // N * 2^SHIFT <= 2^30
        constexpr int ONE = 1 << SHIFT;
        int qv[E];
#pragma unroll
        for (int e = 0; e < E; ++e) qv[e] = __float2int_rn(v[e] * (float)ONE);  // -inf -> INT_MIN
        int tq = -ONE;  // tau >= max - 1
        int cprev = -1, C = 0;
        for (int iter = 0; iter < 64; ++iter) {
            int sq = 0, c = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const bool in = qv[e] > tq;
                sq += in ? qv[e] : 0;
                c += in ? 1 : 0;
            }
            const int S = gsum_fast<G>(sq);
            C = gsum_fast<G>(c);
            if (__all_sync(kFull, C == cprev)) break;  // supports stopped shrinking in every row
            cprev = C;
            const int tn = __float2int_rd(__fdividef((float)(S - ONE), (float)max(C, 1)));
            tq = max(tq, tn);
        }
        int k = max(C, 1);
        float tau;
        bool need_sort;
        {
            double sd = 0.0;
#pragma unroll
            for (int e = 0; e < E; ++e) sd += (qv[e] > tq) ? (double)v[e] : 0.0;
            sd = group_sum<G>(sd);
            // the reference's arithmetic: tau = fl32(fl32(fl32(cumsum_k) - 1) / k)
            const float kf = (float)k;
            tau = __fdiv_rn(__fsub_rn((float)sd, 1.0f), kf);
            // margins (sparsemax): inside needs x - tau > 6*2^-24 (+ tau rounding),
            // outside needs (tau - x) * k > 2^-24 (3.06 K + 1) (+ k * tau rounding); x2 safety
            const float thr_in = tau + 1e-6f;
            const float thr_out = tau - __fdividef((3.06f * (float)K + 1.f) * 1.2e-7f, kf) - 3e-7f;
            bool bad = false;
#pragma unroll
            for (int e = 0; e < E; ++e) bad |= (qv[e] > tq) ? !(v[e] > thr_in) : !(v[e] < thr_out);
            const unsigned gm = (G == 32) ? kFull : (((1u << G) - 1u) << (g * G));
            // (ballot hoisted: `live && ballot()` would short-circuit the collective away
            //  from dead lanes and deadlock the warp)
            const unsigned bal = __ballot_sync(kFull, bad);
            need_sort = live && ((bal & gm) != 0u);
        }

        if (__any_sync(kFull, need_sort)) {
            // ---- exact path: the reference's definition, by sorting -----------------------
            int ks;
            float taus;
            exact_threshold_sorted<E, G>(v, l, Kr, ks, taus);
            if (need_sort) {
                tau = taus;
                k = ks;
            }
        }

        float h = 0.f;
        int nz = 0;
        const float eps = 1.1920928955078125e-07f;  // torch.finfo(float32).eps
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const float pe = fmaxf(__fsub_rn(v[e], tau), 0.f);
            v[e] = pe;
            if (pe > 0.f) {
                ++nz;
                if (ent != nullptr) {
                    const float ps = fminf(fmaxf(pe, eps), 1.f - eps);
                    h = fmaf(ps, __logf(ps), h);  // MUFU.LG2 path: |err| <= 4e-7 on p log p
                }
            }
        }
        if (live) store_row<E, G, VEC>(p + row * (int64_t)K, K, l, v);
        if (nnz != nullptr) {
            // on the verified fast path #(p > 0) == k; only sorted rows need the count
            int nzr = k;
            if (__any_sync(kFull, need_sort)) {
                const int cnt = gsum_fast<G>(nz);
                if (need_sort) nzr = cnt;
            }
            if (live && l == 0) nnz[row] = nzr;
        }
        if (ent != nullptr) {
            h = group_sum<G>(h);
            if (live && l == 0) ent[row] = -h;
        }
        if (supp != nullptr && live && l == 0) supp[row] = k;
    }
}

// U rows per group and iteration: every load of the U rows is issued before the first
// dependent instruction, so a warp keeps 2*U..3*U 128-bit requests per lane in flight
// (ncu r01: the one-row version stalled 19 cycles per issue on the long scoreboard, 71 % of
// the HBM roofline with four streams).
template <int E, int G, bool VEC, int U>
__global__ void __launch_bounds__(kThreads)
sparsemax_bwd_reg(const float* __restrict__ p, const int32_t* __restrict__ supp,
                  const float* __restrict__ dp, const float* __restrict__ loss,
                  const float* __restrict__ loss_scale_dev, float loss_scale_host,
                  const float* __restrict__ dent, float* __restrict__ dx,
                  float* __restrict__ dloss, int64_t rows, int K)
{
    constexpr int RPW = kWarp / G;
    const int lane = threadIdx.x & 31;
    const int l = lane % G;
    const int g = lane / G;
    const int64_t warps_total = (int64_t)gridDim.x * (kThreads / kWarp);
    int64_t w = (int64_t)blockIdx.x * (kThreads / kWarp) + (threadIdx.x >> 5);
    float ls = loss_scale_host;
    if (loss_scale_dev != nullptr) ls *= __ldg(loss_scale_dev);

    for (; w * (RPW * U) < rows; w += warps_total) {
        float pv[U][E], gv[U][E], lv[U][E], de[U];
        int kk[U];
        bool live[U];
        int64_t off[U];
        // ---- all loads first ------------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = (w * U + u) * RPW + g;
            live[u] = row < rows;
            const int Kr = live[u] ? K : 0;
            off[u] = (live[u] ? row : 0) * (int64_t)K;
            load_row<E, G, VEC>(p + off[u], Kr, l, 0.f, pv[u]);
            if (dp != nullptr) {
                load_row<E, G, VEC>(dp + off[u], Kr, l, 0.f, gv[u]);
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) gv[u][e] = 0.f;
            }
            if (loss != nullptr) load_row<E, G, VEC>(loss + off[u], Kr, l, 0.f, lv[u]);
            de[u] = (dent != nullptr && live[u]) ? __ldg(dent + row) : 0.f;
            kk[u] = live[u] ? __ldg(supp + row) : 1;
        }
        // ---- then the arithmetic and the stores -------------------------------------------------
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (dloss != nullptr) {  // decoder side of the weighted loss: d/dL sum(P L) * ls = ls * P
                float dl[E];
#pragma unroll
                for (int e = 0; e < E; ++e) dl[e] = ls * pv[u][e];
                if (live[u]) store_row<E, G, VEC>(dloss + off[u], K, l, dl);
            }
            if (loss != nullptr) {
#pragma unroll
                for (int e = 0; e < E; ++e) gv[u][e] = fmaf(ls, lv[u][e], gv[u][e]);
            }
            if (dent != nullptr) {
#pragma unroll
                for (int e = 0; e < E; ++e) gv[u][e] = fmaf(de[u], dent_dp(pv[u][e]), gv[u][e]);
            }
            float sum = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                gv[u][e] = (pv[u][e] != 0.f) ? gv[u][e] : 0.f;
                sum += gv[u][e];
            }
            sum = group_sum<G>(sum);
            const float vhat = __fdiv_rn(sum, (float)kk[u]);
#pragma unroll
            for (int e = 0; e < E; ++e) gv[u][e] = (pv[u][e] != 0.f) ? __fsub_rn(gv[u][e], vhat) : 0.f;
            if (live[u]) store_row<E, G, VEC>(dx + off[u], K, l, gv[u]);
        }
    }
}

// ---- shared-memory path (1024 < K <= 8192): one CTA per row -------------------

constexpr int kWideThreads = 256;

__global__ void __launch_bounds__(kWideThreads)
sparsemax_fwd_wide(const float* __restrict__ x, float* __restrict__ p, int32_t* __restrict__ supp,
                   int32_t* __restrict__ nnz, float* __restrict__ ent, int64_t* __restrict__ amax,
                   int64_t rows, int K, int N /* pow2 >= K */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);  // [N] shifted row
    float* ss = xs + N;                              // [N] sorted copy
    double* red = reinterpret_cast<double*>(ss + N); // [kWideThreads + 32]
    __shared__ float s_f[2];
    __shared__ int s_i[2];
    const int tid = threadIdx.x;

    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const float* xr = x + row * (int64_t)K;
        float m = -CUDART_INF_F;
        for (int c = tid; c < N; c += kWideThreads) {
            float val = (c < K) ? ldg_stream(xr + c) : -CUDART_INF_F;
            xs[c] = val;
            m = fmaxf(m, val);
        }
        m = group_max<32>(m);
        float* redf = reinterpret_cast<float*>(red);
        __syncthreads();
        if ((tid & 31) == 0) redf[tid >> 5] = m;
        __syncthreads();
        m = redf[0];
        for (int i = 1; i < kWideThreads / 32; ++i) m = fmaxf(m, redf[i]);
        __syncthreads();
        int first = 0x7fffffff;
        for (int c = tid; c < N; c += kWideThreads) {
            if (c < K && xs[c] == m) first = min(first, c);
            float sh = __fsub_rn(xs[c], m);
            xs[c] = sh;
            ss[c] = sh;
        }
        if (amax != nullptr) {
            first = group_reduce<32>(first, [](int a, int b) { return min(a, b); });
            int* redi = reinterpret_cast<int*>(red);
            if ((tid & 31) == 0) redi[tid >> 5] = first;
            __syncthreads();
            if (tid == 0) {
                int f = redi[0];
                for (int i = 1; i < kWideThreads / 32; ++i) f = min(f, redi[i]);
                amax[row] = f;
            }
        }
        __syncthreads();
        // bitonic sort, descending
        for (int k = 2; k <= N; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < N; i += kWideThreads) {
                    int q = i ^ j;
                    if (q > i) {
                        bool desc = (i & k) == 0;
                        float a = ss[i], b = ss[q];
                        float hi = fmaxf(a, b), lo = fminf(a, b);
                        ss[i] = desc ? hi : lo;
                        ss[q] = desc ? lo : hi;
                    }
                }
                __syncthreads();
            }
        }
        // fp64 prefix sums: contiguous chunk per thread, then scan of chunk totals
        const int chunk = N / kWideThreads;  // N >= 2048 > kWideThreads
        const int beg = tid * chunk;
        double run = 0.0;
        for (int i = 0; i < chunk; ++i) run += (double)ss[beg + i];
        red[tid] = run;
        __syncthreads();
        if (tid == 0) {
            double acc = 0.0;
            for (int i = 0; i < kWideThreads; ++i) {
                double t = red[i];
                red[i] = acc;
                acc += t;
            }
        }
        __syncthreads();
        const double excl = red[tid];
        run = 0.0;
        int cnt = 0;
        for (int i = 0; i < chunk; ++i) {
            const int pos = beg + i;
            run += (double)ss[pos];
            float C = __fsub_rn((float)(excl + run), 1.0f);
            float lhs = __fmul_rn((float)(pos + 1), ss[pos]);
            cnt += (pos < K && lhs > C) ? 1 : 0;
        }
        cnt = group_sum<32>(cnt);
        __syncthreads();
        int* redi = reinterpret_cast<int*>(red);
        const double excl_keep = excl;
        if ((tid & 31) == 0) redi[tid >> 5] = cnt;
        __syncthreads();
        if (tid == 0) {
            int kk = 0;
            for (int i = 0; i < kWideThreads / 32; ++i) kk += redi[i];
            s_i[0] = max(kk, 1);
        }
        __syncthreads();
        const int k = s_i[0];
        if ((k - 1) / chunk == tid) {
            double r2 = 0.0;
            for (int i = 0; i <= (k - 1) - beg; ++i) r2 += (double)ss[beg + i];
            float C = __fsub_rn((float)(excl_keep + r2), 1.0f);
            s_f[0] = __fdiv_rn(C, (float)k);
        }
        __syncthreads();
        const float tau = s_f[0];
        const float eps = 1.1920928955078125e-07f;
        int nz = 0;
        float h = 0.f;
        float* pr = p + row * (int64_t)K;
        for (int c = tid; c < K; c += kWideThreads) {
            float pe = fmaxf(__fsub_rn(xs[c], tau), 0.f);
            stg_stream(pr + c, pe);
            if (pe > 0.f) {
                ++nz;
                float ps = fminf(fmaxf(pe, eps), 1.f - eps);
                h = fmaf(ps, logf(ps), h);
            }
        }
        double tot_h = block_sum((double)h, red);
        double tot_nz = block_sum((double)nz, red);
        if (tid == 0) {
            if (supp != nullptr) supp[row] = k;
            if (nnz != nullptr) nnz[row] = (int)tot_nz;
            if (ent != nullptr) ent[row] = -(float)tot_h;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kWideThreads)
sparsemax_bwd_wide(const float* __restrict__ p, const int32_t* __restrict__ supp,
                   const float* __restrict__ dp, const float* __restrict__ loss,
                   const float* __restrict__ loss_scale_dev, float loss_scale_host,
                   const float* __restrict__ dent, float* __restrict__ dx,
                   float* __restrict__ dloss, int64_t rows, int K)
{
    __shared__ double red[kWideThreads / 32 + 1];
    const int tid = threadIdx.x;
    float ls = loss_scale_host;
    if (loss_scale_dev != nullptr) ls *= __ldg(loss_scale_dev);
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t off = row * (int64_t)K;
        const float de = dent != nullptr ? __ldg(dent + row) : 0.f;
        float sum = 0.f;
        for (int c = tid; c < K; c += kWideThreads) {
            float pe = p[off + c];
            if (pe != 0.f) {
                float g = dp != nullptr ? dp[off + c] : 0.f;
                if (loss != nullptr) g = fmaf(ls, loss[off + c], g);
                if (dent != nullptr) g = fmaf(de, dent_dp(pe), g);
                sum += g;
            }
        }
        const float tot = (float)block_sum((double)sum, red);
        const float vhat = __fdiv_rn(tot, (float)__ldg(supp + row));
        for (int c = tid; c < K; c += kWideThreads) {
            float pe = p[off + c];
            float o = 0.f;
            if (dloss != nullptr) stg_stream(dloss + off + c, ls * pe);
            if (pe != 0.f) {
                float g = dp != nullptr ? dp[off + c] : 0.f;
                if (loss != nullptr) g = fmaf(ls, loss[off + c], g);
                if (dent != nullptr) g = fmaf(de, dent_dp(pe), g);
                o = __fsub_rn(g, vhat);
            }
            stg_stream(dx + off + c, o);
        }
        __syncthreads();
    }
}

// ---- backward over the support only (ragged upstream gradient) ---------------------------------
// dx is zero outside the support: the dense matrix is cleared by a memset node and one thread per
// row scatters g_j - mean(g) at its (few) support columns.  Reads 12-20 B per support entry instead
// of 8K-12K B per row.
__global__ void __launch_bounds__(256)
sparsemax_bwd_ragged_kernel(const int64_t* __restrict__ offsets, const int32_t* __restrict__ cols,
                            const float* __restrict__ vals, const int32_t* __restrict__ supp,
                            const float* __restrict__ dp, const float* __restrict__ loss,
                            const float* __restrict__ loss_scale_dev, float loss_scale_host,
                            const float* __restrict__ dent, float* __restrict__ dx, float* __restrict__ dloss,
                            int64_t rows, int K, int64_t n_items)
{
    float ls = loss_scale_host;
    if (loss_scale_dev != nullptr) ls *= __ldg(loss_scale_dev);
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = min(__ldg(offsets + r), n_items), e = min(__ldg(offsets + r + 1), n_items);
        const float de = dent != nullptr ? __ldg(dent + r) : 0.f;
        float sum = 0.f;
        for (int64_t j = b; j < e; ++j) {
            const float pj = __ldg(vals + j);
            float g = dp != nullptr ? __ldg(dp + j) : 0.f;
            if (loss != nullptr) g = fmaf(ls, __ldg(loss + j), g);
            if (dent != nullptr) g = fmaf(de, dent_dp(pj), g);
            sum += g;
            if (dloss != nullptr) dloss[j] = ls * pj;
        }
        const float vhat = __fdiv_rn(sum, (float)__ldg(supp + r));
        float* out = dx + r * (int64_t)K;
        for (int64_t j = b; j < e; ++j) {
            const float pj = __ldg(vals + j);
            float g = dp != nullptr ? __ldg(dp + j) : 0.f;
            if (loss != nullptr) g = fmaf(ls, __ldg(loss + j), g);
            if (dent != nullptr) g = fmaf(de, dent_dp(pj), g);
            out[__ldg(cols + j)] = __fsub_rn(g, vhat);
        }
    }
}

// ---- host dispatch ----------------------------------------------------------------

struct RegShape {
    int E, G;
};

// Smallest (E, G) tile with E*G >= K; E = 4 (one float4 per lane) up to K = 128.
static RegShape pick_shape(int K)
{
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

static int grid_for(int64_t rows, int G)
{
    const int rows_per_cta = (kThreads / kWarp) * (kWarp / G);
    int64_t need = (rows + rows_per_cta - 1) / rows_per_cta;
    int64_t cap = (int64_t)device_sm_count() * 16;  // resident waves; grid-stride beyond
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

template <int E, int G>
static void launch_fwd(bool vec, int grid, cudaStream_t st, const float* x, float* p, int32_t* supp,
                       int32_t* nnz, float* ent, int64_t* amax, int64_t rows, int K)
{
    if (vec)
        sparsemax_fwd_reg<E, G, true><<<grid, kThreads, 0, st>>>(x, p, supp, nnz, ent, amax, rows, K);
    else
        sparsemax_fwd_reg<E, G, false><<<grid, kThreads, 0, st>>>(x, p, supp, nnz, ent, amax, rows, K);
}
template <int E, int G>
static void launch_bwd(bool vec, int grid, cudaStream_t st, const float* p, const int32_t* supp,
                       const float* dp, const float* loss, const float* lsd, float lsh,
                       const float* dent, float* dx, float* dloss, int64_t rows, int K)
{
    constexpr int U = (E <= 8) ? 2 : 1;
    const int64_t rows_per_cta = (int64_t)(kThreads / kWarp) * (kWarp / G) * U;
    const int64_t need = (rows + rows_per_cta - 1) / rows_per_cta;
    grid = (int)(need < grid ? (need > 0 ? need : 1) : grid);
    if (vec)
        sparsemax_bwd_reg<E, G, true, U><<<grid, kThreads, 0, st>>>(p, supp, dp, loss, lsd, lsh, dent, dx, dloss, rows, K);
    else
        sparsemax_bwd_reg<E, G, false, U><<<grid, kThreads, 0, st>>>(p, supp, dp, loss, lsd, lsh, dent, dx, dloss, rows, K);
}

#define SMLVM_DISPATCH_SHAPE(FN, ...)                                   \
    switch (sh.E * 100 + sh.G) {                                        \
        case 401: FN<4, 1>(__VA_ARGS__); break;                         \
        case 402: FN<4, 2>(__VA_ARGS__); break;                         \
        case 404: FN<4, 4>(__VA_ARGS__); break;                         \
        case 408: FN<4, 8>(__VA_ARGS__); break;                         \
        case 416: FN<4, 16>(__VA_ARGS__); break;                        \
        case 432: FN<4, 32>(__VA_ARGS__); break;                        \
        case 832: FN<8, 32>(__VA_ARGS__); break;                        \
        case 1632: FN<16, 32>(__VA_ARGS__); break;                      \
        default: FN<32, 32>(__VA_ARGS__); break;                        \
    }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// sparsemax_bwd_tile.cu
bool bwd_slots_supported(int cols);
int bwd_slots_launch(const float* p, const int32_t* supp, const float* dp, const float* loss, const float* lsd,
                     float lsh, const float* dent, const float* slots, const float* loss_slots, float* dx,
                     float* dloss, int64_t rows, int K, cudaStream_t st);
bool bwd_ragged_tile_supported(int cols);
int bwd_ragged_tile_launch(const int64_t* offsets, const int32_t* cols_idx, const float* vals, const int32_t* supp,
                           const float* dp, const float* loss, const float* lsd, float lsh, const float* dent,
                           float* dx, float* dloss, int64_t rows, int K, int64_t n_items, cudaStream_t st);

// sparsemax_tile.cu
bool sparsemax_tile_supported(int cols);
int sparsemax_fwd_tile_launch(const float* x, float* p, int32_t* supp, int32_t* nnz, float* ent, int64_t* amax,
                              float* support_slots, int64_t rows, int cols, cudaStream_t st);
// sparsemax_coop.cu
bool sparsemax_coop_supported(int cols);
int sparsemax_fwd_coop_launch(const float* x, float* p, int32_t* supp, int32_t* nnz, float* ent, int64_t* amax,
                              int64_t rows, int cols, bool vec, cudaStream_t st);
// compact_tile.cu
int support_slots_launch(const float* p, float* slots, int64_t rows, int cols, cudaStream_t st);

// Forward path selection.  The TMA-staged tile kernel wins once there are enough 32-row tiles to
// occupy the SMs; tiny batches (the B = 64 experiment shapes) are launch-bound and stay on the
// register kernel.  SMLVM_SPARSEMAX_FWD=reg|tile|coop overrides (tests/test_sparsemax_gpu.py runs every path in a subprocess).
static int fwd_path_override()
{
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("SMLVM_SPARSEMAX_FWD");
        cached = (e == nullptr) ? 0 : (strcmp(e, "reg") == 0 ? 1 : (strcmp(e, "tile") == 0 ? 2 : (strcmp(e, "coop") == 0 ? 3 : 0)));
    }
    return cached;
}
constexpr int64_t kTileMinRows = 4096;

}  // namespace smlvm

using namespace smlvm;

extern "C" int smlvm_sparsemax_fwd(const float* x, float* p, int32_t* supp_size, int32_t* nnz,
                                   float* entropy, int64_t* argmax, float* support_slots, int64_t rows,
                                   int32_t cols, void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (x == nullptr || p == nullptr) return SMLVM_ERR_ARG;
    if (cols > SMLVM_SPARSEMAX_MAX_COLS) return SMLVM_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int ovr = fwd_path_override();
    // Three forward kernels.  Large batches: the cooperative register kernel (G lanes per row, cost
    // independent of the support size) unless the caller wants the support slots -- a caller asks for those when it
    // expects peaked rows (supports of a few entries), where the thread-per-row tile kernel is the fastest and the
    // only one that emits them.  Small batches are launch-bound and stay on the warp-per-row register kernel.
    const bool big = rows >= kTileMinRows;
    const bool coop_ok = sparsemax_coop_supported(cols) && ovr != 1 && ovr != 2 && (big || ovr == 3);
    const bool tile_ok = sparsemax_tile_supported(cols) && aligned16(x) && aligned16(p) && ovr != 1 && ovr != 3 &&
                         (big || ovr == 2);
    if (tile_ok && (support_slots != nullptr || !coop_ok))
        return sparsemax_fwd_tile_launch(x, p, supp_size, nnz, entropy, argmax, support_slots, rows, cols, st);
    if (coop_ok) {
        const bool vec = (cols % 4 == 0) && aligned16(x) && aligned16(p);
        const int rc = sparsemax_fwd_coop_launch(x, p, supp_size, nnz, entropy, argmax, rows, cols, vec, st);
        if (rc != SMLVM_OK) return rc;
    } else if (cols <= 1024) {
        RegShape sh = pick_shape(cols);
        const bool vec = (cols % 4 == 0) && aligned16(x) && aligned16(p);
        const int grid = grid_for(rows, sh.G);
        SMLVM_DISPATCH_SHAPE(launch_fwd, vec, grid, st, x, p, supp_size, nnz, entropy, argmax, rows, cols);
    } else {
        int N = 2048;
        while (N < cols) N <<= 1;
        size_t smem = (size_t)N * 8 + (kWideThreads + 32) * sizeof(double);
        static bool attr_set_dev[64] = {false};   // function attributes are per device
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attr_set_dev[dev & 63]) {
            cudaError_t ea = cudaFuncSetAttribute(sparsemax_fwd_wide, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  8192 * 8 + (kWideThreads + 32) * (int)sizeof(double));
            if (ea != cudaSuccess) return (int)ea;
            attr_set_dev[dev & 63] = true;
        }
        int64_t cap = (int64_t)device_sm_count() * 8;
        int grid = (int)(rows < cap ? rows : cap);
        sparsemax_fwd_wide<<<grid, kWideThreads, smem, st>>>(x, p, supp_size, nnz, entropy, argmax,
                                                             rows, cols, N);
    }
    SMLVM_LAUNCH_CHECK();
    if (support_slots != nullptr) return support_slots_launch(p, support_slots, rows, cols, st);
    return SMLVM_OK;
}

extern "C" int smlvm_sparsemax_bwd(const float* p, const int32_t* supp_size, const float* dp,
                                   const float* loss, const float* loss_scale_dev,
                                   float loss_scale_host, const float* dent, const float* support_slots,
                                   const float* loss_slots, float* dx, float* dloss, int64_t rows,
                                   int32_t cols, void* stream)
{
    if (rows < 0 || cols < 1) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (p == nullptr || supp_size == nullptr || dx == nullptr) return SMLVM_ERR_ARG;
    if (cols > SMLVM_SPARSEMAX_MAX_COLS) return SMLVM_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (support_slots != nullptr && bwd_slots_supported(cols) && rows >= kTileMinRows && aligned16(dx) &&
        (dloss == nullptr || aligned16(dloss)))
        return bwd_slots_launch(p, supp_size, dp, loss, loss_scale_dev, loss_scale_host, dent, support_slots,
                                loss_slots, dx, dloss, rows, cols, st);
    if (cols <= 1024) {
        RegShape sh = pick_shape(cols);
        const bool vec = (cols % 4 == 0) && aligned16(p) && aligned16(dx) &&
                         (dp == nullptr || aligned16(dp)) && (loss == nullptr || aligned16(loss)) &&
                         (dloss == nullptr || aligned16(dloss));
        const int grid = grid_for(rows, sh.G);
        SMLVM_DISPATCH_SHAPE(launch_bwd, vec, grid, st, p, supp_size, dp, loss, loss_scale_dev,
                             loss_scale_host, dent, dx, dloss, rows, cols);
    } else {
        int64_t cap = (int64_t)device_sm_count() * 8;
        int grid = (int)(rows < cap ? rows : cap);
        sparsemax_bwd_wide<<<grid, kWideThreads, 0, st>>>(p, supp_size, dp, loss, loss_scale_dev,
                                                          loss_scale_host, dent, dx, dloss, rows, cols);
    }
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

// sparsemax_bwd_tile.cu
namespace smlvm {
int step_fused_launch(const float* p, const int32_t* supp, const float* loss, float scale, const float* dent,
                      const float* slots, const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx,
                      int64_t capacity, float* row_loss, float* total, void* workspace, float* dx, float* dloss,
                      int64_t rows, int K, cudaStream_t st);
// marg_step_dense.cu
bool step_dense_supported(int cols);
int step_dense_launch(const float* p, const int32_t* supp, const float* loss, float scale, const float* dent,
                      const int64_t* offsets, float* vals, int32_t* row_idx, int32_t* col_idx, int64_t capacity,
                      float* row_loss, float* total, void* workspace, float* dx, float* dloss, int64_t rows, int K,
                      cudaStream_t st);
}

extern "C" int smlvm_marg_step_fused(const float* p, const int32_t* supp_size, const float* loss, float scale,
                                     const float* dent, const float* support_slots, const int64_t* offsets,
                                     float* vals, int32_t* row_idx, int32_t* col_idx, int64_t capacity,
                                     float* row_loss, float* total, float* dx, float* dloss, int64_t rows,
                                     int32_t cols, void* workspace, size_t workspace_bytes, void* stream)
{
    if (rows < 0 || cols < 1 || capacity < 0) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (p == nullptr || supp_size == nullptr || loss == nullptr || dx == nullptr) return SMLVM_ERR_ARG;
    if (total != nullptr && (workspace == nullptr || workspace_bytes < smlvm_reduce_workspace_bytes()))
        return SMLVM_ERR_WORKSPACE;
    if (support_slots == nullptr) {   // wide supports: one dense read of P and L (marg_step_dense.cu)
        if (!step_dense_supported(cols) || !aligned16(p) || !aligned16(loss) || !aligned16(dx) ||
            (dloss != nullptr && !aligned16(dloss)))
            return SMLVM_ERR_UNSUPPORTED;
        return step_dense_launch(p, supp_size, loss, scale, dent, offsets, vals, row_idx, col_idx, capacity, row_loss,
                                 total, workspace, dx, dloss, rows, cols, as_stream(stream));
    }
    if (!bwd_slots_supported(cols) || !aligned16(dx) || (dloss != nullptr && !aligned16(dloss)))
        return SMLVM_ERR_UNSUPPORTED;
    return step_fused_launch(p, supp_size, loss, scale, dent, support_slots, offsets, vals, row_idx, col_idx, capacity,
                             row_loss, total, workspace, dx, dloss, rows, cols, as_stream(stream));
}

extern "C" int smlvm_sparsemax_bwd_ragged(const int64_t* offsets, const int32_t* col_idx, const float* vals,
                                          const int32_t* supp_size, const float* dp, const float* loss,
                                          const float* loss_scale_dev, float loss_scale_host, const float* dent,
                                          float* dx, float* dloss, int64_t rows, int32_t cols, int64_t n_items,
                                          void* stream)
{
    if (rows < 0 || cols < 1 || n_items < 0) return SMLVM_ERR_ARG;
    if (rows == 0) return SMLVM_OK;
    if (offsets == nullptr || col_idx == nullptr || vals == nullptr || supp_size == nullptr || dx == nullptr)
        return SMLVM_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    if (bwd_ragged_tile_supported(cols) && aligned16(dx))
        return bwd_ragged_tile_launch(offsets, col_idx, vals, supp_size, dp, loss, loss_scale_dev, loss_scale_host,
                                      dent, dx, dloss, rows, cols, n_items, st);
    cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)rows * cols * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    int64_t need = (rows + 255) / 256;
    int64_t cap = (int64_t)device_sm_count() * 16;
    sparsemax_bwd_ragged_kernel<<<(int)(need < cap ? need : cap), 256, 0, st>>>(
        offsets, col_idx, vals, supp_size, dp, loss, loss_scale_dev, loss_scale_host, dent, dx, dloss, rows, cols, n_items);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}
