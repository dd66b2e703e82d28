// k-sparsemax of the top-k path for sm_100a: the sparsemax over the k re-scored candidates of every row with the
// structured entropy (forward), and its backward fused with the autograd of the re-scoring einsum.
//
// replaces, per batch (lvmhelpers/structmarg.py:47-58):
//   distr = entmax.sparsemax(scores_k, dim=-1)                      [B, k]
//   entropy = -(distr[distr > 0] @ log(distr[distr > 0])) / B       scalar
// and, backward, SparsemaxFunction.backward composed with the gradient of the boolean-mask dot product and of
// torch.einsum("bkj,bj->bk", bit_vector_z, scores) w.r.t. scores: ONE launch that reads [B, k] twice and the packed
// candidates once and writes dscores [B, n].  Round 1 ran this as 12 launches around a support compaction with a host
// read (profiles/r01_topk_path_launches_v3.csv); the compaction is now only built when TopKSparsemaxMarg asks for it.
#include "common.cuh"
#include "reduce_ws.cuh"
#include "sparsemax_list.cuh"

namespace smlvm {

constexpr int kKsMax = 16;   // candidates per row this path holds in registers

// A thread owns a row of k <= 16 scores: the reference's definition on the sorted list (list_threshold: fp64 cumulative
// sums in sorted order, fl32(r * S_r) > fl32(fl32(cumsum_r) - 1), tau = C_k / k), P = max(x' - tau, 0).
__global__ void __launch_bounds__(256)
ksparsemax_fwd_kernel(const float* __restrict__ sk, float* __restrict__ distr, int32_t* __restrict__ supp,
                      int32_t* __restrict__ nnz, float* __restrict__ row_ent, float* __restrict__ ent_total, float scale,
                      int64_t rows, int k, ReduceWs ws)
{
    __shared__ float lv_all[8][kKsMax * 32];
    __shared__ double red[9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* lv = lv_all[warp] + lane;
    const int64_t ngroups = (rows + 31) / 32, stride = (int64_t)gridDim.x * 8;
    double acc = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * 8 + warp; t < ngroups; t += stride) {
        const int64_t r = t * 32 + lane;
        const bool live = r < rows;
        float s[kKsMax];
        float m = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < kKsMax; ++j) {
            s[j] = (live && j < k) ? __ldg(sk + r * k + j) : -CUDART_INF_F;
            m = fmaxf(m, s[j]);
        }
#pragma unroll
        for (int j = 0; j < kKsMax; ++j) {
            s[j] = __fsub_rn(s[j], m);
            if (j < k) lv[j * 32] = s[j];
        }
        int kk;
        float tau;
        list_threshold<kKsMax>(lv, live ? k : 0, kk, tau);   // (warp-wide vote inside: every lane calls it)
        if (live) {
            int nz = 0;
            double h = 0.0;
#pragma unroll
            for (int j = 0; j < kKsMax; ++j) {
                if (j < k) {
                    const float pe = fmaxf(__fsub_rn(s[j], tau), 0.f);
                    distr[r * k + j] = pe;
                    if (pe > 0.f) {
                        ++nz;
                        h += (double)(pe * logf(pe));
                    }
                }
            }
            if (supp != nullptr) supp[r] = kk;
            if (nnz != nullptr) nnz[r] = nz;
            if (row_ent != nullptr) row_ent[r] = (float)(-h);
            acc -= h;
        }
    }
    if (ent_total != nullptr) finish_totals(ws, acc, 0.0, scale, ent_total, nullptr, red);
}

// dscores[r, c] = sum_j ds_j bit(r, j, c) with ds = sparsemax Jacobian (over the k candidates) applied to
// g_j = ddistr_j - gent * scale * (log p_j + 1) on the support.  A thread per CPT columns of one row; the 2k + 1 floats
// and the k configuration words of a row are L1 hits for the n / CPT threads that share them.
template <int CPT>
__global__ void __launch_bounds__(256)
topk_sparsemax_bwd_kernel(const uint32_t* __restrict__ bits, const float* __restrict__ distr,
                          const int32_t* __restrict__ supp, const float* __restrict__ ddistr,
                          const float* __restrict__ gent, float ent_scale, float* __restrict__ dx, int64_t rows, int n, int k)
{
    const int words_n = (n + 31) >> 5, qn = n / CPT;
    const float ge = gent != nullptr ? __ldg(gent) * ent_scale : 0.f;
    const int64_t total = rows * qn, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int64_t r = div_idx(t, qn);
        const int c = (int)(t - r * qn) * CPT;
        float ds[kKsMax];
        unsigned on = 0u;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kKsMax; ++j) {
            ds[j] = 0.f;
            if (j < k) {
                const float p = __ldg(distr + r * k + j);
                float g = ddistr != nullptr ? __ldg(ddistr + r * k + j) : 0.f;
                if (p != 0.f) {
                    if (gent != nullptr && p > 0.f) g -= ge * (logf(p) + 1.f);
                    ds[j] = g;
                    sum += g;
                    on |= 1u << j;
                }
            }
        }
        const float vhat = __fdiv_rn(sum, (float)__ldg(supp + r));
#pragma unroll
        for (int j = 0; j < kKsMax; ++j) ds[j] = ((on >> j) & 1u) ? __fsub_rn(ds[j], vhat) : 0.f;
        const uint32_t* b = bits + r * (int64_t)k * words_n + (c >> 5);
        const int sh = c & 31;
        float a[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) a[e] = 0.f;
        // candidates outside the k-sparsemax support have ds = 0 (typically all but the first three or four of a row):
        // a candidate rank that is zero in every row of the warp is skipped
        const unsigned act = __activemask();
#pragma unroll
        for (int j = 0; j < kKsMax; ++j) {
            if (j < k && __any_sync(act, ds[j] != 0.f)) {
                const uint32_t w = __ldg(b + j * words_n) >> sh;
                const float dj = ds[j];
#pragma unroll
                for (int e = 0; e < CPT; ++e) a[e] += ((w >> e) & 1u) ? dj : 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < CPT; e += 4)
            stg_stream4(dx + t * CPT + e, make_float4(a[e], a[e + 1], a[e + 2], a[e + 3]));
    }
}

}  // namespace smlvm

using namespace smlvm;

extern "C" int smlvm_ksparsemax_fwd(const float* scores_k, float* distr, int32_t* supp_size, int32_t* nnz,
                                    float* row_entropy, float* entropy_total, float scale, int64_t rows, int32_t k,
                                    void* workspace, size_t workspace_bytes, void* stream)
{
    if (rows < 0 || k < 1) return SMLVM_ERR_ARG;
    if (k > kKsMax) return SMLVM_ERR_UNSUPPORTED;
    if (rows == 0) return SMLVM_OK;
    if (scores_k == nullptr || distr == nullptr) return SMLVM_ERR_ARG;
    if (entropy_total != nullptr && (workspace == nullptr || workspace_bytes < smlvm_reduce_workspace_bytes()))
        return SMLVM_ERR_WORKSPACE;
    const int64_t need = (rows + 255) / 256;
    int64_t cap = (int64_t)device_sm_count() * 8;
    if (cap > kMaxCtas) cap = kMaxCtas;
    ksparsemax_fwd_kernel<<<(int)(need < cap ? need : cap), 256, 0, as_stream(stream)>>>(
        scores_k, distr, supp_size, nnz, row_entropy, entropy_total, scale, rows, k, ws_view(workspace));
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}

extern "C" int smlvm_topk_sparsemax_bwd(const uint32_t* bits, const float* distr, const int32_t* supp_size,
                                        const float* ddistr, const float* dentropy_total, float scale, float* dx,
                                        int64_t rows, int32_t n, int32_t k, void* stream)
{
    if (rows < 0 || n < 1 || k < 1) return SMLVM_ERR_ARG;
    if (k > kKsMax || (n & 7) != 0 || (reinterpret_cast<uintptr_t>(dx) & 15u) != 0) return SMLVM_ERR_UNSUPPORTED;
    if (rows == 0) return SMLVM_OK;
    if (bits == nullptr || distr == nullptr || supp_size == nullptr || dx == nullptr) return SMLVM_ERR_ARG;
    const int cpt = (n & 31) == 0 ? 32 : ((n & 15) == 0 ? 16 : 8);
    const int64_t need = (rows * (n / cpt) + 255) / 256, cap = (int64_t)device_sm_count() * 16;
    const int grid = (int)(need < cap ? need : cap);
    cudaStream_t st = as_stream(stream);
    if (cpt == 32)
        topk_sparsemax_bwd_kernel<32><<<grid, 256, 0, st>>>(bits, distr, supp_size, ddistr, dentropy_total, scale, dx, rows, n, k);
    else if (cpt == 16)
        topk_sparsemax_bwd_kernel<16><<<grid, 256, 0, st>>>(bits, distr, supp_size, ddistr, dentropy_total, scale, dx, rows, n, k);
    else
        topk_sparsemax_bwd_kernel<8><<<grid, 256, 0, st>>>(bits, distr, supp_size, ddistr, dentropy_total, scale, dx, rows, n, k);
    SMLVM_LAUNCH_CHECK();
    return SMLVM_OK;
}
