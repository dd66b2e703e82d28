/*
 * smlvm.h -- C ABI of the B200-native sparse-marginalization hot path.
 *
 * One shared library (libsmlvm.so, built by `nvcc -gencode
 * arch=compute_100a,code=sm_100a`) exports exactly the entry points below.
 * They are what the reference's FFI for this path binds today through Cython /
 * torch / lpsmap; each declaration cites the reference interface it replaces
 * (paths relative to the reference repo deep-spin/sparse-marginalization-lvm).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *  - all buffers are allocated and owned by the caller (the torch tensors of
 *    the Python side); the library never allocates, never synchronises the
 *    device and never keeps a pointer after the call returns
 *    (same ownership contract as the Cython memoryviews,
 *    lvmhelpers/pbinary_topk.pyx:23-25);
 *  - `stream` is a `cudaStream_t` passed as `void*` (0 = legacy default stream);
 *    launches are asynchronous on it;
 *  - return value: 0 = OK; <0 = argument error (SMLVM_ERR_*); >0 = the
 *    `cudaError_t` of a failed launch.  Nothing throws across the ABI;
 *  - matrices are row-major, contiguous, fp32 unless stated.
 */
#ifndef SMLVM_H_
#define SMLVM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMLVM_ABI_VERSION 5

#define SMLVM_OK 0
#define SMLVM_ERR_ARG (-1)         /* null pointer, negative size, bad enum        */
#define SMLVM_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are built for */
#define SMLVM_ERR_WORKSPACE (-3)   /* caller-provided workspace too small          */

/* limits (checked; SMLVM_ERR_UNSUPPORTED beyond them) */
#define SMLVM_SPARSEMAX_MAX_COLS 8192 /* register path <=1024, shared-memory path beyond */
#define SMLVM_TOPK_MAX_BITS 512       /* = CODE_MAX_SIZE, lvmhelpers/binary_topk.h:13    */
#define SMLVM_TOPK_MAX_K 32
#define SMLVM_SUPPORT_SLOTS 8          /* (value, column) pairs per row the forward hands to the compaction */
#define SMLVM_SPARSEMAP_MAX_BITS 128  /* (n+1)^2 fp64 inverse must fit 227 KB smem       */

/* factor kinds of the SparseMAP solver */
#define SMLVM_FACTOR_BERNOULLI 0       /* lvmhelpers/bernoulli.h  (PFactorBernoulli)      */
#define SMLVM_FACTOR_BUDGET 1          /* lvmhelpers/budget.h     (PFactorBudget)         */
#define SMLVM_FACTOR_SEQUENCE_BINARY 2 /* lvmhelpers/sequence_binary.h (PFactorSequenceBinary) */
#define SMLVM_FACTOR_SEQUENCE 3        /* lvmhelpers/sequence.h   (PFactorSequence: general chain; smlvm_sparsemap_seq_fwd) */

/* per-sample solver status written by smlvm_sparsemap_fwd */
#define SMLVM_SOLVE_CONVERGED 0
#define SMLVM_SOLVE_MAX_ITER 1
#define SMLVM_SOLVE_REPEATED_CONFIG 2
#define SMLVM_SOLVE_SINGULAR 3

int smlvm_version(void);
/* Static description of an error code returned by this library (never NULL). */
const char* smlvm_error_string(int code);

/* ------------------------------------------------------------------------- *
 * (1) batched sparsemax
 * replaces: entmax.sparsemax(X, dim=-1) as called at lvmhelpers/marg.py:51 and
 *           lvmhelpers/structmarg.py:48 (SparsemaxFunction.forward/backward),
 *           fused with what the callers do to its output row by row:
 *           entropy(p) lvmhelpers/marg.py:6-28 and scores.argmax(-1) marg.py:53.
 * ------------------------------------------------------------------------- */

/* x[rows,cols] -> p[rows,cols]; optional per-row outputs (NULL to skip):
 *   supp_size[rows]  int32  entmax's support size k (count of rho*S_rho > cumsum_rho - 1)
 *   nnz[rows]        int32  number of p > 0 (what the compaction will emit)
 *   entropy[rows]    fp32   -sum_{p>0} clamp(p,eps,1-eps) log clamp(p,eps,1-eps)
 *   argmax[rows]     int64  first index of the row maximum of x
 *   support_slots[rows, SMLVM_SUPPORT_SLOTS, 2] fp32  up to 8 (value, column-as-int-bits) pairs of
 *                    the row's support in any order (column -1 = empty; column -2 in slot 0 = the
 *                    row does not fit); hand it to smlvm_compact_fill so that the compaction does
 *                    not re-read p                                                            */
int smlvm_sparsemax_fwd(const float* x, float* p, int32_t* supp_size, int32_t* nnz,
                        float* entropy, int64_t* argmax, float* support_slots, int64_t rows,
                        int32_t cols, void* stream);

/* Jacobian-vector product over the support, fused with the gradient sources the
 * callers feed it.  Effective upstream gradient per element
 *     g = dp + loss_scale * loss + dent[row] * d entropy / d p
 * where each of dp[rows,cols], loss[rows,cols], dent[rows] may be NULL (term absent);
 * loss_scale_dev is a device scalar (fp32) or NULL (=1).
 *     dx = where(p != 0, g - sum_{p != 0} g / supp_size, 0)
 * dloss[rows,cols] (NULL to skip) receives the decoder side of the weighted loss,
 *     dloss = loss_scale * p    (d/dloss of loss_scale * sum(p * loss)).
 * replaces: SparsemaxFunction.backward (entmax) + the autograd of
 *           `encoder_probs.unsqueeze(1).bmm(losses.unsqueeze(-1))` marg.py:125 and of
 *           entropy() marg.py:6-28 w.r.t. p.                                              */
/* support_slots (NULL = read the dense rows): the pairs written by smlvm_sparsemax_fwd for THIS p;
 * with them the kernel reads 64 B per row plus one sector of loss / dp per non-zero and never reads
 * p, loss or dp in full (rows marked "does not fit" are processed densely).  loss_slots (optional,
 * with support_slots): loss already gathered on the support by smlvm_wloss_dense_fwd
 * ([rows, SMLVM_SUPPORT_SLOTS] fp32, same slot order) -- streamed instead of gathered again.        */
int smlvm_sparsemax_bwd(const float* p, const int32_t* supp_size, const float* dp,
                        const float* loss, const float* loss_scale_dev, float loss_scale_host,
                        const float* dent, const float* support_slots, const float* loss_slots,
                        float* dx, float* dloss, int64_t rows, int32_t cols, void* stream);

/* The same Jacobian-vector product when the upstream gradient lives on the SUPPORT only (the
 * decoder ran on the compacted (row, col) pairs): offsets[rows+1], col_idx[N], vals[N] are the
 * compaction of p (smlvm_compact_fill); dp[N], loss[N] ragged (either may be NULL), dent[rows]
 * optional.  dx[rows,cols] is written in full (zero outside the support); dloss[N] (NULL to skip)
 * = loss_scale * vals.  Reads O(N) instead of O(rows*cols).
 * replaces: SparsemaxFunction.backward composed with the boolean-mask indexing backward of
 *           lvmhelpers/structmarg.py:116-126,140 (`p_flat[mask] @ loss`).                          */
/* n_items = entries the ragged arrays hold; offsets beyond it (a compaction that overflowed its capacity) are
 * clamped, so the kernel never reads past the caller's lists.                                              */
int smlvm_sparsemax_bwd_ragged(const int64_t* offsets, const int32_t* col_idx, const float* vals,
                               const int32_t* supp_size, const float* dp, const float* loss,
                               const float* loss_scale_dev, float loss_scale_host, const float* dent,
                               float* dx, float* dloss, int64_t rows, int32_t cols, int64_t n_items,
                               void* stream);

/* The categorical step behind the forward, in ONE launch, driven by the forward's support slots -- or, with
 * support_slots == NULL (rows with wide supports: 32 < cols <= 512, cols % 4 == 0, 16-byte aligned p / loss / dx /
 * dloss), by ONE dense read of p and loss:
 *   weighted loss      row_loss[r] = sum_c p[r,c]*loss[r,c] (NULL to skip), total[0] = scale * sum_r row_loss[r]
 *                      (NULL to skip; fp64, fixed order; workspace: smlvm_reduce_workspace_bytes())
 *   its backward       dx = sparsemax Jacobian of (scale * loss + dent[row] * d entropy / d p), dloss = scale * p
 *                      (dent, dloss may be NULL) -- upstream gradient 1, as after loss.backward()
 *   support compaction (vals, row_idx, col_idx) of p > 0 at offsets[r] + rank in column order, at most `capacity`
 *                      entries written (offsets NULL: skipped; each list may be NULL)
 * One read of the 64 B of slots and one gather of loss per non-zero serve all three.  SMLVM_ERR_UNSUPPORTED when the
 * slot-driven tile kernel does not cover the shape (cols % 4, cols <= 512, 16-byte aligned dx / dloss): use the
 * separate entry points.
 * With support_slots != NULL, `loss` may also point to PINNED HOST memory (cudaHostAlloc, or registered memory on systems
 * where the host pointer is usable on the device: cudaDevAttrCanUseHostPointerForRegisteredMem), reachable through the
 * unified address space: the kernel reads it only on the support, one PCIe read per non-zero, so a caller
 * whose loss matrix lives on the host need not upload it (hotpath.HostPipelinedMargStep(zero_copy_losses=True)).
 * replaces: marg.py:125-126 forward + autograd backward, entropy() backward marg.py:6-28, and the mask indexing of
 *           structmarg.py:116-126.                                                                              */
int smlvm_marg_step_fused(const float* p, const int32_t* supp_size, const float* loss, float scale,
                          const float* dent, const float* support_slots, const int64_t* offsets, float* vals,
                          int32_t* row_idx, int32_t* col_idx, int64_t capacity, float* row_loss, float* total,
                          float* dx, float* dloss, int64_t rows, int32_t cols, void* workspace,
                          size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * (2) support compaction
 * replaces: `mask = p.view(-1) > 0; t[mask]` lvmhelpers/structmarg.py:116-126 and the
 *           column test of lvmhelpers/marg.py:87.  Order = torch.nonzero(p > 0), i.e.
 *           row-major.
 * ------------------------------------------------------------------------- */

/* counts[rows] int32 -> offsets[rows+1] int64 (exclusive prefix sum, offsets[rows] = total).
 * Single pass (decoupled look-back).  workspace: smlvm_scan_workspace_bytes(rows) bytes.   */
size_t smlvm_scan_workspace_bytes(int64_t rows);
int smlvm_scan_counts(const int32_t* counts, int64_t* offsets, int64_t rows, void* workspace,
                      size_t workspace_bytes, void* stream);

/* counts[rows] = number of p[row,:] > 0 (only needed when smlvm_sparsemax_fwd did not
 * already write `nnz`).                                                                   */
int smlvm_compact_count(const float* p, int32_t* counts, int64_t rows, int32_t cols, void* stream);

/* Emits, for every p[r,c] > 0 in row-major order at position offsets[r] + rank:
 *   vals (fp32), row_idx (int32), col_idx (int32); each output may be NULL.
 * capacity = entries each list can hold: positions >= capacity are NOT written (never out of bounds); the caller
 * detects the overflow as offsets[rows] > capacity.
 * support_slots (NULL = scan p): the pairs written by smlvm_sparsemax_fwd; with them the kernel
 * streams 64 B per row instead of the whole row of p (rows that did not fit are scanned).       */
int smlvm_compact_fill(const float* p, const int64_t* offsets, const float* support_slots,
                       float* vals, int32_t* row_idx, int32_t* col_idx, int64_t rows, int32_t cols,
                       int64_t capacity, void* stream);

/* ------------------------------------------------------------------------- *
 * (3a) k-best bit-vectors
 * replaces: lvmhelpers.pbinary_topk.batched_topk(scores, configs, k)
 *           lvmhelpers/pbinary_topk.pyx:23-34 -> topk()/merge_branch()
 *           lvmhelpers/binary_topk.h:57-129.
 * ------------------------------------------------------------------------- */

/* x[rows,n] -> any of (NULL to skip):
 *   configs[rows,k,n] fp32 in {0,1}      (the reference's in-place output)
 *   bits[rows,k,ceil(n/32)] uint32       (bit i of a config = bit i%32 of word i/32)
 *   dp_scores[rows,k] fp32               (the search's own fp32 scores, descending)
 * Requires 2 <= k <= min(2^n, SMLVM_TOPK_MAX_K), 1 <= n <= SMLVM_TOPK_MAX_BITS.            */
int smlvm_topk_bits(const float* x, float* configs, uint32_t* bits, float* dp_scores,
                    int64_t rows, int32_t n, int32_t k, void* stream);

/* scores_k[rows,k] = sum_j bits[r,k,j] * x[r,j]   (fp64 accumulate, fp32 out)
 * replaces: torch.einsum("bkj,bj->bk", bit_vector_z, scores) lvmhelpers/structmarg.py:47    */
int smlvm_bits_score_fwd(const uint32_t* bits, const float* x, float* scores_k, int64_t rows,
                         int32_t n, int32_t k, void* stream);
/* dx[rows,n] = sum_k dscores_k[r,k] * bits[r,k,j]  (autograd of the einsum w.r.t. scores)   */
int smlvm_bits_score_bwd(const uint32_t* bits, const float* dscores_k, float* dx, int64_t rows,
                         int32_t n, int32_t k, void* stream);
/* Sparsemax over the k <= 16 re-scored candidates of every row, with what the callers take from it:
 *   distr[rows,k]  = entmax.sparsemax(scores_k, dim=-1)                       lvmhelpers/structmarg.py:48
 *   supp_size, nnz [rows] int32 (NULL to skip), row_entropy[rows] = -sum_{p>0} p log p (NULL to skip)
 *   entropy_total[0] = scale * sum_r row_entropy[r]  (fp64, fixed order; NULL to skip): the structured entropy
 *                      -(distr[mask] @ log distr[mask]) / B of structmarg.py:51-58 with scale = 1 / B
 * workspace: smlvm_reduce_workspace_bytes().  SMLVM_ERR_UNSUPPORTED for k > 16 (use smlvm_sparsemax_fwd). */
int smlvm_ksparsemax_fwd(const float* scores_k, float* distr, int32_t* supp_size, int32_t* nnz,
                         float* row_entropy, float* entropy_total, float scale, int64_t rows, int32_t k,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the two above and of smlvm_bits_score_fwd in one launch:
 *   g_j  = ddistr[r,j] - dentropy_total[0] * scale * (log p_j + 1)   on the support (either source may be NULL)
 *   ds   = where(p != 0, g - sum_{p != 0} g / supp_size, 0)           SparsemaxFunction.backward over the k candidates
 *   dx[r,c] = sum_j ds_j * bits[r,j,c]                                autograd of the einsum, structmarg.py:47
 * Needs k <= 16, n % 8 == 0 and a 16-byte aligned dx (else SMLVM_ERR_UNSUPPORTED: compose the separate calls). */
int smlvm_topk_sparsemax_bwd(const uint32_t* bits, const float* distr, const int32_t* supp_size,
                             const float* ddistr, const float* dentropy_total, float scale, float* dx,
                             int64_t rows, int32_t n, int32_t k, void* stream);
/* out[n_out,n] fp32 {0,1} = the n bits of packed row index[e] of bits[*,W] (index NULL: row e).
 * replaces: `bit_vector_z.view(-1, n)[mask]` lvmhelpers/structmarg.py:116-126 (with index = the
 * compacted candidate ids of smlvm_compact_fill) without materialising the dense [B,k,n] tensor. */
int smlvm_bits_expand(const uint32_t* bits, const int64_t* index, float* out, int64_t n_out,
                      int32_t n, void* stream);
/* out[e] = src[idx[e]] for n_out rows of row_bytes bytes (a multiple of 4; 16-byte rows and pointers
 * take 128-bit accesses).  replaces: `x.unsqueeze(1).repeat(1,k,1).view(B*k,-1)[mask]`
 * lvmhelpers/structmarg.py:95-124 (idx = the compacted row ids of smlvm_compact_fill) and
 * `encoder_input[idxs]`, `decoder_input[idxs]`, `labels[idxs]` structmarg.py:236-240.               */
int smlvm_gather_rows(const void* src, const int32_t* idx, void* out, int64_t n_out, int64_t row_bytes,
                      void* stream);

/* ------------------------------------------------------------------------- *
 * (3b) SparseMAP active-set solver, one CTA per sample
 * replaces: SparseMAP.run_sparsemap lvmhelpers/sparsemap.py:13-36 (factor graph
 *           construction, init_active_set_from_scores, solve_qp, get_sparse_solution)
 *           for PFactorBernoulli / PFactorBudget / PFactorSequenceBinary
 *           (lvmhelpers/pbernoulli.pyx:10-22, pbudget.pyx:10-22, psequence.pyx:25-37),
 *           i.e. libad3 GenericFactor::SolveQP with the MAP oracles of
 *           lvmhelpers/bernoulli.h:14-35, budget.h:24-85, sequence.h:74-151.
 * ------------------------------------------------------------------------- */

/* Capacity (vertices per sample) the padded outputs must have: n + 1.                     */
int32_t smlvm_sparsemap_capacity(int32_t n);

/* How smlvm_sparsemap_fwd will run for n bits of factor `kind`: the solver keeps a sample's (|A|+1)^2 fp64 inverse in
 * shared memory and works through up to 3 capacity tiers (largest active set a launch can hold;
 * the last is n + 1); a sample that outgrows a tier is re-solved by the next.  Writes, per tier,
 * the capacity, the dynamic shared memory per CTA and the resident CTAs (= samples in flight) per
 * SM; each output may be NULL, arrays need 3 entries.  Returns the number of tiers.            */
int32_t smlvm_sparsemap_plan(int32_t n, int32_t kind, int32_t* caps, int32_t* smem_bytes,
                             int32_t* ctas_per_sm);

/* x[rows,n] on-scores (off-scores are 0, sparsemap.py:22);
 * t[rows,n-1] 1->1 transition scores (sequence-binary only, else NULL);
 * init_scores[rows,2n] fp64: the uniform draws of sparsemap.py:26 (NULL = init=False).
 * Padded outputs, cap = smlvm_sparsemap_capacity(n):
 *   p_pad[rows,cap]            fp32   distribution (float(double), as torch.tensor(p))
 *   aset_bits[rows,cap,W]      uint32 W = ceil(n/32); active vertices in solver order
 *   counts[rows]               int32  |A|
 *   kept[rows]                 int32  number of p_pad > 0 (NULL to skip)
 *   iters[rows], status[rows]  int32  (NULL to skip)
 * The inverse needed by smlvm_sparsemap_bwd, S = inverse_A[1:,1:] (|A|^2 doubles per sample, row stride |A|; the
 * reference keeps the whole factor object alive for this, sparsemap.py:19,42), is bump-allocated from a caller-owned
 * arena (NULL to skip):
 *   S_arena[S_capacity]        fp64
 *   S_offsets[rows]            int64  element offset of sample r's S in the arena, or -1 if it did not fit
 *   S_cursor[1]                uint64 device counter, ZEROED BY THE CALLER before the (first) call that shares the
 *                                     arena; ends at the number of doubles all samples needed -- > S_capacity means
 *                                     some samples got offset -1 (re-solve with an arena of that size).  Padded to
 *                                     (n+1)^2 per sample the arena cannot overflow.
 * workspace: at least smlvm_sparsemap_workspace_bytes(rows) bytes (worklists of the capacity tiers); rows < 2^31.
 *            Bytes beyond that are used as the pool of hand-off records: a sample whose active set outgrows a tier
 *            is resumed by the next tier from its state (~8 (|A|+1)^2 bytes) when the pool has room, re-solved from
 *            scratch otherwise -- same results either way.                                                          */
size_t smlvm_sparsemap_workspace_bytes(int64_t rows);
int smlvm_sparsemap_fwd(const float* x, const float* t, const double* init_scores, int32_t kind,
                        int32_t budget, int32_t max_iter, int64_t rows, int32_t n, float* p_pad,
                        uint32_t* aset_bits, int32_t* counts, int32_t* kept, int32_t* iters,
                        int32_t* status, double* S_arena, int64_t S_capacity, int64_t* S_offsets,
                        uint64_t* S_cursor, void* workspace, size_t workspace_bytes, void* stream);

/* The general linear chain, FactorSequence of lvmhelpers/sequence.h:25-291 (bound as PFactorSequence.initialize(
 * list[int]), lvmhelpers/psequence.pyx:10-22): `length` positions, position i with S_i <= 32 states, one variable per
 * (position, state) -- n = sum S_i <= 128 -- and num_edges = S_0 + sum S_{i-1} S_i + S_{L-1} additional scores
 * (start -> first, every transition row-major in (previous, current), last -> stop; sequence.h:264-291).
 *   x[rows,n] node scores, edge_scores[rows,num_edges] (NULL: zeros),
 *   layout[2 length + 3] int32 (device) = offsets of the states of position 0..length, then the first edge of
 *   position 0..length+1.
 * A vertex is the one-hot pattern of its state sequence: bit layout[i] + state_i set for every position i; the
 * outputs have the layout of smlvm_sparsemap_fwd (aset_bits over the n variables).  MAP oracle = Viterbi with the
 * reference's tie rule (lower predecessor wins, sequence.h:112,136); Gram entries = equal positions (:212-226).   */
int smlvm_sparsemap_seq_fwd(const float* x, const float* edge_scores, const int32_t* layout, int32_t length,
                            int32_t num_edges, int32_t max_iter, int64_t rows, int32_t n, float* p_pad,
                            uint32_t* aset_bits, int32_t* counts, int32_t* kept, int32_t* iters, int32_t* status,
                            double* S_arena, int64_t S_capacity, int64_t* S_offsets, uint64_t* S_cursor,
                            void* workspace, size_t workspace_bytes, void* stream);
/* its backward: dx[rows,n] = M (S dp), dedge[rows,num_edges] = N (S dp) (NULL to skip)                            */
int smlvm_sparsemap_seq_bwd(const float* dp, const int64_t* offsets, int32_t only_positive, const float* p_pad,
                            const uint32_t* aset_bits, const int32_t* counts, const double* S_arena,
                            const int64_t* S_offsets, const int32_t* layout, int32_t length, int32_t num_edges,
                            float* dx, float* dedge, int64_t rows, int32_t n, void* stream);

/* Ragged expansion in sample order (structmarg.py:181-198): for every kept vertex
 * (all of them if only_positive == 0, those with p_pad > 0 otherwise) at
 * offsets[r] + rank:  p_out (fp32), aset_out[.,n] (fp32 {0,1}), idx_out (int32 sample id). */
int smlvm_sparsemap_expand(const float* p_pad, const uint32_t* aset_bits, const int32_t* counts,
                           const int64_t* offsets, int32_t only_positive, float* p_out,
                           float* aset_out, int32_t* idx_out, int64_t rows, int32_t n,
                           void* stream);

/* Backward: dx[r,:] = (M S dp)[:n], dt[r,:] = N S dp (sequence-binary; else NULL).
 * dp is ragged in the same layout smlvm_sparsemap_expand produced (same offsets,
 * same only_positive); dropped vertices get dp = 0.
 * replaces: SparseMAP.jv / dist_jacobian_vec lvmhelpers/sparsemap.py:38-44,80-87          */
/* S of sample r is read at S_arena + S_offsets[r] (what smlvm_sparsemap_fwd wrote); a sample whose offset is -1
 * gets NaN gradients (its inverse did not fit the forward's arena: the caller should have re-solved).             */
int smlvm_sparsemap_bwd(const float* dp, const int64_t* offsets, int32_t only_positive,
                        const float* p_pad, const uint32_t* aset_bits, const int32_t* counts,
                        const double* S_arena, const int64_t* S_offsets, float* dx, float* dt, int64_t rows,
                        int32_t n, void* stream);

/* ------------------------------------------------------------------------- *
 * (4) probability-weighted loss, fused segmented reduce
 * replaces: marg.py:125 (bmm), structmarg.py:140,242 ((p @ loss) / B) and the
 *           structured entropy -p.log p / B structmarg.py:51-58,200-202.
 * ------------------------------------------------------------------------- */

/* Dense: row_loss[r] = sum_c p[r,c]*loss[r,c] (NULL to skip); total[0] (fp32 device scalar,
 * NULL to skip) = scale * sum_r row_loss[r], accumulated in fp64 in a fixed order.
 * workspace: smlvm_reduce_workspace_bytes() bytes.  support_slots (NULL = read p and loss in
 * full): the forward's (value, column) pairs; only loss entries on the support are then read, and
 * written to loss_slots[rows, SMLVM_SUPPORT_SLOTS] (NULL to skip) for smlvm_sparsemax_bwd.          */
size_t smlvm_reduce_workspace_bytes(void);
int smlvm_wloss_dense_fwd(const float* p, const float* loss, const float* support_slots,
                          float* loss_slots, float* row_loss, float* total, float scale, int64_t rows,
                          int32_t cols, void* workspace, size_t workspace_bytes, void* stream);

/* Ragged: offsets[rows+1] segments over p[N], loss[N].
 *   row_loss[r] = sum_{j in seg r} p_j*loss_j      (NULL to skip; offsets may then be NULL)
 *   total[0]    = scale * sum_j p_j*loss_j          (NULL to skip)
 *   neg_plogp[0]= -scale * sum_j p_j*log(p_j)        (NULL to skip; structured entropy)     */
int smlvm_wloss_ragged_fwd(const float* p, const float* loss, const int64_t* offsets,
                           float* row_loss, float* total, float* neg_plogp, float scale,
                           int64_t rows, int64_t n_items, void* workspace,
                           size_t workspace_bytes, void* stream);

/* dp_j = g*scale*loss_j - gent*scale*(log p_j + 1),  dloss_j = g*scale*p_j;
 * g, gent: fp32 device scalars (upstream gradients of total / neg_plogp; NULL = 0).
 * dp / dloss may be NULL.                                                                  */
int smlvm_wloss_ragged_bwd(const float* p, const float* loss, const float* g, const float* gent,
                           float scale, float* dp, float* dloss, int64_t n_items, void* stream);

/* Segmented log-sum-exp over a ragged batch: out[r] = log sum_{j in [offsets[r], offsets[r+1])}
 * exp(sign * v[j] + shift); -inf for an empty segment.
 * replaces: the per-sample loop `torch.logsumexp(logp_x_given_z[torch.tensor(idxs) == k] + logp_z)`
 *           of compute_importance_sampling, experiments/bit_vector-vae/train.py:320-330 (an
 *           O(B * N) host loop), and the masked_scatter + dense logsumexp of :331-345, with
 *           v = loss_output, sign = -1, shift = logp_z.                                         */
int smlvm_segment_logsumexp(const float* v, const int64_t* offsets, float sign, float shift,
                            float* out, int64_t rows, void* stream);

/* ------------------------------------------------------------------------- *
 * (f)4 the remaining `entmax` call sites
 * ------------------------------------------------------------------------- */
/* p[rows,cols] = entmax.entmax15(x, dim=-1): X' = (x - max x) / 2, p = [(X' - tau)_+]^2 with sum p = 1 -- the
 * reference's DEFAULT normalizer, lvmhelpers/marg.py:39,46.  entropy[rows] = entropy(p) of marg.py:6-28 and
 * argmax[rows] = x.argmax(-1) of marg.py:53 (either may be NULL).  cols <= 1024, else SMLVM_ERR_UNSUPPORTED.    */
int smlvm_entmax15_fwd(const float* x, float* p, float* entropy, int64_t* argmax, int64_t rows, int32_t cols,
                       void* stream);
/* Entmax15Function.backward: g = sqrt(p); dx = dp * g - (sum(dp * g) / sum(g)) * g                              */
int smlvm_entmax15_bwd(const float* p, const float* dp, float* dx, int64_t rows, int32_t cols, void* stream);
/* Fenchel-Young loss of entmax.SparsemaxLoss (alpha15 = 0) / Entmax15Loss (alpha15 = 1),
 * experiments/semi_supervised-vae/train.py:9,312-325, given p = normalizer(x):
 *   loss[r] = (1 - sum_j p^alpha) / (alpha (alpha - 1)) + <p - e_target, x>;   p_minus_y[rows,cols] = p - e_target
 * (the gradient w.r.t. x; NULL to skip).  target int64 [rows] in [0, cols).                                     */
int smlvm_fy_loss_fwd(const float* x, const float* p, const int64_t* target, int32_t alpha15, float* loss,
                      float* p_minus_y, int64_t rows, int32_t cols, void* stream);
/* out[r,:] = scale[r] * a[r,:]  (the loss's backward: grad_output[r] * (p - e_target))                          */
int smlvm_row_scale(const float* a, const float* scale, float* out, int64_t rows, int32_t cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMLVM_H_ */
