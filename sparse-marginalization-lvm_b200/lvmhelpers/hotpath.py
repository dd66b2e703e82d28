"""Pre-allocated pipelines over the C-ABI kernels for steady-state use (training loops,
``bench.py``): no allocation, no host synchronisation and a fixed launch sequence per step,
so a step can be replayed back to back or captured in a CUDA graph.

``SparsemaxMargStep`` is one pass of the categorical hot path over a batch
(lvmhelpers/marg.py:49-54,77-132 without the caller's decoder):

    scores [B,K]  --sparsemax fwd-->  P, support size, nnz, entropy, argmax      (kernel 1)
    nnz           --scan----------->  ragged offsets                              (kernel 2a)
    P, offsets    --compact fill--->  (value, row, col) of every p > 0            (kernel 2b)
    P, losses     --weighted loss-->  per-row loss, mean loss                     (kernel 4)
    P, losses     --fused bwd------>  dscores (sparsemax Jacobian of g*L/B + entropy term),
                                      dlosses = g*P/B                             (kernel 1')
"""
import torch

from . import _native as N
from ._native import lib, check, ptr


class SparsemaxMargStep:
    KERNELS = ("sparsemax_fwd", "scan_counts", "compact_fill", "wloss_dense_fwd", "sparsemax_bwd")

    def __init__(self, rows, cols, device, capacity, entropy_coeff=0.0):
        self.rows, self.cols, self.device = rows, cols, device
        self.capacity = int(capacity)
        self.entropy_coeff = float(entropy_coeff)
        f32, i32, i64 = torch.float32, torch.int32, torch.int64
        e = lambda *s, dtype=f32: torch.empty(*s, dtype=dtype, device=device)
        self.p = e(rows, cols)
        self.supp, self.nnz = e(rows, dtype=i32), e(rows, dtype=i32)
        self.ent, self.amax = e(rows), e(rows, dtype=i64)
        self.offsets = e(rows + 1, dtype=i64)
        self.vals, self.ri, self.ci = e(self.capacity), e(self.capacity, dtype=i32), e(self.capacity, dtype=i32)
        self.row_loss, self.total = e(rows), e(())
        self.dx, self.dloss = e(rows, cols), e(rows, cols)
        self.dent = torch.full((rows,), -self.entropy_coeff / rows, dtype=f32, device=device) \
            if self.entropy_coeff != 0.0 else None
        self.scan_ws = torch.zeros(max(lib.smlvm_scan_workspace_bytes(rows), 64), dtype=torch.uint8, device=device)
        self.red_ws = torch.zeros(lib.smlvm_reduce_workspace_bytes(), dtype=torch.uint8, device=device)

    @staticmethod
    def size_capacity(x):
        """One sizing pass (with a host read) to dimension the ragged outputs for fixed inputs."""
        from . import ops
        out = ops.sparsemax_fwd(x)
        return int(out["nnz"].sum().item())

    def run(self, x, losses, rec=None):
        """Launch the five kernels on the current stream.  ``rec(name)`` (optional) is called
        before and after each launch so that a caller can bracket them with CUDA events."""
        st = torch.cuda.current_stream().cuda_stream
        rows, K = self.rows, self.cols
        mark = rec if rec is not None else (lambda name: None)
        mark("sparsemax_fwd")
        check(lib.smlvm_sparsemax_fwd(ptr(x), ptr(self.p), ptr(self.supp), ptr(self.nnz), ptr(self.ent),
                                      ptr(self.amax), rows, K, st), "smlvm_sparsemax_fwd")
        mark("scan_counts")
        check(lib.smlvm_scan_counts(ptr(self.nnz), ptr(self.offsets), rows, ptr(self.scan_ws),
                                    self.scan_ws.numel(), st), "smlvm_scan_counts")
        mark("compact_fill")
        check(lib.smlvm_compact_fill(ptr(self.p), ptr(self.offsets), ptr(self.vals), ptr(self.ri),
                                     ptr(self.ci), rows, K, st), "smlvm_compact_fill")
        mark("wloss_dense_fwd")
        check(lib.smlvm_wloss_dense_fwd(ptr(self.p), ptr(losses), ptr(self.row_loss), ptr(self.total),
                                        1.0 / rows, rows, K, ptr(self.red_ws), self.red_ws.numel(), st),
              "smlvm_wloss_dense_fwd")
        mark("sparsemax_bwd")
        check(lib.smlvm_sparsemax_bwd(ptr(self.p), ptr(self.supp), None, ptr(losses), None, 1.0 / rows,
                                      ptr(self.dent), ptr(self.dx), ptr(self.dloss), rows, K, st),
              "smlvm_sparsemax_bwd")
        mark(None)
        return self.total

    def algorithmic_bytes(self, total_nnz):
        """Bytes each kernel must move per launch (fp32 = 4 B)."""
        r, K, N_ = self.rows, self.cols, total_nnz
        return {
            "sparsemax_fwd": r * (8 * K + 4 + 4 + 4 + 8),       # X in, P out, supp/nnz/entropy/argmax
            "scan_counts": r * (4 + 8),                         # counts in, offsets out
            "compact_fill": r * (4 * K + 8) + 12 * N_,          # P + offsets in, (val,row,col) out
            "wloss_dense_fwd": r * (8 * K + 4),                 # P, L in, row loss out
            "sparsemax_bwd": r * (8 * K + 4 + 8 * K) + (4 * r if self.dent is not None else 0),
        }
