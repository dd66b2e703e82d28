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

Kernels 2b, 4 and 1' are ONE launch (``smlvm_marg_step_fused``): fwd -> scan -> {loss, lists, gradients} -- driven by
the forward's support slots for large batches of peaked rows, by one dense read of P and L otherwise (32 < K <= 512).

Capacity contract: the ragged lists are pre-allocated for ``capacity`` entries.  The fill never writes past them
(entries beyond the capacity are dropped) and the ragged consumers never read past them; a batch whose support is
larger than the capacity leaves ``offsets[-1] > capacity`` -- ``overflowed()`` reads that one scalar.
"""
import torch

from . import _native as N
from ._native import lib, check, ptr


OVERLAP_MIN_ROWS = 65536


class SparsemaxMargStep:
    KERNELS = ("sparsemax_fwd", "scan_counts", "compact_fill", "wloss_dense_fwd", "sparsemax_bwd")

    def __init__(self, rows, cols, device, capacity, entropy_coeff=0.0, global_rows=None, use_slots=None,
                 overlap_compaction=True, fused=None):
        """``global_rows``: batch size B of the 1/B in the loss and entropy terms when this step
        owns only a shard / micro-batch of the batch (default: ``rows``).
        ``use_slots``: let the weighted loss and the backward read the losses only at the support
        entries the forward handed over (64 B of slots + one sector per non-zero instead of the rows
        of P and L), and the compaction read the slots instead of P.  Default: on for large batches
        (>= 4096 rows) whose expected support (``capacity / rows``) is at most 6 entries per row -- rows
        that do not fit the 8 slots fall back to dense processing one by one inside a warp, which only
        pays when they are rare and there are many warps.
        ``fused``: loss, lists and gradients in one launch (``smlvm_marg_step_fused``); default: whenever the
        slots are on and the shape is covered.  ``False`` keeps the separate kernels (compaction on a side stream).
        """
        if use_slots is None:
            use_slots = capacity <= 6 * rows and rows >= 4096
        self.use_slots = bool(use_slots)
        self.rows, self.cols, self.device = rows, cols, device
        self.global_rows = int(global_rows) if global_rows is not None else rows
        self.capacity = int(capacity)
        self.entropy_coeff = float(entropy_coeff)
        f32, i32, i64 = torch.float32, torch.int32, torch.int64
        e = lambda *s, dtype=f32: torch.empty(*s, dtype=dtype, device=device)
        self.p = e(rows, cols)
        self.supp, self.nnz = e(rows, dtype=i32), e(rows, dtype=i32)
        self.ent, self.amax = e(rows), e(rows, dtype=i64)
        self.slots = e(rows, N.SUPPORT_SLOTS, 2)   # support (value, column) pairs: fwd -> compaction, loss, bwd
        self.loss_slots = e(rows, N.SUPPORT_SLOTS)  # losses gathered on the support: weighted loss -> bwd
        self.offsets = e(rows + 1, dtype=i64)
        self.vals, self.ri, self.ci = e(self.capacity), e(self.capacity, dtype=i32), e(self.capacity, dtype=i32)
        self.row_loss, self.total = e(rows), e(())
        self.dx, self.dloss = e(rows, cols), e(rows, cols)
        self.dent = torch.full((rows,), -self.entropy_coeff / self.global_rows, dtype=f32, device=device) \
            if self.entropy_coeff != 0.0 else None
        self.scan_ws = torch.zeros(max(lib.smlvm_scan_workspace_bytes(rows), 64), dtype=torch.uint8, device=device)
        self.red_ws = torch.zeros(lib.smlvm_reduce_workspace_bytes(), dtype=torch.uint8, device=device)
        # side stream of the compaction branch (run()); small batches are launch-bound: one stream
        self.side = torch.cuda.Stream(device=device) if (overlap_compaction and rows >= OVERLAP_MIN_ROWS) else None
        self._ev_fwd, self._ev_side = torch.cuda.Event(), torch.cuda.Event()
        # loss, lists and gradients in one launch behind the scan: from the support slots (peaked rows) or from one
        # dense read of P and L (wide supports, 32 < K <= 512: smlvm_marg_step_fused with slots = NULL)
        covered = cols % 4 == 0 and cols <= 512 and (self.use_slots or cols > 32)
        self.fused = covered and (fused is None or bool(fused))
        self.kernels = ("sparsemax_fwd", "scan_counts", "marg_step_fused") if self.fused else self.KERNELS

    def overflowed(self):
        """True when the last batch had more non-zeros than ``capacity`` (the lists then hold the first ``capacity``
        entries only; loss and gradients are complete -- they do not depend on the lists).  One host read."""
        return int(self.offsets[-1].item()) > self.capacity

    @staticmethod
    def size_capacity(x):
        """One sizing pass (with a host read) to dimension the ragged outputs for fixed inputs."""
        from . import ops
        out = ops.sparsemax_fwd(x)
        return int(out["nnz"].sum().item())

    def run(self, x, losses, rec=None):
        """Launch the five kernels.  The compaction branch (scan -> fill) depends only on the forward,
        like the loss branch (weighted loss -> backward): for large batches it goes to a side stream
        between two events, so the latency-bound scan and the fill run under the HBM-bound loss /
        backward kernels.  ``rec(name)`` (optional) is called before and after each launch so that a
        caller can bracket the kernels with CUDA events; the launches are then serial on the current
        stream."""
        N.require_cuda(x, "scores")      # also: on the current device (the launches go to its stream)
        # ``losses`` may also be a PINNED HOST tensor when the step is slot-driven and fused: the kernel then reads the
        # loss matrix only on the support (one PCIe read per non-zero, through the unified address space) instead of
        # having all of it uploaded first -- the host-buffer pipeline's ``zero_copy_losses``
        mapped = (not losses.is_cuda) and losses.is_pinned() and self.fused and self.use_slots
        if x.device != self.p.device or (losses.device != self.p.device and not mapped):
            raise RuntimeError("SparsemaxMargStep was built for %s, got inputs on %s / %s%s" % (
                self.p.device, x.device, losses.device,
                "" if losses.is_cuda else " (host losses are read in place only from PINNED memory and only by the "
                                          "slot-driven fused step: pinned=%s, fused=%s, slots=%s)"
                                          % (losses.is_pinned(), self.fused, self.use_slots)))
        fork = rec is None and self.side is not None and not self.fused
        cur = torch.cuda.current_stream() if fork else None
        st = cur.cuda_stream if fork else N.stream()
        rows, K = self.rows, self.cols
        mark = rec if rec is not None else (lambda name: None)
        mark("sparsemax_fwd")
        slots = ptr(self.slots) if self.use_slots else None
        check(lib.smlvm_sparsemax_fwd(ptr(x), ptr(self.p), ptr(self.supp), ptr(self.nnz), ptr(self.ent),
                                      ptr(self.amax), slots, rows, K, st), "smlvm_sparsemax_fwd")
        if self.fused:
            mark("scan_counts")
            check(lib.smlvm_scan_counts(ptr(self.nnz), ptr(self.offsets), rows, ptr(self.scan_ws),
                                        self.scan_ws.numel(), st), "smlvm_scan_counts")
            mark("marg_step_fused")
            check(lib.smlvm_marg_step_fused(ptr(self.p), ptr(self.supp), ptr(losses), 1.0 / self.global_rows,
                                            ptr(self.dent), slots, ptr(self.offsets), ptr(self.vals), ptr(self.ri),
                                            ptr(self.ci), self.capacity, ptr(self.row_loss), ptr(self.total),
                                            ptr(self.dx), ptr(self.dloss), rows, K, ptr(self.red_ws),
                                            self.red_ws.numel(), st), "smlvm_marg_step_fused")
            mark(None)
            return self.total
        st_c = st
        if fork:
            self._ev_fwd.record(cur)
            self.side.wait_event(self._ev_fwd)
            st_c = self.side.cuda_stream
        mark("scan_counts")
        check(lib.smlvm_scan_counts(ptr(self.nnz), ptr(self.offsets), rows, ptr(self.scan_ws),
                                    self.scan_ws.numel(), st_c), "smlvm_scan_counts")
        mark("compact_fill")
        check(lib.smlvm_compact_fill(ptr(self.p), ptr(self.offsets), slots, ptr(self.vals), ptr(self.ri),
                                     ptr(self.ci), rows, K, self.capacity, st_c), "smlvm_compact_fill")
        if fork:
            self._ev_side.record(self.side)
        mark("wloss_dense_fwd")
        lslots = ptr(self.loss_slots) if self.use_slots else None
        check(lib.smlvm_wloss_dense_fwd(ptr(self.p), ptr(losses), slots, lslots, ptr(self.row_loss), ptr(self.total),
                                        1.0 / self.global_rows, rows, K, ptr(self.red_ws), self.red_ws.numel(), st),
              "smlvm_wloss_dense_fwd")
        mark("sparsemax_bwd")
        check(lib.smlvm_sparsemax_bwd(ptr(self.p), ptr(self.supp), None, ptr(losses), None, 1.0 / self.global_rows,
                                      ptr(self.dent), slots, lslots, ptr(self.dx), ptr(self.dloss), rows, K, st),
              "smlvm_sparsemax_bwd")
        if fork:
            cur.wait_event(self._ev_side)
        mark(None)
        return self.total

    def capture(self, x, losses):
        """Record the step (five kernels + the scan's workspace memset) into a CUDA graph bound to
        the buffers ``x`` / ``losses`` and return the graph: ``g.replay()`` re-runs the step with
        ONE launch.  For the launch-bound experiment shapes (B = 64) the step is ~5 launches of
        ~2-3 us of work each; the graph removes the per-launch host cost."""
        self.run(x, losses)                      # warm-up: one-time function attributes, allocations
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.run(x, losses)
        return graph

    def algorithmic_bytes(self, total_nnz):
        """Bytes each kernel must move per launch (fp32 = 4 B)."""
        r, K, N_ = self.rows, self.cols, total_nnz
        d = {
            # X in, P out, supp/nnz/entropy/argmax (+ 64 B of support slots)
            "sparsemax_fwd": r * (8 * K + 4 + 4 + 4 + 8 + (64 if self.use_slots else 0)),
            "scan_counts": r * (4 + 8),                         # counts in, offsets out
            # support slots (64 B) + two offsets in, (val,row,col) out; the survey's figure for a fill
            # that scans P is r * (4 * K + 8) + 12 * N_ -- the slots written by the forward replace that read
            "compact_fill": (r * (64 + 16) if self.use_slots else r * (4 * K + 8)) + 12 * N_,
            # P, L in, row loss out  /  P, L, supp (+dent) in, dX, dL out.  With support slots the reads
            # shrink to the slots + one 32-byte sector of L per non-zero
            "wloss_dense_fwd": (r * (64 + 32 + 4) + 32 * N_) if self.use_slots else r * (8 * K + 4),
            "sparsemax_bwd": (r * (64 + 32 + 4 + 8 * K) if self.use_slots else r * (8 * K + 4 + 8 * K))
                             + (4 * r if self.dent is not None else 0),
            # the three above in one launch: slots + offsets + supp (+dent) in, one sector of L per non-zero,
            # row loss, dX, dL and 12 B of lists per non-zero out; without slots P and L are read densely (8K B/row)
            "marg_step_fused": (r * (64 + 16 + 4 + 4 + 8 * K) + (32 + 12) * N_ if self.use_slots
                                else r * (8 * K + 16 + 4 + 4 + 8 * K) + 12 * N_) + (4 * r if self.dent is not None else 0),
        }
        return {k: d[k] for k in self.kernels}


class SparsemaxMargCompactStep:
    """The categorical step when the decoder runs on the SUPPORT only (the paper's point; the
    reference's structured marginalizers work this way, structmarg.py:116-140): losses arrive as a
    ragged vector over the compacted (row, col) pairs.

        scores [B,K] --sparsemax fwd--> P, supp, nnz, entropy        scan --> offsets
        P, offsets   --compact fill--> (value, row, col)              [decoder: ragged losses]
        vals, losses --ragged weighted loss--> mean loss
        support lists --ragged bwd--> dscores [B,K] (zero off the support), dlosses [N]
    """
    KERNELS = ("sparsemax_fwd", "scan_counts", "compact_fill", "wloss_ragged_fwd", "sparsemax_bwd_ragged")

    def __init__(self, rows, cols, device, capacity, entropy_coeff=0.0, global_rows=None):
        self.base = SparsemaxMargStep(rows, cols, device, capacity, entropy_coeff, global_rows)
        self.dloss_ragged = torch.empty(self.base.capacity, dtype=torch.float32, device=device)

    def run(self, x, losses_ragged, n_items, rec=None):
        b = self.base
        st = N.stream()
        rows, K = b.rows, b.cols
        mark = rec if rec is not None else (lambda name: None)
        mark("sparsemax_fwd")
        slots = ptr(b.slots) if b.use_slots else None
        check(lib.smlvm_sparsemax_fwd(ptr(x), ptr(b.p), ptr(b.supp), ptr(b.nnz), ptr(b.ent), ptr(b.amax), slots,
                                      rows, K, st), "smlvm_sparsemax_fwd")
        mark("scan_counts")
        check(lib.smlvm_scan_counts(ptr(b.nnz), ptr(b.offsets), rows, ptr(b.scan_ws), b.scan_ws.numel(), st),
              "smlvm_scan_counts")
        mark("compact_fill")
        check(lib.smlvm_compact_fill(ptr(b.p), ptr(b.offsets), slots, ptr(b.vals), ptr(b.ri), ptr(b.ci), rows, K,
                                     b.capacity, st), "smlvm_compact_fill")
        mark("wloss_ragged_fwd")
        check(lib.smlvm_wloss_ragged_fwd(ptr(b.vals), ptr(losses_ragged), None, None, ptr(b.total), None,
                                         1.0 / b.global_rows, 0, min(int(n_items), b.capacity), ptr(b.red_ws),
                                         b.red_ws.numel(), st),
              "smlvm_wloss_ragged_fwd")
        mark("sparsemax_bwd_ragged")
        check(lib.smlvm_sparsemax_bwd_ragged(ptr(b.offsets), ptr(b.ci), ptr(b.vals), ptr(b.supp), None,
                                             ptr(losses_ragged), None, 1.0 / b.global_rows, ptr(b.dent), ptr(b.dx),
                                             ptr(self.dloss_ragged), rows, K, min(int(n_items), b.capacity), st),
              "smlvm_sparsemax_bwd_ragged")
        mark(None)
        return b.total

    def algorithmic_bytes(self, total_nnz):
        r, K, N_ = self.base.rows, self.base.cols, total_nnz
        rr = r
        d = {"sparsemax_fwd": rr * (8 * K + 4 + 4 + 4 + 8 + (64 if self.base.use_slots else 0)), "scan_counts": rr * 12,
             "compact_fill": (rr * (64 + 16) if self.base.use_slots else rr * (4 * K + 8)) + 12 * N_}
        return {"sparsemax_fwd": d["sparsemax_fwd"], "scan_counts": d["scan_counts"], "compact_fill": d["compact_fill"],
                "wloss_ragged_fwd": 8 * N_ + 4,
                "sparsemax_bwd_ragged": 4 * r * K + r * (8 + 4 + 4) + N_ * (4 + 4 + 4 + 4)}


class HostPipelinedMargStep:
    """The same step for batches that live in (pinned) HOST memory: scores and losses come over
    PCIe, the gradient and the loss go back.  The batch is cut into ``chunks`` micro-batches
    (each a shard with loss scale 1/B_global, exactly like the multi-GPU sharding) that flow
    through three streams -- H2D copy, the five kernels, D2H copy -- over ``nbuf`` rotating
    device buffer sets, so that the uploads (2 x rows x K x 4 B), the kernels and the downloads
    (rows x K x 4 B) overlap; PCIe is full duplex, the step time tends to the upload time alone.

    The call is asynchronous: ``x_host`` / ``l_host`` (pinned; with ``zero_copy_losses`` the kernels read ``l_host`` itself)
    must not be modified, and ``dx_host`` not read, before the current stream has been synchronised.
    ``run(x_host, l_host, dx_host)`` returns the loss as a 0-d pinned host tensor valid after
    the CURRENT stream has been synchronised (the call itself does not block).  ``dx_host=None``: the gradient stays
    on the device (``steps[b].dx`` of the last micro-batches), only the partial losses come back.
    """

    def __init__(self, rows, cols, device, capacity_per_row, entropy_coeff=0.0, chunks=8, nbuf=3, zero_copy_losses=False):
        """``zero_copy_losses``: do not upload the loss matrix; the fused step kernel reads it from the pinned host
        buffer, and only where the support is (the loss of a latent assignment with probability zero never enters the
        objective or a gradient: marg.py:125-126).  Needs the slot-driven fused step (peaked rows); the upload halves."""
        assert rows % chunks == 0, "rows must divide into equal micro-batches"
        self.zero_copy_losses = bool(zero_copy_losses)
        self.rows, self.cols, self.device = rows, cols, device
        self.chunks, self.nbuf = chunks, min(nbuf, chunks)
        self.crow = rows // chunks
        cap = int(capacity_per_row * self.crow) + 1024
        self.steps = [SparsemaxMargStep(self.crow, cols, device, cap, entropy_coeff, global_rows=rows)
                      for _ in range(self.nbuf)]
        self.xd = [torch.empty(self.crow, cols, device=device) for _ in range(self.nbuf)]
        self.ld = [torch.empty(self.crow, cols, device=device) for _ in range(self.nbuf)]
        self.s_h2d, self.s_cmp, self.s_d2h = (torch.cuda.Stream(device) for _ in range(3))
        self.tot_host = torch.zeros(chunks).pin_memory()
        self.loss_host = torch.zeros(()).pin_memory()
        self.ev_free = [None] * self.nbuf   # D2H of the chunk that last used buffer b
        if self.zero_copy_losses and not all(st.fused and st.use_slots for st in self.steps):
            raise RuntimeError("zero_copy_losses needs the slot-driven fused step (supports of at most ~6 entries per row)")
        self.h2d_bytes = (1 if self.zero_copy_losses else 2) * rows * cols * 4   # copies; + one PCIe read per non-zero
        self.d2h_bytes = rows * cols * 4 + 4 * chunks

    def run(self, x_host, l_host, dx_host):
        cur = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(cur)
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            s.wait_event(start)
        last = None
        for c in range(self.chunks):
            b = c % self.nbuf
            lo, hi = c * self.crow, (c + 1) * self.crow
            with torch.cuda.stream(self.s_h2d):
                if self.ev_free[b] is not None:
                    self.s_h2d.wait_event(self.ev_free[b])
                self.xd[b].copy_(x_host[lo:hi], non_blocking=True)
                if not self.zero_copy_losses:
                    self.ld[b].copy_(l_host[lo:hi], non_blocking=True)
                up = torch.cuda.Event()
                up.record(self.s_h2d)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(up)
                self.steps[b].run(self.xd[b], l_host[lo:hi] if self.zero_copy_losses else self.ld[b])
                done = torch.cuda.Event()
                done.record(self.s_cmp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(done)
                if dx_host is not None:
                    dx_host[lo:hi].copy_(self.steps[b].dx, non_blocking=True)
                self.tot_host[c:c + 1].copy_(self.steps[b].total.reshape(1), non_blocking=True)
                last = torch.cuda.Event()
                last.record(self.s_d2h)
                self.ev_free[b] = last
        cur.wait_event(last)
        return self.tot_host

    def loss(self):
        """Sum of the per-micro-batch partial losses (host side; call after synchronising)."""
        return float(self.tot_host.sum())
