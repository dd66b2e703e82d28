"""Batch sharding of the hot path over the GPUs of one box (one process per GPU).

The reference is single-process (no ``torch.distributed`` import anywhere);
every kernel of this path is row-independent, so the batch shards by contiguous row blocks
with NO data-path collective.  What crosses ranks is only what the north star names: the
gradient all-reduce of the wrapped nets' parameters (NCCL over NVLink/NVSwitch on the GPU
box, gloo in the CPU tests), plus optional scalar statistics for logging.

Two places need the GLOBAL batch size even though the data is local:
  * the weighted loss divides by B (lvmhelpers/structmarg.py:140,242; marg.py:126 ``mean``)
    -> every rank passes ``loss_scale(global_rows)`` so per-shard partial losses SUM to the
    single-process loss and summed gradients equal the single-process gradients;
  * ``idxs`` (structmarg.py:195,202) are shard-local row ids; ``to_global_rows`` offsets them.
"""
import torch
import torch.distributed as td


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if td.is_available() and td.is_initialized():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def shard_bounds(global_rows, rank, world_size):
    """Contiguous row block [start, stop) of ``rank``; the first ``global_rows % world_size``
    ranks take one extra row (ragged tails are legal: a shard may be empty)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    if global_rows < 0:
        raise ValueError("negative row count")
    base, extra = divmod(global_rows, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def loss_scale(global_rows):
    """The 1/B of structmarg.py:140,242 with B the GLOBAL batch."""
    return 1.0 / float(global_rows)


def to_global_rows(local_idx, rank, world_size, global_rows):
    """Shard-local sample ids (``idxs``) -> global sample ids."""
    start, _ = shard_bounds(global_rows, rank, world_size)
    return local_idx + start


class GradAllReducer:
    """All-reduce of parameter gradients in flat buckets, launched asynchronously so the collective overlaps
    whatever runs next.

    Buckets are sized for launch latency, not link count (NVSwitch gives every GPU full bandwidth to every
    peer): one bucket of up to ``bucket_bytes`` per all-reduce, in REVERSE parameter order (the order in which
    backward produces the gradients).

    ``average``: which 1/B the sharded losses carry.
      * ``False`` (sum): every rank scaled its loss by 1/B_GLOBAL (``hotpath.SparsemaxMargStep(global_rows=...)``,
        ``loss_scale``) -- summed gradients are the single-process gradients.
      * ``True`` (mean): every rank used its LOCAL 1/B, which is what the reference-API wrappers ``Marginalizer``,
        ``TopKSparsemaxMarg`` and ``SparseMAPMarg`` hard-code (marg.py:126, structmarg.py:140,242).  With equal
        shards the mean over ranks is the single-process gradient (the DDP convention).

    ``attach()`` registers post-accumulate hooks: a bucket's all-reduce is launched from inside ``backward()`` the
    moment its last gradient has been written, so it runs under the rest of the backward pass; ``finish()`` then
    only waits.  Without ``attach()``, call ``start()`` after backward.
    """

    def __init__(self, params, bucket_bytes=64 << 20, group=None, average=False):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.average = bool(average)
        self.buckets = []
        cur, size = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * p.element_size()
            if cur and (size + nbytes > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._pending = []
        self._hooks = []
        self._ready = None

    def _launch(self, bucket):
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        work = td.all_reduce(flat, op=td.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((bucket, flat, work))

    def start(self):
        """Flatten + launch one async all-reduce per bucket.  Gradients that are None count as 0."""
        _, ws = world()
        self._pending = []
        if ws == 1:
            return self
        for bucket in self.buckets:
            self._launch(bucket)
        return self

    def attach(self):
        """Launch each bucket from backward itself (``register_post_accumulate_grad_hook``)."""
        if self._hooks:
            return self
        self._ready = [0] * len(self.buckets)
        where = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}

        def hook(p):
            _, ws = world()
            if ws == 1:
                return
            bi = where[id(p)]
            self._ready[bi] += 1
            if self._ready[bi] == len(self.buckets[bi]):
                self._ready[bi] = 0
                self._launch(self.buckets[bi])

        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(hook))
        return self

    def detach(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def finish(self):
        """Wait and scatter the reduced buckets back into ``.grad``."""
        _, ws = world()
        for bucket, flat, work in self._pending:
            work.wait()
            if self.average:
                flat.div_(ws)
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self._pending = []

    def __call__(self):
        self.start().finish()


def allreduce_stats(stats, group=None):
    """Sum a dict of python/torch scalars across ranks (logging only: loss, support mass)."""
    _, ws = world()
    keys = sorted(stats)
    if ws == 1 or not keys:
        return {k: float(stats[k]) for k in keys}
    first = stats[keys[0]]
    dev = first.device if isinstance(first, torch.Tensor) else torch.device("cpu")
    if td.get_backend(group) == "nccl" and dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=dev)
    td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return dict(zip(keys, t.tolist()))
