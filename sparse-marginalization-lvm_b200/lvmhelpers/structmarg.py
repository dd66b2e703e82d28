"""lvmhelpers.structmarg counterpart: bit-vector latents, top-k sparsemax and SparseMAP.

Reference: lvmhelpers/structmarg.py.  Same classes, constructor arguments, return tuples
and ``log`` keys.  What changed underneath:

* top-k: the k best bit-vectors come from ``smlvm_topk_bits`` on the device (no
  D2H/H2D round trip, structmarg.py:43-44), the re-scoring einsum, the sparsemax over k,
  the support compaction (boolean-mask indexing, :116-126) and the weighted loss (:140)
  are kernels of ``libsmlvm.so``; inputs are gathered by compacted row index instead of
  being repeated k times and masked (:95-126).  When the wrapper is the encoder of a
  ``TopKSparsemaxMarg`` (the only way the experiments use it) the k candidates stay PACKED
  (``PackedBitVectors``: 4 B per 32 bits) and only the candidates on the support are expanded
  to fp32 for the decoder -- the dense [B,k,n] fp32 tensor is never written.  Called on its
  own the wrapper returns the reference's dense tensor.
* SparseMAP: the per-sample Python loop over factor graphs (:181-196) is one batched
  solver launch (one CTA per sample) + one ragged expansion; ``idxs`` / ``support`` are
  still Python lists (one D2H of the per-sample sizes).
"""
import torch
import torch.nn as nn

from . import ops
from ._native import FACTOR_BERNOULLI, FACTOR_BUDGET
from .sparsemap import draw_init_scores


def entropy(p):
    """reference structmarg.py:9-20 (same as marg.entropy)."""
    nz = p > 0
    eps = torch.finfo(p.dtype).eps
    p_stable = p.clone().clamp(min=eps, max=1 - eps)
    out = torch.where(nz, p_stable * torch.log(p_stable),
                      torch.tensor(0., device=p.device, dtype=torch.float))
    return -(out).sum(-1)


class _CompactValues(torch.autograd.Function):
    """p[mask] with mask = p > 0, from the compaction kernel; backward scatters."""

    @staticmethod
    def forward(ctx, p, comp):
        ctx.flat_idx = comp["flat_idx"]
        ctx.shape = p.shape
        return comp["vals"]

    @staticmethod
    def backward(ctx, dvals):
        d = torch.zeros(ctx.shape, dtype=dvals.dtype, device=dvals.device)
        d.view(-1).index_copy_(0, ctx.flat_idx, dvals)
        return d, None


def _compact_with_flat(p, nnz, slots=None):
    comp = ops.compact(p.detach(), nnz, slots=slots)
    if p.numel() < 2 ** 31:  # two launches (int32 multiply-add, widen) instead of four
        comp["flat_idx"] = torch.add(comp["cols"], comp["rows"], alpha=p.shape[-1]).to(torch.int64)
    else:
        comp["flat_idx"] = comp["rows"].to(torch.int64) * p.shape[-1] + comp["cols"].to(torch.int64)
    return comp


class _IdxList(list):
    """``idxs`` as the Python list the reference returns (structmarg.py:195,202), carrying
    the same indices as a device tensor so that gathers need no host round trip."""
    tensor = None


class _DeviceList:
    """``idxs`` / ``support`` of the static-shape SparseMAP wrapper: the values live on the device (``tensor``, with
    ``count`` valid entries when given); the Python list the reference returns (structmarg.py:195-202) is built on
    first use -- that, and only that, synchronises."""

    def __init__(self, tensor, count=None):
        self.tensor, self.count, self._list = tensor, count, None

    def tolist(self):
        if self._list is None:
            n = self.tensor.numel() if self.count is None else int(self.count.item())
            self._list = self.tensor[:n].tolist()
        return self._list

    def __iter__(self):
        return iter(self.tolist())

    def __len__(self):
        return len(self.tolist())

    def __getitem__(self, i):
        return self.tolist()[i]


class PackedBitVectors:
    """The wrapper's first output when its consumer is ``TopKSparsemaxMarg``: the [B,k,n] {0,1}
    candidates as packed words [B,k,W].  ``dense()`` gives the reference's fp32 tensor,
    ``rows(index)`` the fp32 rows of the flattened [B*k, n] view at ``index``."""

    def __init__(self, bits, n):
        self.bits, self.n = bits, n
        self.shape = torch.Size((bits.shape[0], bits.shape[1], n))
        self.device, self.dtype = bits.device, torch.float32

    def dense(self):
        return ops.bits_expand(self.bits, None, self.n).view(self.shape)

    def rows(self, index):
        return ops.bits_expand(self.bits, index, self.n)


class TopKSparsemaxWrapper(nn.Module):
    """scores -> (k best bit-vectors [B,k,n], sparsemax over their scores [B,k],
    entropy / B)  (reference structmarg.py:23-58)."""

    def __init__(self, agent, k=10):
        super().__init__()
        self.agent = agent
        self.k = k
        self.packed_output = False  # TopKSparsemaxMarg turns it on around its own encoder call

    def forward(self, *args, **kwargs):
        scores = self.agent(*args, **kwargs)
        batch_size, latent_size = scores.shape
        if ops.topk_sparsemax_supported(latent_size, self.k):
            # k-best search (detached, as the numpy round trip of structmarg.py:43), re-scoring (:47), sparsemax over
            # k (:48) and the entropy over the support (:51-58): three launches, no host read.  The support
            # compaction the reference needs further down (TopKSparsemaxMarg, :116-126) is built there, on demand.
            distr, entropy_distr, bits, cfg, nnz, _ = ops.topk_sparsemax(scores, self.k, want_configs=not self.packed_output)
            bit_vector_z = PackedBitVectors(bits, latent_size) if self.packed_output else cfg
            distr._smlvm_nnz = nnz
            return bit_vector_z, distr, entropy_distr
        # general shapes (k > 16 or n % 8 != 0): the separate kernels
        bit_vector_z, bits, _ = ops.topk_bits(scores.detach(), self.k, want_configs=not self.packed_output)
        if self.packed_output:
            bit_vector_z = PackedBitVectors(bits, latent_size)
        # rank them: scores_k = einsum("bkj,bj->bk"), differentiable w.r.t. scores (:47)
        scores_k = ops.bits_score(scores, bits)
        distr, _, _, nnz, _, slots = ops.sparsemax_stats(scores_k)
        # entropy over the support (:51-58)
        comp = _compact_with_flat(distr, nnz, slots)
        vals = _CompactValues.apply(distr, comp)
        entropy_distr = ops.ragged_entropy(vals, 1.0 / batch_size)
        distr._smlvm_compact = comp
        distr._smlvm_vals = vals
        return bit_vector_z, distr, entropy_distr


class TopKSparsemaxMarg(torch.nn.Module):
    """reference structmarg.py:61-153"""

    def __init__(self, encoder, decoder, loss_fun,
                 encoder_entropy_coeff=0.0, decoder_entropy_coeff=0.0):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.loss = loss_fun
        self.encoder_entropy_coeff = encoder_entropy_coeff
        self.decoder_entropy_coeff = decoder_entropy_coeff

    def _encode(self, encoder_input):
        """The encoder call of structmarg.py:83; a TopKSparsemaxWrapper hands its candidates over packed
        for the duration of THIS call only (called on its own it returns the reference's dense tensor)."""
        enc = self.encoder
        if not isinstance(enc, TopKSparsemaxWrapper):
            return enc(encoder_input)
        prev, enc.packed_output = enc.packed_output, True
        try:
            return enc(encoder_input)
        finally:
            enc.packed_output = prev

    def forward(self, encoder_input, decoder_input, labels):
        bit_vector_z, encoder_probs, encoder_entropy = self._encode(encoder_input)
        batch_size = bit_vector_z.shape[0]
        latent_size = bit_vector_z.shape[-1]

        entropy_loss = -(encoder_entropy * self.encoder_entropy_coeff)

        # support compaction (:116-126): row-major order of encoder_probs > 0
        comp = getattr(encoder_probs, "_smlvm_compact", None)
        if comp is None:
            comp = _compact_with_flat(encoder_probs, getattr(encoder_probs, "_smlvm_nnz", None))
            encoder_probs_flat = _CompactValues.apply(encoder_probs, comp)
        else:
            encoder_probs_flat = encoder_probs._smlvm_vals
        rows = comp["rows"]
        # x.unsqueeze(1).repeat(1,k,1).view(B*k,-1)[mask] == x[row of each survivor] (:95-124)
        encoder_input_rep_flat = ops.gather_rows(encoder_input, rows)
        decoder_input_rep_flat = ops.gather_rows(decoder_input, rows)
        labels_rep_flat = ops.gather_rows(labels, rows)
        if isinstance(bit_vector_z, PackedBitVectors):
            bit_vector_z_flat = bit_vector_z.rows(comp["flat_idx"])
        else:
            bit_vector_z_flat = bit_vector_z.view(-1, latent_size).index_select(0, comp["flat_idx"])

        decoder_output = self.decoder(bit_vector_z_flat, decoder_input)

        loss_components, logs = self.loss(
            encoder_input_rep_flat, bit_vector_z_flat, decoder_input_rep_flat,
            decoder_output, labels_rep_flat)

        # (p . loss) / B  (:140)
        loss = ops.ragged_loss(encoder_probs_flat, loss_components, 1.0 / batch_size)
        full_loss = loss + entropy_loss

        for k, v in logs.items():
            if hasattr(v, 'mean'):
                logs[k] = v.mean()

        logs['loss'] = loss.detach()
        logs['encoder_entropy'] = encoder_entropy.detach()
        logs['support'] = (encoder_probs > 0).sum(dim=-1).to(torch.float)
        logs['distr'] = encoder_probs
        logs['loss_output'] = loss_components
        return {'loss': full_loss, 'log': logs}


class SparseMAPWrapper(nn.Module):
    """scores -> (vertices [N,n], distribution [N], entropy / B, idxs, support)
    (reference structmarg.py:156-202)."""

    def __init__(self, agent, budget=0, init=False, max_iter=300, static_shapes=False):
        """``static_shapes`` (extension, default off = the reference's ragged outputs): N = B * (n + 1) entries
        whatever the supports -- the reference's entries in sample order followed by zero rows with probability 0 --
        and ``idxs`` / ``support`` as device-backed lists.  No host synchronisation: a training step over the wrapper
        can be captured in a CUDA graph (the experiments' B = 64, n = 32 steps are launch- and sync-bound)."""
        super().__init__()
        self.agent = agent
        self.budget = budget
        self.init = init
        self.max_iter = max_iter
        self.static_shapes = static_shapes

    def forward(self, *args, **kwargs):
        scores = self.agent(*args, **kwargs)
        batch_size, latent_size = scores.shape
        kind = FACTOR_BUDGET if self.budget > 0 else FACTOR_BERNOULLI
        init_scores = draw_init_scores(batch_size, latent_size) if self.init else None
        distr, sample, aux = ops.sparsemap_batch(
            scores, kind, budget=self.budget, max_iter=self.max_iter,
            init_scores=init_scores, only_positive=True, static=self.static_shapes)
        self.last_solver_state = aux["state"]
        if self.static_shapes:
            entropy_distr = ops.ragged_entropy(distr, 1.0 / batch_size)
            return sample, distr, entropy_distr, _DeviceList(aux["idx"], aux["offsets"][-1]), _DeviceList(aux["sizes"])
        support = aux["sizes"].tolist()
        assert all(s > 0 for s in support)  # structmarg.py:192
        idxs = _IdxList(aux["idx"].tolist())
        idxs.tensor = aux["idx"]
        entropy_distr = ops.ragged_entropy(distr, 1.0 / batch_size)
        self.last_solver_state = aux["state"]
        return sample, distr, entropy_distr, idxs, support


class SparseMAPMarg(torch.nn.Module):
    """reference structmarg.py:205-255"""

    def __init__(self, encoder, decoder, loss_fun,
                 encoder_entropy_coeff=0.0, decoder_entropy_coeff=0.0):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.loss = loss_fun
        self.encoder_entropy_coeff = encoder_entropy_coeff
        self.decoder_entropy_coeff = decoder_entropy_coeff

    def forward(self, encoder_input, decoder_input, labels):
        bit_vector_z, encoder_probs, encoder_entropy, idxs, support = \
            self.encoder(encoder_input)
        batch_size = encoder_input.shape[0]

        entropy_loss = -(encoder_entropy * self.encoder_entropy_coeff)

        decoder_output = self.decoder(bit_vector_z)

        gather = getattr(idxs, "tensor", None)
        if gather is None:
            gather = torch.as_tensor(idxs, device=encoder_input.device)
        loss_components, logs = self.loss(
            ops.gather_rows(encoder_input, gather), bit_vector_z,
            ops.gather_rows(decoder_input, gather), decoder_output,
            ops.gather_rows(labels, gather))

        loss = ops.ragged_loss(encoder_probs, loss_components, 1.0 / batch_size)
        full_loss = loss + entropy_loss

        for k, v in logs.items():
            if hasattr(v, 'mean'):
                logs[k] = v.mean()

        logs['loss'] = loss.detach()
        logs['encoder_entropy'] = encoder_entropy.detach()
        logs['support'] = support.tensor.to(torch.float) if isinstance(support, _DeviceList) \
            else torch.tensor(support).to(torch.float)
        logs['distr'] = encoder_probs.detach()
        logs['loss_output'] = loss_components.detach()
        logs['idxs'] = idxs
        return {'loss': full_loss, 'log': logs}
