"""lvmhelpers.marg counterpart: categorical sparsemax marginalisation.

Reference: lvmhelpers/marg.py.  Same public names, arguments and return
conventions (``entropy``, ``ExplicitWrapper``, ``Marginalizer``); the sparsemax
normaliser, its entropy/argmax, the probability-weighted loss and the whole
backward run as fused sm_100a kernels (``ops.sparsemax_stats``,
``smlvm_wloss_dense_fwd``, ``smlvm_sparsemax_bwd``).
"""
import torch
import torch.nn as nn

from . import ops


def entropy(p: torch.Tensor):
    """Shannon entropy of rows with exact zeros (reference marg.py:6-28).

    Kept in torch ops so it stays usable on any tensor; the sparsemax path of
    ``ExplicitWrapper`` gets the same quantity fused into its forward kernel.
    """
    nz = p > 0
    eps = torch.finfo(p.dtype).eps
    p_stable = p.clone().clamp(min=eps, max=1 - eps)
    out = torch.where(nz, p_stable * torch.log(p_stable),
                      torch.tensor(0., device=p.device, dtype=torch.float))
    return -(out).sum(-1)


class _Entmax15(torch.autograd.Function):
    """1.5-entmax (``entmax.entmax15``, the reference's DEFAULT normalizer argument at
    marg.py:39,46) restated with torch ops: exact sort-based threshold, closed-form backward.
    The definition the kernel (``ops.entmax15``) is tested against, and the path of CUDA tensors the
    kernel does not take (fp64, more than 1024 columns)."""

    @staticmethod
    def forward(ctx, X):
        X = X - X.max(dim=-1, keepdim=True).values
        X = X / 2                                       # p = [(x/2 - tau)_+]^2
        Xsrt, _ = torch.sort(X, dim=-1, descending=True)
        rho = torch.arange(1, X.shape[-1] + 1, device=X.device, dtype=X.dtype)
        mean = Xsrt.cumsum(-1) / rho
        mean_sq = (Xsrt ** 2).cumsum(-1) / rho
        ss = rho * (mean_sq - mean ** 2)
        delta_nz = torch.clamp((1 - ss) / rho, min=0)
        tau = mean - torch.sqrt(delta_nz)
        support = (tau <= Xsrt).sum(dim=-1, keepdim=True)
        tau_star = tau.gather(-1, support - 1)
        Y = torch.clamp(X - tau_star, min=0) ** 2
        ctx.save_for_backward(Y)
        return Y

    @staticmethod
    def backward(ctx, dY):
        (Y,) = ctx.saved_tensors
        gppr = Y.sqrt()
        dX = dY * gppr
        q = dX.sum(-1, keepdim=True) / gppr.sum(-1, keepdim=True)
        return dX - q * gppr


def entmax15(X, dim=-1):
    """``entmax.entmax15``: the sm_100a kernel (``smlvm_entmax15_fwd/_bwd``) for fp32 CUDA scores with at most 1024
    columns; fp64 or wider CUDA tensors take the torch-op restatement above; CPU tensors raise (no CPU path)."""
    if dim not in (-1, X.dim() - 1):
        return entmax15(X.transpose(dim, -1), -1).transpose(dim, -1)
    if not X.is_cuda:
        raise RuntimeError("entmax15: scores must live on a CUDA device: this package has no CPU implementation "
                           "(got device=%s)" % X.device)
    if X.dtype == torch.float32 and X.shape[-1] <= ops.ENTMAX15_MAX_COLS:
        return ops.entmax15(X)
    return _Entmax15.apply(X)       # fp64 scores / more than 1024 columns, on the GPU


class _FusedMarginalObjective(torch.autograd.Function):
    """mean_b sum_z P[b,z] L[b,z]  -  coeff * mean_b H(P[b,:])   (marg.py:83,125-126).

    Forward: one segmented-reduce launch.  Backward: ONE launch that applies the
    sparsemax Jacobian to g*L/B + d(-coeff*mean H)/dP without materialising dP, and
    an elementwise P*g/B for the decoder side.
    """

    @staticmethod
    def forward(ctx, scores, losses, p, supp, ent_rows, coeff, slots=None):
        rows = p.shape[0]
        row_loss, total, lslots = ops.wloss_dense_fwd(p, losses, scale=1.0 / rows, slots=slots, want_loss_slots=True)
        ctx.save_for_backward(p, supp, losses)
        ctx.slots, ctx.loss_slots = slots, lslots
        ctx.coeff = float(coeff)
        ctx.mark_non_differentiable(row_loss)
        full = total - ctx.coeff * ent_rows.mean()
        return full, total.clone(), row_loss

    @staticmethod
    def backward(ctx, g_full, g_total, _g_rows):
        p, supp, losses = ctx.saved_tensors
        rows = p.shape[0]
        g = (g_full + g_total).contiguous().float()
        dent = None
        if ctx.coeff != 0.0:
            dent = (g_full.float() * (-ctx.coeff / rows)).expand(rows).contiguous()
        # one launch: dscores (sparsemax Jacobian of g*L/B + entropy term) and dlosses = g*P/B
        dscores, dlosses = ops.sparsemax_bwd(p, supp, loss=losses, loss_scale=g,
                                             loss_scale_host=1.0 / rows, dent=dent,
                                             want_dloss=True, slots=ctx.slots, loss_slots=ctx.loss_slots)
        return dscores, dlosses, None, None, None, None, None


class _CompactMarginalObjective(torch.autograd.Function):
    """The same objective when the losses exist on the support only (decoder run on the compacted
    (sample, latent) pairs): (p_vals . loss) / B - coeff * mean_b H.  Backward: one launch over the
    support (``smlvm_sparsemax_bwd_ragged``): O(N) reads, dP never materialised, dense dscores."""

    @staticmethod
    def forward(ctx, scores, loss_ragged, comp_offsets, comp_cols, comp_vals, supp, ent_rows, coeff):
        rows, K = scores.shape
        _, total, _ = ops.wloss_ragged_fwd(comp_vals, loss_ragged, scale=1.0 / rows, want_total=True)
        ctx.save_for_backward(loss_ragged, comp_offsets, comp_cols, comp_vals, supp)
        ctx.coeff, ctx.shape = float(coeff), (rows, K)
        return total - ctx.coeff * ent_rows.mean(), total.clone()

    @staticmethod
    def backward(ctx, g_full, g_total):
        loss_ragged, offsets, cols, vals, supp = ctx.saved_tensors
        rows, K = ctx.shape
        g = (g_full + g_total).contiguous().float()
        dent = None
        if ctx.coeff != 0.0:
            dent = (g_full.float() * (-ctx.coeff / rows)).expand(rows).contiguous()
        comp = dict(offsets=offsets, cols=cols, vals=vals)
        dscores, dloss = ops.sparsemax_bwd_ragged(comp, supp, (rows, K), loss=loss_ragged, loss_scale=g,
                                                  loss_scale_host=1.0 / rows, dent=dent, want_dloss=True)
        return dscores, dloss, None, None, None, None, None, None


class ExplicitWrapper(nn.Module):
    """scores -> (sample, distribution, entropy) (reference marg.py:31-54).

    ``normalizer='sparsemax'`` is the B200-native path (one fused kernel for
    sparsemax + entropy + argmax).  ``'softmax'`` and ``'entmax'`` (entmax15, restated above)
    stay on torch ops: they are outside the hot path this package replaces with kernels.
    """

    def __init__(self, agent, normalizer='entmax'):
        super().__init__()
        self.agent = agent
        if normalizer not in ('softmax', 'sparsemax', 'entmax'):
            raise KeyError(normalizer)
        self.normalizer_name = normalizer

    def forward(self, *args, **kwargs):
        scores = self.agent(*args, **kwargs)
        if self.normalizer_name == 'sparsemax':
            shp = scores.shape
            flat = scores.reshape(-1, shp[-1])
            distr, entropy_distr, sample, nnz, supp, slots = ops.sparsemax_stats(flat)
            distr = distr.reshape(shp)
            # side channel for Marginalizer: lets it fuse loss + entropy gradients into
            # one backward launch (dP never materialised)
            distr._smlvm = dict(scores=flat, nnz=nnz, entropy=entropy_distr, supp=supp, slots=slots)
            return sample.reshape(shp[:-1]), distr, entropy_distr.reshape(shp[:-1])
        if self.normalizer_name == 'softmax':
            distr = torch.softmax(scores, dim=-1)
            return scores.argmax(dim=-1), distr, entropy(distr)
        if scores.is_cuda and scores.dtype == torch.float32 and scores.shape[-1] <= ops.ENTMAX15_MAX_COLS:
            distr, entropy_distr, sample = ops.entmax15(scores, want_stats=True)   # one launch
            return sample, distr, entropy_distr
        distr = entmax15(scores, dim=-1)
        return scores.argmax(dim=-1), distr, entropy(distr)


class Marginalizer(torch.nn.Module):
    """Marginalise the downstream loss over the support of a categorical latent
    (reference marg.py:57-132).  Encoder: ``ExplicitWrapper``; decoder:
    ``utils.DeterministicWrapper``.

    Differences in mechanics only: the K per-column host syncs of marg.py:87 become one
    read of the column-activity vector, and with a sparsemax encoder the final
    ``bmm`` + entropy term + their backward run as fused kernels.
    """

    def __init__(self, encoder, decoder, loss_fun,
                 encoder_entropy_coeff=0.0, decoder_entropy_coeff=0.0, compact_support=False):
        """``compact_support`` (extension, default off = the reference's loop): run the decoder
        ONCE on the compacted (sample, latent) pairs of the support instead of once per active
        latent value on the full batch.  Same loss and gradients for a
        decoder that treats rows independently; a decoder that draws random numbers consumes them
        in a different order (experiments/semi_supervised-vae/archs.py:185-187)."""
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder
        self.loss = loss_fun
        self.encoder_entropy_coeff = encoder_entropy_coeff
        self.decoder_entropy_coeff = decoder_entropy_coeff
        self.compact_support = compact_support

    def _forward_compact(self, encoder_input, decoder_input, labels, discrete_latent_z, encoder_probs,
                         encoder_entropy, side):
        """decoder calls proportional to the support: one ragged batch of N = sum_b |supp(b)| rows."""
        batch_size = encoder_probs.shape[0]
        comp = ops.compact(encoder_probs.detach(), side["nnz"], slots=side["slots"])
        rows, cols = comp["rows"], comp["cols"].long()
        take = lambda t: ops.gather_rows(t, rows)
        decoder_input_rep = take(decoder_input)
        decoder_output = self.decoder(cols, decoder_input_rep)
        loss_components, logs = self.loss(take(encoder_input), take(discrete_latent_z), decoder_input_rep,
                                          decoder_output, take(labels))
        p_vals = comp["vals"]
        full_loss, loss_mean = _CompactMarginalObjective.apply(
            side["scores"], loss_components, comp["offsets"], comp["cols"], comp["vals"], side["supp"],
            side["entropy"].detach(), float(self.encoder_entropy_coeff))
        for k, v in logs.items():
            if hasattr(v, 'mean'):
                # sum_z mean_b p[b,z] v[b,z]  ==  (p_vals . v) / B
                logs[k] = (p_vals * v.detach().reshape(-1).to(p_vals.dtype)).sum() / batch_size
        logs['loss'] = loss_mean
        logs['encoder_entropy'] = encoder_entropy.mean()
        logs['support'] = (encoder_probs != 0).sum(-1).to(torch.float).mean()
        logs['distr'] = encoder_probs
        logs['decoder_rows'] = comp["total"]
        return {'loss': full_loss, 'log': logs}

    def forward(self, encoder_input, decoder_input, labels):
        discrete_latent_z, encoder_probs, encoder_entropy = self.encoder(encoder_input)
        batch_size, latent_size = encoder_probs.shape
        if self.compact_support and getattr(encoder_probs, "_smlvm", None) is not None:
            return self._forward_compact(encoder_input, decoder_input, labels, discrete_latent_z,
                                         encoder_probs, encoder_entropy, encoder_probs._smlvm)

        # marg.py:87 -- columns whose total mass is exactly zero are skipped
        active = (encoder_probs.detach().sum(0) != 0).tolist()

        losses = torch.zeros_like(encoder_probs)
        logs_global = None
        logs = {}
        for z in range(latent_size):
            if not active[z]:
                continue
            z_ = torch.full((batch_size,), z, dtype=torch.long, device=encoder_probs.device)
            decoder_output = self.decoder(z_, decoder_input)
            loss_sum_term, logs = self.loss(
                encoder_input, discrete_latent_z, decoder_input, decoder_output, labels)
            losses[:, z] += loss_sum_term
            if not logs_global:
                logs_global = {k: 0.0 for k in logs.keys()}
            for k, v in logs.items():
                if hasattr(v, 'mean'):
                    # expectation of the per-sample statistic under p(z|x)
                    logs_global[k] += (encoder_probs[:, z] * v).mean()
        for k, v in logs.items():
            if hasattr(v, 'mean'):
                logs[k] = logs_global[k]

        side = getattr(encoder_probs, "_smlvm", None)
        if side is not None and encoder_probs.is_cuda:
            full_loss, loss_mean, _ = _FusedMarginalObjective.apply(
                side["scores"], losses, encoder_probs.detach(), side["supp"],
                side["entropy"].detach(), float(self.encoder_entropy_coeff), side.get("slots"))
        else:
            entropy_loss = -(encoder_entropy.mean() * self.encoder_entropy_coeff)
            loss = (encoder_probs * losses).sum(-1)
            loss_mean = loss.mean()
            full_loss = loss_mean + entropy_loss

        logs['loss'] = loss_mean
        logs['encoder_entropy'] = encoder_entropy.mean()
        logs['support'] = (encoder_probs != 0).sum(-1).to(torch.float).mean()
        logs['distr'] = encoder_probs
        return {'loss': full_loss, 'log': logs}
