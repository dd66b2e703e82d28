"""lvmhelpers.pbinary_topk counterpart (reference: lvmhelpers/pbinary_topk.pyx:18-34).

``batched_topk(scores, configs, k)`` fills ``configs`` in place and returns None, like
the Cython function.  Host buffers (numpy arrays / CPU tensors, what the reference
passes at lvmhelpers/structmarg.py:43) are staged to the current CUDA device and the
result copied back into ``configs``; CUDA tensors are processed in place with no copy.
"""
import numpy as np
import torch

from . import ops


def _check_args(rows, n, k):
    if k < 2:
        raise ValueError("k must be > 1 (binary_topk.h:102)")
    if n < 1:
        raise ValueError("scores must have at least one column")
    if n < 31 and (1 << n) < k:
        raise ValueError("only 2**%d bit-vectors exist, k=%d requested" % (n, k))


def batched_topk(scores, configs, k):
    if isinstance(scores, torch.Tensor) and scores.is_cuda:
        rows, n = scores.shape
        _check_args(rows, n, k)
        if not (isinstance(configs, torch.Tensor) and configs.is_cuda and configs.is_contiguous()
                and configs.dtype == torch.float32 and tuple(configs.shape) == (rows, k, n)):
            raise ValueError("configs must be a contiguous CUDA float32 tensor of shape [rows, k, n]")
        from ._native import lib, check, ptr, stream
        x = scores.detach().float().contiguous()
        check(lib.smlvm_topk_bits(ptr(x), ptr(configs), None, None, rows, n, k, stream()),
              "smlvm_topk_bits")
        return None
    # host buffers: memoryview semantics of the Cython signature (float[:, :], float[:, :, :])
    s = np.asarray(scores)
    if s.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float' but got '%s'" % s.dtype)
    rows, n = s.shape
    _check_args(rows, n, k)
    c = configs.numpy() if isinstance(configs, torch.Tensor) else np.asarray(configs)
    if c.dtype != np.float32 or c.shape != (rows, k, n):
        raise ValueError("configs must be float32 of shape [rows, k, n]")
    dev = torch.device("cuda", torch.cuda.current_device())
    x = torch.from_numpy(np.ascontiguousarray(s)).to(dev, non_blocking=True)
    out, _, _ = ops.topk_bits(x, k, want_configs=True, want_bits=False)
    c[...] = out.cpu().numpy()
    return None


def binary_topk(scores, k):
    """[(score, [bit]*n)] for one row (pbinary_topk.pyx:18-20)."""
    s = np.asarray(scores, dtype=np.float32).reshape(1, -1)
    n = s.shape[1]
    _check_args(1, n, k)
    dev = torch.device("cuda", torch.cuda.current_device())
    cfg, _, sc = ops.topk_bits(torch.from_numpy(s).to(dev), k, want_configs=True, want_bits=False,
                               want_scores=True)
    cfg = cfg[0].cpu().numpy().astype(np.int64)
    sc = sc[0].cpu().numpy()
    return [(float(sc[j]), [int(b) for b in cfg[j]]) for j in range(k)]
