"""ctypes binding of the C-ABI library ``libsmlvm.so`` (see ``include/smlvm.h``).

This is the only door between the Python host side and the sm_100a kernels.  There is
no CPU fallback: if the library is missing, fails to load or lacks a symbol, importing
this module raises, and every op raises if it is handed a problem the kernels reject.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMLVM_LIB_PATH") or os.path.join(_HERE, "libsmlvm.so")   # override: development builds only

ABI_VERSION = 5

FACTOR_BERNOULLI, FACTOR_BUDGET, FACTOR_SEQUENCE_BINARY, FACTOR_SEQUENCE = 0, 1, 2, 3
SOLVE_STATUS = {0: "converged", 1: "max_iter", 2: "repeated_config", 3: "singular"}

SPARSEMAX_MAX_COLS = 8192
TOPK_MAX_BITS = 512
TOPK_MAX_K = 32
SPARSEMAP_MAX_BITS = 128
SUPPORT_SLOTS = 8

_vp = ctypes.c_void_p
_i32, _i64, _f32, _sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/smlvm.h one to one
_PROTOTYPES = {
    "smlvm_version": (ctypes.c_int, []),
    "smlvm_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "smlvm_sparsemax_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_sparsemax_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_sparsemax_bwd_ragged": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _i64, _i32, _i64, _vp]),
    "smlvm_marg_step_fused": (ctypes.c_int, [_vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp,
                                             _i64, _i32, _vp, _sz, _vp]),
    "smlvm_scan_workspace_bytes": (_sz, [_i64]),
    "smlvm_scan_counts": (ctypes.c_int, [_vp, _vp, _i64, _vp, _sz, _vp]),
    "smlvm_compact_count": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "smlvm_compact_fill": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i64, _vp]),
    "smlvm_topk_bits": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "smlvm_bits_score_fwd": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "smlvm_bits_score_bwd": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "smlvm_ksparsemax_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp, _sz, _vp]),
    "smlvm_topk_sparsemax_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _f32, _vp, _i64, _i32, _i32, _vp]),
    "smlvm_entmax15_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_entmax15_bwd": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_fy_loss_fwd": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_row_scale": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_bits_expand": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_gather_rows": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "smlvm_sparsemap_capacity": (_i32, [_i32]),
    "smlvm_sparsemap_plan": (_i32, [_i32, _i32, _vp, _vp, _vp]),
    "smlvm_sparsemap_workspace_bytes": (_sz, [_i64]),
    "smlvm_sparsemap_fwd": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "smlvm_sparsemap_expand": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_sparsemap_seq_fwd": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                               _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "smlvm_sparsemap_seq_bwd": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64,
                                               _i32, _vp]),
    "smlvm_sparsemap_bwd": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "smlvm_reduce_workspace_bytes": (_sz, []),
    "smlvm_wloss_dense_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i32, _vp, _sz, _vp]),
    "smlvm_wloss_ragged_fwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _i64, _i64, _vp, _sz, _vp]),
    "smlvm_wloss_ragged_bwd": (ctypes.c_int, [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _i64, _vp]),
    "smlvm_segment_logsumexp": (ctypes.c_int, [_vp, _vp, _f32, _f32, _vp, _i64, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libsmlvm.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C sparse-marginalization-lvm_b200/csrc`. There is no CPU fallback."
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.smlvm_version() != ABI_VERSION:
        raise ImportError("libsmlvm.so ABI version %d != expected %d"
                          % (lib.smlvm_version(), ABI_VERSION))
    return lib


lib = _load()


class SmlvmError(RuntimeError):
    pass


def check(rc, what):
    if rc == 0:
        return
    msg = lib.smlvm_error_string(rc).decode()
    if rc == -1:
        raise ValueError("%s: %s" % (what, msg))
    raise SmlvmError("%s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """cudaStream_t of torch's current stream on the current device.  The two C entry points cost well
    under a microsecond; ``torch.cuda.current_stream()`` walks through a dozen Python frames (0.3 ms per
    wrapper step at the launch-bound B = 64 shapes, 18 launches per step)."""
    if _raw_stream is not None and _cur_device is not None:
        return _raw_stream(_cur_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name):
    """Every op funnels its primary operand through here: CUDA tensors only, and on the CURRENT device -- the
    kernels are launched on the current device's stream, so an operand of another GPU would be read by the wrong
    device (or fault without peer access).  Wrap the call in ``torch.cuda.device(t.device)``."""
    if not t.is_cuda:
        raise RuntimeError(
            "%s must live on a CUDA device: the sparse-marginalization path has no CPU "
            "implementation (got device=%s)" % (name, t.device))
    cur = _cur_device() if _cur_device is not None else torch.cuda.current_device()
    if t.device.index != cur:
        raise RuntimeError(
            "%s lives on %s but the current CUDA device is cuda:%d: the kernels launch on the current device's "
            "stream -- wrap the call in `with torch.cuda.device(%r):`" % (name, t.device, cur, str(t.device)))


# A small cache of zero-initialised reduction / scan workspaces per (device, stream).
_WS = {}


def reduce_workspace(device):
    key = ("reduce", device, stream())
    ws = _WS.get(key)
    if ws is None:
        ws = torch.zeros(lib.smlvm_reduce_workspace_bytes(), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def scan_workspace(device, rows):
    need = lib.smlvm_scan_workspace_bytes(rows)
    key = ("scan", device, stream())
    ws = _WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 4096), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws
