"""CPU suite (no GPU): the C-ABI library loads, exports exactly what include/smlvm.h declares,
validates arguments without touching a device, and the Python host side refuses CPU tensors
(no CPU fallback).  No compute kernel is launched here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smlvm.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smlvm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lvmhelpers import _native
    declared = header_symbols()
    assert len(declared) >= 19
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libsmlvm.so lacks %s declared in include/smlvm.h" % name
    # the ctypes prototypes mirror the header one to one
    assert sorted(_native.EXPORTED_SYMBOLS) == declared


def test_dynamic_symbol_table_has_no_stray_exports():
    from lvmhelpers import _native
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("nm unavailable")
    exported = {ln.split()[-1] for ln in out.stdout.splitlines() if " T " in ln}
    ours = {s for s in exported if s.startswith("smlvm_")}
    assert ours == set(header_symbols())


def test_library_is_sm100a_only():
    from lvmhelpers import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_version_and_error_strings():
    from lvmhelpers import _native
    lib = _native.lib
    assert lib.smlvm_version() == _native.ABI_VERSION == 5
    assert b"argument" in lib.smlvm_error_string(-1).lower()
    for code in (0, -1, -2, -3, 1, 700, 12345):
        assert lib.smlvm_error_string(code) is not None


def test_argument_validation_needs_no_device():
    """Argument errors are reported before any CUDA call (negative sizes, NULL pointers,
    shapes outside the documented limits)."""
    from lvmhelpers._native import lib
    assert lib.smlvm_sparsemax_fwd(None, None, None, None, None, None, None, -1, 4, None) == -1
    assert lib.smlvm_sparsemax_fwd(None, None, None, None, None, None, None, 4, 0, None) == -1
    assert lib.smlvm_sparsemax_fwd(None, None, None, None, None, None, None, 0, 4, None) == 0   # empty batch
    assert lib.smlvm_sparsemax_fwd(None, None, None, None, None, None, None, 4, 4, None) == -1  # NULL x
    assert lib.smlvm_sparsemax_fwd(1 << 20, 1 << 21, None, None, None, None, None, 4, 10000, None) == -2
    assert lib.smlvm_sparsemax_bwd(None, None, None, None, None, 1.0, None, None, None, None, None, 0, 4, None) == 0
    assert lib.smlvm_topk_bits(1 << 20, 1 << 21, None, None, 4, 3, 9, None) == -1   # k > 2^n
    assert lib.smlvm_topk_bits(1 << 20, 1 << 21, None, None, 4, 8, 1, None) == -1   # k < 2 (binary_topk.h:102)
    assert lib.smlvm_topk_bits(1 << 20, 1 << 21, None, None, 4, 600, 4, None) == -2  # n > CODE_MAX_SIZE
    assert lib.smlvm_sparsemap_capacity(32) == 33
    tail = (None, None, None, None, 0, None, None, 1 << 24, 1 << 20, None)
    assert lib.smlvm_sparsemap_fwd(1 << 20, None, None, 7, 0, 10, 4, 32, 1 << 21, 1 << 22, 1 << 23, *tail) == -1   # bad factor kind
    assert lib.smlvm_sparsemap_fwd(1 << 20, None, None, 0, 0, 10, 4, 4096, 1 << 21, 1 << 22, 1 << 23, *tail) == -2  # n too large
    assert lib.smlvm_sparsemap_fwd(1 << 20, None, None, 0, 0, 10, 4, 32, 1 << 21, 1 << 22, 1 << 23,
                                   None, None, None, None, 0, None, None, 1 << 24, 8, None) == -3                  # workspace too small
    assert lib.smlvm_sparsemap_fwd(1 << 20, None, None, 0, 0, 10, 4, 32, 1 << 21, 1 << 22, 1 << 23,
                                   None, None, None, 1 << 25, 100, None, None, 1 << 24, 1 << 20, None) == -1       # arena without offsets
    assert lib.smlvm_sparsemap_workspace_bytes(1000) >= 2 * 4 * 1000
    assert lib.smlvm_scan_workspace_bytes(1 << 20) > 0 and lib.smlvm_reduce_workspace_bytes() > 0
    assert lib.smlvm_gather_rows(None, None, None, 4, 16, None) == -1                 # NULL buffers
    assert lib.smlvm_gather_rows(None, None, None, 0, 16, None) == 0                  # empty selection
    assert lib.smlvm_gather_rows(1 << 20, 1 << 21, 1 << 22, 4, 6, None) == -2         # rows not a multiple of 4 bytes
    assert lib.smlvm_bits_expand(None, None, None, 4, 32, None) == -1
    assert lib.smlvm_bits_expand(None, None, None, 0, 32, None) == 0
    assert lib.smlvm_sparsemap_bwd(1 << 20, 1 << 21, 1, 1 << 22, 1 << 23, 1 << 24, None, None, 1 << 25, None, 4, 32, None) == -1


def test_host_side_refuses_cpu_tensors():
    from lvmhelpers import ops
    from lvmhelpers.marg import ExplicitWrapper
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.sparsemax_fwd(x)
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.topk_bits(x, 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        ops.sparsemap_solve(x, 0)
    with pytest.raises(RuntimeError):
        ExplicitWrapper(torch.nn.Identity(), normalizer="sparsemax")(x)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparse-marginalization-lvm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle/_ref" not in src, f


def test_pbinary_topk_argument_errors():
    from lvmhelpers import pbinary_topk
    s = np.zeros((2, 3), dtype=np.float32)
    with pytest.raises(ValueError):
        pbinary_topk.batched_topk(s, np.zeros((2, 1, 3), dtype=np.float32), 1)
    with pytest.raises(ValueError):
        pbinary_topk.batched_topk(s, np.zeros((2, 9, 3), dtype=np.float32), 9)
    with pytest.raises(ValueError):
        pbinary_topk.batched_topk(s.astype(np.float64), np.zeros((2, 4, 3), dtype=np.float32), 4)
    with pytest.raises(ValueError):
        pbinary_topk.batched_topk(s, np.zeros((2, 4, 4), dtype=np.float32), 4)


def test_wrapper_api_surface_matches_reference():
    """Names and constructor signatures the experiments import."""
    import inspect
    from lvmhelpers import marg, structmarg, sparsemap, utils, pbinary_topk
    assert list(inspect.signature(marg.ExplicitWrapper.__init__).parameters) == ["self", "agent", "normalizer"]
    sig = ["self", "encoder", "decoder", "loss_fun", "encoder_entropy_coeff", "decoder_entropy_coeff"]
    for cls in (marg.Marginalizer, structmarg.TopKSparsemaxMarg, structmarg.SparseMAPMarg):
        assert list(inspect.signature(cls.__init__).parameters)[:6] == sig
    # the only extension: an opt-in keyword, default = the reference's behaviour
    extra = inspect.signature(marg.Marginalizer.__init__).parameters
    assert list(extra)[6:] == ["compact_support"] and extra["compact_support"].default is False
    p = inspect.signature(structmarg.TopKSparsemaxWrapper.__init__).parameters
    assert list(p) == ["self", "agent", "k"] and p["k"].default == 10
    p = inspect.signature(structmarg.SparseMAPWrapper.__init__).parameters
    assert [(k, v.default) for k, v in p.items()][2:5] == [("budget", 0), ("init", False), ("max_iter", 300)]
    assert [(k, v.default) for k, v in p.items()][5:] == [("static_shapes", False)]   # opt-in extension, default off
    p = inspect.signature(sparsemap.budget_smap).parameters
    assert [(k, v.default) for k, v in p.items()][1:] == [("budget", 5), ("max_iter", 100), ("init", True)]
    p = inspect.signature(sparsemap.bernoulli_smap).parameters
    assert [(k, v.default) for k, v in p.items()][1:] == [("max_iter", 100), ("init", True)]
    p = inspect.signature(sparsemap.sequence_smap).parameters
    assert list(p) == ["x", "t", "max_iter", "init"]
    assert callable(marg.entropy) and callable(structmarg.entropy)
    assert callable(pbinary_topk.batched_topk) and callable(pbinary_topk.binary_topk)
    assert hasattr(utils, "DeterministicWrapper")


def test_sparsemap_plan_needs_no_device():
    """Capacity tiers of the solver: ascending, last = n + 1, within the 227 KB of an SM; the budget
    factor starts with a larger first tier than Bernoulli (its active sets run larger)."""
    from lvmhelpers import ops
    from lvmhelpers._native import lib
    for kind in (0, 1, 2):
        for n in (1, 5, 32, 64, 100, 128):
            tiers = ops.sparsemap_plan(n, kind)
            assert 1 <= len(tiers) <= 3
            caps = [c for c, _, _ in tiers]
            assert caps == sorted(set(caps)) and caps[-1] == n + 1 == lib.smlvm_sparsemap_capacity(n)
            for cap, smem, ctas in tiers:
                assert 0 < smem <= 227 * 1024 and 1 <= ctas <= 32
                assert ctas * (smem + 1024) <= 228 * 1024
    assert ops.sparsemap_plan(128, 1)[0][0] > ops.sparsemap_plan(128, 0)[0][0]
    assert ops.sparsemap_plan(128, 0)[0][2] == 6 and ops.sparsemap_plan(128, 1)[0][2] == 4
    assert ops.sparsemap_plan(129, 0) == [] and ops.sparsemap_plan(8, 7) == []
