"""Randomised parity sweep of the categorical pipeline (forward, slots, compaction, slot-driven loss and
backward) against the CPU oracle: random shapes and scales plus adversarial rows (exact ties, constant rows,
gaps of exactly 1, -inf masks, huge magnitudes, one-hot rows).  tools/fuzz_sparsemax.py is the long form."""
import importlib.util
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fuzz(name="fuzz_sparsemax"):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("seed", [3, 11])
def test_fuzz_categorical_pipeline(seed):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    assert _fuzz().run(seed=seed, n_cases=12, verbose=False) == 0


def test_fuzz_topk_bits():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    assert _fuzz("fuzz_solver").run_topk(seed=5, n_cases=15, verbose=False) == 0


def test_fuzz_sparsemap_solver():
    """Every sample's active-set size and iteration count against the oracle's batch loop (capacity tiers
    included: n = 128 rows that outgrow a tier are re-solved), full solutions on a sample of rows."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    assert _fuzz("fuzz_solver").run_solver(seed=5, n_cases=8, rows=320, verbose=False) == 0


def test_fuzz_sparsemap_sequence_factor():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    assert _fuzz("fuzz_solver").run_sequence(seed=5, n_cases=4, rows=32, verbose=False) == 0


@pytest.mark.parametrize("seed", [4, 12])
def test_fuzz_wide_support_kernels(seed):
    """Round-2 kernels (cooperative forward, dense fused step, k-sparsemax, entmax15) on random shapes / scales /
    adversarial rows; the long form (tools/fuzz_wide.py, 150 cases) ran with 0 mismatches on the B200."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    assert _fuzz("fuzz_wide").run(seed=seed, n_cases=10, verbose=False) == 0
