"""Drop-in boundary on the CPU: the names of SURVEY.md §8(b) and the experiments' own import lines.

The hot-path modules resolve to this package, the estimators that are out of scope (sfe, nvil, gumbel,
sum_and_sample) to the reference checkout placed AFTER it on sys.path (lvmhelpers/__init__.py)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sparse-marginalization-lvm_b200")
REF = "/root/reference"


def test_factor_handles_and_class_tree():
    """PFactor*.initialize signatures, SparseMAP / SequenceSparseMAP bases with make_factor."""
    from lvmhelpers import sparsemap, _native
    from lvmhelpers.pbernoulli import PFactorBernoulli
    from lvmhelpers.pbudget import PFactorBudget
    from lvmhelpers.psequence import PFactorSequence, PFactorSequenceBinary
    f = PFactorBernoulli(); f.initialize(32)
    assert (f.kind, f.length, f.budget) == (_native.FACTOR_BERNOULLI, 32, 0)
    f = PFactorBudget(); f.initialize(32, 5)
    assert (f.kind, f.length, f.budget) == (_native.FACTOR_BUDGET, 32, 5)
    f = PFactorSequenceBinary(); f.initialize(7)
    assert (f.kind, f.length) == (_native.FACTOR_SEQUENCE_BINARY, 7)
    f = PFactorSequence(); f.initialize([2, 3, 2])
    # sequence.h:274-289: start->2, 2x3, 3x2, 2->stop
    assert (f.kind, f.length, f.num_variables, f.num_additionals) == (_native.FACTOR_SEQUENCE, 3, 7, 2 + 6 + 6 + 2)
    with pytest.raises(RuntimeError):
        PFactorBernoulli()._require_initialized()
    assert issubclass(sparsemap.BernSparseMAP, sparsemap.SparseMAP)
    assert issubclass(sparsemap.BudgetSparseMAP, sparsemap.SparseMAP)
    assert issubclass(sparsemap.SequenceBinarySparseMAP, sparsemap.SequenceSparseMAP)
    for cls in (sparsemap.SparseMAP, sparsemap.SequenceSparseMAP):
        assert hasattr(cls, "run_sparsemap") and hasattr(cls, "jv") and hasattr(cls, "make_factor")

    class Ctx:
        n, budget = 9, 4
    assert isinstance(sparsemap.BernSparseMAP.make_factor(Ctx), PFactorBernoulli)
    b = sparsemap.BudgetSparseMAP.make_factor(Ctx)
    assert isinstance(b, PFactorBudget) and b.budget == 4 and b.length == 9
    assert isinstance(sparsemap.SequenceBinarySparseMAP.make_factor(Ctx), PFactorSequenceBinary)


def test_utils_without_reference_names_what_is_missing():
    code = ("import sys; sys.path.insert(0, %r); import lvmhelpers.utils as u\n"
            "assert hasattr(u, 'DeterministicWrapper')\n"
            "try:\n    u.populate_common_params\nexcept AttributeError as e:\n    assert 'reference checkout' in str(e)\n"
            "else:\n    raise SystemExit('expected AttributeError')\n") % PKG
    subprocess.run([sys.executable, "-c", code], check=True, cwd="/tmp")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "experiments")), reason="reference checkout not present")
def test_experiment_import_lines_resolve():
    """Every `from lvmhelpers.X import ...` of experiments/*/train.py, executed with this package first and the
    reference second on the path: hot-path modules come from here, the other estimators from the reference."""
    lines = []
    for exp in ("bit_vector-vae", "semi_supervised-vae", "signal-game"):
        src = open(os.path.join(REF, "experiments", exp, "train.py")).read().replace("\\\n", " ")
        lines += [l.strip() for l in src.splitlines() if re.match(r"\s*from lvmhelpers\.", l)]
    assert len(lines) >= 15
    code = "import sys; sys.path[:0] = [%r, %r]\n" % (PKG, REF) + "\n".join(sorted(set(lines))) + "\n"
    code += ("import lvmhelpers.marg, lvmhelpers.structmarg, lvmhelpers.utils, lvmhelpers.sfe, lvmhelpers.nvil\n"
             "here = lambda m: m.__file__.startswith(%r)\n" % PKG +
             "assert here(lvmhelpers.marg) and here(lvmhelpers.structmarg) and here(lvmhelpers.utils)\n"
             "assert lvmhelpers.sfe.__file__.startswith(%r) and lvmhelpers.nvil.__file__.startswith(%r)\n" % (REF, REF) +
             "import argparse; p = populate_common_params(argparse.ArgumentParser())\n"
             "a = p.parse_args(['--mode', 'sparsemap', '--budget', '16'])\n"
             "assert a.mode == 'sparsemap' and a.budget == 16 and a.topksparse == 10\n")
    subprocess.run([sys.executable, "-c", code], check=True, cwd="/tmp")
