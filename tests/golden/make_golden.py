#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ (run HERE, in the build
container; the GPU box has no /root/reference).  TEST INFRASTRUCTURE.

Two kinds of fixtures:

1. ``kernels.npz`` -- inputs + outputs of the pieces of the REFERENCE ITSELF that compile or
   import here:
     * top-k bit-vectors from the reference's own Cython module (lvmhelpers/pbinary_topk.pyx +
       binary_topk.h, built into oracle/_ref by oracle/ref/Makefile);
     * factor MAP oracles from the reference's own headers (bernoulli.h, budget.h, sequence.h,
       sequence_binary.h, built into oracle/_ref/libref_factors.so);
   and of the oracle restatement where the reference's arithmetic lives in absent third-party
   packages (entmax.sparsemax; libad3 SolveQP driven through the reference's own factors).

2. ``wrappers.npz`` -- losses, log entries and parameter gradients produced by the reference's
   own Python wrappers (lvmhelpers/marg.py, structmarg.py, sparsemap.py imported from
   /root/reference, unmodified) on small seeded problems.  The three third-party / Cython
   imports those files need are satisfied by stubs built on the oracle: ``entmax.sparsemax``,
   ``lpsmap.ad3qp.factor_graph.PFactorGraph`` and the ``PFactor*`` classes (oracle solver),
   and ``lvmhelpers.pbinary_topk`` (the reference's real compiled module from oracle/_ref).

    python tests/golden/make_golden.py
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

oracle.build(ref=True)


# ------------------------------------------------------------------------------------------
# stubs for the reference's third-party imports, built on the oracle
# ------------------------------------------------------------------------------------------
def install_reference_package():
    entmax = types.ModuleType("entmax")
    entmax.sparsemax = oracle.sparsemax
    entmax.entmax15 = None  # never reached: normalizer='sparsemax'
    sys.modules["entmax"] = entmax

    class PFactorGraph:  # lpsmap.ad3qp.factor_graph.PFactorGraph as used at sparsemap.py:15-20
        def set_verbosity(self, v):
            pass

        def create_binary_variable(self):
            return object()

        def declare_factor(self, f, variables):
            f.n_vars = len(variables)

    class _PFactor:  # the PGenericFactor surface used at sparsemap.py:27,29-30,42,63
        kind = None

        def __init__(self):
            self.budget, self.t, self.init, self._last = 0, None, None, None

        def set_additional_log_potentials(self, t):
            self.t = np.asarray(t, dtype=np.float64)

        def init_active_set_from_scores(self, eta_u, eta_v):
            self.init = np.asarray(eta_u, dtype=np.float64)

        def solve_qp(self, eta_u, eta_v, max_iter=10):
            eta_u = np.asarray(eta_u, dtype=np.float64)
            n = eta_u.shape[0] // 2
            assert np.all(eta_u[n:] == 0)
            self._args = dict(x=eta_u[:n], kind=self.kind, budget=self.budget, t=self.t,
                              init=self.init, max_iter=max_iter)
            self._last = oracle.sparsemap_fwd(which="ref_factors", **self._args)
            return self._last["u"], None

        def get_sparse_solution(self):
            # vector[double] -> list of Python floats in the Cython wrapper, hence float32 after
            # torch.tensor(p) at sparsemap.py:34
            return ([list(map(int, row)) for row in self._last["aset"]],
                    [float(v) for v in self._last["p"]])

        def dist_jacobian_vec(self, dp, d_eta_u, d_eta_v):
            dx, dt = oracle.sparsemap_bwd(dp=np.asarray(dp, dtype=np.float64), which="ref_factors",
                                          **self._args)
            n = dx.shape[0]
            d_eta_u[:n] = dx
            d_eta_u[n:] = 0
            if dt is not None and len(d_eta_v):
                d_eta_v[:] = dt

    class PFactorBernoulli(_PFactor):
        kind = oracle.KIND_BERNOULLI

        def initialize(self, n):
            self.n = n

    class PFactorBudget(_PFactor):
        kind = oracle.KIND_BUDGET

        def initialize(self, n, budget):
            self.n, self.budget = n, budget

    class PFactorSequenceBinary(_PFactor):
        kind = oracle.KIND_SEQUENCE_BINARY

        def initialize(self, n):
            self.n = n

    for name in ("lpsmap", "lpsmap.ad3qp", "lpsmap.ad3qp.factor_graph"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["lpsmap.ad3qp.factor_graph"].PFactorGraph = PFactorGraph

    pkg = types.ModuleType("lvmhelpers")
    pkg.__path__ = [os.path.join(REF, "lvmhelpers")]  # the reference's own .py files
    sys.modules["lvmhelpers"] = pkg
    for mod, cls in (("pbernoulli", PFactorBernoulli), ("pbudget", PFactorBudget),
                     ("psequence", PFactorSequenceBinary)):
        m = types.ModuleType("lvmhelpers." + mod)
        setattr(m, cls.__name__, cls)
        sys.modules["lvmhelpers." + mod] = m
    # the reference's real compiled top-k module
    so = [f for f in os.listdir(os.path.join(ROOT, "oracle", "_ref")) if f.startswith("pbinary_topk.")
          and f.endswith(".so")][0]
    loader = importlib.machinery.ExtensionFileLoader("lvmhelpers.pbinary_topk",
                                                     os.path.join(ROOT, "oracle", "_ref", so))
    spec = importlib.util.spec_from_loader("lvmhelpers.pbinary_topk", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules["lvmhelpers.pbinary_topk"] = mod
    return mod


sys.path.insert(0, HERE)
from toy import toy_problem, make_modules, loss_fun, CatDecoder, BitDecoder  # noqa: E402


def tonp(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def main():
    ref_topk_mod = install_reference_package()
    out_k, out_w = {}, {}
    rng = np.random.default_rng(2024)

    # ---- 1. kernel-level fixtures --------------------------------------------------------
    for name, rows, K, scale in (("c1", 64, 10, 3.0), ("c2", 64, 256, 0.1), ("c5", 32, 128, 1.0),
                                 ("ties", 16, 16, 1.0)):
        x = (rng.standard_normal((rows, K)) * scale).astype(np.float32)
        if name == "ties":
            x = np.round(x * 2) / 2  # many exact ties / entries exactly at the threshold
        g = rng.standard_normal((rows, K)).astype(np.float32)
        P, k = oracle.sparsemax_fwd(torch.from_numpy(x))
        dx = oracle.sparsemax_bwd(P, k, torch.from_numpy(g))
        out_k.update({f"sparsemax_{name}_x": x, f"sparsemax_{name}_g": g, f"sparsemax_{name}_p": tonp(P),
                      f"sparsemax_{name}_k": tonp(k), f"sparsemax_{name}_dx": tonp(dx),
                      f"sparsemax_{name}_entropy": tonp(oracle.entropy(P))})
    for name, rows, n, k in (("c3", 64, 32, 10), ("n128", 16, 128, 10), ("small", 16, 5, 4), ("k32", 8, 40, 32)):
        s = rng.standard_normal((rows, n)).astype(np.float32)
        cfg = np.empty((rows, k, n), dtype=np.float32)
        ref_topk_mod.batched_topk(s, cfg, k)          # the reference's own compiled code
        out_k.update({f"topk_{name}_x": s, f"topk_{name}_cfg": cfg})
    s = rng.integers(-2, 3, (32, 12)).astype(np.float32)
    cfg = np.empty((32, 8, 12), dtype=np.float32)
    ref_topk_mod.batched_topk(s, cfg, 8)
    out_k.update({"topk_ties_x": s, "topk_ties_cfg": cfg})
    for name, n, kind, budget in (("bern32", 32, 0, 0), ("budget32", 32, 1, 16), ("budget32_3", 32, 1, 3),
                                  ("bern128", 128, 0, 0), ("budget128", 128, 1, 64), ("seq24", 24, 2, 0)):
        rows = 8
        x = rng.standard_normal((rows, n)).astype(np.float32)
        t = rng.standard_normal((rows, n - 1)).astype(np.float32) if kind == 2 else None
        cap = n + 1
        p = np.zeros((rows, cap), dtype=np.float64)
        aset = np.zeros((rows, cap, n), dtype=np.int8)
        cnt = np.zeros(rows, dtype=np.int32)
        st = np.zeros(rows, dtype=np.int32)
        dx = np.zeros((rows, n), dtype=np.float64)
        maps = np.zeros((rows, n), dtype=np.int8)
        for r in range(rows):
            res = oracle.sparsemap_fwd(x[r], kind, budget, t=None if t is None else t[r], max_iter=300,
                                       which="ref_factors")
            m = len(res["p"])
            p[r, :m], aset[r, :m], cnt[r], st[r] = res["p"], res["aset"], m, res["status"]
            w = np.cos(np.arange(m))
            dx[r], _ = oracle.sparsemap_bwd(x[r], w, kind, budget, t=None if t is None else t[r],
                                            max_iter=300, which="ref_factors")
            eta = np.concatenate([x[r], rng.standard_normal(n)]).astype(np.float64)
            y, _ = oracle.factor_map(eta, kind, budget, t=None if t is None else t[r], which="ref_factors")
            maps[r] = y
        out_k.update({f"smap_{name}_x": x, f"smap_{name}_p": p, f"smap_{name}_aset": aset,
                      f"smap_{name}_count": cnt, f"smap_{name}_status": st, f"smap_{name}_dx": dx})
        if t is not None:
            out_k[f"smap_{name}_t"] = t
    # the reference's print-only demo input (sparsemap.py:163-178): seed 42, randn(5), budget 3
    torch.manual_seed(42)
    xd = torch.randn(5).numpy()
    out_k["smap_demo_x"] = xd
    for nm, kind, budget in (("bern", 0, 0), ("budget3", 1, 3)):
        res = oracle.sparsemap_fwd(xd, kind, budget, max_iter=100, which="ref_factors")
        out_k[f"smap_demo_{nm}_p"], out_k[f"smap_demo_{nm}_aset"] = res["p"], res["aset"]
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out_k)

    # ---- 2. the reference's own wrappers ---------------------------------------------------
    from lvmhelpers.marg import ExplicitWrapper, Marginalizer          # /root/reference files
    from lvmhelpers.structmarg import (TopKSparsemaxWrapper, TopKSparsemaxMarg, SparseMAPWrapper,
                                       SparseMAPMarg)
    from lvmhelpers.utils import DeterministicWrapper

    def run(case, method, params, extra_logs):
        res = method(case["x"], case["x"], case["labels"])
        res["loss"].backward()
        out_w[f"{case['name']}_loss"] = tonp(res["loss"])
        for k, v in res["log"].items():
            if k in extra_logs or k in ("loss", "encoder_entropy", "support", "distr", "acc"):
                out_w[f"{case['name']}_log_{k}"] = tonp(v)
        for pname, prm in params.items():
            out_w[f"{case['name']}_grad_{pname}"] = tonp(prm.grad)

    # categorical marginalisation (config 1 shape: K = 10, B = 64)
    prob = toy_problem(1, 64, 20, 10, 6)
    enc, dec = make_modules(prob, categorical=True)
    m = Marginalizer(ExplicitWrapper(enc, normalizer="sparsemax"), DeterministicWrapper(CatDecoder(dec)),
                     loss_fun, encoder_entropy_coeff=0.1)
    run(dict(name="marg", **prob), m, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight}, ())

    # top-k sparsemax (config 3 shape: n = 32, k = 10, B = 64)
    prob = toy_problem(2, 64, 20, 32, 6)
    enc, dec = make_modules(prob, categorical=False)
    m = TopKSparsemaxMarg(TopKSparsemaxWrapper(enc, k=10), DeterministicWrapper(BitDecoder(dec)), loss_fun,
                          encoder_entropy_coeff=1.0)
    run(dict(name="topk", **prob), m, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight},
        ("loss_output",))

    # the same case with the reference's re-scoring einsum (structmarg.py:47) accumulated in fp64 and rounded
    # once: how far the REFERENCE's own results move under the rounding of that one fp32 contraction.  The GPU
    # path accumulates the re-scoring in fp64; tests/test_golden_gpu.py holds it to this variant tightly and to
    # the fp32 one within the band between the two.
    prob = toy_problem(2, 64, 20, 32, 6)
    enc, dec = make_modules(prob, categorical=False)
    m = TopKSparsemaxMarg(TopKSparsemaxWrapper(enc, k=10), DeterministicWrapper(BitDecoder(dec)), loss_fun,
                          encoder_entropy_coeff=1.0)
    einsum32 = torch.einsum
    torch.einsum = lambda eq, *ops: einsum32(eq, *[o.double() for o in ops]).float()
    try:
        run(dict(name="topk64", **prob), m, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight},
            ("loss_output",))
    finally:
        torch.einsum = einsum32

    # SparseMAP, bernoulli and budget 16, init off and on (config 4 shape: n = 32, B = 64)
    for nm, budget, init in (("smap_bern", 0, False), ("smap_budget", 16, False), ("smap_budget_init", 16, True)):
        prob = toy_problem(3, 64, 20, 32, 6)
        enc, dec = make_modules(prob, categorical=False)
        m = SparseMAPMarg(SparseMAPWrapper(enc, budget=budget, init=init, max_iter=300),
                          DeterministicWrapper(BitDecoder(dec)), loss_fun, encoder_entropy_coeff=1.0)
        torch.manual_seed(123)  # consumed by the init=True draws (sparsemap.py:26)
        run(dict(name=nm, **prob), m, {"enc_w": enc.weight, "enc_b": enc.bias, "dec": dec.weight},
            ("loss_output", "idxs"))
    np.savez_compressed(os.path.join(HERE, "wrappers.npz"), **out_w)
    print("wrote", os.path.join(HERE, "kernels.npz"), len(out_k), "arrays;",
          os.path.join(HERE, "wrappers.npz"), len(out_w), "arrays")


if __name__ == "__main__":
    main()
