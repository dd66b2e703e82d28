"""Wrapper-level latency of the four experiment shapes (BASELINE configs 1-4, B = 64): one
forward + backward of the drop-in Method with toy encoder/decoder nets (tests/golden/toy.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sparse-marginalization-lvm_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from toy import toy_problem, make_modules, loss_fun, CatDecoder, BitDecoder
from lvmhelpers.marg import ExplicitWrapper, Marginalizer
from lvmhelpers.structmarg import TopKSparsemaxWrapper, TopKSparsemaxMarg, SparseMAPWrapper, SparseMAPMarg
from lvmhelpers.utils import DeterministicWrapper


def bench(name, make):
    m, x, labels, params = make()
    def step():
        for p in params: p.grad = None
        res = m(x, x, labels)
        res["loss"].backward()
    for _ in range(10): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n): step()
    torch.cuda.synchronize()
    print("%-34s %.3f ms per fwd+bwd (host wall clock, B=64)" % (name, (time.perf_counter() - t0) / n * 1e3), flush=True)


def cat(latent, compact):
    def f():
        prob = toy_problem(1, 64, 20, latent, 6)
        enc, dec = make_modules(prob, True)
        enc, dec = enc.cuda(), dec.cuda()
        m = Marginalizer(ExplicitWrapper(enc, normalizer="sparsemax"), DeterministicWrapper(CatDecoder(dec)), loss_fun,
                         encoder_entropy_coeff=0.1, compact_support=compact)
        return m, prob["x"].cuda(), prob["labels"].cuda(), list(enc.parameters()) + list(dec.parameters())
    return f


def topk():
    prob = toy_problem(2, 64, 20, 32, 6)
    enc, dec = make_modules(prob, False)
    enc, dec = enc.cuda(), dec.cuda()
    m = TopKSparsemaxMarg(TopKSparsemaxWrapper(enc, k=10), DeterministicWrapper(BitDecoder(dec)), loss_fun, encoder_entropy_coeff=1.0)
    return m, prob["x"].cuda(), prob["labels"].cuda(), list(enc.parameters()) + list(dec.parameters())


def smap(budget, init):
    def f():
        prob = toy_problem(3, 64, 20, 32, 6)
        enc, dec = make_modules(prob, False)
        enc, dec = enc.cuda(), dec.cuda()
        m = SparseMAPMarg(SparseMAPWrapper(enc, budget=budget, init=init, max_iter=300), DeterministicWrapper(BitDecoder(dec)),
                          loss_fun, encoder_entropy_coeff=1.0)
        return m, prob["x"].cuda(), prob["labels"].cuda(), list(enc.parameters()) + list(dec.parameters())
    return f


bench("C1 marg sparsemax K=10", cat(10, False))
bench("C1 marg sparsemax K=10 compact", cat(10, True))
bench("C2 marg sparsemax K=256", cat(256, False))
bench("C2 marg sparsemax K=256 compact", cat(256, True))
bench("C3 top-k sparsemax n=32 k=10", topk)
bench("C4 SparseMAP bernoulli n=32", smap(0, False))
bench("C4 SparseMAP budget16 n=32 init", smap(16, True))
