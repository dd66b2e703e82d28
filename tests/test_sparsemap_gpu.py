"""GPU parity: SparseMAP active-set solver vs the CPU oracle (restated AD3) and vs
restatement-independent known answers (closed-form marginals, KKT)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BERN, BUDGET, SEQ = 0, 1, 2


def _solve_gpu(x, kind, budget=0, t=None, init=None, max_iter=300):
    from lvmhelpers import ops
    xs = torch.from_numpy(np.asarray(x, dtype=np.float32)).cuda()
    tt = None if t is None else torch.from_numpy(np.asarray(t, dtype=np.float32)).cuda()
    ini = None if init is None else torch.from_numpy(np.asarray(init, dtype=np.float64))
    st = ops.sparsemap_solve(xs, kind, budget=budget, t=tt, init_scores=ini, max_iter=max_iter)
    torch.cuda.synchronize()
    return st


def _unpack(st, r):
    n = st["n"]
    m = int(st["counts"][r])
    p = st["p_pad"][r, :m].cpu().numpy()
    b = st["bits"][r, :m].cpu().numpy().astype(np.uint32)
    idx = np.arange(n)
    aset = ((b[:, idx // 32] >> (idx % 32)) & 1).astype(np.int32)
    return p, aset


@pytest.mark.parametrize("n,kind,budget", [(32, BERN, 0), (32, BUDGET, 16), (32, BUDGET, 3), (5, BERN, 0),
                                           (5, BUDGET, 3), (128, BERN, 0), (128, BUDGET, 64), (1, BERN, 0),
                                           (33, BERN, 0), (64, BUDGET, 10), (100, BERN, 0)])
@pytest.mark.parametrize("scale", [0.3, 1.0, 3.0])
def test_forward_matches_oracle(oracle, n, kind, budget, scale):
    rows = 24 if n > 64 else 64
    rng = np.random.default_rng(n * 7 + kind + budget)
    x = (rng.standard_normal((rows, n)) * scale).astype(np.float32)
    st = _solve_gpu(x, kind, budget)
    flagged = 0
    for r in range(rows):
        ref = oracle.sparsemap_fwd(x[r], kind, budget, max_iter=300)
        p, aset = _unpack(st, r)
        assert int(st["status"][r]) == ref["status"]
        flagged += ref["status"] >= 2
        # active sets identical (same vertices, same order), p within 1e-5
        assert aset.shape == ref["aset"].shape, (r, aset.shape, ref["aset"].shape)
        assert np.array_equal(aset, ref["aset"])
        assert np.abs(p - ref["p"]).max() <= 1e-5
        assert int(st["iters"][r]) == ref["iters"]
        assert int(st["kept"][r]) == int((ref["p"].astype(np.float32) > 0).sum())
        # restatement-independent: marginals equal the closed form, p on the simplex
        u = aset.T.astype(np.float64) @ p.astype(np.float64)
        cf = oracle.bernoulli_marginals(x[r]) if kind == BERN else oracle.budget_marginals(x[r], budget)
        if ref["status"] == 0:
            # (1e-4: libad3's rank-one inverse updates drift over 100+ pivots -- the CPU
            #  oracle shows the same |sum p - 1| ~ 2e-5 on the long runs)
            assert np.abs(u - cf).max() <= 1e-4
            assert abs(p.sum() - 1) <= 1e-4 and p.min() >= -1e-7
        if kind == BUDGET:
            assert aset.sum(1).max() <= budget
    assert flagged == 0  # no repeated-config / singular exit among these few rows (they do occur: a few per 10^5
    # samples at n = 128, see test_n128_flagged_samples_match_oracle)


@pytest.mark.parametrize("n,kind,budget", [(32, BERN, 0), (32, BUDGET, 16), (128, BUDGET, 64)])
def test_forward_with_random_init(oracle, n, kind, budget):
    """init=True: seed vertex = MAP of torch.rand(2n, double) (sparsemap.py:25-27)."""
    rows = 32
    torch.manual_seed(42)
    init = torch.rand(rows, 2 * n, dtype=torch.double).numpy()
    x = torch.randn(rows, n).numpy()
    st = _solve_gpu(x, kind, budget, init=init)
    for r in range(rows):
        ref = oracle.sparsemap_fwd(x[r], kind, budget, init=init[r], max_iter=300)
        p, aset = _unpack(st, r)
        assert np.array_equal(aset, ref["aset"])
        assert np.abs(p - ref["p"]).max() <= 1e-5


def test_sequence_binary_matches_oracle(oracle):
    rows, n = 48, 24
    rng = np.random.default_rng(3)
    x = rng.standard_normal((rows, n)).astype(np.float32)
    t = rng.standard_normal((rows, n - 1)).astype(np.float32)
    st = _solve_gpu(x, SEQ, t=t, max_iter=100)
    for r in range(rows):
        ref = oracle.sparsemap_fwd(x[r], SEQ, t=t[r], max_iter=100)
        p, aset = _unpack(st, r)
        assert np.array_equal(aset, ref["aset"])
        assert np.abs(p - ref["p"]).max() <= 1e-5


def test_max_iter_cutoff(oracle):
    """Small max_iter: same truncated (p, aset), status = max_iter, zeros allowed in p."""
    rows, n = 32, 32
    x = np.random.default_rng(11).standard_normal((rows, n)).astype(np.float32)
    for mi in (1, 2, 5, 10):
        st = _solve_gpu(x, BERN, max_iter=mi)
        for r in range(rows):
            ref = oracle.sparsemap_fwd(x[r], BERN, max_iter=mi)
            p, aset = _unpack(st, r)
            assert np.array_equal(aset, ref["aset"]) and np.abs(p - ref["p"]).max() <= 1e-5
            assert int(st["status"][r]) == ref["status"]


@pytest.mark.parametrize("n,kind,budget", [(32, BERN, 0), (32, BUDGET, 16), (128, BERN, 0)])
def test_backward_matches_oracle(oracle, n, kind, budget):
    from lvmhelpers import ops
    rows = 16
    rng = np.random.default_rng(n + kind)
    x = rng.standard_normal((rows, n)).astype(np.float32)
    xs = torch.from_numpy(x).cuda().requires_grad_(True)
    p, aset, aux = ops.sparsemap_batch(xs, kind, budget=budget, max_iter=300, only_positive=False)
    w = torch.randn(p.shape[0], generator=torch.Generator().manual_seed(0))
    (p * w.cuda()).sum().backward()
    sizes = aux["sizes"].tolist()
    off = 0
    for r in range(rows):
        dx_ref, _ = oracle.sparsemap_bwd(x[r], w[off:off + sizes[r]].numpy(), kind, budget, max_iter=300)
        off += sizes[r]
        assert np.abs(xs.grad[r].cpu().numpy() - dx_ref).max() <= 1e-5 * max(1.0, np.abs(dx_ref).max())


def test_backward_sequence_and_finite_differences(oracle):
    from lvmhelpers.sparsemap import sequence_smap, bernoulli_smap
    n = 12
    torch.manual_seed(0)
    x = torch.randn(n, device="cuda", requires_grad=True)
    t = torch.randn(n - 1, device="cuda", requires_grad=True)
    p, aset = sequence_smap(x, t, max_iter=100, init=False)
    w = torch.randn(p.shape[0], device="cuda")
    (p * w).sum().backward()
    dx_ref, dt_ref = oracle.sparsemap_bwd(x.detach().cpu().numpy(), w.cpu().numpy(), 2,
                                          t=t.detach().cpu().numpy(), max_iter=100)
    assert np.abs(x.grad.cpu().numpy() - dx_ref).max() <= 1e-5
    assert np.abs(t.grad.cpu().numpy() - dt_ref).max() <= 1e-5
    # finite differences on the (unique) marginals u = aset' p for the Bernoulli factor
    x0 = torch.randn(8, dtype=torch.float32)
    xg = x0.cuda().requires_grad_(True)
    p, aset = bernoulli_smap(xg, max_iter=100, init=False)
    c = torch.randn(8, device="cuda")
    ((aset.t() @ p) * c).sum().backward()
    inside = (x0.abs() < 1).float()  # du_i/dx_i = 1/2 strictly inside the box, 0 when clipped
    assert torch.allclose(xg.grad.cpu(), 0.5 * inside * c.cpu(), atol=1e-5)


def test_structured_api_matches_reference_loop(oracle):
    """SparseMAPWrapper output == the reference's per-sample loop (structmarg.py:181-202)
    run over the oracle solver: ragged order, p > 0 filter, idxs, support, entropy."""
    from lvmhelpers.structmarg import SparseMAPWrapper
    B, n = 64, 32
    torch.manual_seed(42)
    x = torch.randn(B, n)
    for budget, init in [(0, False), (16, False), (0, True), (16, True)]:
        wrap = SparseMAPWrapper(torch.nn.Identity(), budget=budget, init=init, max_iter=300)
        torch.manual_seed(7)
        sample, distr, ent, idxs, support = wrap(x.cuda())
        torch.manual_seed(7)
        ref_p, ref_s, ref_idx, ref_supp = [], [], [], []
        for k in range(B):
            ini = torch.rand(2 * n, dtype=torch.double).numpy() if init else None
            r = oracle.sparsemap_fwd(x[k].numpy(), 1 if budget > 0 else 0, budget, init=ini, max_iter=300)
            pf = r["p"].astype(np.float32)
            keep = pf > 0
            ref_p.append(pf[keep]); ref_s.append(r["aset"][keep].astype(np.float32))
            ref_idx += [k] * int(keep.sum()); ref_supp.append(int(keep.sum()))
        ref_p = np.concatenate(ref_p); ref_s = np.concatenate(ref_s)
        assert list(idxs) == ref_idx and support == ref_supp
        assert np.array_equal(sample.cpu().numpy(), ref_s)
        assert np.abs(distr.detach().cpu().numpy() - ref_p).max() <= 1e-5
        ref_ent = -(ref_p.astype(np.float64) @ np.log(ref_p.astype(np.float64))) / B
        assert abs(ent.item() - ref_ent) <= 1e-5 * abs(ref_ent)


def test_demo_vector(oracle):
    """The reference's print-only smoke demo (sparsemap.py:163-178): seed 42, randn(5)."""
    from lvmhelpers.sparsemap import bernoulli_smap, budget_smap
    torch.manual_seed(42)
    x = torch.randn(5)
    xg = x.cuda().requires_grad_(True)
    p, aset = bernoulli_smap(xg, init=False)
    ref = oracle.sparsemap_fwd(x.numpy(), 0, max_iter=100)
    assert np.array_equal(aset.cpu().numpy(), ref["aset"].astype(np.float32))
    assert np.abs(p.detach().cpu().numpy() - ref["p"]).max() <= 1e-5
    p, aset = budget_smap(xg, budget=3, init=False)
    ref = oracle.sparsemap_fwd(x.numpy(), 1, 3, max_iter=100)
    assert np.array_equal(aset.cpu().numpy(), ref["aset"].astype(np.float32))
    g0 = torch.autograd.grad(p[0], xg, retain_graph=True)[0]
    e0 = np.zeros(len(ref["p"])); e0[0] = 1
    dx_ref, _ = oracle.sparsemap_bwd(x.numpy(), e0, 1, 3, max_iter=100)
    assert np.abs(g0.cpu().numpy() - dx_ref).max() <= 1e-5


@pytest.mark.parametrize("lanes", [1, 2])
@pytest.mark.parametrize("n,kind,budget", [(32, BERN, 0), (33, BUDGET, 9), (20, 2, 0)])
def test_arena_of_inverses(n, kind, budget, lanes):
    """The inverses of the backward are bump-allocated from one arena (8|A|^2 bytes per sample instead of the padded
    8(n+1)^2), also when the batch is split over two streams: same forward state and the same backward whatever the
    arena layout; a too-small arena is reported through the cursor, the samples that did not fit get offset -1 and
    NaN gradients, and re-solving with the reported size fits exactly."""
    from lvmhelpers import ops
    rows = 101
    g = torch.Generator().manual_seed(n)
    x = torch.randn(rows, n, generator=g).cuda()
    t = torch.randn(rows, n - 1, generator=g).cuda() if kind == 2 else None
    ref = ops.sparsemap_solve(x, kind, budget=budget, t=t, max_iter=100)     # worst-case arena, one launch sequence
    m = ref["counts"].long()
    need = int((m * m).sum())
    assert ops.sparsemap_arena_overflow(ref) == 0 and int(ref["S_cursor"]) == need
    assert int(ref["S_offsets"].min()) >= 0
    old = (ops.SMAP_SPLIT_MIN_ROWS, ops.SMAP_CHUNK_LANES)
    ops.SMAP_SPLIT_MIN_ROWS, ops.SMAP_CHUNK_LANES = 8, lanes
    try:
        st = ops.sparsemap_solve(x, kind, budget=budget, t=t, max_iter=100, arena_doubles=need)     # exact fit
        small = ops.sparsemap_solve(x, kind, budget=budget, t=t, max_iter=100, arena_doubles=need // 2)
    finally:
        ops.SMAP_SPLIT_MIN_ROWS, ops.SMAP_CHUNK_LANES = old
    for key in ("p_pad", "bits", "counts", "kept", "iters", "status"):
        assert torch.equal(st[key], ref[key]) and torch.equal(small[key], ref[key]), key
    assert ops.sparsemap_arena_overflow(st) == 0 and int(st["S_offsets"].min()) >= 0
    # every sample owns a disjoint |A|^2 slice
    o = st["S_offsets"].cpu().numpy(); mm = (m * m).cpu().numpy()
    order = np.argsort(o)
    assert np.all(o[order][1:] >= (o[order] + mm[order])[:-1]) and o.max() + mm[np.argmax(o)] <= need
    offsets = ops.scan_counts(ref["kept"])
    total = int(offsets[-1])
    dp = torch.randn(total, generator=g).cuda()
    dx_ref, dt_ref = ops.sparsemap_bwd(ref, dp, offsets, True, want_dt=kind == 2)
    dx, dt = ops.sparsemap_bwd(st, dp, offsets, True, want_dt=kind == 2)
    assert torch.equal(dx, dx_ref)
    if kind == 2:
        assert torch.equal(dt, dt_ref)
    # overflow: reported, flagged per sample, loud in the backward
    assert ops.sparsemap_arena_overflow(small) == need
    lost = small["S_offsets"] < 0
    assert 0 < int(lost.sum()) < rows
    dx_s, _ = ops.sparsemap_bwd(small, dp, offsets, True, want_dt=kind == 2)
    assert torch.isnan(dx_s[lost]).all() and torch.equal(dx_s[~lost], dx_ref[~lost])


def test_batch_autograd_resolves_on_arena_overflow():
    """ops.sparsemap_batch: when the estimated arena is too small the batch is solved again with the exact size."""
    from lvmhelpers import ops
    x = torch.randn(300, 32, generator=torch.Generator().manual_seed(3)).cuda().requires_grad_(True)
    p, aset, aux = ops.sparsemap_batch(x, BERN, max_iter=100)
    p.sum().backward()
    g_ref = x.grad.clone()
    old = ops.sparsemap_arena_doubles
    ops.sparsemap_arena_doubles = lambda rows, n, kind, budget=0: 64        # far too small
    try:
        x.grad = None
        p2, aset2, aux2 = ops.sparsemap_batch(x, BERN, max_iter=100)
        p2.sum().backward()
    finally:
        ops.sparsemap_arena_doubles = old
    assert torch.equal(p2, p) and torch.equal(aset2, aset) and torch.equal(x.grad, g_ref)
    assert int(aux2["state"]["S_offsets"].min()) >= 0


@pytest.mark.parametrize("kind,budget,lo_hi_maxiter", [(BUDGET, 64, (0.03, 0.09)), (BERN, 0, (0.0, 0.005))])
def test_n128_flagged_samples_match_oracle(oracle, kind, budget, lo_hi_maxiter):
    """The samples that do NOT end in plain convergence at the benchmark shape (n = 128, max_iter = 300,
    N(0,1) scores, structmarg.py:172-196): every sample with status >= 2 (repeated configuration / singular
    insertion) of a 131072-sample batch, and every max_iter sample of its first 16384 rows, compared with the
    oracle in full -- (p, aset, iters, status).  Round 1 measured status_hist = [61844, 3690, 2, 0] on 65536
    budget-64 samples and [65498, 38, 0, 0] for Bernoulli: about 5.6 % / 0.06 % of the samples stop at max_iter
    and a handful per 10^5 stop because the oracle returns a vertex that is already active (libad3 would clear
    its cache there and the reference trip `assert supp > 0`, structmarg.py:192; we keep the solution -- the
    oracle restates exactly that, oracle/ad3_min/GenericFactor.cpp:9-16).  First run of THIS test on the B200, 131072
    budget-64 samples: [123714, 7348, 7, 3] -- three singular insertions as well."""
    rows, head = 131072, 16384
    g = torch.Generator().manual_seed(20261020 + kind)
    x = torch.randn(rows, 128, generator=g)
    from lvmhelpers import ops
    st = ops.sparsemap_solve(x.cuda(), kind, budget=budget, max_iter=300, keep_S=False)
    status = st["status"].cpu().numpy()
    hist = np.bincount(status, minlength=4)
    # singular insertions (|Schur complement| < 1e-9) do occur at the budget factor: 3 of 131072 samples when this
    # test was first run on the B200.  libad3 evicts a vertex there through an eigendecomposition (third-party
    # code, absent here); kernel and oracle both stop with the current iterate and flag the sample
    assert hist.sum() == rows and hist[3] <= 20, hist
    assert lo_hi_maxiter[0] * rows <= hist[1] <= lo_hi_maxiter[1] * rows, hist
    assert hist[2] <= 40, hist                                      # repeated configurations: a few per 10^5
    pick = np.concatenate([np.nonzero(status >= 2)[0], np.nonzero(status[:head] == 1)[0]])
    pick = pick[:1500]
    assert len(pick) > 0
    xn = x.numpy()
    for r in pick:
        ref = oracle.sparsemap_fwd(xn[r], kind, budget, max_iter=300)
        p, aset = _unpack(st, int(r))
        assert int(status[r]) == ref["status"], (r, status[r], ref["status"])
        assert int(st["iters"][r]) == ref["iters"]
        assert aset.shape == ref["aset"].shape and np.array_equal(aset, ref["aset"]), r
        assert np.abs(p - ref["p"]).max() <= 1e-5
        assert int(st["kept"][r]) == int((ref["p"].astype(np.float32) > 0).sum()) >= 1
    # and the cheap invariants on ALL samples: a max_iter / repeated-config sample still carries a usable
    # distribution (what SparseMAPWrapper hands on after dropping p <= 0, structmarg.py:189-192)
    assert int(st["kept"].min()) >= 1
    cnt = st["counts"].cpu().numpy()
    assert cnt.min() >= 1 and cnt.max() <= 129


@pytest.mark.parametrize("budget", [0, 16])
def test_static_shape_wrapper_equals_ragged_and_replays_as_graph(budget):
    """SparseMAPWrapper(static_shapes=True): the reference's ragged outputs in sample order followed by zero padding
    (N = B (n + 1)), device-backed idxs / support, no host synchronisation -- same loss, entropy and gradients as the
    ragged wrapper through SparseMAPMarg, and the whole step replays as ONE CUDA graph."""
    from lvmhelpers.structmarg import SparseMAPWrapper, SparseMAPMarg
    B, n = 64, 32
    torch.manual_seed(3)
    enc = torch.nn.Linear(20, n).cuda()
    dec = torch.nn.Linear(n, 12).cuda()
    x = torch.randn(B, 20, device="cuda")
    labels = torch.randn(B, 12, device="cuda")

    def loss_fun(enc_in, z, dec_in, dec_out, lab):
        return ((dec_out - lab) ** 2).sum(-1), {"acc": dec_out.mean(-1)}

    params = list(enc.parameters()) + list(dec.parameters())
    outs = {}
    for static in (False, True):
        for q in params:
            q.grad = None
        marg = SparseMAPMarg(SparseMAPWrapper(enc, budget=budget, max_iter=300, static_shapes=static), dec, loss_fun,
                             encoder_entropy_coeff=0.7)
        res = marg(x, x, labels)
        res["loss"].backward()
        outs[static] = (res, [q.grad.clone() for q in params])
    (r0, g0), (r1, g1) = outs[False], outs[True]
    N = len(r0["log"]["idxs"])
    assert r1["log"]["distr"].numel() == B * (n + 1)
    assert torch.equal(r1["log"]["distr"][:N], r0["log"]["distr"]) and bool((r1["log"]["distr"][N:] == 0).all())
    assert list(r1["log"]["idxs"]) == list(r0["log"]["idxs"])
    assert torch.equal(r1["log"]["support"].cpu(), r0["log"]["support"].cpu())
    assert abs(r1["loss"].item() - r0["loss"].item()) <= 1e-5 * max(1.0, abs(r0["loss"].item()))
    for a, b in zip(g0, g1):
        assert (a - b).abs().max().item() <= 1e-5 * max(1.0, a.abs().max().item())

    # one CUDA graph for forward + backward (fresh copies of the nets: autograd nodes of the eager passes above, made
    # on the default stream and still referenced by their results, would pull that stream into the capture)
    import copy
    del outs, r0, r1, res
    enc_e, dec_e = enc, dec
    enc, dec = copy.deepcopy(enc_e), copy.deepcopy(dec_e)
    params = list(enc.parameters()) + list(dec.parameters())
    marg = SparseMAPMarg(SparseMAPWrapper(enc, budget=budget, max_iter=300, static_shapes=True), dec, loss_fun,
                         encoder_entropy_coeff=0.7)
    xs = x.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            for q in params:
                q.grad = None
            marg(xs, xs, labels)["loss"].backward()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    for q in params:
        q.grad = None
    with torch.cuda.graph(graph):
        out = marg(xs, xs, labels)
        out["loss"].backward()
    xs.copy_(torch.randn(B, 20, device="cuda"))
    graph.replay()
    torch.cuda.synchronize()
    got_loss, got_grads = out["loss"].item(), [q.grad.clone() for q in params]
    params_e = list(enc_e.parameters()) + list(dec_e.parameters())
    for q in params_e:
        q.grad = None
    ref = SparseMAPMarg(SparseMAPWrapper(enc_e, budget=budget, max_iter=300), dec_e, loss_fun, encoder_entropy_coeff=0.7)(xs, xs, labels)
    ref["loss"].backward()
    assert abs(got_loss - ref["loss"].item()) <= 1e-5 * max(1.0, abs(ref["loss"].item()))
    for a, q in zip(got_grads, params_e):
        assert (a - q.grad).abs().max().item() <= 1e-5 * max(1.0, q.grad.abs().max().item())


SEQ_SHAPES = [[3, 4, 2, 5, 3], [2] * 12, [4] * 8, [1, 3, 1, 2], [32, 32, 32, 32], [5], [6, 2, 7, 3, 4, 4, 2, 8], [3] * 40]


@pytest.mark.parametrize("num_states", SEQ_SHAPES)
def test_general_sequence_factor_matches_oracle(oracle, num_states):
    """The general chain factor (sequence.h:25-291, PFactorSequence): batched solve against the oracle's restatement
    (pinned to the reference's own header compiled in oracle/_ref, tests/test_oracle_cpu.py): identical active sets in
    the solver's order, p within 1e-5, iterations and status; backward (d x, d edge scores) against
    dist_jacobian_vec of the oracle."""
    from lvmhelpers import ops
    V, E, L = sum(num_states), oracle.seq_num_edges(num_states), len(num_states)
    rng = np.random.default_rng(V * 7 + L)
    rows = 24
    x = rng.standard_normal((rows, V)).astype(np.float32)
    t = (rng.standard_normal((rows, E)) * 0.5).astype(np.float32)
    x[3] = np.round(x[3] * 2) / 2          # ties in the Viterbi recursion
    t[3] = 0.0
    state = ops.sparsemap_solve_sequence(torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), num_states, max_iter=200)
    offsets = ops.scan_counts(state["counts"])
    total = int(offsets[-1])
    p, onehot, idx = ops.sparsemap_expand(state, False, offsets, total)
    aset = ops.onehot_to_states(onehot, num_states).cpu().numpy().astype(np.int32)
    assert bool((onehot.sum(-1) == L).all())
    p, off = p.cpu().numpy(), offsets.cpu().numpy()
    g = torch.Generator().manual_seed(1)
    dp = torch.randn(total, generator=g)
    dx, de = ops.sparsemap_bwd_sequence(state, dp.cuda(), offsets, False)
    for r in range(rows):
        ref = oracle.seq_fwd(x[r], t[r], num_states, max_iter=200)
        sl = slice(off[r], off[r + 1])
        assert state["counts"][r].item() == len(ref["p"])
        assert np.array_equal(aset[sl], ref["aset"])
        assert np.abs(p[sl] - ref["p"]).max() <= 1e-5
        assert state["iters"][r].item() == ref["iters"] and state["status"][r].item() == ref["status"]
        dx_ref, de_ref = oracle.seq_bwd(x[r], t[r], num_states, dp[sl].numpy(), max_iter=200)
        scale = max(1.0, np.abs(dx_ref).max(), np.abs(de_ref).max())
        assert np.abs(dx[r].cpu().numpy() - dx_ref).max() <= 1e-4 * scale
        assert np.abs(de[r].cpu().numpy() - de_ref).max() <= 1e-4 * scale


def test_general_sequence_autograd_function_and_binary_equivalence(oracle):
    """general_sequence_smap: (p, state sequences) + gradients; with two states per position, zero start / stop edges
    and a score on 1 -> 1 only it is the binary chain of sequence_smap (sparsemap.py:130-148)."""
    from lvmhelpers.sparsemap import general_sequence_smap, sequence_smap
    torch.manual_seed(5)
    n = 9
    x = torch.randn(n)
    t = torch.randn(n - 1)
    p_b, aset_b = sequence_smap(x.cuda(), t.cuda(), max_iter=200, init=False)
    ns = [2] * n
    xg = torch.zeros(2 * n)
    xg[1::2] = x
    add = torch.zeros(oracle.seq_num_edges(ns))
    for i in range(1, n):
        add[2 + 4 * (i - 1) + 3] = t[i - 1]       # edge (previous = 1, current = 1) of position i
    xg = xg.cuda().requires_grad_(True)
    add = add.cuda().requires_grad_(True)
    p_g, aset_g = general_sequence_smap(xg, add, ns, max_iter=200)
    assert torch.equal(aset_g, aset_b)
    assert (p_g.detach() - p_b).abs().max().item() <= 1e-6
    w = torch.randn(p_g.numel(), device="cuda")
    (p_g * w).sum().backward()
    xb = x.cuda().requires_grad_(True)
    tb = t.cuda().requires_grad_(True)
    p_b2, _ = sequence_smap(xb, tb, max_iter=200, init=False)
    (p_b2 * w).sum().backward()
    # the binary chain scores "off" states 0: its d x is the gradient of the "on" node scores
    assert (xg.grad[1::2] - xb.grad).abs().max().item() <= 1e-5
    got_t = torch.stack([add.grad[2 + 4 * (i - 1) + 3] for i in range(1, n)])
    assert (got_t - tb.grad).abs().max().item() <= 1e-5
