"""GPU tests of the round-2 additions: compact count read-out, read-out-only contexts (any n), deterministic gene
sums, per-gene geneMeanLevels, duplicate cell ids, the cached context of the MatriSeq module, Poisson(T) member steps in
the prepared mode, and the extension flags of the read-out (negative-binomial draws, dropout)."""
import random

import numpy as np
import pytest

from conftest import seed_all
from oracle import fast_oracle as fo

pytestmark = pytest.mark.gpu


def sparse_net(n, seed, gain=1.0):
    rng = np.random.RandomState(seed)
    W = np.zeros((n, n))
    regs = rng.choice(n, size=max(2, n // 3), replace=False)
    deg = np.minimum((0.5 * (1 - rng.rand(n)) ** (-1 / 1.05) + 0.5).astype(int), len(regs))
    for r in range(n):
        W[r, rng.choice(regs, size=deg[r], replace=False)] = rng.beta(2, 3, size=deg[r]) * rng.choice([-1, 1])
    return W * gain


def engine(W, cells, prec="f32", offset=0):
    from scrutiny_b200 import engine as em
    e = em.Engine(0, prec)
    e.set_network(W)
    e.alloc_cells(cells, offset)
    return e


# ---------------------------------------------------------------------------------------------
# compact read-out
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_compact_counts_are_the_int32_counts(prec):
    """uint8 / uint16 rows + overflow list widen to exactly the integers scr_counts_range returns (same Philox stream);
    rates from 1e-3 to ~3000 so that both escape codes are exercised; row blocks, device output, capacity retry."""
    import torch
    from scrutiny_b200 import artefacts
    n, cells = 200, 777
    rng = np.random.RandomState(31)
    e = engine(sparse_net(n, 5), cells, prec, offset=2 ** 33 + 1)
    e.set_states(np.tanh(rng.randn(n, cells)))
    shift = np.concatenate([np.log(rng.gamma(0.35, 1 / 0.19, size=n - 40)), np.linspace(3.0, 8.0, 40)])[rng.permutation(n)]
    size = rng.normal(1, 0.14, size=cells) ** 3
    want = e.counts(shift, size, seed=77)
    assert (want >= 255).sum() > 500 and (want >= 65535).sum() == 0 and want.max() > 1000
    # the float32-screened sampler (small-lambda and PTRS screens, default) against the float64 sampler alone
    import os
    os.environ["SCR_COUNTS_EXACT"] = "1"
    try:
        assert np.array_equal(e.counts(shift, size, seed=77), want)
        assert np.array_equal(e.counts_compact(shift, size, seed=77, dtype="uint8")[0], np.minimum(want, 255))
    finally:
        del os.environ["SCR_COUNTS_EXACT"]
    for dtype, esc in (("uint8", 255), ("uint16", 65535)):
        got, ovf = e.counts_compact(shift, size, seed=77, dtype=dtype)
        assert got.dtype == np.dtype(dtype) and ovf.shape[1] == 3
        assert np.array_equal(got, np.minimum(want, esc))
        rows, genes = np.nonzero(want >= esc)                               # row-major order = sorted by (row, gene)
        assert np.array_equal(ovf, np.stack([rows, genes, want[rows, genes]], axis=1))
        m = artefacts.CountMatrix([got], [ovf])
        assert np.array_equal(np.asarray(m), want.T.astype(np.float64))
    # a capacity that is too small is reported, and the engine wrapper retries with the reported size
    small, ovf_small = e.counts_compact(shift, size, seed=77, dtype="uint8", overflow_capacity=3)
    full, ovf_full = e.counts_compact(shift, size, seed=77, dtype="uint8")
    assert np.array_equal(small, full) and np.array_equal(ovf_small, ovf_full) and len(ovf_full) > 3
    # ragged row blocks into one device buffer
    dev = torch.zeros((cells, n), dtype=torch.uint8, device="cuda:0")
    parts = []
    for r0, r1 in ((0, 1), (1, 300), (300, 777)):
        _, o = e.counts_compact(shift, size[r0:r1], seed=77, dtype="uint8", cell_begin=r0, cell_count=r1 - r0,
                                out_device_ptr=dev.data_ptr() + n * r0)
        o[:, 0] += r0
        parts.append(o)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), full) and np.array_equal(np.concatenate(parts), ovf_full)


def test_transform_data_compact_equals_dense(golden):
    import scrutiny_b200.MatriSeq as ms
    from scrutiny_b200 import artefacts
    S = golden("pipeline")["final_states"]
    n, cells = S.shape
    outs = []
    for compact in (False, True):
        ms.init(outputDir=None, seed=11, verbose=False)
        counts, size = ms.transformData(n, cells, S.copy(), compact=compact)
        outs.append((counts, size))
    assert isinstance(outs[0][0], np.ndarray) and isinstance(outs[1][0], artefacts.CountMatrix)
    assert np.array_equal(outs[0][0], np.asarray(outs[1][0])) and np.array_equal(outs[0][1], outs[1][1])


# ---------------------------------------------------------------------------------------------
# read-out context without a network: any n
# ---------------------------------------------------------------------------------------------
def test_readout_context_takes_a_gene_count_beyond_the_steppers_budget():
    """n = 20 000 (the reference's driver mentions n = 19 027): transformData needs no network at all, so the read-out
    context takes any n; the stepper takes it too, through its global-memory fallback."""
    from scrutiny_b200 import engine as em
    n, cells = 20000, 24
    rng = np.random.RandomState(8)
    S = np.tanh(rng.randn(n, cells))
    e = em.Engine(0, "f64")
    e.set_genes(n)
    e.alloc_cells(cells)
    e.set_states(S)
    assert np.array_equal(e.get_states(), S)
    assert np.allclose(e.gene_sums(), S.sum(axis=1), rtol=1e-13, atol=1e-12)
    shift = rng.normal(0.0, 1.5, size=n)
    size = rng.normal(1, 0.14, size=cells) ** 3
    counts, lam = e.counts(shift, size, seed=3, want_lambda=True)
    assert np.allclose(lam, fo.poisson_lambda(S, shift, 1.0, size).T, rtol=1e-14)
    got, ovf = e.counts_compact(shift, size, seed=3, dtype="uint8")
    assert np.array_equal(np.minimum(counts, 255), got)
    with pytest.raises(em.ScrutinyError) as err:
        e.run_phase([0], [1], np.ones(n), np.zeros(n))
    assert err.value.code == -3
    big = em.Engine(0, "f32")                                             # the stepper takes this n with its groups in global memory
    big.set_network(csr=(np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), np.ones(n)), n=n)
    assert big.layout()["state_in_global"] == 1
    huge = 40000                                                          # beyond 32 768 genes: a clear error
    with pytest.raises(em.ScrutinyError) as err:
        big.set_network(csr=(np.arange(huge + 1, dtype=np.int32), np.arange(huge, dtype=np.int32), np.ones(huge)), n=huge)
    assert err.value.code == -4 and "n too large" in str(err.value)


def test_gene_sums_are_deterministic():
    n, cells = 300, 50000
    e = engine(sparse_net(n, 2), cells)
    rng = np.random.RandomState(1)
    e.set_states_broadcast(rng.uniform(-1, 1, size=n))
    e.run_phase(np.arange(cells), rng.randint(1, 6, size=cells), np.ones(n), np.zeros(n), seed=4)
    a, b = e.gene_sums(), e.gene_sums()
    assert np.array_equal(a, b)
    assert np.allclose(a, e.get_states().sum(axis=1), rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# stepping
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,prec,tol", [("sparse", "f64", 1e-10), ("sparse", "f32", 1e-4), ("dense", "f32", 1e-4)])
def test_per_gene_mean_levels_vector(kind, prec, tol):
    """developCell broadcasts ``x - geneMeanLevels`` (MS:530): a per-gene vector must work like in numpy."""
    rng = np.random.RandomState(41)
    n, cells, steps = 260, 12, 30
    if kind == "dense":
        W = rng.beta(2, 3, size=(n, n)) * rng.choice([-1.0, 1.0], size=(n, 1)) * (rng.rand(n, n) < 0.4) / 6.0
    else:
        W = sparse_net(n, 9, gain=4.0)
    mu = 0.3 * rng.randn(n)
    X0 = rng.uniform(-1, 1, size=(cells, n)) + mu
    noise = (0.05 * rng.randn(cells, steps, n)).astype(np.float32).astype(np.float64)
    proj = (rng.rand(n) > 0.1).astype(float)
    const = np.where(proj == 0, 0.5, 0.0)
    st = rng.randint(0, steps + 1, size=cells)
    e = engine(W, cells, prec)
    assert e.layout()["dense_stepper"] == int(kind == "dense")
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), st, proj, const, mu=mu, noise=noise)
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    fo.run_phase(fo.as_operator(W), want, st, proj, const, 0.01, noise, mu=mu)
    assert np.max(np.abs(e.get_states().T - want)) <= tol
    # a uniform vector is the scalar; a vector of the wrong length is a clear error
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), st, proj, const, mu=np.full(n, 0.2), noise=noise)
    a = e.get_states()
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), st, proj, const, mu=0.2, noise=noise)
    assert np.array_equal(a, e.get_states())
    with pytest.raises(ValueError):
        e.run_phase([0], [1], proj, const, mu=np.zeros(n + 1))
    if kind == "sparse":
        it_vec = e.probe(np.arange(4), proj, const, mu=mu, max_iter=40, noise=np.zeros((4, 40, n)))
        want_it, _, _ = fo.probe(fo.as_operator(W), e.get_states()[:, :4].T.copy(), proj, const, 0.01, np.zeros((4, 40, n)),
                                 0.1, 40, 20, mu=mu)
        assert np.max(np.abs(it_vec - want_it)) <= (0 if prec == "f64" else 2)


def test_duplicate_cell_ids_are_refused():
    from scrutiny_b200 import engine as em
    n = 64
    e = engine(sparse_net(n, 3), 9)
    with pytest.raises(em.ScrutinyError) as err:
        e.run_phase([1, 5, 1], [3, 3, 3], np.ones(n), np.zeros(n))
    assert err.value.code == -1 and "twice" in str(err.value)


def test_module_keeps_the_context_between_calls():
    """developCell / findCellNextState per cell (MS:503-537, 683-736) reuse the packed operator; an edited matrix does not."""
    import scrutiny_b200.MatriSeq as ms
    ms.init(outputDir=None, seed=5, verbose=False, precision="f64")
    n = 90
    gc = sparse_net(n, 12, gain=3.0)
    x = np.random.RandomState(0).uniform(-1, 1, size=n)
    a = ms.developCell(n, gc, 0.01, x, np.ones(n), np.zeros(n), 0.0, 0)
    eng = ms._engine_cache[(0, "f64")][1]
    b = ms.findCellNextState(gc, n, 0.01, 0, np.ones(n), np.zeros(n), x, 1, 0.0, 0, False)
    assert ms._engine_cache[(0, "f64")][1] is eng and np.array_equal(a, b)
    want = x + 0.01 * (np.tanh(gc @ x) - x)
    assert np.max(np.abs(a - want)) <= 1e-14
    gc /= 2.0                                                             # same buffer, new contents (EX:49 style)
    c = ms.developCell(n, gc, 0.01, x, np.ones(n), np.zeros(n), 0.0, 0)
    assert ms._engine_cache[(0, "f64")][1] is not eng
    assert np.max(np.abs(c - (x + 0.01 * (np.tanh(gc @ x) - x)))) <= 1e-14
    ms.init(outputDir=None, seed=5, verbose=False)
    assert not ms._engine_cache


def test_prepared_mode_with_poisson_member_steps():
    """cellDevelopmentMode=False (MS:590-591) in run_all: member cells step k ~ Poisson(T), descendants T."""
    import bench
    from scrutiny_b200 import engine as em, network
    from scrutiny_b200.simulate import LineageSimulation
    seed_all(1337)
    gc, tf = network.generate_gene_correlation_matrix(400, "SimulatedPowerLaw", powerLawExponent=2.5)
    cells, T = 20000, 25
    tree = bench.build_tree(400, cells, tf)
    e = em.Engine(0, "f32")
    e.set_network(gc / 0.4)
    e.alloc_cells(cells)
    e.set_states_broadcast(np.random.RandomState(1).uniform(-1, 1, size=400))
    sim = LineageSimulation(e, cells, seed=9)
    res = sim.run_all(tree, fixed_iterations=T, do_counts=False, seed=9, cellDevelopmentMode=False)
    root = np.asarray(tree[0].cellTypeMembers)
    k = res["iterations"][root]
    assert abs(k.mean() - T) < 4 * np.sqrt(T / len(root)) and abs(k.var() / T - 1) < 0.15 and k.max() > T
    leaf = np.asarray(tree[3].cellTypeMembers)                            # root T + type-1 T + own Poisson(T)
    assert abs(res["iterations"][leaf].mean() - 3 * T) < 0.5


# ---------------------------------------------------------------------------------------------
# default-off extensions of the read-out: negative-binomial draws, dropout
# ---------------------------------------------------------------------------------------------
def test_extension_draws_bit_exact_given_uniforms_and_default_off():
    """scr_counts_ext against its own C oracle (oracle_counts_ext_batch) on the same uniforms: identical integers for NB
    (dispersion below and above 1), dropout and both; with both switches at 0 it returns scr_counts_range's integers."""
    from oracle import poisson_ref
    n, cells = 96, 60
    rng = np.random.RandomState(51)
    e = engine(sparse_net(n, 4), cells, "f64", offset=123)
    S = np.tanh(rng.randn(n, cells))
    e.set_states(S)
    shift = rng.normal(0.6, 1.6, size=n)
    size = rng.normal(1, 0.14, size=cells) ** 3
    draws = 120
    U = rng.random_sample((cells, n, draws))
    base = e.counts(shift, size, uniforms=U[:, :, :48].copy())
    off, lam = e.counts_ext(shift, size, 0.0, 0.0, uniforms=U[:, :, :48].copy(), want_lambda=True)
    assert np.array_equal(base, off)
    assert np.allclose(lam, fo.poisson_lambda(S, shift, 1.0, size).T, rtol=1e-14)
    for phi, p in ((0.3, 0.0), (3.0, 0.0), (0.0, 0.4), (0.7, 0.2)):
        got = e.counts_ext(shift, size, phi, p, uniforms=U)
        want, used = poisson_ref.counts_ext_batch(lam, U, phi, p)
        ok = want >= 0
        assert ok.mean() > 0.999 and np.array_equal(got[ok], want[ok]), (phi, p)
        assert not np.array_equal(got, base)
    # free-running stream: switches off == the ordinary read-out
    a = e.counts(shift, size, seed=9)
    b = e.counts_ext(shift, size, 0.0, 0.0, seed=9)
    assert np.array_equal(a, b)


def test_extension_moments_free_running_and_module_flags(golden):
    n, cells = 48, 30000
    rng = np.random.RandomState(52)
    e = engine(sparse_net(n, 6), cells, "f32")
    S = np.repeat(np.tanh(rng.randn(n, 1)), cells, axis=1)              # identical cells: pure sampling noise
    e.set_states(S)
    shift = np.linspace(-1.0, 3.0, n)
    lam = np.exp(S[:, 0].astype(np.float32).astype(np.float64) + shift)
    phi, p = 0.5, 0.3
    got = e.counts_ext(shift, np.ones(cells), phi, p, seed=77).astype(np.float64)
    mean, var = got.mean(axis=0), got.var(axis=0)
    want_mean = (1 - p) * lam
    want_var = (1 - p) * (lam + phi * lam ** 2) + p * (1 - p) * lam ** 2     # zero-inflated negative binomial
    assert np.all(np.abs(mean - want_mean) < 6 * np.sqrt(want_var / cells) + 1e-3)
    assert np.all(np.abs(var / want_var - 1) < 0.12)
    assert np.all(np.abs((got == 0).mean(axis=0) - (p + (1 - p) * (1 + phi * lam) ** (-1 / phi))) < 0.012)
    # through the module: the reference's dropoutProportion stays unused unless applyDropout is set (MS:819)
    import scrutiny_b200.MatriSeq as ms
    Sg = golden("pipeline")["final_states"]
    ng, cg = Sg.shape
    outs = {}
    for key, kw in (("plain", {}), ("ignored", dict(dropoutProportion=0.5)), ("drop", dict(dropoutProportion=0.5, applyDropout=True)),
                    ("nb", dict(nbDispersion=1.0))):
        ms.init(outputDir=None, seed=21, verbose=False)
        outs[key] = ms.transformData(ng, cg, Sg.copy(), **kw)[0]
    assert np.array_equal(outs["plain"], outs["ignored"])
    assert abs((outs["drop"] == 0).mean() - (0.5 + 0.5 * (outs["plain"] == 0).mean())) < 0.02
    assert outs["nb"].var() > 1.5 * outs["plain"].var() and abs(outs["nb"].mean() / outs["plain"].mean() - 1) < 0.15


# ---------------------------------------------------------------------------------------------
# large n: cell groups in global memory (step_kernel.cuh kModeGlobal)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("f32", 1e-4), ("f64", 1e-10)])
def test_large_gene_count_falls_back_to_global_state(prec, tol):
    """n = 9000 (a float32 cell group of that size needs 290 KB: more than an SM has): the stepper keeps the groups in
    global memory instead of refusing the network.  Ragged supplied-noise phase against the oracle, free-running noise
    against its own dump, the probe against the oracle; SCR_GLOBAL_STATE=1 forces the same path on a small network, where
    it must reproduce the shared-memory path bit for bit."""
    import os
    import scipy.sparse as sp
    n, cells, smax = 9000, 21, 12
    rng = np.random.RandomState(61)
    deg = np.minimum((0.5 * (1 - rng.rand(n)) ** (-1 / 1.3) + 0.5).astype(int), 400)
    regs = rng.choice(n, size=3000, replace=False)
    rows = np.repeat(np.arange(n), deg)
    cols = np.concatenate([rng.choice(regs, size=k, replace=False) for k in deg])
    W = sp.csr_matrix((rng.beta(2, 3, size=len(rows)) * rng.choice([-1.0, 1.0], size=len(rows)) * 2.0, (rows, cols)), shape=(n, n))
    from scrutiny_b200 import engine as em
    e = em.Engine(0, prec)
    e.set_network(W)
    e.alloc_cells(cells, 7)
    lay = e.layout()
    assert lay["state_in_global"] == 1 and lay["chunks"] > 0
    X0 = rng.uniform(-1, 1, size=(cells, n))
    proj = (rng.rand(n) > 0.05).astype(float)
    const = np.where(proj == 0, -1.0, 0.0)
    steps = rng.randint(0, smax + 1, size=cells)
    steps[:3] = [smax, 0, smax]
    noise = (0.05 * rng.randn(cells, smax, n)).astype(np.float32).astype(np.float64)
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), steps, proj, const, noise=noise)
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    fo.run_phase(W, want, steps, proj, const, 0.01, noise)
    assert np.max(np.abs(e.get_states().T - want)) <= tol
    # free-running == supplied(dump of the free-running noise)
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), np.full(cells, 3), proj, const, seed=5, step_base=np.full(cells, 40))
    free = e.get_states()
    dump = np.stack([[e.debug_noise(7 + c, 40 + s, 0.05, 5) for s in range(3)] for c in range(cells)])
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), np.full(cells, 3), proj, const, noise=dump)
    assert np.max(np.abs(free - e.get_states())) <= (1e-12 if prec == "f64" else 1e-6)
    it = e.probe(np.arange(6), proj, const, max_iter=30, avg_num=5, thresh=0.3, noise=np.zeros((6, 30, n)))
    want_it, _, _ = fo.probe(W, e.get_states()[:, :6].T.copy(), proj, const, 0.01, np.zeros((6, 30, n)), 0.3, 30, 5)
    assert np.max(np.abs(it - want_it)) <= (0 if prec == "f64" else 1)
    # forced on a small network: identical to the shared-memory path
    Ws = sparse_net(300, 17, gain=3.0)
    st = rng.randint(0, 30, size=50)
    outs = []
    for force in ("1", "0"):
        os.environ["SCR_GLOBAL_STATE"] = force
        try:
            s = engine(Ws, 50, prec, offset=3)
        finally:
            del os.environ["SCR_GLOBAL_STATE"]
        assert s.layout()["state_in_global"] == int(force)
        s.set_states_broadcast(np.linspace(-1, 1, 300))
        s.run_phase(np.arange(50), st, np.ones(300), np.zeros(300), seed=8)
        outs.append(s.get_states())
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_probes_of_sibling_nodes_side_by_side_equal_separate_probes(prec):
    """scr_probe_multi: the probes of several lineage nodes (own proj / const / seed each) launched together return the
    iteration counts of separate scr_probe calls, sample for sample; a dense context takes the sequential fallback."""
    from scrutiny_b200 import engine
    from test_gpu_parity import sparse_net
    n, cells = 700, 400
    rng = np.random.RandomState(12)
    W = sparse_net(n, 5, gain=1.0 / 0.3)
    e = engine.Engine(0, prec)
    e.set_network(W)
    e.alloc_cells(cells, 2 ** 32 + 9)
    e.set_states(rng.uniform(-1, 1, size=(n, cells)))
    samples = [rng.choice(cells, size=k, replace=False) for k in (50, 50, 7, 33)]
    projs, consts = [], []
    for _ in samples:
        proj = (rng.rand(n) > 0.05).astype(float)
        projs.append(proj)
        consts.append(np.where(proj == 0, rng.choice([-1.0, 1.0], size=n), 0.0))
    seeds = [1337 + 7919 * (i + 1) for i in range(len(samples))]
    kw = dict(thresh=0.25, max_iter=300, avg_num=10)
    want = [e.probe(s, p, c, seed=sd, **kw) for s, p, c, sd in zip(samples, projs, consts, seeds)]
    e.probe_kernel_stats(reset=True)
    got = e.probe_multi(samples, projs, consts, seeds, **kw)
    assert e.probe_kernel_stats()[1] == len(samples)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert max(int(w.max()) for w in want) > 20 and len({int(w.sum()) for w in want}) > 1
    again = [e.probe(s, p, c, seed=sd, **kw) for s, p, c, sd in zip(samples, projs, consts, seeds)]   # per-launch buffers were restored
    for g, w in zip(again, want):
        assert np.array_equal(g, w)
