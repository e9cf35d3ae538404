"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs.

Tolerances (stated per test):
  * float64 context: <= 1e-9 absolute on states in [-1, 1] (libm tanh/summation-order noise
    amplified by the dynamics); iteration counts exact.
  * float32 context: one teacher-forced step <= 5e-7; multi-step runs <= 1e-4 in the regimes
    SURVEY 7.3 H2 lists (alpha >= 0.2, or <= 200 steps at gain 50).  Long runs at gain 50 are
    checked on the float64 path and the float32 drift is only bounded loosely (chaotic amplification).
  * integer counts: bit-exact given the same uniforms.
"""
import numpy as np
import pytest

from oracle import fast_oracle as fo
from oracle import matriseq_port as port
from oracle import philox_ref, poisson_ref
from oracle_utils import noise_table, tree_schedule
from test_oracle_golden import replay_pipeline, replay_probe

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from scrutiny_b200 import engine
    return engine


def make_engine(eng_mod, prec, W, cells, offset=0):
    e = eng_mod.Engine(device=0, precision=prec)
    e.set_network(W)
    e.alloc_cells(cells, offset)
    return e


def sparse_net(n, seed, gain=1.0, heavy=1):
    rng = np.random.RandomState(seed)
    W = np.zeros((n, n))
    deg = np.minimum((0.5 * (1 - rng.rand(n)) ** (-1 / 1.05) + 0.5).astype(int), n)
    regs = rng.choice(n, size=min(n, max(2, n // 3)), replace=False)
    for r in range(n):
        k = min(deg[r], len(regs))
        W[r, rng.choice(regs, size=k, replace=False)] = rng.beta(2, 3, size=k) * rng.choice([-1, 1])
    for _ in range(heavy):
        r = rng.randint(n)
        W[r, :] = rng.beta(2, 3, size=n) * (rng.rand(n) < 0.3)
    return W * gain


# ---------------------------------------------------------------------------------------------
def test_philox_words_match_reference_generator(eng_mod):
    e = eng_mod.Engine(device=0)
    kat = np.array([[0, 0, 0, 0], [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], dtype=np.uint32)
    assert [hex(v) for v in e.debug_philox(kat[:1], 0, 0)[0]] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(v) for v in e.debug_philox(kat[1:2], 0xFFFFFFFF, 0xFFFFFFFF)[0]] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(v) for v in e.debug_philox(kat[2:3], 0xA4093822, 0x299F31D0)[0]] == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']
    rng = np.random.RandomState(0)
    ctr = rng.randint(0, 2 ** 32, size=(5000, 4), dtype=np.uint64).astype(np.uint32)
    got = e.debug_philox(ctr, 123456789, 987654321)
    want = np.stack(philox_ref.philox4x32_10(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], 123456789, 987654321), axis=1)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("prec,tol", [("f64", 1e-13), ("f32", 5e-7)])
def test_single_step_against_reference_output(eng_mod, golden, prec, tol):
    """developCell (MS:503-537): golden x1 produced by the unmodified reference."""
    g = golden("step")
    e = make_engine(eng_mod, prec, g["gc"], 3)
    e.set_states(np.stack([g["x0"]] * 3, axis=1))
    noise = np.stack([g["noise"][None, :]] * 3)
    e.run_phase([0, 1, 2], [1, 1, 0], g["proj"], g["const"], dt=float(g["dt"]), sigma=float(g["sigma"]),
                mu=float(g["mu"]), noise=noise)
    out = e.get_states()
    assert np.max(np.abs(out[:, 0] - g["x1"])) <= tol
    assert np.max(np.abs(out[:, 1] - g["x1"])) <= tol
    f = np.float32 if prec == "f32" else np.float64
    assert np.array_equal(out[:, 2], g["x0"].astype(f).astype(np.float64))   # zero steps: untouched


@pytest.mark.parametrize("prec,tol", [("f64", 1e-10), ("f32", 1e-4)])
def test_forty_steps_against_reference_output(eng_mod, golden, prec, tol):
    """findCellNextState (MS:683-736), 40 steps at gain 50 with the reference's own noise."""
    g = golden("cellsteps")
    e = make_engine(eng_mod, prec, g["gc"], 1)
    e.set_states(g["x0"][:, None])
    e.run_phase([0], [int(g["steps"])], g["proj"], g["const"], noise=g["noise"][None])
    assert np.max(np.abs(e.get_states()[:, 0] - g["xT"])) <= tol


@pytest.mark.parametrize("prec,tol", [("f64", 1e-10), ("f32", 1e-4)])
@pytest.mark.parametrize("n,cells,seed", [(300, 37, 1), (1000, 9, 2), (130, 4, 3)])
def test_ragged_phase_against_oracle(eng_mod, prec, tol, n, cells, seed):
    """Ragged per-cell step counts incl. zeros, arbitrary cell subsets and step bases (MS:587-605)."""
    rng = np.random.RandomState(seed)
    W = sparse_net(n, seed, gain=1.0 / 0.25)
    X0 = rng.uniform(-1, 1, size=(cells, n))
    proj = (rng.rand(n) > 0.1).astype(float)
    const = np.where(proj == 0, rng.choice([-1.0, 1.0], size=n), 0.0)
    ids = rng.permutation(cells)[: max(1, cells - 3)]
    steps = rng.randint(0, 31, size=len(ids))
    steps[0] = 30
    if len(ids) > 2:
        steps[1] = 0
    noise = (0.05 * rng.randn(len(ids), 30, n)).astype(np.float32).astype(np.float64)
    e = make_engine(eng_mod, prec, W, cells)
    e.set_states(X0.T.copy())
    e.run_phase(ids, steps, proj, const, step_base=rng.randint(0, 1000, size=len(ids)), noise=noise)
    got = e.get_states().T
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    Xp = want[ids]
    fo.run_phase(fo.as_operator(W), Xp, steps, proj, const, 0.01, noise)
    want[ids] = Xp
    assert np.max(np.abs(got - want)) <= tol


def test_pipeline_tree_against_reference_output(eng_mod, golden):
    """The whole lineage run (EX:62-68) with the reference's noise, float64 context, against the
    final states the unmodified reference produced; float32 context bounded loosely."""
    g = golden("pipeline")
    rec = port.RecordingDraws()
    log = []
    o = replay_pipeline(draws=rec, log=log)
    n, cells, tree = o["n"], o["cells"], o["tree"]
    noise = noise_table(rec.normals, "step", cells, n)
    phases = tree_schedule(tree, log, o["iters"], cells)
    for prec, tol in (("f64", 1e-9), ("f32", 1e-4)):   # alpha = 0.27 here: float32 meets 1e-4
        e = make_engine(eng_mod, prec, o["gc"], cells)
        e.set_states(o["init"])
        for t, ids, steps, base in phases:
            nz = np.zeros((len(ids), int(steps.max()), n))
            for i, (c, b, s) in enumerate(zip(ids, base, steps)):
                nz[i, :s] = noise[c, b:b + s]
            e.run_phase(ids, steps, tree[t].proj, tree[t].const, step_base=base, noise=nz)
        got = e.get_states()
        err = np.max(np.abs(got - g["final_states"]))
        assert err <= tol, (prec, err)


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_probe_against_oracle(eng_mod, golden, prec):
    """findOptimalIterations' sampled loop (MS:650-669): per-cell iteration counts and traces."""
    g = golden("probe")
    n, gc, states = replay_probe(g)
    rec = port.RecordingDraws()
    sampled = []
    T = port.optimal_iterations(n, gc, 0.01, list(g["members"]), g["proj"], g["const"], 0, float(g["thresh"]),
                                0.05, states, int(g["maxIter"]), int(g["avgNum"]), int(g["numToSample"]), rec,
                                sampled=sampled)
    assert T == int(g["T"])
    max_iter, avg = int(g["maxIter"]), int(g["avgNum"])
    noise = noise_table(rec.normals, "probe", states.shape[1], n)
    noise = np.pad(noise, ((0, 0), (0, max_iter - noise.shape[1]), (0, 0)))[sampled]
    want_it, want_nd, _ = fo.probe(fo.as_operator(gc), states[:, sampled].T.copy(), g["proj"], g["const"], 0.01, noise,
                                   float(g["thresh"]), max_iter, avg)
    e = make_engine(eng_mod, prec, gc, states.shape[1])
    e.set_states(states)
    it, nd = e.probe(sampled, g["proj"], g["const"], thresh=float(g["thresh"]), max_iter=max_iter, avg_num=avg,
                     noise=noise, want_trace=True)
    if prec == "f64":
        assert list(it) == list(want_it)
        assert fo.iterations_from_probe(it, avg) == T
        for i in range(len(sampled)):
            assert np.max(np.abs(nd[i, :it[i]] - want_nd[i, :it[i]])) <= 1e-9
    else:
        assert np.max(np.abs(it - want_it)) <= 2          # threshold ties may move by an iteration
        for i in range(len(sampled)):
            m = min(it[i], want_it[i])
            assert np.max(np.abs(nd[i, :m] - want_nd[i, :m])) <= 1e-3
    assert np.array_equal(e.get_states(), states.astype(np.float32 if prec == "f32" else np.float64))  # untouched


def test_probe_alpha_ladder(eng_mod):
    """w_div rescales the network per sample group (MS:461): compare with separately scaled oracles."""
    rng = np.random.RandomState(4)
    n, k = 96, 8
    W = sparse_net(n, 7)
    X = rng.uniform(-1, 1, size=(k, n))
    noise = (0.05 * rng.randn(k, 60, n)).astype(np.float32).astype(np.float64)
    div = np.repeat([0.05, 0.4, 0.4, 2.0], 2).astype(float)
    e = make_engine(eng_mod, "f64", W, k)
    e.set_states(X.T.copy())
    it = e.probe(np.arange(k), np.ones(n), np.zeros(n), thresh=0.5, max_iter=60, avg_num=5, noise=noise, w_div=div)
    want, _, _ = fo.probe(fo.as_operator(W), X, np.ones(n), np.zeros(n), 0.01, noise, 0.5, 60, 5, scale=div)
    assert list(it) == list(want)


def test_free_running_noise_is_the_documented_generator(eng_mod):
    """Free-running mode = Philox4x32-10 keyed (seed, stream | cell, step, gene block, lane pair) + Box-Muller.
    (1) the dumped noise matches the numpy restatement; (2) a free-running run equals a
    supplied-noise run fed with the dump (ties both modes together); (3) moments / KS."""
    from scipy import stats
    n, cells, steps, seed, sigma = 700, 6, 25, 0xDEADBEEF12345, 0.05
    W = sparse_net(n, 11, gain=4.0)
    rng = np.random.RandomState(5)
    X0 = rng.uniform(-1, 1, size=(cells, n))
    off = 2 ** 33 + 5                                  # global ids beyond 32 bits
    e = make_engine(eng_mod, "f64", W, cells, offset=off)
    dump = np.stack([[e.debug_noise(off + c, 100 + s, sigma, seed) for s in range(steps)] for c in range(cells)])
    # (1) restatement (oracle/philox_ref.py stepper_noise: one call per lane pair, 16-bit radius / angle halves of each
    # word, float64 maths); the device uses MUFU approximations: typically <= 5e-6 * sigma, worst case (tiny radius,
    # where lg2 cancels) <= 1e-3 * sigma
    assert e.noise_stream_version() == 2
    lay = e.layout()
    perm = e.gene_order()
    for c, s in ((2, 3), (5, 0), (0, 24)):
        want_noise = philox_ref.stepper_noise(off + c, 100 + s, lay["n_pad"], perm, sigma, seed)
        near = np.abs(dump[c, s] - want_noise)
        assert np.quantile(near, 0.99) <= 5e-6 * sigma and np.max(near) <= 1e-3 * sigma
    # (2) free-running == supplied(dump)
    e.set_states(X0.T.copy())
    e.run_phase(np.arange(cells), np.full(cells, steps), np.ones(n), np.zeros(n), step_base=np.full(cells, 100),
                sigma=sigma, seed=seed)
    free = e.get_states()
    e.set_states(X0.T.copy())
    e.run_phase(np.arange(cells), np.full(cells, steps), np.ones(n), np.zeros(n), step_base=np.full(cells, 100),
                sigma=sigma, seed=seed, noise=dump)
    assert np.max(np.abs(free - e.get_states())) <= 1e-12
    # and both equal the oracle on that noise
    want = X0.copy()
    fo.run_phase(fo.as_operator(W), want, np.full(cells, steps), np.ones(n), np.zeros(n), 0.01, dump)
    assert np.max(np.abs(free.T - want)) <= 1e-10
    # (3) distribution
    z = dump.ravel() / sigma
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.std() - 1) < 0.01
    assert abs(stats.skew(z)) < 0.03 and abs(stats.kurtosis(z)) < 0.06
    assert stats.kstest(z, "norm").pvalue > 1e-3
    # independent across cells / steps
    assert abs(np.corrcoef(dump[0, 0], dump[1, 0])[0, 1]) < 0.15
    assert abs(np.corrcoef(dump[0, 0], dump[0, 1])[0, 1]) < 0.15


def test_results_do_not_depend_on_sharding(eng_mod):
    """Noise is keyed on the GLOBAL cell id: one context with 24 cells == two contexts (offsets 0 and
    16) with 16 + 8 cells, bit for bit."""
    n, seed = 260, 99
    W = sparse_net(n, 21, gain=5.0)
    rng = np.random.RandomState(6)
    X0 = rng.uniform(-1, 1, size=(24, n)).astype(np.float32).astype(np.float64)
    steps = rng.randint(1, 40, size=24)
    base = rng.randint(0, 50, size=24)
    outs = []
    for prec in ("f32", "f64"):
        whole = make_engine(eng_mod, prec, W, 24)
        whole.set_states(X0.T.copy())
        whole.run_phase(np.arange(24), steps, np.ones(n), np.zeros(n), step_base=base, seed=seed)
        full = whole.get_states()
        a = make_engine(eng_mod, prec, W, 16, offset=0)
        b = make_engine(eng_mod, prec, W, 8, offset=16)
        a.set_states(X0[:16].T.copy())
        b.set_states(X0[16:].T.copy())
        a.run_phase(np.arange(16)[::-1], steps[:16][::-1], np.ones(n), np.zeros(n), step_base=base[:16][::-1], seed=seed)
        b.run_phase(np.arange(8), steps[16:], np.ones(n), np.zeros(n), step_base=base[16:], seed=seed)
        assert np.array_equal(full[:, :16], a.get_states())
        assert np.array_equal(full[:, 16:], b.get_states())
        outs.append(full)
    # same noise in both precisions: float32 tracks float64 over <= 40 steps at gain 5
    assert np.max(np.abs(outs[0] - outs[1])) <= 1e-4


def test_states_roundtrip_and_broadcast(eng_mod):
    n, cells = 333, 50
    W = sparse_net(n, 31)
    rng = np.random.RandomState(8)
    S = rng.randn(n, cells)
    for prec, f in (("f32", np.float32), ("f64", np.float64)):
        e = make_engine(eng_mod, prec, W, cells)
        e.set_states(S)
        assert np.array_equal(e.get_states(), S.astype(f).astype(np.float64))
        big = rng.randn(n, cells + 7)
        e.set_states(big[:, 3:13], cell0=20)                 # strided view, sub-range
        part = e.get_states(cell0=20, count=10)
        assert np.array_equal(part, big[:, 3:13].astype(f).astype(np.float64))
        e.set_states_broadcast(S[:, 0])
        out = e.get_states()
        assert np.array_equal(out, np.repeat(S[:, :1].astype(f).astype(np.float64), cells, axis=1))
        sums = e.gene_sums()
        assert np.allclose(sums, out.sum(axis=1), rtol=1e-12, atol=1e-9)


def test_empty_and_degenerate_phases(eng_mod):
    n = 64
    W = sparse_net(n, 41)
    e = make_engine(eng_mod, "f32", W, 5)
    S = np.random.RandomState(1).uniform(-1, 1, size=(n, 5)).astype(np.float32).astype(np.float64)
    e.set_states(S)
    e.run_phase([], [], np.ones(n), np.zeros(n))
    e.run_phase([0, 1, 2], [0, 0, 0], np.ones(n), np.zeros(n))
    assert np.array_equal(e.get_states(), S)
    with pytest.raises(eng_mod.ScrutinyError):
        e.run_phase([7], [1], np.ones(n), np.zeros(n))
    # proj = 0, const = c  =>  state becomes exactly c after one step
    c = np.linspace(-1, 1, n)
    e.run_phase([4], [3], np.zeros(n), c)
    assert np.array_equal(e.get_states()[:, 4], c.astype(np.float32).astype(np.float64))
    assert np.array_equal(e.get_states()[:, :4], S[:, :4])


@pytest.mark.parametrize("prec,tol", [("f64", 1e-10), ("f32", 1e-4)])
def test_trrust_sized_network(eng_mod, prec, tol):
    """n = 2895, nnz = 8444 (config c2's network), 10 cells x 60 steps at gain 50."""
    from scrutiny_b200.network import trrust_adjacency
    adj = trrust_adjacency()
    n = adj.shape[0]
    rng = np.random.RandomState(12)
    vals = np.stack([rng.beta(rng.randint(1, 11), rng.randint(1, 11), size=n) * rng.choice([-1.0, 1.0]) for _ in range(n)])
    W = adj * vals / 0.02
    cells, steps = 10, 60
    X0 = np.repeat(rng.uniform(-1, 1, size=(1, n)), cells, axis=0)
    noise = (0.05 * rng.randn(cells, steps, n)).astype(np.float32).astype(np.float64)
    e = make_engine(eng_mod, prec, W, cells)
    lay = e.layout()
    assert lay["chunks"] > 0                                   # long rows exercise phase A
    e.set_states(X0.T.copy())
    st = np.array([60, 60, 60, 60, 59, 31, 30, 7, 1, 60])
    e.run_phase(np.arange(cells), st, np.ones(n), np.zeros(n), noise=noise)
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    fo.run_phase(fo.as_operator(W), want, st, np.ones(n), np.zeros(n), 0.01, noise)
    assert np.max(np.abs(e.get_states().T - want)) <= tol


@pytest.mark.parametrize("prec,tol", [("f64", 1e-10), ("f32", 1e-4)])
def test_dense_network_c1_shape(eng_mod, prec, tol):
    """Random n = 500, 99 % dense (config c1), alpha 2.68: operator streamed from L2."""
    rng = np.random.RandomState(13)
    n, cells, steps = 500, 11, 50
    W = np.stack([rng.beta(rng.randint(1, 11), rng.randint(1, 11), size=n) * rng.choice([-1.0, 1.0]) for _ in range(n)])
    W *= rng.rand(n, n) < 0.99
    W /= 2.68
    X0 = rng.uniform(-1, 1, size=(cells, n))
    noise = (0.05 * rng.randn(cells, steps, n)).astype(np.float32).astype(np.float64)
    e = make_engine(eng_mod, prec, W, cells)
    e.set_states(X0.T.copy())
    e.run_phase(np.arange(cells), np.full(cells, steps), np.ones(n), np.zeros(n), noise=noise)
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    fo.run_phase(fo.as_operator(W), want, np.full(cells, steps), np.ones(n), np.zeros(n), 0.01, noise)
    assert np.max(np.abs(e.get_states().T - want)) <= tol


# ---------------------------------------------------------------------------------------------
# read-out
# ---------------------------------------------------------------------------------------------
def test_counts_bit_exact_given_uniforms(eng_mod):
    """MS:873-885: lambda and Poisson draws.  Same uniforms -> identical integers as the C oracle,
    for supplied uniforms and for the free-running Philox stream (restated in numpy)."""
    n, cells = 120, 33
    rng = np.random.RandomState(17)
    W = sparse_net(n, 51)
    S = np.tanh(rng.randn(n, cells))
    shift = rng.normal(0.5, 1.8, size=n)                       # lambda from ~0.01 to ~100: both sampler branches
    size = rng.normal(1, 0.14, size=cells) ** 3
    size[3] = -0.2                                             # negative size factor -> lambda clamps to 0 (MS:882-883)
    e = make_engine(eng_mod, "f64", W, cells, offset=1000)
    e.set_states(S)
    draws = 48
    U = rng.random_sample((cells, n, draws))
    U[:, :, 24:] = np.clip(U[:, :, 24:], 0.0, 0.5)             # make exhaustion impossible for lambda < 10
    got, lam = e.counts(shift, size, exp_scale=1.0, uniforms=U, want_lambda=True)
    want_lam = fo.poisson_lambda(S, shift, 1.0, size).T
    assert np.allclose(lam, want_lam, rtol=1e-14, atol=0)
    want, used = poisson_ref.poisson_batch(lam, U)
    assert (want >= 0).all()
    assert np.array_equal(got, want)
    assert (got[3] == 0).all() and (lam >= 10).sum() > 50 and ((lam > 0) & (lam < 10)).sum() > 1000
    # free-running stream: uniforms = (word + 0.5) / 2^32 from Philox(gene, call, cell_lo, cell_hi)
    seed = 0xABCDEF987
    got2, lam2 = e.counts(shift, size, exp_scale=1.0, seed=seed, want_lambda=True)
    calls = 24
    cell = (1000 + np.arange(cells))[:, None, None]
    gene = np.arange(n)[None, :, None]
    call = np.arange(calls)[None, None, :]
    w = philox_ref.philox4x32_10(gene + 0 * cell + 0 * call, call + 0 * cell + 0 * gene, cell + 0 * gene + 0 * call,
                                 0 * cell + 0 * gene + 0 * call, seed & 0xFFFFFFFF, (seed >> 32) ^ 2)
    U2 = ((np.stack(w, axis=-1).astype(np.float64) + 0.5) / 2.0 ** 32).reshape(cells, n, calls * 4)
    want2, _ = poisson_ref.poisson_batch(lam2, U2)
    ok = want2 >= 0
    assert ok.mean() > 0.999
    assert np.array_equal(got2[ok], want2[ok])


def test_counts_float32_context_and_statistics(eng_mod):
    from scipy import stats
    n, cells = 64, 4000
    rng = np.random.RandomState(19)
    W = sparse_net(n, 61)
    S = np.repeat(np.tanh(rng.randn(n, 1)), cells, axis=1)     # identical cells: pure Poisson sampling noise
    shift = np.linspace(-2, 3.2, n)
    size = np.ones(cells)
    e = make_engine(eng_mod, "f32", W, cells)
    e.set_states(S)
    got, lam = e.counts(shift, size, seed=5, want_lambda=True)
    lam0 = lam[0]
    assert np.allclose(lam, lam0[None, :])
    mean, var = got.mean(axis=0), got.var(axis=0)
    assert np.all(np.abs(mean - lam0) < 5 * np.sqrt(lam0 / cells) + 1e-3)
    assert np.all(np.abs(var - lam0) < 0.15 * lam0 + 0.02)
    for g in (5, 30, 60):                                      # chi-square against the Poisson pmf
        k = np.arange(0, got[:, g].max() + 1)
        obs = np.bincount(got[:, g], minlength=len(k))
        exp = stats.poisson.pmf(k, lam0[g]) * cells
        m = exp > 5
        chi = ((obs[m] - exp[m]) ** 2 / exp[m]).sum()
        assert stats.chi2.sf(chi, m.sum() - 1) > 1e-4
    # device-resident output gives the same integers as the host path
    import torch
    buf = torch.empty((cells, n), dtype=torch.int32, device="cuda:0")
    e.counts(shift, size, seed=5, out_device_ptr=buf.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy(), got)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_counts_float32_screening_equals_float64_sampler(eng_mod, prec):
    """The free-running read-out screens lambda < 10 entries in float32 with a guard band (aux_kernels.cuh,
    poisson_small_f32) and redoes everything it cannot decide in float64: the integers must be those of the float64
    sampler alone (forced by a lambda dump or SCR_COUNTS_EXACT=1) entry for entry -- lambda from 1e-12 to 200, size
    factors inside and outside the screened range, 2 M entries (~400 of them land in the guard band)."""
    import os
    n, cells = 512, 4000
    rng = np.random.RandomState(29)
    e = make_engine(eng_mod, prec, sparse_net(n, 91), cells, offset=2 ** 33 + 5)        # cell ids beyond 32 bits
    e.set_states(np.tanh(rng.randn(n, cells)))
    shift = np.concatenate([np.log(rng.gamma(0.35, 1 / 0.19, size=n - 64)), np.linspace(-28, 5.3, 64)])[rng.permutation(n)]
    size = rng.normal(1, 0.14, size=cells) ** 3
    size[:8] = [0.01, 0.049, 0.051, 7.9, 8.1, -0.2, 0.0, 30.0]
    exact, lam = e.counts(shift, size, seed=41, want_lambda=True)
    fast = e.counts(shift, size, seed=41)
    assert np.array_equal(fast, exact)
    os.environ["SCR_COUNTS_EXACT"] = "1"
    try:
        forced = e.counts(shift, size, seed=41)
    finally:
        del os.environ["SCR_COUNTS_EXACT"]
    assert np.array_equal(forced, exact)
    assert ((lam > 0) & (lam < 1e-6)).sum() > 1000 and ((lam > 5) & (lam < 10)).sum() > 10000 and (lam >= 10).sum() > 10000
    assert (exact[5] == 0).all() and (exact[6] == 0).all()
    other = e.counts(shift, size, seed=42, exp_scale=0.7)
    exact_other = e.counts(shift, size, seed=42, exp_scale=0.7, want_lambda=True)[0]
    assert np.array_equal(other, exact_other) and not np.array_equal(other, fast)


def test_counts_by_row_blocks_equal_one_call(eng_mod):
    """scr_counts_range: sampling the matrix in ragged row blocks (host and device output, Philox and supplied
    uniforms, lambda dump) gives exactly the integers of one call over all cells; bad ranges are refused."""
    import torch
    n, cells = 96, 203
    rng = np.random.RandomState(23)
    e = make_engine(eng_mod, "f32", sparse_net(n, 81), cells, offset=777)
    e.set_states(np.tanh(rng.randn(n, cells)))
    shift = rng.normal(0.3, 1.5, size=n)
    size = rng.normal(1, 0.14, size=cells) ** 3
    U = rng.random_sample((cells, n, 40))
    U[:, :, 20:] = np.clip(U[:, :, 20:], 0.0, 0.5)
    whole, lam = e.counts(shift, size, seed=11, want_lambda=True)
    whole_u = e.counts(shift, size, uniforms=U)
    edges = [0, 1, 64, 65, 130, 203]
    host = np.full((cells, n), -1, dtype=np.int32)
    dev = torch.full((cells, n), -1, dtype=torch.int32, device="cuda:0")
    for r0, r1 in zip(edges[:-1], edges[1:]):
        got, lam_b = e.counts(shift, size[r0:r1], seed=11, cell_begin=r0, cell_count=r1 - r0, out=host[r0:r1], want_lambda=True)
        assert np.array_equal(lam_b, lam[r0:r1])
        e.counts(shift, size[r0:r1], seed=11, cell_begin=r0, cell_count=r1 - r0, out_device_ptr=dev.data_ptr() + 4 * n * r0)
        assert np.array_equal(e.counts(shift, size[r0:r1], uniforms=U[r0:r1], cell_begin=r0, cell_count=r1 - r0), whole_u[r0:r1])
    torch.cuda.synchronize()
    assert np.array_equal(host, whole) and np.array_equal(dev.cpu().numpy(), whole)
    for bad in ((-1, 4), (200, 4), (0, cells + 1)):
        with pytest.raises(eng_mod.ScrutinyError) as err:
            e.counts(shift, np.ones(bad[1]), cell_begin=bad[0], cell_count=bad[1])
        assert err.value.code == -1


def test_counts_uniform_exhaustion_is_an_error(eng_mod):
    n, cells = 32, 2
    e = make_engine(eng_mod, "f64", sparse_net(n, 71), cells)
    e.set_states(np.zeros((n, cells)))
    with pytest.raises(eng_mod.ScrutinyError) as err:
        e.counts(np.full(n, 3.0), np.ones(cells), uniforms=np.full((cells, n, 2), 0.9))
    assert err.value.code == -5


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_partly_resident_operator_is_bit_identical(eng_mod, prec):
    """When the operator does not fit in shared memory next to the cell groups, its head is staged there and the
    rest read through L1 (StepParams::w_prefix).  Where an entry is read from must not change a single bit: compare
    SCR_W_PREFIX_KB=0 (all global), 4 / 24 (some blocks staged) and the default policy on a network big enough
    that the default is partial, with long rows (chunk table beyond the prefix) and ragged steps."""
    import os
    n, cells, seed = 3000, 1300, 99   # > 2 x 148 groups: two groups per CTA, so the default prefix is partial
    rng = np.random.RandomState(4)
    W = np.zeros((n, n))
    regs = rng.choice(n, size=2000, replace=False)
    for r in range(n):
        k = 1 + int(rng.rand() < 0.6) + 40 * int(rng.rand() < 0.01)
        W[r, rng.choice(regs, size=k, replace=False)] = rng.beta(2, 3, size=k) * rng.choice([-1, 1]) * 3.0
    X0 = rng.uniform(-1, 1, size=(n, cells))
    steps = rng.randint(0, 25, size=cells)
    proj = (rng.rand(n) > 0.05).astype(np.float64)
    const = 0.1 * rng.randn(n) * (1 - proj)
    outs, iters = [], []
    for kb in ("0", "4", "24", None):
        if kb is not None:
            os.environ["SCR_W_PREFIX_KB"] = kb
        try:
            e = make_engine(eng_mod, prec, W, cells, offset=123)
            lay = e.layout()
            e.set_states(X0)
            e.run_phase(np.arange(cells), steps, proj, const, seed=seed)
            it = e.probe(np.arange(0, cells, 11), proj, const, max_iter=60, seed=seed)
        finally:
            os.environ.pop("SCR_W_PREFIX_KB", None)
        if kb == "0":
            assert lay["operator_prefix_bytes"] == 0
        elif kb == "4":
            assert lay["operator_prefix_bytes"] == 4096
        elif kb == "24":
            assert 4096 < lay["operator_prefix_bytes"] <= 24576 and not lay["operator_in_smem"]   # capped by the cap or by the room left
        elif not lay["operator_in_smem"]:   # the default policy: all that fits, or nothing when the rest would overflow L1
            assert 0 <= lay["operator_prefix_bytes"] < lay["operator_bytes"]
        outs.append(e.get_states())
        iters.append(it)
    for o, it in zip(outs[1:], iters[1:]):
        assert np.array_equal(o, outs[0])
        assert np.array_equal(it, iters[0])
