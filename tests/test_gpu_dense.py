"""Dense networks: the tcgen05 split-precision GEMM stepper (float32 contexts, nnz >= 5 % of n^2) against the
float64 oracle.  Tolerance 1e-4 absolute on states in [-1, 1]: both operand encodings (fp16 split, 3xTF32) carry
~22 significant bits per operand, the test runs <= 60 steps at moderate gain."""
import os

import numpy as np
import pytest

from oracle import fast_oracle as fo

pytestmark = pytest.mark.gpu


def dense_net(n, seed, density=0.5, alpha=3.0):
    rng = np.random.RandomState(seed)
    W = np.stack([rng.beta(rng.randint(1, 11), rng.randint(1, 11), size=n) * rng.choice([-1.0, 1.0]) for _ in range(n)])
    W *= rng.rand(n, n) < density
    return W / alpha


def engine_for(W, cells, prec="f32", offset=0):
    from scrutiny_b200 import engine
    e = engine.Engine(0, prec)
    e.set_network(W)
    e.alloc_cells(cells, offset)
    return e


@pytest.mark.parametrize("env", [{}, {"SCR_DENSE_RESB": "0"}, {"SCR_DENSE_INSIDE": "0"}, {"SCR_DENSE_KIND": "tf32"},
                                 {"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0"},
                                 {"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0", "SCR_DENSE_KIND": "tf32"},
                                 {"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0", "SCR_DENSE_PAIR": "0"}],
                         ids=["f16-resident-W", "f16-narrow-inside", "f16-per-step", "tf32", "f16-cta-pairs", "tf32-cta-pairs",
                              "f16-wide-single-cta"])
@pytest.mark.parametrize("n,cells,smax,density", [(500, 300, 40, 0.99), (200, 131, 25, 0.5), (700, 64, 12, 0.2)])
def test_dense_stepper_ragged_against_oracle(n, cells, smax, density, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.RandomState(n)
    W = dense_net(n, n + 1, density, alpha=0.02 * n * density)
    X0 = rng.uniform(-1, 1, size=(cells, n))
    proj = (rng.rand(n) > 0.1).astype(float)
    const = np.where(proj == 0, rng.choice([-1.0, 1.0], size=n), 0.0)
    ids = rng.permutation(cells)[: cells - 5]
    steps = rng.randint(0, smax + 1, size=len(ids))
    steps[0], steps[1] = smax, 0
    noise = (0.05 * rng.randn(len(ids), smax, n)).astype(np.float32).astype(np.float64)
    e = engine_for(W, cells)
    assert e.layout()["dense_stepper"] == 1
    e.set_states(X0.T.copy())
    e.run_phase(ids, steps, proj, const, noise=noise, mu=0.1)
    got = e.get_states().T
    want = X0.astype(np.float32).astype(np.float64)
    Xp = want[ids]
    fo.run_phase(fo.as_operator(W), Xp, steps, proj, const, 0.01, noise, mu=0.1)
    want[ids] = Xp
    err = np.max(np.abs(got - want))
    assert err <= 1e-4, err
    untouched = np.setdiff1d(np.arange(cells), ids[steps > 0])
    assert np.array_equal(got[untouched], X0.astype(np.float32).astype(np.float64)[untouched])


def test_dense_free_running_equals_supplied_dump_and_is_shard_invariant():
    n, cells, steps, seed = 256, 140, 9, 0x1234567
    W = dense_net(n, 3, 0.6, alpha=4.0)
    rng = np.random.RandomState(1)
    X0 = rng.uniform(-1, 1, size=(cells, n)).astype(np.float32).astype(np.float64)
    off = 2 ** 32 + 64
    e = engine_for(W, cells, offset=off)
    dump = np.stack([[e.debug_noise(off + c, 7 + s, 0.05, seed) for s in range(steps)] for c in range(cells)])
    e.set_states(X0.T.copy())
    e.run_phase(np.arange(cells), np.full(cells, steps), np.ones(n), np.zeros(n), step_base=np.full(cells, 7), seed=seed)
    free = e.get_states()
    e.set_states(X0.T.copy())
    e.run_phase(np.arange(cells), np.full(cells, steps), np.ones(n), np.zeros(n), noise=dump)
    assert np.max(np.abs(free - e.get_states())) <= 1e-6
    # two shards with global offsets reproduce the single context bit for bit
    a = engine_for(W, 100, offset=off)
    b = engine_for(W, 40, offset=off + 100)
    a.set_states(X0[:100].T.copy())
    b.set_states(X0[100:].T.copy())
    a.run_phase(np.arange(100), np.full(100, steps), np.ones(n), np.zeros(n), step_base=np.full(100, 7), seed=seed)
    b.run_phase(np.arange(40)[::-1], np.full(40, steps), np.ones(n), np.zeros(n), step_base=np.full(40, 7), seed=seed)
    assert np.array_equal(free[:, :100], a.get_states()) and np.array_equal(free[:, 100:], b.get_states())
    z = dump.ravel() / 0.05
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


def test_dense_and_sparse_kernels_agree_on_the_same_network():
    n, cells, steps = 320, 20, 30
    W = dense_net(n, 9, 0.3, alpha=2.0)
    rng = np.random.RandomState(2)
    X0 = rng.uniform(-1, 1, size=(cells, n))
    noise = (0.05 * rng.randn(cells, steps, n)).astype(np.float32).astype(np.float64)
    outs = []
    for force in ("1", "0"):
        os.environ["SCR_DENSE"] = force
        try:
            e = engine_for(W, cells)
        finally:
            del os.environ["SCR_DENSE"]
        assert e.layout()["dense_stepper"] == int(force)
        e.set_states(X0.T.copy())
        e.run_phase(np.arange(cells), np.full(cells, steps), np.ones(n), np.zeros(n), noise=noise)
        outs.append(e.get_states())
    assert np.max(np.abs(outs[0] - outs[1])) <= 2e-5


def test_probe_works_on_dense_networks():
    n, k = 192, 8
    W = dense_net(n, 4, 0.5, alpha=6.0)
    rng = np.random.RandomState(3)
    X = rng.uniform(-1, 1, size=(k, n))
    noise = (0.05 * rng.randn(k, 80, n)).astype(np.float32).astype(np.float64)
    e = engine_for(W, k)
    e.set_states(X.T.copy())
    it = e.probe(np.arange(k), np.ones(n), np.zeros(n), thresh=0.4, max_iter=80, avg_num=6, noise=noise)
    want, _, _ = fo.probe(fo.as_operator(W), X.astype(np.float32).astype(np.float64), np.ones(n), np.zeros(n), 0.01,
                          noise, 0.4, 80, 6)
    assert np.max(np.abs(it - want)) <= 1


def test_states_outside_the_fp16_operand_range_are_refused_not_saturated():
    """The default operand encoding scales the state by 64 into fp16: |x| >= 1023 cannot be represented.  The call fails
    (SCR_E_RESOURCE); the tf32 encoding takes the same input."""
    from scrutiny_b200 import engine
    n, cells = 128, 4
    W = dense_net(n, 5, 0.5, alpha=2.0)
    X0 = np.zeros((cells, n))
    X0[2, 17] = 2000.0
    e = engine_for(W, cells)
    e.set_states(X0.T.copy())
    with pytest.raises(engine.ScrutinyError) as ei:
        e.run_phase(np.arange(cells), np.full(cells, 3), np.ones(n), np.zeros(n), noise=np.zeros((cells, 3, n)))
    assert "fp16 operand range" in str(ei.value)
    os.environ["SCR_DENSE_KIND"] = "tf32"
    try:
        e2 = engine_for(W, cells)
    finally:
        del os.environ["SCR_DENSE_KIND"]
    e2.set_states(X0.T.copy())
    e2.run_phase(np.arange(cells), np.full(cells, 3), np.ones(n), np.zeros(n), noise=np.zeros((cells, 3, n)))
    want = X0.copy()
    fo.run_phase(fo.as_operator(W), want, np.full(cells, 3), np.ones(n), np.zeros(n), 0.01, np.zeros((cells, 3, n)))
    assert np.max(np.abs(e2.get_states().T - want)) <= 1e-3   # (values of 2000: fp32 ulp 1.2e-4)


@pytest.mark.parametrize("kind", ["f16", "tf32"])
def test_dense_stepper_against_the_reference_outputs(golden, kind, monkeypatch):
    """The committed outputs of the UNMODIFIED reference (tests/golden, oracle/make_golden.py) through the GEMM stepper
    (forced with SCR_DENSE=1: the golden networks are small), both operand encodings: one developCell step with
    non-trivial proj / const / geneMeanLevels (MS:503-537) and 40 steps of findCellNextState at gain 50 with the
    reference's own noise (MS:683-736)."""
    from scrutiny_b200 import engine
    monkeypatch.setenv("SCR_DENSE", "1")
    monkeypatch.setenv("SCR_DENSE_KIND", kind)
    g = golden("step")
    e = engine.Engine(0, "f32")
    e.set_network(g["gc"])
    assert e.layout()["dense_stepper"] == 1 and e.layout()["dense_encoding"] == (1 if kind == "f16" else 2)
    e.alloc_cells(3)
    e.set_states(np.stack([g["x0"]] * 3, axis=1))
    e.run_phase([0, 1, 2], [1, 1, 0], g["proj"], g["const"], dt=float(g["dt"]), sigma=float(g["sigma"]),
                mu=float(g["mu"]), noise=np.stack([g["noise"][None, :]] * 3))
    out = e.get_states()
    assert np.max(np.abs(out[:, 0] - g["x1"])) <= 2e-6 and np.max(np.abs(out[:, 1] - g["x1"])) <= 2e-6
    assert np.array_equal(out[:, 2], g["x0"].astype(np.float32).astype(np.float64))   # zero steps: untouched
    g = golden("cellsteps")
    e = engine.Engine(0, "f32")
    e.set_network(g["gc"])
    assert e.layout()["dense_stepper"] == 1
    e.alloc_cells(1)
    e.set_states(g["x0"][:, None])
    e.run_phase([0], [int(g["steps"])], g["proj"], g["const"], noise=g["noise"][None])
    assert np.max(np.abs(e.get_states()[:, 0] - g["xT"])) <= 1e-4
