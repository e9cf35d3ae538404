"""Randomised differential test of the stepper against the oracle: network shapes the fixed cases do not hit
(n from 1 gene up, n not a multiple of 32 / 128, empty rows, self loops, rows longer than the cap, rows reading
every regulator, isolated genes, denser-than-5 % networks on the tcgen05 path), tiny and ragged cell sets, non-zero
gene mean level, pinned genes -- with the noise stream supplied identically to both sides.
Tolerances: float64 context 1e-10 absolute (states in [-1, 1]); float32 context 1e-4 (gain <= 4, <= 12 steps)."""
import numpy as np
import pytest

from oracle import fast_oracle as fo

pytestmark = pytest.mark.gpu


def random_network(rng, n):
    kind = int(rng.choice([0, 1, 1, 2, 3, 4, 4, 5]))
    W = np.zeros((n, n))
    if kind == 0:                                   # very sparse, many empty rows
        m = rng.randint(0, n + 1)
        W[rng.randint(n, size=m), rng.randint(n, size=m)] = rng.randn(m)
    elif kind == 1:                                 # power-law-ish in-degree with a few very long rows
        regs = rng.choice(n, size=max(1, n // 2), replace=False)
        for r in range(n):
            k = min(len(regs), int(rng.pareto(1.2)) + int(rng.rand() < 0.7))
            W[r, rng.choice(regs, size=k, replace=False)] = rng.beta(2, 3, size=k) * rng.choice([-1, 1])
        for r in rng.choice(n, size=min(n, 3), replace=False):
            W[r, regs] = rng.beta(2, 3, size=len(regs)) * 0.2
    elif kind == 2:                                 # one regulator drives everything, plus self loops
        W[:, rng.randint(n)] = rng.randn(n)
        W[np.arange(n), np.arange(n)] += rng.randn(n) * (rng.rand(n) < 0.3)
    elif kind == 3:                                 # banded
        for d in range(1, min(n, 1 + rng.randint(1, 12))):
            W[np.arange(n - d), np.arange(d, n)] = rng.randn(n - d) * 0.5
    elif kind == 4:                                 # dense-ish: above the 5 % switch for n >= 64
        W = rng.randn(n, n) * (rng.rand(n, n) < rng.uniform(0.06, 0.6)) / np.sqrt(n)
    else:                                           # empty operator
        pass
    gain = rng.uniform(0.5, 4.0)
    scale = np.abs(W).sum(axis=1).max()
    return W * (gain / scale) if scale > 0 else W


@pytest.mark.parametrize("seed", range(160))
def test_random_case_against_oracle(seed):
    from scrutiny_b200 import engine
    rng = np.random.RandomState(1000 + seed)
    n = int(rng.choice([1, 2, 5, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 200, 255, 257, 300, 511, 640, 700]))
    cells = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 31, 50, 150]))
    prec, tol = (("f64", 1e-10), ("f32", 1e-4))[seed % 2]
    W = random_network(rng, n)
    smax = int(rng.randint(1, 13))
    X0 = rng.uniform(-1, 1, size=(cells, n))
    proj = (rng.rand(n) > 0.15).astype(float)
    const = np.where(proj == 0, rng.uniform(-1, 1, size=n), 0.0)
    mu = float(rng.choice([0.0, 0.0, 0.1, -0.05]))
    ids = rng.permutation(cells)[: rng.randint(1, cells + 1)]
    steps = rng.randint(0, smax + 1, size=len(ids))
    steps[rng.randint(len(ids))] = smax
    noise = (0.05 * rng.randn(len(ids), smax, n)).astype(np.float32).astype(np.float64)
    e = engine.Engine(device=0, precision=prec)
    e.set_network(W)
    off = int(rng.randint(0, 10 ** 6))
    e.alloc_cells(cells, off)
    e.set_states(X0.T.copy())
    e.run_phase(ids, steps, proj, const, step_base=rng.randint(0, 1000, size=len(ids)), noise=noise, mu=mu)
    got = e.get_states().T
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    Xp = want[ids]
    fo.run_phase(fo.as_operator(W), Xp, steps, proj, const, 0.01, noise, mu=mu)
    want[ids] = Xp
    err = np.max(np.abs(got - want))
    assert err <= tol, (n, cells, prec, e.layout(), err)
    # free-running mode on the same case: equals a supplied-noise run fed the generator's own dump (both kernels' noise
    # is keyed by global cell, step and gene only)
    if seed % 3 == 0:
        sd = 4242
        e.set_states(X0.T.copy())
        base = rng.randint(0, 1000, size=len(ids))
        e.run_phase(ids, steps, proj, const, step_base=base, seed=sd, mu=mu)
        free = e.get_states().T
        dump = np.zeros((len(ids), smax, n))
        for i, c in enumerate(ids):
            for s in range(int(steps[i])):
                dump[i, s] = e.debug_noise(off + int(c), int(base[i]) + s, 0.05, sd)
        e.set_states(X0.T.copy())
        e.run_phase(ids, steps, proj, const, step_base=base, noise=dump, mu=mu)
        assert np.max(np.abs(e.get_states().T - free)) <= (1e-12 if prec == "f64" else 2e-6)


DENSE_SCHEDULES = [
    {},                                                                        # resident-W / narrow tiles, steps inside the launch
    {"SCR_DENSE_RESB": "0"},                                                   # narrow (or wide) tiles, steps inside
    {"SCR_DENSE_INSIDE": "0"},                                                 # one launch per timestep (narrow tiles or CTA pairs)
    {"SCR_DENSE_NARROW": "0"},                                                 # wide tiles, steps inside
    {"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0"},                        # CTA pairs
    {"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0", "SCR_DENSE_PAIR": "0"},  # single-CTA wide tiles per timestep
]


@pytest.mark.parametrize("seed", range(48))
def test_random_dense_case_on_every_schedule_against_oracle(seed, monkeypatch):
    """The GEMM stepper on shapes the fixed cases do not hit -- n and cell counts around the tile sizes (32 / 64 / 128 / 256),
    one cell, odd numbers of cell tiles (the CTA-pair kernel's half-empty last pair), K extents shorter than the ring --
    cycling through the schedules and both operand encodings, supplied noise, ragged steps; then the probe of the same
    context against the oracle's (iteration counts within one step: float32 traces against float64)."""
    from scrutiny_b200 import engine
    rng = np.random.RandomState(7000 + seed)
    for k, v in DENSE_SCHEDULES[seed % len(DENSE_SCHEDULES)].items():
        monkeypatch.setenv(k, v)
    if (seed // len(DENSE_SCHEDULES)) % 2:
        monkeypatch.setenv("SCR_DENSE_KIND", "tf32")
    n = int(rng.choice([64, 65, 127, 128, 200, 256, 257, 500, 513, 700, 1000, 1100]))
    cells = int(rng.choice([1, 50, 128, 129, 255, 256, 257, 400, 700]))
    W = rng.randn(n, n) * (rng.rand(n, n) < rng.uniform(0.06, 0.9))
    W *= rng.uniform(0.5, 4.0) / np.abs(W).sum(axis=1).max()
    smax = int(rng.randint(1, 9))
    X0 = rng.uniform(-1, 1, size=(cells, n))
    proj = (rng.rand(n) > 0.15).astype(float)
    const = np.where(proj == 0, rng.uniform(-1, 1, size=n), 0.0)
    mu = float(rng.choice([0.0, 0.0, 0.1]))
    ids = rng.permutation(cells)[: rng.randint(1, cells + 1)]
    steps = rng.randint(0, smax + 1, size=len(ids))
    steps[rng.randint(len(ids))] = smax
    noise = (0.05 * rng.randn(len(ids), smax, n)).astype(np.float32).astype(np.float64)
    e = engine.Engine(device=0, precision="f32")
    e.set_network(W)
    assert e.layout()["dense_stepper"] == 1
    e.alloc_cells(cells, int(rng.randint(0, 10 ** 6)))
    e.set_states(X0.T.copy())
    e.run_phase(ids, steps, proj, const, step_base=rng.randint(0, 1000, size=len(ids)), noise=noise, mu=mu)
    got = e.get_states().T
    want = X0.astype(np.float32).astype(np.float64)
    Xp = want[ids]
    op = fo.as_operator(W)
    fo.run_phase(op, Xp, steps, proj, const, 0.01, noise, mu=mu)
    want[ids] = Xp
    err = np.max(np.abs(got - want))
    assert err <= 1e-4, (n, cells, len(ids), smax, e.layout(), err)
    if seed % 4 == 0:
        k = min(cells, 9)
        pn = (0.05 * rng.randn(k, 40, n)).astype(np.float32).astype(np.float64)
        e.set_states(X0.T.copy())
        it = e.probe(np.arange(k), proj, const, thresh=0.5, max_iter=40, avg_num=5, noise=pn, mu=mu)
        ref, _, _ = fo.probe(op, X0[:k].astype(np.float32).astype(np.float64), proj, const, 0.01, pn, 0.5, 40, 5, mu)
        assert np.max(np.abs(it - ref)) <= 1, (it, ref)
