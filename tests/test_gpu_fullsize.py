"""Size-independent properties at BASELINE scale (n = 3000 power-law network, 10^5 cells), where the
oracle would take hours: determinism, independence of cell grouping / call partitioning, pinned genes,
boundedness, noise-free collapse, and the count read-out's moments.  (Bit-level parity against the
oracle is covered at oracle-sized inputs in test_gpu_parity.py.)"""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from scrutiny_b200 import engine, network
    np.random.seed(1337)
    random.seed(1337)
    gc, tf = network.generate_gene_correlation_matrix(3000, "SimulatedPowerLaw", powerLawExponent=2.5)
    gc /= 0.4
    n, cells = 3000, 100000
    rng = np.random.RandomState(3)
    proj = np.ones(n)
    const = np.zeros(n)
    frozen = rng.choice(tf, size=100, replace=False)
    proj[frozen] = 0
    const[frozen] = rng.choice([-1.0, 1.0], size=100)
    x0 = rng.uniform(-1, 1, size=n)
    steps = rng.randint(0, 41, size=cells).astype(np.int32)
    base = rng.randint(0, 500, size=cells)
    e = engine.Engine(0, "f32")
    e.set_network(gc)
    e.alloc_cells(cells)
    return dict(e=e, n=n, cells=cells, proj=proj, const=const, frozen=frozen, x0=x0, steps=steps, base=base, gc=gc)


def run(big, ids_list, seed=77, sigma=0.05):
    e = big["e"]
    e.set_states_broadcast(big["x0"])
    for ids in ids_list:
        e.run_phase(ids, big["steps"][ids], big["proj"], big["const"], step_base=big["base"][ids], seed=seed, sigma=sigma)
    return e


def sample_states(e, idx):
    return np.stack([e.get_states(int(i), 1)[:, 0] for i in idx], axis=1)


def test_determinism_and_independence_of_partitioning(big):
    cells = big["cells"]
    probe = np.random.RandomState(0).choice(cells, size=64, replace=False)
    a = sample_states(run(big, [np.arange(cells)]), probe)
    b = sample_states(run(big, [np.arange(cells)]), probe)
    assert np.array_equal(a, b)                                           # same call twice
    perm = np.random.RandomState(1).permutation(cells)
    c = sample_states(run(big, [perm]), probe)
    assert np.array_equal(a, c)                                           # different grouping of cells
    parts = np.array_split(perm, 7)
    d = sample_states(run(big, parts), probe)
    assert np.array_equal(a, d)                                           # seven launches instead of one
    import os
    for threads in ("1", "3"):                                            # host-side group building: 4 threads by default
        os.environ["SCR_HOST_THREADS"] = threads
        try:
            t = sample_states(run(big, [perm]), probe)
        finally:
            del os.environ["SCR_HOST_THREADS"]
        assert np.array_equal(a, t)
    assert not np.array_equal(a[:, 0], a[:, 1])


def test_invariants_of_the_update(big):
    e = run(big, [np.arange(big["cells"])])
    idx = np.arange(0, big["cells"], 997)
    S = sample_states(e, idx)
    st = big["steps"][idx]
    assert np.isfinite(S).all() and np.abs(S).max() <= 1.0 + 1e-6       # tanh-bounded Euler blend
    moved = st > 0
    f = big["frozen"]
    assert np.array_equal(S[np.ix_(f, np.nonzero(moved)[0])],
                          np.repeat(big["const"][f][:, None], moved.sum(), axis=1))   # proj = 0 => x = const (MS:535)
    x0f = big["x0"].astype(np.float32).astype(np.float64)
    assert np.array_equal(S[:, ~moved], np.repeat(x0f[:, None], (~moved).sum(), axis=1))   # zero steps: untouched
    # gene sums over all cells agree with a float64 host sum of downloaded states on a slice
    sl = e.get_states(0, 2048)
    part = big["e"].gene_sums()
    assert part.shape == (big["n"],) and np.isfinite(part).all()
    assert np.allclose(sl.sum(axis=1), sample_states(e, np.arange(2048)).sum(axis=1))


def test_noise_free_cells_with_equal_steps_stay_identical(big):
    e = big["e"]
    cells = big["cells"]
    e.set_states_broadcast(big["x0"])
    steps = np.full(cells, 25, dtype=np.int32)
    e.run_phase(np.arange(cells), steps, big["proj"], big["const"], sigma=0.0, seed=1)
    S = sample_states(e, [0, 1, 31337, cells - 1])
    assert np.array_equal(S[:, 0], S[:, 1]) and np.array_equal(S[:, 0], S[:, 2]) and np.array_equal(S[:, 0], S[:, 3])
    # and that common trajectory is the deterministic map iterated 25 times (float64 host evaluation)
    x = big["x0"].astype(np.float32).astype(np.float64)
    W = big["gc"]
    for _ in range(25):
        x = big["proj"] * (x + 0.01 * (np.tanh(W @ x) - x)) + big["const"]
    assert np.max(np.abs(S[:, 0] - x)) <= 2e-5


def test_count_readout_moments_at_scale(big):
    e = run(big, [np.arange(big["cells"])])
    n, cells = big["n"], big["cells"]
    rng = np.random.RandomState(5)
    shift = np.sort(np.log(rng.gamma(0.35, 1 / 0.19, size=n)))[rng.permutation(n)]
    size = rng.normal(1, 0.14, size=cells) ** 3
    import torch
    buf = torch.empty((cells, n), dtype=torch.int32, device="cuda:0")
    e.counts(shift, size, seed=9, out_device_ptr=buf.data_ptr())
    torch.cuda.synchronize()
    assert int(buf.min()) >= 0
    sub = np.arange(0, cells, 50)
    S = sample_states(e, sub)
    lam = np.maximum(np.exp(S + shift[:, None]) * size[sub][None, :], 0)
    got = buf[torch.as_tensor(sub, device="cuda:0")].cpu().numpy().T.astype(np.float64)
    # E[count] = lambda: compare totals per gene over 2000 cells (relative, for genes with enough mass)
    mass = lam.sum(axis=1)
    big_genes = mass > 200
    rel = np.abs(got.sum(axis=1)[big_genes] - mass[big_genes]) / np.sqrt(mass[big_genes])
    assert rel.max() < 6.0, rel.max()                                    # Poisson: sd of the sum = sqrt(mass)
    assert abs((got == 0).mean() - np.exp(-lam).mean()) < 0.005          # zero fraction = E[exp(-lambda)]
    # the float32-screened sampler (default) and the float64 sampler alone give the same 3e8 integers
    import os
    os.environ["SCR_COUNTS_EXACT"] = "1"
    try:
        ref = torch.empty((cells, n), dtype=torch.int32, device="cuda:0")
        e.counts(shift, size, seed=9, out_device_ptr=ref.data_ptr())
        torch.cuda.synchronize()
    finally:
        del os.environ["SCR_COUNTS_EXACT"]
    assert torch.equal(buf, ref)
