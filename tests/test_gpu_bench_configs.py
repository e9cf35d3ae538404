"""Oracle parity ON THE CONFIGURATIONS THE BENCH NUMBERS ARE QUOTED ON:

  (a) the bench network itself -- SimulatedPowerLaw n = 3000, gamma = 2.5, seed 1337, divided by alpha = 0.4 -- and the
      reference's default exponent 2.05, with enough cells (1216 -> 304 / 608 groups > 2 x 148) that the launch takes the
      layout bench.py runs: two cell groups per CTA and the operator's head staged in shared memory;
        - supplied noise (step_kernel<.., SUPPLIED>), ragged 0..30 steps, every cell against oracle/fast_oracle.py;
        - free-running noise (the packed fast path bench.py times), ragged 0..60 steps, a sample of cells replayed through
          the oracle with the noise scr_debug_noise dumps for them;
  (b) config c4 -- Random n = 3000, sparsity 0.5, alpha 2.21 -- 256 cells x 20 steps on every schedule and operand encoding
      of the tcgen05 stepper, and 4 000 free-running cells on the CTA-pair schedule bench.py's c4 block runs (sampled cells
      replayed through the oracle, shard invariance, agreement with the single-CTA kernel);
  (c) a c2-shaped prepared run (TRRUST, 7-type tree, 10 000 cells, run_all with fixed iterations): sampled cells replayed
      phase by phase through the oracle with their dumped noise;
  (d) transformData through the public module: lambda against the port with the same seed.

Tolerances: float32 <= 1e-4 absolute on states in [-1, 1] (north_star), float64 <= 1e-10 (1e-9 over the multi-phase run).
"""
import random

import numpy as np
import pytest

from conftest import seed_all
from oracle import fast_oracle as fo
from oracle import matriseq_port as port

pytestmark = pytest.mark.gpu

CELLS = 1216


def bench_network(exponent):
    from scrutiny_b200 import network
    seed_all(1337)
    gc, tf = network.generate_gene_correlation_matrix(3000, "SimulatedPowerLaw", powerLawExponent=exponent)
    return gc / 0.4, tf


def node_vectors(n, tf, rng):
    proj, const = np.ones(n), np.zeros(n)
    frozen = rng.choice(tf, size=max(1, len(tf) // 20), replace=False)
    proj[frozen] = 0
    const[frozen] = rng.choice([-1.0, 1.0], size=len(frozen))
    return proj, const


@pytest.fixture(scope="module", params=[2.5, 2.05], ids=["gamma2.5", "gamma2.05"])
def net(request):
    gc, tf = bench_network(request.param)
    n = gc.shape[0]
    rng = np.random.RandomState(7)
    proj, const = node_vectors(n, tf, rng)
    x0 = rng.uniform(-1, 1, size=n)
    return dict(gc=gc, W=fo.as_operator(gc), n=n, proj=proj, const=const, x0=x0, exponent=request.param)


def engine_on(net, prec):
    from scrutiny_b200 import engine
    e = engine.Engine(0, prec)
    e.set_network(net["gc"])
    e.alloc_cells(CELLS, 2 ** 32 + 3)                        # odd offset: parity placement must use GLOBAL ids
    lay = e.layout()
    if prec == "f32":   # two cell groups per CTA and a partly resident operator, as in bench.py (float64 rows leave room for one group)
        assert lay["groups_per_cta"] == 2 and not lay["operator_in_smem"] and lay["operator_prefix_bytes"] > 0
    if net["exponent"] == 2.05:
        assert lay["chunks"] > 300                           # max in-degree ~300: many long-row chunks
    return e


@pytest.mark.parametrize("prec,tol", [("f32", 1e-4), ("f64", 1e-10)])
def test_bench_network_supplied_noise_every_cell(net, prec, tol):
    n, smax = net["n"], 30
    rng = np.random.RandomState(11)
    X0 = np.repeat(net["x0"][None, :], CELLS, axis=0) + 0.05 * rng.randn(CELLS, n)
    steps = rng.randint(0, smax + 1, size=CELLS)
    steps[:8] = [smax, 0, smax, 1, smax - 1, smax, 2, smax]
    noise = (0.05 * rng.standard_normal((CELLS, smax, n)).astype(np.float32)).astype(np.float64)
    e = engine_on(net, prec)
    e.set_states(np.ascontiguousarray(X0.T))
    ids = rng.permutation(CELLS)                             # noise row i belongs to cell ids[i]
    e.run_phase(ids, steps[ids], net["proj"], net["const"], noise=noise)
    got = e.get_states().T
    want = X0.astype(np.float32).astype(np.float64) if prec == "f32" else X0.copy()
    fo.run_phase(net["W"], want, steps, net["proj"], net["const"], 0.01, noise, rows=np.argsort(ids))
    err = np.max(np.abs(got - want))
    assert err <= tol, (prec, err)
    assert e.path_launches()["sparse_persistent"] == 1


@pytest.mark.parametrize("prec,tol", [("f32", 1e-4), ("f64", 1e-10)])
def test_bench_network_free_running_fast_path_sampled_cells(net, prec, tol):
    """The kernel instantiation bench.py times (free-running Philox noise, packed update, parity-placed groups)."""
    n, smax, seed, sigma = net["n"], 60, 0x5EED1234ABCD, 0.05
    rng = np.random.RandomState(13)
    steps = rng.randint(0, smax + 1, size=CELLS).astype(np.int32)
    base = rng.randint(0, 900, size=CELLS)
    e = engine_on(net, prec)
    e.set_states_broadcast(net["x0"])
    e.run_phase(np.arange(CELLS), steps, net["proj"], net["const"], step_base=base, seed=seed, sigma=sigma)
    sample = np.concatenate([np.arange(8), rng.choice(np.arange(8, CELLS), size=40, replace=False)])
    got = np.stack([e.get_states(int(c), 1)[:, 0] for c in sample])
    off = 2 ** 32 + 3
    noise = np.zeros((len(sample), smax, n))
    for i, c in enumerate(sample):
        for s in range(steps[c]):
            noise[i, s] = e.debug_noise(off + int(c), int(base[c]) + s, sigma, seed)
    x0 = net["x0"].astype(np.float32).astype(np.float64) if prec == "f32" else net["x0"]
    want = np.repeat(x0[None, :], len(sample), axis=0)
    fo.run_phase(net["W"], want, steps[sample], net["proj"], net["const"], 0.01, noise)
    err = np.max(np.abs(got - want))
    assert err <= tol, (prec, err)
    z = noise[steps[sample][:, None] > np.arange(smax)[None, :]].ravel() / sigma
    assert abs(z.mean()) < 5 / np.sqrt(z.size) and abs(z.std() - 1) < 5e-3


@pytest.mark.parametrize("env,launches", [
    ({}, 1),                                                       # default: fp16 split, narrow tiles, steps inside the launch
    ({"SCR_DENSE_NARROW": "0"}, 1),                                # wide tiles, steps inside
    ({"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0"}, 20),      # CTA pairs (cta_group::2), one launch per timestep (what 200 k cells run)
    ({"SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0", "SCR_DENSE_PAIR": "0"}, 20),   # single-CTA wide tiles per timestep
    ({"SCR_DENSE_KIND": "tf32"}, 1),                               # north_star's 3xTF32 encoding
    ({"SCR_DENSE_KIND": "tf32", "SCR_DENSE_NARROW": "0", "SCR_DENSE_INSIDE": "0"}, 20),
])
def test_c4_dense_network_on_the_tcgen05_stepper(env, launches, monkeypatch):
    """Config c4: Random n = 3000, networkSparsity 0.5 (nnz ~4.5 M), alpha 2.21 (what the search returns for it:
    profiles/r1_final_configs_c1_c4.jsonl); full K extent (48 / 96 K slices) and all gene tiles, on every schedule and
    operand encoding of the dense stepper."""
    from scrutiny_b200 import engine, network
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    seed_all(1337)
    gc, tf = network.generate_gene_correlation_matrix(3000, "Random", networkSparsity=0.5)
    gc /= 2.21
    n, cells, steps_max = 3000, 256, 20
    rng = np.random.RandomState(17)
    proj, const = node_vectors(n, tf, rng)
    X0 = rng.uniform(-1, 1, size=(cells, n))
    steps = rng.randint(0, steps_max + 1, size=cells)
    steps[:4] = [steps_max, 0, steps_max, 7]
    noise = (0.05 * rng.standard_normal((cells, steps_max, n)).astype(np.float32)).astype(np.float64)
    e = engine.Engine(0, "f32")
    e.set_network(gc)
    assert e.layout()["dense_stepper"] == 1
    e.alloc_cells(cells)
    e.set_states(np.ascontiguousarray(X0.T))
    e.run_phase(np.arange(cells), steps, proj, const, noise=noise)
    assert e.path_launches()["dense_gemm"] == launches
    want = X0.astype(np.float32).astype(np.float64)
    fo.run_phase(gc, want, steps, proj, const, 0.01, noise)
    err = np.max(np.abs(e.get_states().T - want))
    assert err <= 1e-4, err


def test_c4_large_batch_takes_the_cta_pair_schedule_free_running(monkeypatch):
    """What bench.py's dense_c4 block runs, at a size the oracle can follow: 4 000 cells (32 cell tiles > SMs / 12 gene tiles,
    so the launch per timestep takes CTA pairs; 4 000 is not a multiple of 128 or 256: padding rows), ragged 0..12 steps,
    free-running Philox noise.  (i) a sample of cells replayed through the float64 oracle with the noise scr_debug_noise
    dumps for them; (ii) two shards with global offsets reproduce the single context bit for bit; (iii) the single-CTA
    kernel (SCR_DENSE_PAIR=0) agrees to rounding."""
    from scrutiny_b200 import engine, network
    seed_all(1337)
    gc, tf = network.generate_gene_correlation_matrix(3000, "Random", networkSparsity=0.5)
    gc /= 2.21
    n, cells, smax, seed, sigma, off = 3000, 4000, 12, 0xC4C4, 0.05, 2 ** 33 + 5
    rng = np.random.RandomState(23)
    proj, const = node_vectors(n, tf, rng)
    x0 = rng.uniform(-1, 1, size=n)
    steps = rng.randint(0, smax + 1, size=cells)
    steps[:3] = [smax, 0, smax]
    base = rng.randint(0, 50, size=cells)

    def run(lo, hi):
        e = engine.Engine(0, "f32")
        e.set_network(gc)
        assert e.layout()["dense_stepper"] == 1 and e.layout()["dense_encoding"] == 1
        e.alloc_cells(hi - lo, off + lo)
        e.set_states_broadcast(x0)
        e.run_phase(np.arange(hi - lo), steps[lo:hi], proj, const, step_base=base[lo:hi], sigma=sigma, seed=seed)
        return e, e.get_states().T

    e, full = run(0, cells)
    # one CTA-pair launch per timestep while more than 12 cell tiles are live (steps 0..6 here); the steps after that share
    # one steps-inside launch of wide tiles
    assert 6 <= e.path_launches()["dense_gemm"] <= smax
    assert np.isfinite(full).all() and np.abs(full).max() <= 2.0
    sample = np.concatenate([[0, 1, 2], rng.choice(cells, size=5, replace=False)])
    noise = np.zeros((len(sample), smax, n))
    for i, c in enumerate(sample):
        for s in range(steps[c]):
            noise[i, s] = e.debug_noise(off + int(c), int(base[c]) + s, sigma, seed)
    want = np.repeat(x0.astype(np.float32).astype(np.float64)[None, :], len(sample), axis=0)
    fo.run_phase(gc, want, steps[sample], proj, const, 0.01, noise)
    err = np.max(np.abs(full[sample] - want))
    assert err <= 1e-4, err
    _, a = run(0, 2432)
    _, b = run(2432, cells)
    assert np.array_equal(full[:2432], a) and np.array_equal(full[2432:], b)
    monkeypatch.setenv("SCR_DENSE_PAIR", "0")
    _, single = run(0, cells)
    assert np.max(np.abs(single - full)) <= 1e-6


@pytest.mark.parametrize("prec,tol", [("f32", 1e-4), ("f64", 1e-9)])
def test_c2_shaped_prepared_run_replayed_through_the_oracle(prec, tol):
    """run_all (the scalable path bench.py uses) on TRRUST, the 7-type / 4-leaf tree, 10 000 cells, fixed iterations:
    for a sample of cells the whole multi-phase trajectory is recomputed by the oracle from the noise the free-running
    stepper used (scr_debug_noise), the member step counts and the node vectors."""
    import bench
    from scrutiny_b200 import engine, network
    from scrutiny_b200.simulate import LineageSimulation
    seed_all(1337)
    gc, tf = network.generate_gene_correlation_matrix(0, "TRRUST")
    n, cells, T = gc.shape[0], 10000, 30
    tree = bench.build_tree(n, cells, tf)
    x0 = 2.0 * np.random.rand(n) - 1.0
    gc = gc / 0.1                                             # c2's searched alpha
    e = engine.Engine(0, prec)
    e.set_network(gc)
    e.alloc_cells(cells)
    e.set_states_broadcast(x0)
    sim = LineageSimulation(e, cells, seed=4242)
    res = sim.run_all(tree, fixed_iterations=T, do_counts=False, seed=4242)
    rng = np.random.RandomState(3)
    sample = np.sort(rng.choice(cells, size=24, replace=False))
    got = np.stack([e.get_states(int(c), 1)[:, 0] for c in sample])
    W = fo.as_operator(gc)
    x0r = x0.astype(np.float32).astype(np.float64) if prec == "f32" else x0
    want = np.repeat(x0r[None, :], len(sample), axis=0)
    done = np.zeros(len(sample), dtype=np.int64)
    for p in sim.plan:                                        # depth-first node order = phase order
        node = p["node"]
        if not len(p["members"]):
            continue
        u = sim._member_uniforms(p)
        mem_steps = dict(zip(p["mem_loc"].tolist(), np.minimum((u * (T + 1)).astype(np.int64), T).tolist()))
        below = set(p["below_loc"].tolist())
        st = np.array([mem_steps.get(int(c), T if int(c) in below else 0) for c in sample])
        noise = np.zeros((len(sample), T, n))
        for i, c in enumerate(sample):
            for s in range(st[i]):
                noise[i, s] = e.debug_noise(int(c), int(done[i]) + s, 0.05, 4242)
        fo.run_phase(W, want, st, node.proj, node.const, 0.01, noise)
        done += st
    assert np.array_equal(done, res["iterations"][sample])
    err = np.max(np.abs(got - want))
    assert err <= tol, (prec, err)
    assert done.max() > 2 * T                                 # leaves went through three phases


def test_transform_data_lambda_through_the_public_module(golden):
    """MS:851-878 at API level: same seed -> same gamma targets and size factors as the port; the Poisson rate matrix
    the GPU read-out samples from equals the port's to 1e-12 (relative), the remap included."""
    import scrutiny_b200.MatriSeq as ms
    ms.init(outputDir=None, seed=1337, verbose=False, precision="f64")
    S = golden("pipeline")["final_states"]
    n, cells = S.shape
    for kw in (dict(), dict(gammaDistMean=False, geneScale=1.5, expLambda=0.8, expScale=0.7)):
        seed_all(5)
        cap_port = {}
        _, want_size = port.transform_data(n, cells, S.copy(), capture=cap_port, **kw)
        seed_all(5)
        cap = {}
        counts, size = ms.transformData(n, cells, S.copy(), capture=cap, **kw)
        assert np.array_equal(size, want_size)
        lam, want = cap["lambda"], np.maximum(cap_port["lambda"], 0.0)
        assert lam.shape == want.shape == (n, cells)
        assert np.max(np.abs(lam - want) / np.maximum(want, 1e-300)) <= 1e-12
        assert counts.shape == (n, cells) and counts.min() >= 0
