"""GPU parity on the round-2 reference fixtures (tests/golden/pipeline7.npz, two_roots.npz, mu_vector.npz; written by
oracle/make_golden.py from the UNMODIFIED reference): whole lineage runs with the reference's own noise, through the C ABI,
against the final states the reference produced.

  * pipeline7: config c2's tree shape (7 types, depth 3, 'Random' network, gain 10), up to 1227 steps per cell;
  * two_roots: two root types (the walk of MS:607-612 over both subtrees), row-normalised power-law network;
  * mu_vector: findAllNextStates started at a type node with a per-gene geneMeanLevels vector (MS:530, 535), dt = 0.02.

Tolerances: float64 context <= 1e-9 (a numpy replay in another summation order lands at 1e-14); float32 context <= 1e-4 where
the run is short or the gain low (two_roots, mu_vector: a float32 numpy emulation lands at 1e-6), <= 5e-4 for pipeline7 (1200
steps at gain 10: the emulation lands at 1.5e-5, the device's approximate tanh adds to it; the float64 context is the parity path
for such runs).  Measured on B200 (profiles/r2s3_pytest_golden_lineages.log): float64 1.8e-14 / 3.3e-16 /
1.7e-16, float32 4.7e-5 / 3.1e-6 / 1.5e-6.  The module-level run (free-running device noise) is checked on what does not depend on the
noise: types, the tree bookkeeping, and iteration counts inside the reference's range."""
import numpy as np
import pytest

from oracle import matriseq_port as port
from oracle_utils import noise_table, tree_schedule
import test_oracle_golden as tg

pytestmark = pytest.mark.gpu


def _replay(o, rec, log, final, tols, start=None, dt=0.01, mu=0.0):
    from scrutiny_b200 import engine
    n, cells, tree = o["n"], o["cells"], o["tree"]
    noise = noise_table(rec.normals, "step", cells, n)
    if start is None:
        phases = tree_schedule(tree, log, o["iters"], cells)
    else:                                                     # one node: its members step their own counts from 0
        mem = np.asarray(tree[start].cellTypeMembers, dtype=np.int64)
        phases = [(start, mem, o["iters"][mem].astype(np.int64), np.zeros(len(mem), dtype=np.int64))]
    errs = {}
    for prec, tol in tols:
        e = engine.Engine(device=0, precision=prec)
        e.set_network(o["gc"])
        e.alloc_cells(cells)
        e.set_states(o["init"])
        for t, ids, steps, base in phases:
            if not len(ids) or int(steps.max()) == 0:
                continue
            nz = np.zeros((len(ids), int(steps.max()), n))
            for i, (c, b, s) in enumerate(zip(ids, base, steps)):
                nz[i, :s] = noise[c, b:b + s]
            e.run_phase(ids, steps, tree[t].proj, tree[t].const, step_base=base, dt=dt, mu=mu, noise=nz)
        errs[prec] = float(np.max(np.abs(e.get_states() - final)))
        e.close()
    print("lineage replay errors:", errs)
    for prec, tol in tols:
        assert errs[prec] <= tol, errs


def test_seven_type_lineage_against_reference_output(golden):
    g = golden("pipeline7")
    rec, log = port.RecordingDraws(), []
    o = tg.replay_pipeline7(draws=rec, log=log, read_out=False)
    assert np.array_equal(o["final_states"], g["final_states"])          # the port run that recorded the noise IS the reference run
    _replay(o, rec, log, g["final_states"], (("f64", 1e-9), ("f32", 5e-4)))


def test_two_root_lineage_against_reference_output(golden):
    g = golden("two_roots")
    rec, log = port.RecordingDraws(), []
    o = tg.replay_two_roots(draws=rec, log=log)
    assert np.array_equal(o["final_states"], g["final_states"])
    _replay(o, rec, log, g["final_states"], (("f64", 1e-9), ("f32", 1e-4)))


def test_per_gene_mean_levels_lineage_against_reference_output(golden):
    g = golden("mu_vector")
    rec, log = port.RecordingDraws(), []
    o = tg.replay_mu_vector(draws=rec, log=log)
    assert np.array_equal(o["final_states"], g["final"])
    _replay(o, rec, log, g["final"], (("f64", 1e-9), ("f32", 1e-4)), start=0, dt=0.02, mu=g["mu"])


def test_module_run_on_the_seven_type_tree(golden):
    """The drop-in module on the fixture's inputs (device noise, so states differ draw for draw): the alpha ladder ends near
    the reference's alpha, every cell gets the reference's type, leaves have stepped at least as often as any root-type member,
    and the in-place contract holds (MS:540-612)."""
    import scrutiny_b200.MatriSeq as ms
    g = golden("pipeline7")
    ms.init(outputDir=None, seed=77, verbose=False, precision="f64")
    spec = tg.TREE7
    n, cells = spec["n"], sum(spec["num"])
    gc, tf = ms.generateGeneCorrelationMatrix(n, "Random", networkSparsity=0.85)
    tree, projM, constM = ms.formCellTypeTree(n, cells, spec["parents"], spec["num"], spec["props"], tf, [0] * 7)
    states = ms.generateInitialStates(n, cells, geneMeanLevels=g["mu"])
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=8, maxNumIterations=300, convDistAvgNum=10,
                                convergenceThreshold=0.25, geneMeanLevels=g["mu"])
    assert 0.02 <= alpha <= float(g["alpha"]) + 0.1            # same ladder; device noise, so possibly another rung
    gc = gc / float(g["alpha"])
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    alias = states
    ms.findAllNextStates(gc, tree.head(), tree, states, types, iters)
    assert np.array_equal(types, g["types"])
    assert alias is states and not np.array_equal(states, g["init"]) and np.isfinite(states).all()
    assert iters.max() <= 3 * 480 and iters.min() >= 0                 # three levels, T <= maxNumIterations - convDistAvgNum
    root = np.asarray(tree[0].cellTypeMembers)
    for leaf in (3, 4, 5, 6):                                          # a leaf steps T_root + T_parent + its own share
        assert iters[np.asarray(tree[leaf].cellTypeMembers)].min() >= iters[root].max()
    assert np.max(np.abs(states)) <= 1.0 + 0.05                        # tanh dynamics, proj / const in {0, +-1}
    ms.init(outputDir=None, seed=1337, verbose=False)
