"""The MatriSeq-compatible module on the GPU: deterministic parity where noiseDisp = 0 makes the
reference deterministic, distributional checks elsewhere, and the example.py call sequence."""
import random

import numpy as np
import pytest

from conftest import seed_all
from oracle import matriseq_port as port

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ms():
    import scrutiny_b200.MatriSeq as m
    m.init(outputDir=None, seed=1337, verbose=False, precision="f64")
    return m


def small_setup(ms, n=40, cells=166):
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw", powerLawExponent=2.3)
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1, 0, 0], [56, 50, 60], [0.1, 0.15, 0.2], tf, [0, 0, 0])
    init = ms.generateInitialStates(n, cells)
    return gc, tf, tree, projM, constM, init


def test_noise_free_functions_equal_the_port(ms):
    """developCell / findCellNextState / findOptimalIterations with noiseDisp = 0 are deterministic
    in the reference: the GPU float64 path must reproduce the port to rounding."""
    gc, tf, tree, _, _, init = small_setup(ms)
    gc = gc / 0.27
    node = tree[1]
    x = init[:, 0]
    got = ms.developCell(40, gc, 0.01, x, node.proj, node.const, 0.0, 0.1)
    want = port.develop_cell(40, gc, 0.01, x, node.proj, node.const, 0.0, 0.1)
    assert np.max(np.abs(got - want)) <= 1e-14
    got = ms.findCellNextState(gc, 40, 0.01, 0, node.proj, node.const, x, 300, 0.0, 0, False)
    want = port.cell_next_state(gc, 40, 0.01, node.proj, node.const, x, 300, 0.0, 0)
    assert np.max(np.abs(got - want)) <= 1e-11
    members = list(range(60))
    state_g, state_p = random.getstate(), None
    T_gpu = ms.findOptimalIterations(40, gc, 0.01, members, node.proj, node.const, 0, 0.1, 0.0, init, 500, 20, 50)
    random.setstate(state_g)
    T_port = port.optimal_iterations(40, gc, 0.01, members, node.proj, node.const, 0, 0.1, 0.0, init, 500, 20, 50)
    assert T_gpu == T_port


def test_alpha_search_lands_on_the_reference_value(ms, golden):
    g = golden("pipeline")
    gc, tf, tree, _, _, init = small_setup(ms)
    alpha = ms.findOptimalAlpha(40, gc, tree, init, numToSample=10, maxNumIterations=120, convDistAvgNum=10,
                                convergenceThreshold=0.3, batch=8)
    assert abs(alpha - float(g["alpha"])) <= 0.045, alpha       # different noise stream: +-2 rungs
    k = round(alpha * 100)
    assert abs(alpha * 100 - k) < 1e-9                            # stays on the 0.01 grid of MS:449-452


def test_alpha_search_noise_free_is_exact(ms):
    """noiseDisp = 0: the sequential port and the batched GPU ladder must return the same alpha and
    leave ``random`` in the same state."""
    gc, tf, tree, _, _, init = small_setup(ms)
    st = random.getstate()
    a_gpu = ms.findOptimalAlpha(40, gc, tree, init, numToSample=6, noiseDisp=0.0, maxNumIterations=150,
                                convDistAvgNum=10, convergenceThreshold=0.3, batch=5)
    after_gpu = random.getstate()
    random.setstate(st)
    a_port = port.optimal_alpha(40, gc, tree, init, 6, noiseDisp=0.0, maxNumIterations=150, convDistAvgNum=10,
                                convergenceThreshold=0.3)
    assert a_gpu == a_port
    assert after_gpu == random.getstate()


def test_lineage_run_bookkeeping_and_statistics(ms, golden):
    g = golden("pipeline")
    gc, tf, tree, projM, constM, init = small_setup(ms)
    gc = gc / float(g["alpha"])
    cells = 166
    states = init.copy()
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    assert ms.findAllNextStates(gc, tree.head(), tree, states, types, iters) is None
    assert np.array_equal(types, g["types"])                      # same host RNG sequence -> same assignment
    T0 = iters[np.array(tree[1].cellTypeMembers)].min()
    assert iters[np.array(tree[0].cellTypeMembers)].max() <= T0   # members stop inside the root phase
    assert abs(T0 - 480) <= 40                                    # reference: T(type 0) = 480
    assert np.isfinite(states).all() and np.abs(states).max() <= 1.0 + 1e-9
    assert not np.array_equal(states, init)
    # clamped genes carry their constant exactly (proj = 0 => x = const, MS:535)
    for t in (1, 2):
        mem = np.array(tree[t].cellTypeMembers)
        done = mem[iters[mem] > T0]
        frozen = np.nonzero(projM[t] == 0)[0]
        assert np.array_equal(states[np.ix_(frozen, done)], np.repeat(constM[t][frozen][:, None], len(done), axis=1))
    # per-type gene means agree with the reference run (different noise, same dynamics)
    ref = g["final_states"]
    for t in range(3):
        mem = np.array(tree[t].cellTypeMembers)
        c = np.corrcoef(states[:, mem].mean(axis=1), ref[:, mem].mean(axis=1))[0, 1]
        assert c > 0.9, (t, c)


def test_transform_data_matches_reference_statistics(ms, golden):
    g = golden("pipeline")
    S = g["final_states"]
    n, cells = S.shape
    seed_all(5)
    want_counts, want_size = port.transform_data(n, cells, S.copy())
    seed_all(5)
    counts, size = ms.transformData(n, cells, S.copy())
    assert counts.shape == (n, cells) and counts.dtype == np.float64
    assert np.array_equal(size, want_size)                        # same host draws in the same order
    assert np.array_equal(counts, np.round(counts)) and counts.min() >= 0
    lib_c = np.corrcoef(counts.sum(axis=0), want_counts.sum(axis=0))[0, 1]
    gene_c = np.corrcoef(np.log1p(counts.mean(axis=1)), np.log1p(want_counts.mean(axis=1)))[0, 1]
    assert lib_c > 0.8 and gene_c > 0.95, (lib_c, gene_c)
    assert abs((counts == 0).mean() - (want_counts == 0).mean()) < 0.03


def test_example_sequence_runs_end_to_end(tmp_path):
    import scrutiny_b200.MatriSeq as ms
    ms.init(outputDir=str(tmp_path / "MatriSeq"), seed=7, verbose=False)
    n, cells = 120, 200
    gc, tf = ms.generateGeneCorrelationMatrix(n=n, networkStructure="SimulatedPowerLaw")
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1, -1], [100, 100], [0.0, 0.0], tf, [0, 0])
    states = ms.generateInitialStates(n, cells)
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=50)
    assert abs(alpha - 0.02) < 1e-12                              # two roots => 0.02
    gc /= alpha
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree.head(), tree, states, types, iters)
    pseudo = 0.01 * iters
    counts, size = ms.transformData(n, cells, states)
    ms.saveData(counts, types, gc, projM, constM, size, pseudo)
    out = np.load(tmp_path / "MatriSeq" / "synthscRNAseq.npy")
    assert out.shape == (n, cells) and (np.bincount(types) == 100).all() and iters.max() > 50


def test_example_artefacts_resemble_the_ones_the_reference_ships(tmp_path):
    """the reference pins no values, but it SHIPS the output of its example.py (500 genes, 200 cells, two
    root types); tests/golden/example_fix_summary.json (oracle/make_fix_summary.py) holds shapes, dtypes and summary statistics
    of those files.  The same call sequence through this module must write artefacts of the same shapes, dtypes and
    orientation whose statistics sit inside the spread the reference's own algorithm shows from seed to seed (port, 5 seeds:
    zero fraction 0.50-0.55, library size 890-1062, T 440-480, 1027-1313 links)."""
    import json
    import os
    import scrutiny_b200.MatriSeq as ms
    from conftest import GOLDEN
    fix = json.load(open(os.path.join(GOLDEN, "example_fix_summary.json")))
    st = fix["stats"]
    out = tmp_path / "MatriSeq"
    ms.init(outputDir=str(out), seed=1337, verbose=False)
    n, cells = 500, 200
    gc, tf = ms.generateGeneCorrelationMatrix(n=n, networkStructure="SimulatedPowerLaw")
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1, -1], [100, 100], [0.0, 0.0], tf, [0, 0])
    states = ms.generateInitialStates(n, cells)
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=50)
    gc /= alpha
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    ms.findAllNextStates(gc, tree.head(), tree, states, types, iters)
    counts, size = ms.transformData(n, cells, states)
    ms.saveData(counts, types, gc, projM, constM, size, 0.01 * iters)
    for name in ms.ARTEFACTS:
        a = np.load(out / (name + ".npy"))
        assert list(a.shape) == fix[name]["shape"] and a.dtype.str == fix[name]["dtype"], name
        assert (out / (name + ".gzip")).is_file()
    c = np.load(out / "synthscRNAseq.npy")
    assert np.array_equal(c, np.round(c)) and c.min() >= 0 and st["counts_integer_valued"]
    assert abs((c == 0).mean() - st["zero_fraction"]) < 0.1
    assert 0.6 * st["mean_library_size"] < c.sum(axis=0).mean() < 1.6 * st["mean_library_size"]
    assert 0.3 * st["max_count"] < c.max() < 5 * st["max_count"]
    assert 0.4 * st["median_gene_mean"] < np.median(c.mean(axis=1)) < 2.0 * st["median_gene_mean"]
    pt = np.load(out / "synthPseudotimes.npy")
    assert pt.min() >= 0 and 3.5 <= pt.max() <= 4.8 and abs(pt.max() - st["pseudotime_max"]) < 1.0
    g = np.load(out / "gc.npy")
    assert 45.0 < np.abs(g).max() <= 50.0 and abs(np.abs(g).max() - st["gc_abs_max"]) < 5.0        # alpha = 0.02 on both sides
    assert abs(np.count_nonzero(g) - st["gc_nnz"]) < 0.35 * st["gc_nnz"]
    assert np.bincount(np.load(out / "cellTypesRecord.npy")).tolist() == st["cells_per_type"]
    assert (np.load(out / "projMatrix.npy") == 1).all() and (np.load(out / "constMatrix.npy") == 0).all()
    sf = np.cbrt(np.load(out / "cellSizeFactors.npy"))
    assert abs(sf.mean() - 1.0) < 0.04 and abs(sf.std() - 0.14) < 0.03
