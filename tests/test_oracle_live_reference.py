"""Build container only: the port in oracle/matriseq_port.py against the UNMODIFIED reference imported live
(oracle/reference_harness.py), on seeds and shapes that are NOT among the committed fixtures -- every public function of the
simulation path, bit for bit.  Skipped where /root/reference does not exist (the GPU box); the committed fixtures under
tests/golden/ are what travels."""
import random

import numpy as np
import pytest

from oracle import matriseq_port as port
from oracle import reference_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


def both(seed):
    ms = rh.load_reference()
    ms.init(outputDir=None, seed=seed, verbose=False)
    return ms


def reseed(seed):
    np.random.seed(seed)
    random.seed(seed)


@pytest.mark.parametrize("seed,structure,kw", [(201, "SimulatedPowerLaw", dict(powerLawExponent=2.05)),
                                               (202, "SimulatedPowerLaw", dict(powerLawExponent=3.0, normalizeWRows=True)),
                                               (203, "Random", dict(networkSparsity=0.6, maxAlphaBeta=4)),
                                               (204, "Random", dict(networkSparsity=0.97, normalizeWRows=True))])
def test_networks(seed, structure, kw):
    ms = both(seed)
    want_gc, want_tf = ms.generateGeneCorrelationMatrix(37, structure, **kw)
    reseed(seed)
    gc, tf = port.generate_network(37, structure, **kw)
    assert np.array_equal(gc, want_gc) and np.array_equal(tf, want_tf)


@pytest.mark.parametrize("seed,parents,num,props", [(211, [-1, 0, 0, 1], [7, 9, 5, 11], [0.3, 0.2, 0.1, 0.5]),
                                                    (212, [-1, -1, -1], [4, 4, 6], [0.0, 1.0, 0.5]),
                                                    (213, [-1, 0, 1, 2, 3], [3, 3, 3, 3, 3], [0.2] * 5)])
def test_trees_and_initial_states(seed, parents, num, props):
    ms = both(seed)
    n, cells = 25, sum(num)
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw")
    want_tree, want_p, want_c = ms.formCellTypeTree(n, cells, parents, num, props, tf, [0] * len(num))
    want_a = ms.generateInitialStates(n, cells)
    want_b = ms.generateInitialStates(n, cells, sameInitialCell=False, geneMeanLevels=0.5)
    reseed(seed)
    gc2, tf2 = port.generate_network(n, "SimulatedPowerLaw")
    tree, p, c = port.form_tree(n, cells, parents, num, props, tf2, [0] * len(num))
    assert np.array_equal(p, want_p) and np.array_equal(c, want_c)
    for t in range(len(num)):
        assert list(tree[t].cellTypeMembers) == list(want_tree[t].cellTypeMembers)
        assert list(tree[t].childCellTypeMembers) == list(want_tree[t].childCellTypeMembers)
        assert list(tree[t].children) == list(want_tree[t].children)
        assert np.array_equal(tree[t].proj, want_tree[t].proj) and np.array_equal(tree[t].const, want_tree[t].const)
    assert np.array_equal(port.initial_states(n, cells), want_a)
    assert np.array_equal(port.initial_states(n, cells, sameInitialCell=False, geneMeanLevels=0.5), want_b)


@pytest.mark.parametrize("seed,mode", [(221, True), (222, False)])
def test_lineage_run_alpha_and_read_out(seed, mode):
    """Alpha search, the walk from the head over two 50-cell types (children run with the reference's defaults), the
    read-out -- one seeded sequence, compared after every stage."""
    ms = both(seed)
    n, num = 14, [50, 50]
    cells = sum(num)

    def run(m, net, tree_fn, init_fn, alpha_fn, walk_fn, read_fn):
        gc, tf = net(n, "Random", networkSparsity=0.75)
        tree, _, _ = tree_fn(n, cells, [-1, 0], num, [0.1, 0.3], tf, [0, 0])
        states = init_fn(n, cells, sameInitialCell=False)
        alpha = alpha_fn(n, gc, tree, states, 5, maxNumIterations=150, convDistAvgNum=8, convergenceThreshold=0.3)
        gc = gc / alpha
        types = np.zeros(cells, dtype="int64")
        iters = np.zeros(cells, dtype="int64")
        walk_fn(gc, tree[0], tree, states, types, iters, cellDevelopmentMode=mode, maxNumIterations=120, convDistAvgNum=6,
                cellsToSample=10, convergenceThreshold=0.35)
        final = states.copy()
        counts, size = read_fn(n, cells, states, expScale=1.5)
        return alpha, types, iters, final, counts, size

    want = run(ms, ms.generateGeneCorrelationMatrix, ms.formCellTypeTree, ms.generateInitialStates, ms.findOptimalAlpha,
               ms.findAllNextStates, ms.transformData)
    reseed(seed)
    got = run(port, port.generate_network, port.form_tree, port.initial_states, port.optimal_alpha, port.all_next_states,
              port.transform_data)
    assert got[0] == want[0]
    for a, b, what in zip(got[1:], want[1:], ("types", "iters", "final states", "counts", "size factors")):
        assert np.array_equal(a, b), what
    assert want[2].max() > 20


def test_single_cell_functions():
    ms = both(231)
    n = 19
    gc, tf = ms.generateGeneCorrelationMatrix(n, "Random", networkSparsity=0.5)
    gc = gc / 0.8
    x0 = 2 * np.random.rand(n) - 1
    proj = (np.random.rand(n) > 0.3).astype(float)
    const = np.where(proj == 0, 1.0, 0.0)
    mu = 0.002 * np.arange(n)
    st, pst = np.random.get_state(), random.getstate()
    want1 = ms.developCell(n, gc, 0.02, x0, proj, const, 0.04, mu)
    want2 = ms.findCellNextState(gc, n, 0.02, 0, proj, const, x0, 33, 0.04, mu, False)
    states = np.tile(x0[:, None], (1, 12)) + 0.01 * np.arange(12)
    wantT = ms.findOptimalIterations(n, gc, 0.02, list(range(12)), proj, const, 0, 0.3, 0.04, states, 140, 7, 5)
    np.random.set_state(st)
    random.setstate(pst)
    assert np.array_equal(port.develop_cell(n, gc, 0.02, x0, proj, const, 0.04, mu), want1)
    assert np.array_equal(port.cell_next_state(gc, n, 0.02, proj, const, x0, 33, 0.04, mu), want2)
    assert port.optimal_iterations(n, gc, 0.02, list(range(12)), proj, const, 0, 0.3, 0.04, states, 140, 7, 5) == wantT


@pytest.mark.parametrize("seed,structure,kw", [(241, "SimulatedPowerLaw", dict(powerLawExponent=2.4)),
                                               (242, "Random", dict(networkSparsity=0.8, normalizeWRows=True)),
                                               (243, "TRRUST", dict(maxAlphaBeta=7))])
def test_module_input_producers_against_the_live_reference(seed, structure, kw):
    """The product's own host functions (scrutiny_b200.MatriSeq), not the port: network as CSR -> dense, tree, initial
    states, from the same seeds as the unmodified reference."""
    import scrutiny_b200.MatriSeq as mine
    ms = both(seed)
    n = 31
    want_gc, want_tf = ms.generateGeneCorrelationMatrix(n, structure, **kw)
    n = want_gc.shape[0]
    parents, num, props = [-1, 0, 0], [6, 9, 8], [0.2, 0.4, 0.1]
    want_tree, want_p, want_c = ms.formCellTypeTree(n, sum(num), parents, num, props, want_tf, [0, 0, 0])
    want_s = ms.generateInitialStates(n, sum(num), sameInitialCell=False)
    mine.init(outputDir=None, seed=seed, verbose=False)
    gc, tf = mine.generateGeneCorrelationMatrix(n if structure != "TRRUST" else 31, structure, **kw)
    assert np.array_equal(gc, want_gc) and np.array_equal(tf, want_tf)
    tree, p, c = mine.formCellTypeTree(n, sum(num), parents, num, props, tf, [0, 0, 0])
    assert np.array_equal(p, want_p) and np.array_equal(c, want_c)
    for t in range(3):
        assert list(tree[t].cellTypeMembers) == list(want_tree[t].cellTypeMembers)
        assert list(tree[t].childCellTypeMembers) == list(want_tree[t].childCellTypeMembers)
    assert np.array_equal(mine.generateInitialStates(n, sum(num), sameInitialCell=False), want_s)
    mine.init(outputDir=None, seed=1337, verbose=False)
