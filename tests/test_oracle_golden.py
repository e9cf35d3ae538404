"""The port in oracle/matriseq_port.py must reproduce, bit for bit, the outputs the
unmodified reference produced for the same seeded call sequence (tests/golden/*.npz,
made by oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import matriseq_port as port
from conftest import seed_all


def csr_dense(n, r, c, v):
    m = np.zeros((n, n))
    m[r, c] = v
    return m


def test_networks_match_reference(golden):
    g = golden("networks")
    seed_all(11)
    gc, tf = port.generate_network(60, "SimulatedPowerLaw", powerLawExponent=2.5)
    assert np.array_equal(gc, g["pl_gc"]) and np.array_equal(tf, g["pl_tf"])
    seed_all(12)
    gc, tf = port.generate_network(50, "Random", networkSparsity=0.9, normalizeWRows=True)
    assert np.array_equal(gc, g["rnd_gc"]) and np.array_equal(tf, g["rnd_tf"])


def test_trrust_network_matches_reference(golden):
    from scrutiny_b200.network import trrust_adjacency
    g = golden("networks")
    seed_all(13)
    gc, tf = port.generate_network(5, "TRRUST", maxAlphaBeta=6, trrust_adjacency=trrust_adjacency())
    n = int(g["tr_n"])
    assert gc.shape == (n, n)
    r, c = np.nonzero(gc)
    assert np.array_equal(r, g["tr_rows"]) and np.array_equal(c, g["tr_cols"])
    assert np.array_equal(gc[r, c], g["tr_vals"]) and np.array_equal(tf, g["tr_tf"])


def test_single_step(golden):
    g = golden("step")
    class Fixed(port.Draws):
        def normal_vec(self, scale, n, tag=None):
            return g["noise"]
    x1 = port.develop_cell(48, g["gc"], float(g["dt"]), g["x0"], g["proj"], g["const"],
                           float(g["sigma"]), float(g["mu"]), Fixed())
    assert np.array_equal(x1, g["x1"])


def test_cell_steps(golden):
    g = golden("cellsteps")
    noise = g["noise"]
    class Fixed(port.Draws):
        def uniform_scalar(self):
            return 0.0
        def normal_vec(self, scale, n, tag=None):
            return noise[tag[2]]
    xT = port.cell_next_state(g["gc"], 96, 0.01, g["proj"], g["const"], g["x0"], int(g["steps"]),
                              0.05, 0, Fixed(), cell=0)
    assert np.array_equal(xT, g["xT"])


def replay_probe(g):
    seed_all(int(g["seed"]))
    n, cells = 64, 12
    gc, tf = port.generate_network(n, "SimulatedPowerLaw")
    gc = gc / 0.3
    states = 2 * np.random.rand(n, cells) - 1
    return n, gc, states


def test_probe(golden):
    g = golden("probe")
    n, gc, states = replay_probe(g)
    assert np.array_equal(gc, g["gc"]) and np.array_equal(states, g["states"])
    T = port.optimal_iterations(n, gc, 0.01, list(g["members"]), g["proj"], g["const"], 0,
                                float(g["thresh"]), 0.05, states, int(g["maxIter"]),
                                int(g["avgNum"]), int(g["numToSample"]))
    assert T == int(g["T"])


def replay_pipeline(draws=None, log=None):
    draws = draws or port.Draws()
    seed_all(1337)
    n, cells = 40, 166
    gc, tf = port.generate_network(n, "SimulatedPowerLaw", powerLawExponent=2.3)
    tree, projM, constM = port.form_tree(n, cells, [-1, 0, 0], [56, 50, 60], [0.1, 0.15, 0.2], tf,
                                         [0, 0, 0])
    init = port.initial_states(n, cells)
    alpha = port.optimal_alpha(n, gc, tree, init, 10, maxNumIterations=120, convDistAvgNum=10,
                               convergenceThreshold=0.3)
    out = dict(n=n, cells=cells, gc_unscaled=gc.copy(), tf=tf, tree=tree, projM=projM, constM=constM,
               init=init.copy(), alpha=alpha)
    gc = gc / alpha
    states = init
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    port.all_next_states(gc, tree.head(), tree, states, types, iters, draws=draws, log=log)
    out.update(gc=gc, final_states=states.copy(), types=types, iters=iters)
    counts, size = port.transform_data(n, cells, states)
    out.update(counts=counts, size=size)
    return out


def test_pipeline(golden):
    g = golden("pipeline")
    o = replay_pipeline()
    assert np.array_equal(o["gc_unscaled"], g["gc_unscaled"])
    assert np.array_equal(o["projM"], g["projM"]) and np.array_equal(o["constM"], g["constM"])
    for t in range(3):
        assert list(o["tree"][t].cellTypeMembers) == list(g["members%d" % t])
    assert list(o["tree"][0].childCellTypeMembers) == list(g["child_members0"])
    assert o["alpha"] == float(g["alpha"])
    assert np.array_equal(o["types"], g["types"]) and np.array_equal(o["iters"], g["iters"])
    assert np.array_equal(o["final_states"], g["final_states"])
    assert np.array_equal(o["size"], g["size"])
    assert np.array_equal(o["counts"], g["counts"])


def test_transform_branches(golden):
    g = golden("transform_alt")
    seed_all(51)
    n, cells = 20, 15
    S = np.tanh(np.random.normal(size=(n, cells)))
    assert np.array_equal(S, g["S"])
    c1, s1 = port.transform_data(n, cells, S.copy(), expScale=4.0, cellRadiusDisp=0.3)
    c2, s2 = port.transform_data(n, cells, S.copy(), gammaDistMean=False, geneScale=1.5, expLambda=0.7)
    assert np.array_equal(c1, g["c1"]) and np.array_equal(s1, g["s1"])
    assert np.array_equal(c2, g["c2"]) and np.array_equal(s2, g["s2"])


def test_poisson_development_mode(golden):
    g = golden("poisson_mode")
    seed_all(61)
    n, cells = 24, 52
    gc, tf = port.generate_network(n, "Random", networkSparsity=0.8)
    gc = gc / 1.5
    tree, _, _ = port.form_tree(n, cells, [-1], [52], [0.0], tf, [0])
    states = port.initial_states(n, cells, sameInitialCell=False)
    assert np.array_equal(states, g["init"])
    types = np.zeros(cells, dtype="int64"); iters = np.zeros(cells, dtype="int64")
    port.all_next_states(gc, tree[0], tree, states, types, iters, cellDevelopmentMode=False,
                         maxNumIterations=90, convDistAvgNum=6, cellsToSample=7,
                         convergenceThreshold=0.4, noiseDisp=0.02, timeStep=0.05)
    assert np.array_equal(iters, g["iters"]) and np.array_equal(states, g["final"])


# ---- round 2 fixtures: deeper lineage (config c2's tree shape), two roots, per-gene mean levels ---------------------------
TREE7 = dict(n=30, parents=[-1, 0, 0, 1, 1, 2, 2], num=[50, 52, 51, 55, 50, 53, 50], props=[0.05, 0.1, 0.1, 0.15, 0.15, 0.2, 0.2])
TWO_ROOTS = dict(n=26, parents=[-1, -1, 0, 1], num=[51, 50, 52, 50], props=[0.1, 0.1, 0.2, 0.2])


def replay_pipeline7(draws=None, log=None, read_out=True):
    """oracle/make_golden.py case_pipeline7 through the port."""
    draws = draws or port.Draws()
    seed_all(77)
    n, cells = TREE7["n"], sum(TREE7["num"])
    gc, tf = port.generate_network(n, "Random", networkSparsity=0.85)
    tree, projM, constM = port.form_tree(n, cells, TREE7["parents"], TREE7["num"], TREE7["props"], tf, [0] * 7)
    mu = 0.001 * np.cos(np.arange(n))
    init = port.initial_states(n, cells, geneMeanLevels=mu)
    alpha = port.optimal_alpha(n, gc, tree, init, 8, maxNumIterations=300, convDistAvgNum=10, convergenceThreshold=0.25,
                               geneMeanLevels=mu)
    out = dict(n=n, cells=cells, gc_unscaled=gc.copy(), tf=tf, tree=tree, projM=projM, constM=constM, mu=mu, init=init.copy(),
               alpha=alpha)
    gc = gc / alpha
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    port.all_next_states(gc, tree.head(), tree, init, types, iters, draws=draws, log=log)
    out.update(gc=gc, final_states=init.copy(), types=types, iters=iters)
    if read_out:
        out["counts"], out["size"] = port.transform_data(n, cells, init)
    return out


def replay_two_roots(draws=None, log=None):
    """oracle/make_golden.py case_two_roots through the port."""
    draws = draws or port.Draws()
    seed_all(88)
    n, cells = TWO_ROOTS["n"], sum(TWO_ROOTS["num"])
    gc, tf = port.generate_network(n, "SimulatedPowerLaw", powerLawExponent=2.2, normalizeWRows=True)
    tree, projM, constM = port.form_tree(n, cells, TWO_ROOTS["parents"], TWO_ROOTS["num"], TWO_ROOTS["props"], tf, [0] * 4)
    init = port.initial_states(n, cells, sameInitialCell=False)
    alpha = port.optimal_alpha(n, gc, tree, init, 6, maxNumIterations=200, convDistAvgNum=20)
    out = dict(n=n, cells=cells, gc_unscaled=gc.copy(), tf=tf, tree=tree, projM=projM, constM=constM, init=init.copy(), alpha=alpha)
    gc = gc / 0.3
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    port.all_next_states(gc, tree.head(), tree, init, types, iters, draws=draws, log=log)
    out.update(gc=gc, final_states=init.copy(), types=types, iters=iters)
    return out


MU_KW = dict(maxNumIterations=160, convDistAvgNum=8, cellsToSample=9, convergenceThreshold=0.18, noiseDisp=0.03, timeStep=0.02)


def replay_mu_vector(draws=None, log=None):
    """oracle/make_golden.py case_mu_vector through the port."""
    draws = draws or port.Draws()
    seed_all(99)
    n, cells = 28, 40
    gc, tf = port.generate_network(n, "Random", networkSparsity=0.8)
    gc = gc / 1.2
    tree, _, _ = port.form_tree(n, cells, [-1], [40], [0.2], tf, [0])
    mu = np.linspace(-0.003, 0.003, n)
    states = port.initial_states(n, cells, sameInitialCell=False, geneMeanLevels=mu)
    out = dict(n=n, cells=cells, gc=gc, tree=tree, mu=mu, init=states.copy())
    types = np.zeros(cells, dtype="int64")
    iters = np.zeros(cells, dtype="int64")
    port.all_next_states(gc, tree[0], tree, states, types, iters, geneMeanLevels=mu, draws=draws, log=log, **MU_KW)
    out.update(final_states=states.copy(), types=types, iters=iters)
    return out


def _same_tree(o, g, kinds):
    assert np.array_equal(o["gc_unscaled"], g["gc_unscaled"]) and np.array_equal(o["tf"], g["tf"])
    assert np.array_equal(o["projM"], g["projM"]) and np.array_equal(o["constM"], g["constM"])
    assert list(o["tree"].head().children) == list(g["head_children"])
    for t in range(kinds):
        assert list(o["tree"][t].cellTypeMembers) == list(g["members%d" % t])
        assert list(o["tree"][t].childCellTypeMembers) == list(g["below%d" % t])
    assert np.array_equal(o["init"], g["init"])
    assert o["alpha"] == float(g["alpha"])


def test_seven_type_pipeline(golden):
    g = golden("pipeline7")
    o = replay_pipeline7()
    _same_tree(o, g, 7)
    assert np.array_equal(o["types"], g["types"]) and np.array_equal(o["iters"], g["iters"])
    assert np.array_equal(o["final_states"], g["final_states"])
    assert np.array_equal(o["size"], g["size"]) and np.array_equal(o["counts"], g["counts"])


def test_two_root_lineage(golden):
    g = golden("two_roots")
    o = replay_two_roots()
    _same_tree(o, g, 4)
    assert abs(o["alpha"] - 0.02) < 1e-15                      # quirk R2: proj = const = 0 converges on the first rung
    assert np.array_equal(o["types"], g["types"]) and np.array_equal(o["iters"], g["iters"])
    assert np.array_equal(o["final_states"], g["final_states"])


def test_per_gene_mean_levels(golden):
    g = golden("mu_vector")
    o = replay_mu_vector()
    assert np.array_equal(o["gc"], g["gc"]) and np.array_equal(o["init"], g["init"]) and np.array_equal(o["mu"], g["mu"])
    assert np.array_equal(o["tree"][0].proj, g["proj"]) and np.array_equal(o["tree"][0].const, g["const"])
    assert np.array_equal(o["iters"], g["iters"]) and np.array_equal(o["final_states"], g["final"])
