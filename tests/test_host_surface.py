"""Host-side pieces of the MatriSeq-compatible module against the reference's golden outputs
(tree construction, initial states) and the multi-rank bookkeeping of the scheduler."""
import numpy as np
import pytest

from conftest import seed_all


def test_tree_and_initial_states_match_reference(golden):
    import scrutiny_b200.MatriSeq as ms
    g = golden("pipeline")
    ms.init(outputDir=None, seed=1337, verbose=False)
    n, cells = 40, 166
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw", powerLawExponent=2.3)
    assert np.array_equal(gc, g["gc_unscaled"]) and np.array_equal(tf, g["tf"])
    tree, projM, constM = ms.formCellTypeTree(n, cells, [-1, 0, 0], [56, 50, 60], [0.1, 0.15, 0.2], tf, [0, 0, 0])
    assert np.array_equal(projM, g["projM"]) and np.array_equal(constM, g["constM"])
    for t in range(3):
        assert list(tree[t].cellTypeMembers) == list(g["members%d" % t])
    assert list(tree.head().children) == list(g["head_children"])
    assert list(tree[0].childCellTypeMembers) == list(g["child_members0"])
    assert list(tree.head().childCellTypeMembers) == list(tree[0].childCellTypeMembers) + list(tree[0].cellTypeMembers)
    init = ms.generateInitialStates(n, cells)
    assert np.array_equal(init, g["init"])
    order = [nd.identifier for nd in tree.walk()]
    assert order == [-1, 0, 1, 2]


def test_invalid_structure_raises_like_reference():
    import pytest
    import scrutiny_b200.MatriSeq as ms
    ms.init(outputDir=None, verbose=False)
    with pytest.raises(Exception, match="Invalid network structure"):
        ms.generateGeneCorrelationMatrix(10, "Nope")


def test_initial_states_per_cell_branch():
    import scrutiny_b200.MatriSeq as ms
    from oracle import matriseq_port as port
    seed_all(3)
    a = ms.generateInitialStates(7, 5, sameInitialCell=False, geneMeanLevels=0.25)
    seed_all(3)
    b = port.initial_states(7, 5, sameInitialCell=False, geneMeanLevels=0.25)
    assert np.array_equal(a, b)


def test_shard_bounds_cover_all_cells():
    from scrutiny_b200.simulate import shard_bounds
    for cells in (1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(cells, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == cells
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(lo % 4 == 0 for lo, hi in spans if lo < cells)


def test_save_data_writes_the_seven_artefacts(tmp_path):
    import scrutiny_b200.MatriSeq as ms
    ms.init(outputDir=str(tmp_path / "out"), verbose=False)
    n, cells = 4, 3
    ms.saveData(np.ones((n, cells)), np.zeros(cells, dtype="int64"), np.eye(n), np.ones((2, n)), np.zeros((2, n)),
                np.ones(cells), np.arange(cells) * 0.01)
    for name in ("synthscRNAseq", "cellTypesRecord", "gc", "projMatrix", "constMatrix", "cellSizeFactors",
                 "synthPseudotimes"):
        assert (tmp_path / "out" / (name + ".npy")).exists()
        assert (tmp_path / "out" / (name + ".gzip")).exists()
    assert np.load(tmp_path / "out" / "synthscRNAseq.npy").shape == (n, cells)
    from scrutiny_b200.rdata import read_rdata
    name, back = read_rdata(str(tmp_path / "out" / "projMatrix.gzip"))
    assert name == "projMatrix" and np.array_equal(back, np.ones((2, n)))
    name, back = read_rdata(str(tmp_path / "out" / "cellTypesRecord.gzip"))
    assert back.dtype == np.int64 and back.shape == (cells,)


def test_r_workspace_writer_reproduces_files_written_by_r():
    """tests/golden/r_*.gzip were written by R itself (shipped by the reference under R_files/):
    reading them and re-serialising must give the same bytes."""
    import gzip
    import os
    from scrutiny_b200 import rdata
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, shape, dtype in (("cellTypesRecord", (100,), np.int64), ("cellSizeFactors", (100,), np.float64),
                               ("projMatrix", (3, 500), np.float64)):
        path = os.path.join(here, "r_%s.gzip" % name)
        got_name, value = rdata.read_rdata(path)
        assert got_name == name and value.shape == shape and value.dtype == dtype
        assert rdata.serialize(name, value) == gzip.open(path).read()
    _, proj = rdata.read_rdata(os.path.join(here, "r_projMatrix.gzip"))
    assert set(np.unique(proj)) <= {0.0, 1.0} and (proj == 0).sum(axis=1).tolist() == sorted((proj == 0).sum(axis=1).tolist())


def test_network_as_csr_equals_the_dense_reference_matrix(golden):
    """generate_csr draws exactly what the reference draws (MS:66-127) without forming the two dense (n, n) arrays; the
    sparse=True result of the module is the same matrix as the dense one for every structure."""
    import scipy.sparse as sp
    import scrutiny_b200.MatriSeq as ms
    from scrutiny_b200 import network
    for kind, kw, n in (("SimulatedPowerLaw", dict(powerLawExponent=2.3), 40), ("Random", dict(networkSparsity=0.7), 60),
                        ("TRRUST", {}, 0), ("SimulatedPowerLaw", dict(normalizeWRows=True), 80)):
        ms.init(outputDir=None, seed=1337, verbose=False)
        dense, tf = ms.generateGeneCorrelationMatrix(n, kind, **kw)
        state = np.random.get_state()[1][:5].copy()
        ms.init(outputDir=None, seed=1337, verbose=False)
        sparse, tf2 = ms.generateGeneCorrelationMatrix(n, kind, sparse=True, **kw)
        assert sp.issparse(sparse) and np.array_equal(sparse.toarray(), dense) and np.array_equal(tf, tf2)
        assert np.array_equal(np.random.get_state()[1][:5], state)            # same stream position afterwards
        rowptr, col, val = network.dense_to_csr(sparse)
        r2, c2, v2 = network.dense_to_csr(dense)
        assert np.array_equal(rowptr, r2) and np.array_equal(col, c2) and np.array_equal(val, v2)
        assert np.array_equal(tf, np.nonzero((dense != 0).sum(axis=0))[0]) or kind == "TRRUST"
    g = golden("pipeline")
    ms.init(outputDir=None, seed=1337, verbose=False)
    gc, tf = ms.generateGeneCorrelationMatrix(40, "SimulatedPowerLaw", powerLawExponent=2.3, alpha=0.5, sparse=True)
    assert np.array_equal(gc.toarray(), g["gc_unscaled"] / 0.5)


def test_scalable_tree_assignment():
    """formCellTypeTree(scalable=True): same tree shape, node vectors and type sizes as the reference's assignment, members
    a uniformly random partition held as sorted arrays."""
    import scrutiny_b200.MatriSeq as ms
    n, cells = 30, 5000
    parents, sizes, props = [-1, 0, 0, 1, 1], [500, 1000, 1500, 1000, 1000], [0.1, 0.1, 0.2, 0.0, 0.3]
    ms.init(outputDir=None, seed=4, verbose=False)
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw")
    ref_tree, projR, constR = ms.formCellTypeTree(n, cells, parents, sizes, props, tf, [0] * 5)
    ms.init(outputDir=None, seed=4, verbose=False)
    gc, tf = ms.generateGeneCorrelationMatrix(n, "SimulatedPowerLaw")
    state = __import__("random").getstate()
    tree, projM, constM = ms.formCellTypeTree(n, cells, parents, sizes, props, tf, [0] * 5, scalable=True)
    assert projM.shape == projR.shape and [nd.identifier for nd in tree.walk()] == [nd.identifier for nd in ref_tree.walk()]
    allm = np.concatenate([tree[k].cellTypeMembers for k in range(5)])
    assert np.array_equal(np.sort(allm), np.arange(cells))
    for k in range(5):
        m = tree[k].cellTypeMembers
        assert m.dtype == np.int64 and len(m) == sizes[k] and np.all(np.diff(m) > 0)
        assert abs(m.mean() - cells / 2) < 6 * cells / np.sqrt(12 * sizes[k])
        assert len(tree[k].childCellTypeMembers) == len(ref_tree[k].childCellTypeMembers)
    assert sorted(tree[0].childCellTypeMembers.tolist()) == sorted(np.concatenate([tree[k].cellTypeMembers for k in (1, 2, 3, 4)]).tolist())
    assert np.array_equal((projM == 0).sum(axis=1) >= 0, np.ones(5, bool)) and np.all((projM == 0) | (projM == 1))
    assert np.array_equal(constM[projM == 1], np.zeros((projM == 1).sum()))


def test_module_surface_matches_the_reference():
    """Drop-in boundary: every public function of the reference's MatriSeq module (tests/golden/
    matriseq_signatures.json, written from the unmodified module by oracle/make_signatures.py) exists here; the functions of
    the simulation path take the reference's parameters first, in order, with the reference's defaults, and anything added
    comes after them with a default.  The plotting functions are accepted with any arguments (documented no-ops)."""
    import inspect
    import json
    import os
    import scrutiny_b200.MatriSeq as ms
    table = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "matriseq_signatures.json")))
    no_ops = {"analyzeDataBeforeTransformation", "analyzeSCRNAseqData", "plotSCRNAseqData"}
    assert len(table) == 16
    for name, ref in table.items():
        assert hasattr(ms, name), name
        if name in no_ops:
            assert getattr(ms, name)(*range(len(ref["params"]))) is None
            continue
        ours = list(inspect.signature(getattr(ms, name)).parameters.values())
        for i, (pname, pdefault) in enumerate(ref["params"]):
            assert ours[i].name == pname, (name, i, ours[i].name, pname)
            if pdefault is None:
                assert ours[i].default is inspect.Parameter.empty, (name, pname)
            else:
                assert repr(ours[i].default) == pdefault, (name, pname, ours[i].default, pdefault)
        for extra in ours[len(ref["params"]):]:
            assert extra.default is not inspect.Parameter.empty, (name, extra.name)


class _ReadoutDouble:
    """Stand-in for engine.Engine on the read-out path only (no GPU here): what the host code of
    MatriSeq.transformData asks of the library -- per-gene sums in, lambda = exp(expScale (S + shift)) size out."""

    def __init__(self, device=0, precision="f32"):
        self.cells, self.n = 0, 0

    def set_genes(self, n):
        self.n = n

    def alloc_cells(self, cells, global_offset=0):
        self.cells = cells

    def set_states(self, S, cell0=0):
        self.X = np.array(S, dtype=np.float64).T.copy()

    def gene_sums(self):
        return self.X.sum(axis=0)

    def counts(self, shift, size_factors, exp_scale=1.0, seed=0, out=None, out_device_ptr=None, want_lambda=False, **kw):
        lam = np.maximum(np.exp(exp_scale * (self.X + shift[None, :])) * np.asarray(size_factors)[:, None], 0.0)
        got = np.floor(lam).astype(np.int32)
        return (got, lam) if want_lambda else got

    def close(self):
        pass


def test_transform_data_host_side_follows_the_reference(monkeypatch):
    """Host half of transformData (MS:851-878) through the module function with the library replaced by a double:
    the same gamma / normal draws from numpy's stream, the same per-gene shift, the same Poisson rates, and the
    reference's side effect on its argument (MS:858-862 shift the caller's matrix in place)."""
    import scrutiny_b200.MatriSeq as ms
    from oracle import matriseq_port as port
    monkeypatch.setattr(ms, "Engine", _ReadoutDouble)
    n, cells = 23, 17
    rng = np.random.RandomState(5)
    S = rng.uniform(-1, 1, size=(n, cells))
    for kw in (dict(), dict(gammaDistMean=False, geneScale=0.3, expLambda=0.7), dict(expScale=1.7, cellRadiusDisp=0.3)):
        ms.init(outputDir=None, seed=11, verbose=False)
        mine, cap = S.copy(), {}
        counts, size = ms.transformData(n, cells, mine, capture=cap, **kw)
        seed_all(11)
        theirs, ref = S.copy(), {}

        class NoDraw(port.Draws):
            def poisson(self, lam):
                return np.floor(lam)
        ref_counts, ref_size = port.transform_data(n, cells, theirs, draws=NoDraw(), capture=ref, **kw)
        assert np.array_equal(size, ref_size)
        assert np.allclose(mine, theirs, rtol=0, atol=1e-12) and not np.allclose(mine, S)       # in-place shift
        assert np.allclose(cap["lambda"], ref["lambda"], rtol=1e-12)
        assert counts.shape == (n, cells) and counts.dtype == np.float64
        assert np.abs(counts - ref_counts).max() <= 1                                           # floor of rates 1e-12 apart
    ms.init(outputDir=None, seed=1337, verbose=False)                                           # drops the double


def test_deeper_trees_match_reference(golden):
    """formCellTypeTree / generateInitialStates of the module on config c2's tree shape (7 types, depth 3, 'Random'
    network, per-gene mean levels) and on a two-root tree with a row-normalised power-law network: gene for gene and
    member for member what the unmodified reference built (tests/golden/pipeline7.npz, two_roots.npz)."""
    import scrutiny_b200.MatriSeq as ms
    from test_oracle_golden import TREE7, TWO_ROOTS
    for name, seed, spec, net in (("pipeline7", 77, TREE7, dict(networkStructure="Random", networkSparsity=0.85)),
                                  ("two_roots", 88, TWO_ROOTS, dict(networkStructure="SimulatedPowerLaw", powerLawExponent=2.2,
                                                                    normalizeWRows=True))):
        g = golden(name)
        ms.init(outputDir=None, seed=seed, verbose=False)
        n, cells, kinds = spec["n"], sum(spec["num"]), len(spec["num"])
        gc, tf = ms.generateGeneCorrelationMatrix(n, **net)
        assert np.array_equal(gc, g["gc_unscaled"]) and np.array_equal(tf, g["tf"])
        tree, projM, constM = ms.formCellTypeTree(n, cells, spec["parents"], spec["num"], spec["props"], tf, [0] * kinds)
        assert np.array_equal(projM, g["projM"]) and np.array_equal(constM, g["constM"])
        assert list(tree.head().children) == list(g["head_children"])
        for t in range(kinds):
            assert list(tree[t].cellTypeMembers) == list(g["members%d" % t])
            assert list(tree[t].childCellTypeMembers) == list(g["below%d" % t])
        if name == "pipeline7":
            init = ms.generateInitialStates(n, cells, geneMeanLevels=g["mu"])
        else:
            init = ms.generateInitialStates(n, cells, sameInitialCell=False)
        assert np.array_equal(init, g["init"])
        want = [-1, 0, 1, 3, 4, 2, 5, 6] if kinds == 7 else [-1, 0, 2, 1, 3]             # depth first, MS:607-612
        assert [nd.identifier for nd in tree.walk()] == want


def _ladder_noise(rung, pos, step, n, sigma):
    rs = np.random.RandomState((rung * 1000003 + pos * 7919 + step * 31 + 5) % (2 ** 32))
    return sigma * rs.randn(n)


class _LadderDouble:
    """Stand-in for engine.Engine on the alpha search only: probes every slot with the port's own loop on gc / w_div, noise
    a pure function of (rung of the ladder, position in the rung's sample, step) -- what the device generator is keyed by."""

    def __init__(self, device=0, precision="f32"):
        self.cells = 0

    def set_network(self, gc):
        self.gc = np.array(gc, dtype=np.float64)
        self.n = self.gc.shape[0]

    def alloc_cells(self, cells, global_offset=0):
        self.cells = cells

    def set_states(self, S, cell0=0):
        self.X = np.array(S, dtype=np.float64)

    def probe(self, sample_ids, proj, const, dt=0.01, sigma=0.05, mu=0.0, thresh=0.1, max_iter=500, avg_num=20, seed=0,
              stream=0, w_div=None, **kw):
        from oracle import matriseq_port as port
        from scrutiny_b200.engine import STREAM_ALPHA
        per = self.per
        out = np.zeros(len(sample_ids), dtype=np.int32)
        for s in sample_ids:
            rung, pos = (stream - STREAM_ALPHA) + s // per, s % per

            class Keyed(port.Draws):
                def normal_vec(self, scale, n, tag=None):
                    return _ladder_noise(rung, pos, tag[2], n, scale)
            out[s], _ = port.probe_one_cell(self.n, self.gc / w_div[s], dt, self.X[:, s].copy(), proj, const, sigma, mu, thresh,
                                            max_iter, avg_num, Keyed(), cell=s, stream="alpha")
        return out

    def close(self):
        pass


@pytest.mark.parametrize("batch", [1, 3, 24])
def test_batched_alpha_ladder_is_the_sequential_search(monkeypatch, batch):
    """findOptimalAlpha evaluates ``batch`` rungs per device call; the candidates, the per-rung samples, the switch to 0.01
    steps after the first partial success (MS:449-451), the stopping rung and the final state of ``random`` must be those of
    the reference's one-rung-at-a-time loop (MS:438-497) -- here with per-cell initial states so that partial success occurs."""
    import random
    import scrutiny_b200.MatriSeq as ms
    from oracle import matriseq_port as port
    n, cells, num = 24, 40, 7
    per = -(-num // 4) * 4
    _LadderDouble.per = per
    monkeypatch.setattr(ms, "Engine", _LadderDouble)
    kw = dict(maxNumIterations=140, convDistAvgNum=10, convergenceThreshold=0.3)

    def inputs(mod_init):
        mod_init()
        gc, tf = port.generate_network(n, "Random", networkSparsity=0.7)
        tree, _, _ = port.form_tree(n, cells, [-1], [cells], [0.1], tf, [0])
        return gc, tree, port.initial_states(n, cells, sameInitialCell=False)

    gc, tree, states = inputs(lambda: seed_all(123))
    rungs, samples = [], []

    class Seq(port.Draws):
        def sample(self, population, k):
            samples.append(random.sample(population, k))
            return samples[-1]

        def normal_vec(self, scale, n_, tag=None):
            return _ladder_noise(len(samples) - 1, samples[-1].index(tag[1]), tag[2], n_, scale)
    want = port.optimal_alpha(n, gc, tree, states, num, draws=Seq(), log=rungs, **kw)
    want_state = random.getstate()
    assert any(0 < r < 1 for _, r in rungs), rungs                    # the case exercises the 0.01 steps
    assert abs(rungs[-1][0] - rungs[-2][0] - 0.01) < 1e-12

    gc2, tree2, states2 = inputs(lambda: ms.init(outputDir=None, seed=123, verbose=False))

    class ModTree:                                                    # the module reads head().children / [child].proj, .const
        def head(self):
            return tree2.head()

        def __getitem__(self, k):
            return tree2[k]
    got = ms.findOptimalAlpha(n, gc2, ModTree(), states2, num, batch=batch, **kw)
    assert got == want
    assert random.getstate() == want_state
    ms.init(outputDir=None, seed=1337, verbose=False)


def _keyed_noise(stream, cell, step, n, sigma):
    rs = np.random.RandomState(({"step": 1, "probe": 2}[stream] * 2000003 + int(cell) * 7919 + int(step) * 13) % (2 ** 32))
    return sigma * rs.randn(n)


class _StepperDouble:
    """Stand-in for engine.Engine on the stepping path: the port's own per-cell loops, noise a pure function of
    (stream, cell, steps the cell has done) like the device generator's key."""

    def __init__(self, device=0, precision="f32"):
        self.cells = 0

    def set_network(self, gc):
        self.gc = np.array(gc, dtype=np.float64)
        self.n = self.gc.shape[0]

    def alloc_cells(self, cells, global_offset=0):
        self.cells = cells

    def set_states(self, S, cell0=0):
        self.X = np.array(S, dtype=np.float64)

    def get_states(self, cell0=0, count=None, out=None):
        if out is None:
            return self.X.copy()
        out[...] = self.X
        return out

    def _draws(self):
        from oracle import matriseq_port as port

        class Keyed(port.Draws):
            def uniform_scalar(self):
                return 0.0

            def normal_vec(self, scale, n, tag=None):          # tag = (stream, cell, step) as the port passes it
                return _keyed_noise(tag[0], tag[1], tag[2], n, scale)
        return Keyed()

    def run_phase(self, cell_ids, steps, proj, const, step_base=None, dt=0.01, sigma=0.05, mu=0.0, seed=0, noise=None):
        from oracle import matriseq_port as port
        for c, k, b in zip(cell_ids, steps, step_base if step_base is not None else np.zeros(len(cell_ids), dtype=int)):
            self.X[:, c] = port.cell_next_state(self.gc, self.n, dt, proj, const, self.X[:, c], int(k), sigma, mu,
                                                self._draws(), cell=int(c), step0=int(b))

    def probe(self, sample_ids, proj, const, dt=0.01, sigma=0.05, mu=0.0, thresh=0.1, max_iter=500, avg_num=20, seed=0,
              stream=1, **kw):
        from oracle import matriseq_port as port
        return np.array([port.probe_one_cell(self.n, self.gc, dt, self.X[:, c].copy(), proj, const, sigma, mu, thresh, max_iter,
                                             avg_num, self._draws(), cell=int(c))[0] for c in sample_ids], dtype=np.int32)

    def close(self):
        pass


@pytest.mark.parametrize("compat", [False, True])
def test_lineage_walk_host_side_follows_the_reference(monkeypatch, compat):
    """findAllNextStates of the module with the library replaced by a double built on the port's per-cell loops: the walk
    order, the probe samples (``random.sample``), the member step draws (``np.random.randint``), the type record, the
    iteration bookkeeping and the in-place update of a NON-contiguous state matrix are those of the reference's recursion
    (MS:540-612) -- with ``reference_compat=True`` including its habit of dropping the keyword arguments below the node
    the caller passes (MS:611-612), otherwise forwarding them."""
    import scrutiny_b200.MatriSeq as ms
    from oracle import matriseq_port as port
    monkeypatch.setattr(ms, "Engine", _StepperDouble)
    n, num = 12, [50, 51, 50]
    cells = sum(num)
    user = dict(timeStep=0.02, noiseDisp=0.03, convergenceThreshold=0.3, maxNumIterations=120, convDistAvgNum=6,
                cellsToSample=11)

    def inputs():
        gc, tf = port.generate_network(n, "Random", networkSparsity=0.75)
        tree, _, _ = port.form_tree(n, cells, [-1, 0, 0], num, [0.1, 0.2, 0.2], tf, [0, 0, 0])
        return gc / 1.5, tree, port.initial_states(n, cells, sameInitialCell=False)

    seed_all(321)
    gc, tree, want = inputs()
    wt, wi = np.zeros(cells, dtype="int64"), np.zeros(cells, dtype="int64")
    port.all_next_states(gc, tree[0], tree, want, wt, wi, forward_kwargs=not compat, draws=_StepperDouble()._draws(),
                         **user)
    ms.init(outputDir=None, seed=321, verbose=False)
    gc2, tree2, init = inputs()
    wide = np.zeros((n, 2 * cells))
    states = wide[:, ::2]                                                     # strided columns: not the fast download path
    states[...] = init
    gt, gi = np.zeros(cells, dtype="int64"), np.zeros(cells, dtype="int64")

    class ModTree:
        def walk(self, start):
            yield start
            for ch in start.children:
                yield from self.walk(tree2[ch])
    assert ms.findAllNextStates(gc2, tree2[0], ModTree(), states, gt, gi, reference_compat=compat, **user) is None
    assert np.array_equal(gt, wt) and np.array_equal(gi, wi)
    assert np.array_equal(states, want) and np.array_equal(wide[:, 1::2], np.zeros((n, cells)))
    assert gi.max() > 100 if compat else gi.max() <= 3 * 120                  # children ran with the defaults / the user's limits
    ms.init(outputDir=None, seed=1337, verbose=False)


def test_single_cell_wrappers_host_side(monkeypatch):
    """developCell / findCellNextState / findOptimalIterations of the module (MS:503-537, 683-736, 616-680) with the library
    replaced by the stepping double: argument order, the ``random.sample`` of the probe, the returned shapes and that the
    caller's arrays stay untouched."""
    import random
    import scrutiny_b200.MatriSeq as ms
    from oracle import matriseq_port as port
    monkeypatch.setattr(ms, "Engine", _StepperDouble)
    ms.init(outputDir=None, seed=5, verbose=False)
    n, cells = 14, 20
    gc, tf = port.generate_network(n, "Random", networkSparsity=0.7)
    gc = gc / 1.3
    states = port.initial_states(n, cells, sameInitialCell=False)
    keep = states.copy()
    proj = np.ones(n)
    proj[tf[:2]] = 0
    const = np.zeros(n)
    const[tf[:2]] = [1, -1]
    draws = _StepperDouble()._draws()
    x0 = states[:, 3]
    one = ms.developCell(n, gc, 0.02, x0, proj, const, 0.04, 0.001)
    assert np.array_equal(one, port.develop_cell(n, gc, 0.02, x0, proj, const, 0.04, 0.001, draws, tag=("step", 0, 0)))
    many = ms.findCellNextState(gc, n, 0.02, 7, proj, const, x0, 25, 0.04, 0, False)
    assert np.array_equal(many, port.cell_next_state(gc, n, 0.02, proj, const, x0, 25, 0.04, 0, draws, cell=0))
    assert many.shape == (n,) and np.array_equal(states, keep)
    members = list(range(2, cells))
    st = random.getstate()
    T = ms.findOptimalIterations(n, gc, 0.02, members, proj, const, 0, 0.3, 0.04, states, 150, 6, 5)
    after = random.getstate()
    random.setstate(st)
    cols = random.sample(members, 5)
    assert random.getstate() == after
    # the module probes COPIES placed in slots 0..4 of its context: the double keys the noise by slot
    its = [port.probe_one_cell(n, gc, 0.02, states[:, c].copy(), proj, const, 0.04, 0, 0.3, 150, 6, draws, cell=slot)[0]
           for slot, c in enumerate(cols)]
    assert T == int(sum(its) / 5) - 6 and np.array_equal(states, keep)
    ms.init(outputDir=None, seed=1337, verbose=False)


def test_prepared_mode_draws_are_keyed_by_the_global_cell_id():
    """The scalable scheduler's per-cell draws (member step counts U{0..T} / Poisson(T), size factors N(1, sd)^3: the
    counterparts of MS:589-591 and MS:875-876) are pure functions of (seed, node, GLOBAL cell id): any sharding reproduces
    them, and they have the reference's distributions."""
    from scipy import stats
    from scrutiny_b200.simulate import LineageSimulation, hashed_uniform, poisson_quantile
    ids = np.arange(200000, dtype=np.int64)
    u = hashed_uniform(1337, 11, ids)
    assert u.min() >= 0.0 and u.max() < 1.0 and stats.kstest(u, "uniform").pvalue > 1e-3
    assert np.array_equal(u[5000:9000], hashed_uniform(1337, 11, ids[5000:9000]))                  # any slice: same values
    assert abs(np.corrcoef(u, hashed_uniform(1337, 13, ids))[0, 1]) < 0.01                          # another node
    assert abs(np.corrcoef(u, hashed_uniform(1338, 11, ids))[0, 1]) < 0.01                          # another seed
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 0.01                                            # neighbouring cells
    big = hashed_uniform(1337, 11, ids + 2 ** 40)
    assert stats.kstest(big, "uniform").pvalue > 1e-3 and abs(np.corrcoef(u, big)[0, 1]) < 0.01    # ids beyond 32 bits
    T = 37
    k = np.minimum((u * (T + 1)).astype(np.int32), T)                                              # run_all's U{0..T}
    assert k.min() == 0 and k.max() == T and stats.chisquare(np.bincount(k, minlength=T + 1)).pvalue > 1e-3
    kp = poisson_quantile(u, T)                                                                    # run_all's Poisson(T)
    assert abs(kp.mean() - T) < 0.1 and abs(kp.var() - T) < 0.5
    cells = 100000
    whole = LineageSimulation(None, cells, 0, cells, seed=5)._size_factors(0.14)
    parts = [LineageSimulation(None, cells, lo, hi, seed=5)._size_factors(0.14) for lo, hi in ((0, 33336), (33336, 66672), (66672, cells))]
    assert np.array_equal(np.concatenate(parts), whole)
    z = (np.cbrt(whole) - 1.0) / 0.14
    assert stats.kstest(z, "norm").pvalue > 1e-3 and abs(z.std() - 1) < 0.01


def test_all_seven_r_files_of_the_reference_round_trip_when_present():
    """Build container only (the reference tree is absent on the GPU box): every ``R_files/<artefact>.gzip`` the reference
    ships -- a saved 3-type run, written by R's save(compress=TRUE) through MS:1062-1065 -- is read by rdata.read_rdata
    with the artefact's name, shape, dtype and orientation, and written back byte for byte."""
    import gzip
    import os
    from oracle import reference_harness as rh
    from scrutiny_b200 import rdata
    import scrutiny_b200.MatriSeq as ms
    folder = os.path.join(rh.REFERENCE_ROOT, "R_files")
    if not os.path.isfile(os.path.join(folder, "gc.gzip")):
        pytest.skip("reference tree not present")
    shapes = {"synthscRNAseq": (500, 100), "cellTypesRecord": (100,), "gc": (500, 500), "projMatrix": (3, 500),
              "constMatrix": (3, 500), "cellSizeFactors": (100,), "synthPseudotimes": (100,)}
    got = {}
    for name in ms.ARTEFACTS:
        path = os.path.join(folder, name + ".gzip")
        label, value = rdata.read_rdata(path)
        assert label == name and value.shape == shapes[name], name
        assert value.dtype == (np.int64 if name == "cellTypesRecord" else np.float64)
        assert rdata.serialize(name, value) == gzip.open(path).read(), name
        got[name] = value
    counts = got["synthscRNAseq"]
    assert np.array_equal(counts, np.round(counts)) and counts.min() >= 0           # (genes, cells), integer-valued
    assert set(np.unique(got["constMatrix"])) <= {-1.0, 0.0, 1.0} and set(np.unique(got["projMatrix"])) <= {0.0, 1.0}
    assert ((got["projMatrix"] == 0) >= (got["constMatrix"] != 0)).all()              # a constant is only set where proj is 0
    assert np.bincount(got["cellTypesRecord"]).sum() == 100 and got["synthPseudotimes"].min() >= 0


def test_tree_display_and_traverse_like_the_reference(capsys):
    """tree.py:56-99: ``display`` prints the identifiers indented by depth, ``traverse`` yields the per-type description
    lists depth first (the stepping order) or breadth first -- compared with the reference's own class when it is present."""
    import importlib
    import os
    import sys
    from oracle import reference_harness as rh
    from scrutiny_b200 import lineage
    n = 4
    spec = [(0, -1), (1, 0), (2, 0), (3, 1), (4, 1), (5, 2), (6, -1)]

    def build(mod):
        tree = mod.Tree()
        tree.add_head(-1, [], np.zeros(n), np.ones(n), 0)
        tree[-1].setChildCellTypeMembers(list(range(9)))              # the reference's Node has no default for it
        for k, parent in spec:
            tree.add_node(k, list(range(k + 1)), np.full(n, float(k)), np.ones(n), k * 0.5, parent)
            tree[k].setChildCellTypeMembers(list(range(2 * k)))
        return tree

    mine = build(lineage)
    order = lambda tree, mode: [int(lines[0].split(": ")[1]) for lines in tree.traverse(-1, mode)]     # noqa: E731
    assert order(mine, lineage._DEPTH) == [-1, 0, 1, 3, 4, 2, 5, 6] == [nd.identifier for nd in mine.walk()]
    assert order(mine, lineage._BREADTH) == [-1, 0, 6, 1, 2, 3, 4, 5]
    assert order(mine, lineage._DEPTH)[:1] == [-1] and [int(l[0].split(": ")[1]) for l in mine.traverse(1)] == [1, 3, 4]
    first = next(iter(mine.traverse(3)))
    assert first[4] == "Num cellType Members: 4" and first[5] == "Num cellType Child Members: 6" and len(first) == 6
    mine.display(-1)
    shown = capsys.readouterr().out
    assert shown.splitlines()[:4] == ["-1", "\t 0", "\t\t 1", "\t\t\t 3"]
    if rh.available():
        if rh._PKG_DIR not in sys.path:
            sys.path.insert(0, rh._PKG_DIR)
        theirs = build(importlib.import_module("tree"))
        for mode in (lineage._DEPTH, lineage._BREADTH):
            assert list(mine.traverse(-1, mode)) == list(theirs.traverse(-1, mode))
        assert list(mine.traverse(2)) == list(theirs.traverse(2))
        theirs.display(-1)
        assert capsys.readouterr().out == shown


def test_add_tree_children_is_the_reference_helper(golden):
    """MS:243-297 as a public function: driven the way formCellTypeTree drives it (MS:217-232), it builds the tree of the
    7-type fixture -- and, in the build container, returns and records exactly what the reference's own helper does."""
    import random
    import scrutiny_b200.MatriSeq as ms
    from scrutiny_b200.lineage import Tree
    from oracle import reference_harness as rh
    from test_oracle_golden import TREE7
    g = golden("pipeline7")
    n, kinds = TREE7["n"], 7
    members = [list(g["members%d" % t]) for t in range(kinds)]

    def drive(mod, tree_cls, seed):
        random.seed(seed)
        tree = tree_cls()
        head = tree.add_head(-1, [], np.zeros(n), np.ones(n), 0)
        head.setChildCellTypeMembers(list(range(sum(TREE7["num"]))))
        pm, cm = np.zeros((kinds, n)), np.zeros((kinds, n))
        got = mod.addTreeChildren(n, tree, TREE7["parents"], members, TREE7["props"], np.ones(n), np.zeros(n), pm, cm, [0] * kinds,
                                  -1, g["tf"])
        return tree, pm, cm, got

    tree, pm, cm, got = drive(ms, Tree, 4)
    for t in range(kinds):
        assert list(tree[t].cellTypeMembers) == members[t] and list(tree[t].childCellTypeMembers) == list(g["below%d" % t])
        assert np.array_equal(tree[t].proj, pm[t]) and np.array_equal(tree[t].const, cm[t])
    assert list(tree.head().children) == [0] and got == list(tree.head().childCellTypeMembers) + members[-1]
    for child, parent in ((1, 0), (3, 1), (6, 2)):                     # constants are inherited down the lineage
        assert ((pm[child] == 0) >= (pm[parent] == 0)).all()
        assert np.array_equal(cm[child][pm[parent] == 0], cm[parent][pm[parent] == 0])
    if rh.available():
        import importlib
        import sys
        ref = rh.load_reference()
        theirs, rpm, rcm, rgot = drive(ref, importlib.import_module("tree").Tree, 4)
        assert got == rgot and np.array_equal(pm, rpm) and np.array_equal(cm, rcm)
        for t in [-1] + list(range(kinds)):
            assert list(tree[t].children) == list(theirs[t].children)
            assert list(tree[t].childCellTypeMembers) == list(theirs[t].childCellTypeMembers)
