"""World-size-2 (gloo, CPU) test of the multi-rank scheduler: cells sharded by contiguous global id,
the probe's iteration sum and the per-gene sums all-reduced, every per-cell draw keyed by the global
cell id.  The CUDA engine is replaced by a test double built on the oracle (tests may use it), whose
"device noise" is also a pure function of the global cell id -- so two ranks must reproduce the
single-rank run exactly, which is the property the real engine is tested for on the GPU
(test_results_do_not_depend_on_sharding)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleEngine:
    """Stand-in for scrutiny_b200.engine.Engine over oracle/fast_oracle.py (float64)."""

    def __init__(self, W, cells, offset):
        from oracle import fast_oracle as fo
        self.fo, self.W, self.n = fo, fo.as_operator(W), W.shape[0]
        self.cells, self.offset = cells, offset
        self.X = np.zeros((cells, self.n))

    def _noise(self, gid, step, sigma, seed, salt):
        rs = np.random.RandomState((int(gid) * 1000003 + int(step) * 7919 + int(seed) + salt) % (2 ** 32))
        return sigma * rs.randn(self.n)

    def set_states_broadcast(self, x0):
        self.X[:] = x0

    def run_phase(self, cell_ids, steps, proj, const, step_base=None, dt=0.01, sigma=0.05, mu=0.0, seed=0, noise=None):
        for i, (c, k) in enumerate(zip(cell_ids, steps)):
            for s in range(int(k)):
                nz = self._noise(self.offset + c, step_base[i] + s, sigma, seed, 0)
                self.X[c] = self.fo.step_batch(self.W, self.X[c][None], proj, const, dt, nz[None], mu)[0]

    def probe(self, sample_ids, proj, const, dt=0.01, sigma=0.05, mu=0.0, thresh=0.1, max_iter=500, avg_num=20,
              seed=0, stream=1, **kw):
        noise = np.stack([[self._noise(self.offset + c, s, sigma, seed, 17) for s in range(max_iter)] for c in sample_ids])
        it, _, _ = self.fo.probe(self.W, self.X[np.asarray(sample_ids)], proj, const, dt, noise, thresh, max_iter, avg_num, mu)
        return it.astype(np.int32)

    def probe_multi(self, samples, projs, consts, seeds, **kw):
        self.multi_calls = getattr(self, "multi_calls", 0) + 1
        return [self.probe(s, p, c, seed=sd, **kw) for s, p, c, sd in zip(samples, projs, consts, seeds)]

    def gene_sums(self):
        return self.X.sum(axis=0)

    def counts(self, shift, size, exp_scale=1.0, seed=0, out=None, out_device_ptr=None, cell_begin=0, cell_count=None):
        r1 = len(self.X) if cell_count is None else cell_begin + cell_count
        if cell_begin == 0:
            self.lam = np.zeros_like(self.X)
        self.lam[cell_begin:r1] = self.fo.poisson_lambda(self.X[cell_begin:r1].T, shift, exp_scale, size).T
        return self.lam[cell_begin:r1]


    def counts_compact(self, shift, size, exp_scale=1.0, seed=0, dtype="uint8", out=None, out_device_ptr=None, cell_begin=0,
                       cell_count=None):
        """Compact read-out double: integers drawn per GLOBAL cell id (rates scaled up so that some exceed a byte)."""
        r1 = len(self.X) if cell_count is None else cell_begin + cell_count
        lam = self.fo.poisson_lambda(self.X[cell_begin:r1].T, shift, exp_scale, size).T
        full = np.stack([np.random.RandomState((self.offset + cell_begin + i) * 31 + int(seed)).poisson(60.0 * row)
                         for i, row in enumerate(lam)])
        esc = 255 if dtype == "uint8" else 65535
        rows, genes = np.nonzero(full >= esc)
        out[...] = np.minimum(full, esc)
        return out, np.stack([rows, genes, full[rows, genes]], axis=1).astype(np.int64)


def build_case():
    from scrutiny_b200.lineage import Tree
    np.random.seed(5)
    n, cells = 24, 44
    W = (np.random.rand(n, n) < 0.15) * np.random.randn(n, n) / 0.6
    tree = Tree()
    tree.add_head(-1, [], np.zeros(n), np.ones(n), 0)
    members = [np.arange(0, 12), np.arange(12, 27), np.arange(27, 44)]
    proj1 = np.ones(n)
    proj1[3] = 0
    const1 = np.zeros(n)
    const1[3] = 1.0
    tree.add_node(0, members[0], np.zeros(n), np.ones(n), 0, -1)
    tree.add_node(1, members[1], const1, proj1, 0, 0)
    tree.add_node(2, members[2], np.zeros(n), np.ones(n), 0, 0)
    tree[0].setChildCellTypeMembers(np.concatenate([members[1], members[2]]))
    tree[-1].setChildCellTypeMembers(np.arange(cells))
    x0 = np.random.uniform(-1, 1, n)
    return n, cells, W, tree, x0


def build_case7():
    """Config c2's tree shape (7 types, 4 leaves, depth 3) over 70 cells in contiguous blocks of 10: with three ranks most
    nodes have no member -- and so no probe sample -- on some rank."""
    from scrutiny_b200.lineage import Tree
    rng = np.random.RandomState(8)
    n, cells = 20, 70
    W = (rng.rand(n, n) < 0.2) * rng.randn(n, n) / 0.7
    parents = [-1, 0, 0, 1, 1, 2, 2]
    bounds = [0, 10, 20, 30, 40, 50, 60, 70]
    tree = Tree()
    tree.add_head(-1, [], np.zeros(n), np.ones(n), 0)
    for k, par in enumerate(parents):
        proj = np.ones(n)
        const = np.zeros(n)
        proj[k] = 0
        const[k] = 1.0 if k % 2 else -1.0
        tree.add_node(k, np.arange(bounds[k], bounds[k + 1]), const, proj, 0, par)
    below = {3: [], 4: [], 5: [], 6: [], 1: [3, 4], 2: [5, 6], 0: [1, 2, 3, 4, 5, 6]}
    for k, kids in below.items():
        tree[k].setChildCellTypeMembers(np.concatenate([np.arange(bounds[c], bounds[c + 1]) for c in kids]) if kids
                                        else np.zeros(0, dtype=np.int64))
    tree[-1].setChildCellTypeMembers(np.arange(cells))
    x0 = rng.uniform(-1, 1, n)
    return n, cells, W, tree, x0


def run_rank(rank, world, port_no, out_path, case="small"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from scrutiny_b200.simulate import Comm, LineageSimulation, shard_bounds
    if world > 1:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=rank, world_size=world)
    n, cells, W, tree, x0 = build_case() if case == "small" else build_case7()
    lo, hi = shard_bounds(cells, rank, world)
    eng = OracleEngine(W, hi - lo, lo)
    eng.set_states_broadcast(x0)
    sim = LineageSimulation(eng, cells, lo, hi, comm=Comm(), seed=99)
    blocks = []   # the two-rank run samples its counts in three row blocks (bench.py's gather overlap), the one-rank run in one
    kw = (dict(cellsToSample=6, maxNumIterations=60, convDistAvgNum=5, convergenceThreshold=0.6) if case == "small" else
          dict(cellsToSample=4, maxNumIterations=40, convDistAvgNum=4, convergenceThreshold=0.7))
    res = sim.run_all(tree, seed=99, counts_blocks=3 if world > 1 else 1, on_counts_block=lambda r0, r1: blocks.append((r0, r1)),
                      **kw)
    assert blocks[0][0] == 0 and blocks[-1][1] == hi - lo and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert len(blocks) == (3 if world > 1 else 1)
    np.savez(out_path, lo=lo, hi=hi, X=eng.X, iters=res["iterations"], types=res["types"], size=res["size_factors"],
             lam=eng.lam, glob=res["global_cell_steps"], T=np.array(sorted(sim.node_iterations.items())),
             multi_calls=getattr(eng, "multi_calls", 0))
    if world > 1:
        dist.destroy_process_group()


def run_rank_shards(rank, world, port_no, out_dir):
    """The compact, sharded read-out of bench.py / a 10^6-cell run: every rank samples its own rows as unsigned bytes and
    writes its own shard; nothing is gathered."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from scrutiny_b200 import artefacts
    from scrutiny_b200.simulate import Comm, LineageSimulation, shard_bounds
    if world > 1:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=rank, world_size=world)
    n, cells, W, tree, x0 = build_case()
    lo, hi = shard_bounds(cells, rank, world)
    eng = OracleEngine(W, hi - lo, lo)
    eng.set_states_broadcast(x0)
    sim = LineageSimulation(eng, cells, lo, hi, comm=Comm(), seed=99)
    host = np.zeros((hi - lo, n), dtype=np.uint8)
    res = sim.run_all(tree, seed=99, fixed_iterations=7, counts_out=host, counts_dtype="uint8", counts_blocks=2,
                      cellDevelopmentMode=False)                        # member steps ~ Poisson(T) (MS:590-591), keyed by global id
    assert res["overflow"].shape[1] == 3 and len(res["overflow"]) > 0
    os.makedirs(out_dir, exist_ok=True)
    artefacts.save_shard(out_dir, host, res["overflow"], rank)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        artefacts.write_manifest(out_dir, n, [shard_bounds(cells, r, world)[1] - shard_bounds(cells, r, world)[0] for r in range(world)],
                                 np.uint8)


def test_two_ranks_write_count_shards_that_load_as_one_matrix(tmp_path):
    import torch.multiprocessing as mp
    from scrutiny_b200 import artefacts
    run_rank_shards(0, 1, 0, str(tmp_path / "one"))
    port_no = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=run_rank_shards, args=(r, 2, port_no, str(tmp_path / "two"))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    one, two = artefacts.load(str(tmp_path / "one")), artefacts.load(str(tmp_path / "two"))
    assert len(two.blocks) == 2 and one.shape == two.shape == (24, 44)
    a = np.asarray(one)
    assert np.array_equal(a, np.asarray(two)) and a.max() > 255 and a.dtype == np.float64
    assert np.array_equal(two[:, 20:30], a[:, 20:30])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_ranks_reproduce_one_rank(tmp_path):
    import torch.multiprocessing as mp
    single = str(tmp_path / "single.npz")
    run_rank(0, 1, 0, single)
    port_no = _free_port()
    ctx = mp.get_context("spawn")
    procs = []
    for r in range(2):
        p = ctx.Process(target=run_rank, args=(r, 2, port_no, str(tmp_path / ("rank%d.npz" % r))))
        p.start()
        procs.append(p)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    one = np.load(single)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(2)]
    assert int(parts[0]["lo"]) == 0 and int(parts[0]["hi"]) == int(parts[1]["lo"]) and int(parts[1]["hi"]) == 44
    assert np.array_equal(parts[0]["T"], one["T"]) and np.array_equal(parts[1]["T"], one["T"])
    for key in ("X", "iters", "types", "size", "lam"):
        assert np.allclose(np.concatenate([parts[0][key], parts[1][key]]), one[key], rtol=1e-13, atol=1e-13), key
    assert float(parts[0]["glob"]) == float(one["glob"]) == float(one["iters"].sum())
    assert one["iters"].max() > 0 and len(np.unique(one["types"])) == 3


def test_three_ranks_on_the_seven_type_tree_reproduce_one_rank(tmp_path):
    """World size 3 on config c2's tree shape, phases in level order with a level's probes batched: ranks that own no sample
    of a node (or of a whole level) still take part in that round's all-reduce, and the shards put together are the
    single-rank run -- states, iteration counts, types, the probes' T, the Poisson rates of the read-out."""
    import torch.multiprocessing as mp
    single = str(tmp_path / "single.npz")
    run_rank(0, 1, 0, single, "tree7")
    port_no = _free_port()
    ctx = mp.get_context("spawn")
    procs = []
    for r in range(3):
        p = ctx.Process(target=run_rank, args=(r, 3, port_no, str(tmp_path / ("rank%d.npz" % r)), "tree7"))
        p.start()
        procs.append(p)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    one = np.load(single)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(3)]
    assert [int(q["lo"]) for q in parts] == [0, 24, 48] and int(parts[2]["hi"]) == 70
    assert len(one["T"]) == 7 and all(np.array_equal(q["T"], one["T"]) for q in parts)
    for key in ("X", "iters", "types", "size", "lam"):
        assert np.allclose(np.concatenate([q[key] for q in parts]), one[key], rtol=1e-13, atol=1e-13), key
    assert all(float(q["glob"]) == float(one["glob"]) for q in parts) and float(one["glob"]) == float(one["iters"].sum())
    assert int(one["multi_calls"]) == 2                                   # levels {1, 2} and {3, 4, 5, 6}: one batched call each
    # rank 0 (cells 0..23) holds samples of nodes 1 and 2 but of no leaf; rank 1 (24..47) of node 2 alone on level 1 (a plain
    # probe) and of leaves 3, 4; rank 2 (48..69) of leaves only: one batched call each, in different rounds
    assert [int(q["multi_calls"]) for q in parts] == [1, 1, 1]


def test_batched_sibling_probes_give_the_sequential_schedule(tmp_path):
    """run_all probes the children of a node together right after the parent's phase (scr_probe_multi).  An engine without
    probe_multi is probed node by node in the same order; an engine with it gets the siblings in one call.  Iteration
    counts, states and bookkeeping are identical, and both equal probing each node right before its own phase (run_node,
    the reference's order) in the T they find."""
    sys.path.insert(0, ROOT)
    from scrutiny_b200.simulate import Comm, LineageSimulation
    n, cells, W, tree, x0 = build_case()
    kw = dict(seed=99, cellsToSample=6, maxNumIterations=60, convDistAvgNum=5, convergenceThreshold=0.6)

    class NodeByNodeEngine(OracleEngine):
        probe_multi = None

    def run(with_multi):
        eng = (OracleEngine if with_multi else NodeByNodeEngine)(W, cells, 0)
        eng.set_states_broadcast(x0)
        sim = LineageSimulation(eng, cells, 0, cells, comm=Comm(single=True), seed=99)
        return eng, sim, sim.run_all(tree, **kw)

    e1, s1, r1 = run(True)
    e2, s2, r2 = run(False)
    assert getattr(e1, "multi_calls", 0) == 1 and getattr(e2, "multi_calls", 0) == 0   # nodes 1 and 2 are siblings
    assert s1.node_iterations == s2.node_iterations and len(s1.node_iterations) == 3
    assert np.array_equal(e1.X, e2.X) and np.array_equal(r1["iterations"], r2["iterations"])


def test_level_order_of_the_phases_gives_the_depth_first_result(monkeypatch):
    """run_all steps the nodes level by level (all probes of a level side by side) instead of the reference's depth-first
    walk (MS:607-612): subtrees touch disjoint cells and the noise is keyed by (seed, node, cell, iterations done), so
    states, iteration counts, types and the probes' T are those of the depth-first order (SCR_LEVEL_ORDER=0)."""
    sys.path.insert(0, ROOT)
    from scrutiny_b200.simulate import Comm, LineageSimulation
    n, cells, W, tree, x0 = build_case7()
    kw = dict(seed=5, cellsToSample=4, maxNumIterations=40, convDistAvgNum=4, convergenceThreshold=0.7, do_counts=False)

    def run(level):
        monkeypatch.setenv("SCR_LEVEL_ORDER", "1" if level else "0")
        eng = OracleEngine(W, cells, 0)
        eng.set_states_broadcast(x0)
        sim = LineageSimulation(eng, cells, 0, cells, comm=Comm(single=True), seed=5)
        res = sim.run_all(tree, **kw)
        return eng, sim, res

    e1, s1, r1 = run(True)
    e0, s0, r0 = run(False)
    assert getattr(e1, "multi_calls", 0) == 2 and getattr(e0, "multi_calls", 0) == 3    # {1,2},{3,4,5,6} vs {1,2},{3,4},{5,6}
    assert s1.node_iterations == s0.node_iterations and len(s1.node_iterations) == 7
    assert np.array_equal(e1.X, e0.X) and np.array_equal(r1["iterations"], r0["iterations"])
    assert np.array_equal(r1["types"], r0["types"]) and r1["local_cell_steps"] == r0["local_cell_steps"]
