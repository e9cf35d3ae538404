"""oracle/fast_oracle.py (batched, noise looked up by (cell, step)) must agree with the pinned
port when fed the noise the port drew.  Tolerance 1e-12: only summation order differs."""
import numpy as np

from oracle import fast_oracle as fo
from oracle import matriseq_port as port
from oracle_utils import noise_table, tree_schedule
from test_oracle_golden import replay_pipeline, replay_probe


def test_tree_run_matches_port():
    rec = port.RecordingDraws()
    log = []
    o = replay_pipeline(draws=rec, log=log)
    n, cells = o["n"], o["cells"]
    noise = noise_table(rec.normals, "step", cells, n)
    W = fo.as_operator(o["gc"])
    X = np.ascontiguousarray(o["init"].T)
    tree = o["tree"]
    total = 0
    for t, ids, steps, base in tree_schedule(tree, log, o["iters"], cells):
        Xp = X[ids]
        fo.run_phase(W, Xp, steps, tree[t].proj, tree[t].const, 0.01, noise, step_base=base, rows=ids)
        X[ids] = Xp
        total += steps.sum()
    assert total == o["iters"].sum()
    assert np.max(np.abs(X.T - o["final_states"])) < 1e-12


def test_probe_matches_port(golden):
    g = golden("probe")
    n, gc, states = replay_probe(g)
    rec = port.RecordingDraws()
    sampled, trace = [], []
    T = port.optimal_iterations(n, gc, 0.01, list(g["members"]), g["proj"], g["const"], 0,
                                float(g["thresh"]), 0.05, states, int(g["maxIter"]), int(g["avgNum"]),
                                int(g["numToSample"]), rec, trace=trace, sampled=sampled)
    assert T == int(g["T"])
    noise = noise_table(rec.normals, "probe", states.shape[1], n)
    noise = np.pad(noise, ((0, 0), (0, int(g["maxIter"]) - noise.shape[1]), (0, 0)))
    iters, nd, _ = fo.probe(fo.as_operator(gc), states[:, sampled].T.copy(), g["proj"], g["const"], 0.01,
                            noise[sampled], float(g["thresh"]), int(g["maxIter"]), int(g["avgNum"]))
    per_cell = {}
    for cell, it, d, dist in trace:
        per_cell[cell] = it + 1
    assert [per_cell[c] for c in sampled] == list(iters)
    assert fo.iterations_from_probe(iters, int(g["avgNum"])) == T


def test_gene_shift_and_lambda_match_port(golden):
    g = golden("pipeline")
    S = g["final_states"].copy()
    cap = {}
    from conftest import seed_all
    seed_all(5)
    counts, size = port.transform_data(S.shape[0], S.shape[1], S.copy(), capture=cap)
    shift = fo.gene_shift(g["final_states"], cap["target_sorted"])
    lam = fo.poisson_lambda(g["final_states"], shift, 1.0, size)
    assert np.allclose(lam, cap["lambda"], rtol=1e-12, atol=0)
