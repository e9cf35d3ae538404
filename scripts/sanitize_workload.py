"""CI-sized workload that launches every kernel of the library once, small: meant for
    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_workload.py [sparse|dense|readout|regnet ...]
on a machine that allows it (the round-1 GPU pool refuses compute-sanitizer: the runs leave GPUs needing a reset), and usable
as a plain every-kernel smoke run otherwise."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scrutiny_b200 import engine  # noqa: E402


def sparse_net(n, rng, long_rows=2):
    W = np.zeros((n, n))
    regs = rng.choice(n, size=max(2, 2 * n // 3), replace=False)
    for r in range(n):
        k = 1 + int(rng.rand() < 0.5)
        W[r, rng.choice(regs, size=k, replace=False)] = rng.beta(2, 3, size=k) * rng.choice([-1, 1]) * 2.0
    for r in rng.choice(n, size=long_rows, replace=False):
        k = min(len(regs), 40)
        W[r, rng.choice(regs, size=k, replace=False)] = rng.beta(2, 3, size=k)
    return W


def sparse(prec):
    rng = np.random.RandomState(1)
    n, cells = 700, 330                      # 6 blocks, long rows -> chunks; > 148 groups at f64 -> 2 groups per CTA
    e = engine.Engine(0, prec)
    e.set_network(sparse_net(n, rng))
    e.alloc_cells(cells)
    e.set_states(rng.uniform(-1, 1, size=(n, cells)))
    steps = rng.randint(0, 7, size=cells).astype(np.int32)
    proj = (rng.rand(n) > 0.1).astype(np.float64)
    e.run_phase(np.arange(cells), steps, proj, 0.1 * (1 - proj), seed=3)                       # free-running, ragged
    e.run_phase(np.arange(8), np.full(8, 3, np.int32), proj, 0.1 * (1 - proj), seed=3,
                noise=0.05 * rng.randn(8, 3, n))                                                # supplied noise
    e.probe(np.arange(0, cells, 30), proj, 0.1 * (1 - proj), max_iter=30, avg_num=5, seed=4)    # probe
    print(prec, "sparse ok", float(np.abs(e.get_states(0, 8)).max()), flush=True)
    return e


def readout(e):
    n = e.n
    sums = e.gene_sums()
    out = e.counts(np.zeros(n) - sums / e.cells, np.ones(e.cells), seed=5)
    packed, ovf = e.counts_compact(np.zeros(n) - sums / e.cells + 3.0, np.ones(e.cells), seed=5, dtype="uint8")   # some entries overflow a byte
    ext = e.counts_ext(np.zeros(n) - sums / e.cells, np.ones(e.cells), nb_dispersion=0.5, dropout=0.2, seed=5)
    print("readout ok", int(np.asarray(out).sum()), int(packed.sum()), len(ovf), int(ext.sum()), flush=True)


def dense():
    rng = np.random.RandomState(2)
    n, cells = 300, 200
    W = rng.beta(2, 3, size=(n, n)) * rng.choice([-1, 1], size=(n, 1)) * (rng.rand(n, n) < 0.5) / 5.0
    e = engine.Engine(0, "f32")
    e.set_network(W)
    assert e.layout()["dense_stepper"] == 1
    e.alloc_cells(cells)
    e.set_states(rng.uniform(-1, 1, size=(n, cells)))
    e.run_phase(np.arange(cells), rng.randint(0, 5, size=cells).astype(np.int32), np.ones(n), np.zeros(n), seed=3)
    e.probe(np.arange(0, cells, 20), np.ones(n), np.zeros(n), max_iter=25, avg_num=5, seed=4)
    print("dense ok", float(np.abs(e.get_states(0, 8)).max()), flush=True)


def regnet():
    rng = np.random.RandomState(3)
    e = engine.Engine(0, "f32")
    S = rng.poisson(2.0, size=(60, 500)).astype(np.float64)
    e.rank_transform(S)
    S2 = rng.randn(40, 300)
    e.rank_transform(S2)
    pairs = e.cell_pairs(S2, rng.rand(300), rng.randint(0, 3, size=300), 0.02)
    print("regnet ok", float(S.sum()), len(pairs), flush=True)


def main():
    what = sys.argv[1:] or ["sparse", "dense", "readout", "regnet"]
    e = None
    if "sparse" in what:
        e = sparse("f32")
        sparse("f64")
    if "readout" in what:
        readout(e or sparse("f32"))
    if "dense" in what:
        dense()
    if "regnet" in what:
        regnet()


if __name__ == "__main__":
    main()
