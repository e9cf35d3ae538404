"""Developer micro-benchmark of the stepping kernel alone (not the contract bench; see bench.py).
usage: python scripts/quick_bench.py [--net powerlaw|trrust|dense500] [--cells N] [--steps T] [--prec f32]"""
import argparse
import os
import random
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scrutiny_b200 import engine, network  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="powerlaw")
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--prec", default="f32")
    ap.add_argument("--alpha", type=float, default=0.02)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ragged", action="store_true")
    a = ap.parse_args()
    np.random.seed(1337)
    random.seed(1337)
    if a.net == "powerlaw":
        gc, tf = network.generate_gene_correlation_matrix(3000, "SimulatedPowerLaw", powerLawExponent=2.5)
    elif a.net == "powerlaw205":
        gc, tf = network.generate_gene_correlation_matrix(3000, "SimulatedPowerLaw")
    elif a.net == "trrust":
        gc, tf = network.generate_gene_correlation_matrix(0, "TRRUST")
    elif a.net == "dense500":
        gc, tf = network.generate_gene_correlation_matrix(500, "Random", networkSparsity=0.01)
    elif a.net == "dense3000":
        gc, tf = network.generate_gene_correlation_matrix(3000, "Random", networkSparsity=0.5)
    else:
        raise SystemExit("unknown net")
    n = gc.shape[0]
    e = engine.Engine(0, a.prec)
    e.set_network(gc / a.alpha)
    print("layout", e.layout(), "nnz", np.count_nonzero(gc), flush=True)
    e.alloc_cells(a.cells)
    x0 = 2 * np.random.rand(n) - 1
    ids = np.arange(a.cells)
    steps = np.full(a.cells, a.steps, dtype=np.int32)
    if a.ragged:
        steps = np.random.randint(0, a.steps + 1, size=a.cells).astype(np.int32)
    for r in range(a.reps):
        e.set_states_broadcast(x0)
        e.step_kernel_stats(reset=True)
        t0 = time.time()
        e.run_phase(ids, steps, np.ones(n), np.zeros(n), seed=r)
        wall = time.time() - t0
        ms, _ = e.step_kernel_stats()
        cs = float(steps.sum())
        print("rep %d: kernel %.2f ms  wall %.2f ms  %.1f M cell-steps/s  algorithmic %.0f GB/s (%.1f%% of 6556)  "
              "3-product-issued %.1f TFLOP/s" % (r, ms, wall * 1e3, cs / ms / 1e3, cs * 8 * n / ms / 1e6,
                                            cs * 8 * n / ms / 1e6 / 65.562, cs * 6.0 * n * n / ms / 1e9), flush=True)
    s = e.get_states(0, 4)
    print("finite:", np.isfinite(s).all(), "range", s.min(), s.max())


if __name__ == "__main__":
    main()
