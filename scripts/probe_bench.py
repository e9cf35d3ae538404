"""Developer micro-benchmark of the convergence probe (scr_probe) alone: one node's 50 sampled cells on the contract network.
usage: python scripts/probe_bench.py [--net powerlaw|trrust|dense500|dense3000] [--k 50] [--reps 3]"""
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
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--alpha", type=float, default=0.4)
    a = ap.parse_args()
    np.random.seed(1337)
    random.seed(1337)
    if a.net == "powerlaw":
        gc, tf = network.generate_gene_correlation_matrix(3000, "SimulatedPowerLaw", powerLawExponent=2.5)
    elif a.net == "trrust":
        gc, tf = network.generate_gene_correlation_matrix(0, "TRRUST")
    elif a.net == "dense500":
        gc, tf = network.generate_gene_correlation_matrix(500, "Random", networkSparsity=0.01)
    else:
        gc, tf = network.generate_gene_correlation_matrix(3000, "Random", networkSparsity=0.5)
    n = gc.shape[0]
    e = engine.Engine(0, "f32")
    e.set_network(gc / a.alpha)
    e.alloc_cells(1000)
    e.set_states_broadcast(2 * np.random.rand(n) - 1)
    ids = np.arange(a.k)
    for r in range(a.reps):
        e.probe_kernel_stats(reset=True)
        t0 = time.perf_counter()
        it = e.probe(ids, np.ones(n), np.zeros(n), seed=5 + r)
        wall = time.perf_counter() - t0
        ms, launches = e.probe_kernel_stats()
        steps = int(it.max())
        print("rep %d: probe wall %.3f ms, kernel %.3f ms (%d launches), longest cell %d iterations -> %.2f us per sequential step"
              % (r, wall * 1e3, ms, launches, steps, (ms if ms > 0 else wall * 1e3) * 1e3 / max(steps, 1)), flush=True)


if __name__ == "__main__":
    main()
