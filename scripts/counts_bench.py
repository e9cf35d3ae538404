"""Developer micro-benchmark of the count read-out kernel alone (scr_counts with a device-resident int32 output).
usage: python scripts/counts_bench.py [--cells N] [--reps R]     (SCR_COUNTS_EXACT=1: float64 sampler only)
Prints wall time of the synchronous call and entries/s; states come from a short free-running phase so that
lambda varies per cell like in the bench pass."""
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
    ap.add_argument("--cells", type=int, default=200000)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    import torch
    np.random.seed(1337)
    random.seed(1337)
    n = 3000
    gc, tf = network.generate_gene_correlation_matrix(n, "SimulatedPowerLaw", powerLawExponent=2.5)
    e = engine.Engine(0, "f32")
    e.set_network(gc / 0.4)
    e.alloc_cells(a.cells)
    e.set_states_broadcast(2 * np.random.rand(n) - 1)
    e.run_phase(np.arange(a.cells), np.full(a.cells, 30, dtype=np.int32), np.ones(n), np.zeros(n), seed=3)
    rng = np.random.RandomState(5)
    shift = np.log(rng.gamma(0.35, 1 / 0.19, size=n))       # transformData's log-Gamma gene levels (MS:855-857)
    size = rng.normal(1, 0.14, size=a.cells) ** 3
    buf = torch.empty((a.cells, n), dtype=torch.int32, device="cuda:0")
    mode = "float64 sampler only" if os.environ.get("SCR_COUNTS_EXACT") == "1" else "float32 screening"
    for r in range(a.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e.counts(shift, size, seed=9, out_device_ptr=buf.data_ptr())
        dt = time.perf_counter() - t0
        print("%s rep %d: %.2f ms  %.1f G entries/s  (%.0f GB/s algorithmic at 8 B per entry)"
              % (mode, r, dt * 1e3, a.cells * n / dt / 1e9, a.cells * n * 8 / dt / 1e9), flush=True)
    print("checksum", int(buf.sum(dtype=torch.int64)), "max", int(buf.max()))


if __name__ == "__main__":
    main()
