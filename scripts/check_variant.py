"""Bit-for-bit comparison of stepper variants selected by environment variables read when the network is set.
usage: python scripts/check_variant.py SCR_HALF=6 [SCR_HALF=8 ...]   (each compared with the default kernel)"""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scrutiny_b200 import engine, network  # noqa: E402


def run(net, env):
    for k, v in env.items():
        os.environ[k] = v
    np.random.seed(1337)
    random.seed(1337)
    if net == "trrust":
        gc, tf = network.generate_gene_correlation_matrix(0, "TRRUST")
    else:
        gc, tf = network.generate_gene_correlation_matrix(3000, "SimulatedPowerLaw", powerLawExponent=2.05 if net == "powerlaw205" else 2.5)
    n = gc.shape[0]
    e = engine.Engine(0, "f32")
    e.set_network(gc / 0.3)
    cells = 6001
    e.alloc_cells(cells)
    e.set_states_broadcast(2 * np.random.rand(n) - 1)
    ids = np.arange(cells)
    steps = np.random.randint(0, 41, size=cells).astype(np.int32)
    proj = (np.random.rand(n) > 0.05).astype(np.float64)
    const = 0.1 * np.random.randn(n) * (1 - proj)
    e.run_phase(ids, steps, proj, const, seed=5)
    e.run_phase(ids[::3], steps[::3], proj, const, step_base=steps[::3].astype(np.int64), seed=5)
    out = e.get_states(0, cells)
    for k in env:
        del os.environ[k]
    return out


def main():
    for net in ("powerlaw", "trrust", "powerlaw205"):
        ref = run(net, {})
        for spec in sys.argv[1:]:
            env = dict(kv.split("=") for kv in spec.split(","))
            got = run(net, env)
            same = np.array_equal(ref, got)
            print(net, spec, "bit-identical" if same else "DIFFERENT max|d| = %g" % np.abs(ref - got).max(), flush=True)
            if not same:
                sys.exit(1)


if __name__ == "__main__":
    main()
