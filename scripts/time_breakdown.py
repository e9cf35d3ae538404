"""Wall-time breakdown of one bench pass by API call (developer tool)."""
import os, sys, time, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from scrutiny_b200 import engine as eng_mod
from scrutiny_b200.network import dense_to_csr
from scrutiny_b200.simulate import LineageSimulation

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
host = len(sys.argv) > 2 and sys.argv[2] == "host"      # e2e variant: counts to a pinned host matrix
n = 3000
gc, tf, x0, tree = bench.build_workload(n, "powerlaw", cells)
gcs = gc / 0.4
eng = eng_mod.Engine(0, "f32")
eng.set_network(csr=dense_to_csr(gcs), n=n)
eng.alloc_cells(cells)
sim = LineageSimulation(eng, cells, seed=1)
sim.prepare(tree)
import torch
dev = torch.empty((cells, n), dtype=torch.uint8, device="cuda:0") if not host else None
hostbuf = torch.empty((cells, n), dtype=torch.uint8, pin_memory=True).numpy() if host else None
acc = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
for name in ("probe", "probe_multi", "run_phase", "gene_sums", "counts", "counts_compact"):
    wrap(eng, name)
for rep in range(4):
    acc.clear()
    eng.set_states_broadcast(x0)
    eng.step_kernel_stats(reset=True)
    t0 = time.perf_counter()
    if host:
        t1 = time.perf_counter(); eng.set_network(csr=dense_to_csr(gcs), n=n); acc["set_network"] = time.perf_counter() - t1
        res = sim.run_all(tree, counts_out=hostbuf, seed=5 + rep, counts_dtype="uint8")
    else:
        res = sim.run_all(tree, counts_dev_ptr=dev.data_ptr(), seed=5 + rep, counts_dtype="uint8")
    wall = time.perf_counter() - t0
    kms, _ = eng.step_kernel_stats()
    pms, _ = eng.probe_kernel_stats(reset=True)
    print("rep %d wall %.1f ms | stepper kernel %.1f ms | probe kernel %.1f ms | " % (rep, wall * 1e3, kms, pms) +
          " ".join("%s %.1f" % (k, v * 1e3) for k, v in acc.items()) + " | other host %.1f ms" % ((wall - sum(acc.values())) * 1e3))
