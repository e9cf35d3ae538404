#!/usr/bin/env python
"""Contract benchmark of the MatriSeq simulation hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (BASELINE.json config 5): SimulatedPowerLaw network, n = 3000 genes, powerLawExponent 2.5
(nnz ~4.6k), alpha from the reference's own search rule run on the GPU, the 4-leaf / 7-type lineage
tree of config 2 (parents [-1,0,0,1,1,2,2]), 1,000,000 cells PER GPU (weak scaling), default
dt / noise / thresholds of MatriSeq.py:546-554, then the transformData count read-out
(MatriSeq.py:816-891).  One "step" = one full pass: convergence probe per tree node, ragged
stepping of members and descendants, per-gene mean all-reduce, Poisson counts and -- for N > 1 --
the NCCL gather of the count matrix to rank 0.  metric = cell-steps/s = sum of
totalCellIterations over all cells / time.

`value`: inputs resident in HBM.  `e2e`: the same pass through the C ABI with HOST buffers
(network, initial state and schedules copied in, count matrix copied out to pinned memory).
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARENTS = [-1, 0, 0, 1, 1, 2, 2]
SHARES = [0.04, 0.08, 0.08, 0.20, 0.20, 0.20, 0.20]
CONST_PROPS = [0.05] * 7
GENES = 3000
EXPONENT = 2.5
HBM_FALLBACK_GBS = 6650.0
ALPHA_SEARCHED = 0.4   # what findOptimalAlpha returns for this seeded network (our arm re-derives it on the GPU)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1000000, help="cells per GPU")
    ap.add_argument("--genes", type=int, default=GENES)
    ap.add_argument("--network", default="powerlaw", choices=["powerlaw", "trrust"],
                    help="powerlaw = the contract workload; trrust = the same pass on the TRRUST network (n = 2895)")
    ap.add_argument("--alpha", type=float, default=None, help="skip the search")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workload construction (host, deterministic, identical on every rank)
# ---------------------------------------------------------------------------------------------
def build_network(n, kind="powerlaw"):
    """kind "powerlaw" (the default workload) or "trrust" (the other network BASELINE.json's config 5 names;
    n is then TRRUST's own 2895, MS:99-100)."""
    from scrutiny_b200 import network
    np.random.seed(1337)
    random.seed(1337)
    if kind == "trrust":
        return network.generate_gene_correlation_matrix(0, "TRRUST")
    gc, tf = network.generate_gene_correlation_matrix(n, "SimulatedPowerLaw", powerLawExponent=EXPONENT)
    return gc, tf


def build_tree(n, cells, tf):
    """Same structure formCellTypeTree builds (MatriSeq.py:175-297): every cell gets its final type
    up front (a random permutation split by SHARES), proj/const inherited down the tree."""
    from scrutiny_b200.MatriSeq import getNextProjConst
    from scrutiny_b200.lineage import Tree
    counts = [int(round(s * cells)) for s in SHARES]
    counts[-1] = cells - sum(counts[:-1])
    perm = np.random.permutation(cells)
    bounds = np.cumsum([0] + counts)
    members = [np.sort(perm[bounds[t]:bounds[t + 1]]) for t in range(len(PARENTS))]
    tree = Tree()
    tree.add_head(-1, np.zeros(0, dtype=np.int64), np.zeros(n), np.ones(n), 0)

    def attach(parent, proj_up, const_up):
        below = []
        for t in [i for i, p in enumerate(PARENTS) if p == parent]:
            proj, const = getNextProjConst(n, proj_up, const_up, CONST_PROPS[t], tf)
            tree.add_node(t, members[t], const, proj, 0, parent)
            below += attach(t, proj, const) + [members[t]]
        tree[parent].setChildCellTypeMembers(np.concatenate(below) if below else np.zeros(0, dtype=np.int64))
        return below

    attach(-1, np.ones(n), np.zeros(n))
    return tree


def search_alpha(gc, tree, x0, n):
    """findOptimalAlpha (MatriSeq.py:368-500) on the GPU over 50 sampled copies of the shared
    initial state; setup, not timed."""
    import scrutiny_b200.MatriSeq as ms
    st_np, st_py = np.random.get_state(), random.getstate()
    ms.init(outputDir=None, seed=1337, verbose=False, device=int(os.environ.get("LOCAL_RANK", 0)))
    states = np.repeat(x0[:, None], 64, axis=1)
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=50)
    np.random.set_state(st_np)
    random.setstate(st_py)
    return alpha


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(gc, proj, const, x0):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _CPU["limit"] = threadpool_limits(1)
    except Exception:
        pass
    _CPU.update(gc=gc, proj=proj, const=const, x0=x0)


def _cpu_work(args):
    """One worker: findCellNextState (MatriSeq.py:683-736, dense float64 gemv per step like
    developCell at MatriSeq.py:530) for `cells` cells x `iters` iterations."""
    seed, cells, iters = args
    from oracle import matriseq_port as port
    np.random.seed(seed)
    n = _CPU["gc"].shape[0]
    t0 = time.perf_counter()
    for _ in range(cells):
        port.cell_next_state(_CPU["gc"], n, 0.01, _CPU["proj"], _CPU["const"], _CPU["x0"], iters, 0.05, 0)
    return time.perf_counter() - t0


def cpu_baseline(gc, proj, const, x0, seconds):
    """cell-steps/s of the oracle port with one process per host core (cells are independent, so this
    is how the reference's per-cell loop parallelises), on a sample sized for ~`seconds`."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = min(cores, 64)
    ctx = mp.get_context("fork")
    _cpu_init(gc, proj, const, x0)
    t = _cpu_work((0, 1, 20))                       # calibrate: one core, 20 steps
    per_step = t / 20
    iters = 50
    cells_each = max(1, int(seconds / (per_step * iters * 1.6)))   # 1.6: expected slowdown when all cores share DRAM
    with ctx.Pool(workers, initializer=_cpu_init, initargs=(gc, proj, const, x0)) as pool:
        t0 = time.perf_counter()
        pool.map(_cpu_work, [(100 + w, cells_each, iters) for w in range(workers)])
        wall = time.perf_counter() - t0
    steps = workers * cells_each * iters
    return {"value": steps / wall, "unit": "cell-steps/s", "cores": workers, "kind": "port",
            "sample": "%d cells x %d iterations of findCellNextState (dense float64 gemv, n=%d) on %d processes, "
                      "%.1f s wall; single-process rate %.0f cell-steps/s" % (workers * cells_each, iters, gc.shape[0],
                                                                          workers, wall, 1.0 / per_step)}


def run_reference(args, rank):
    if rank != 0:
        return
    gc, tf = build_network(args.genes, args.network)
    n = gc.shape[0]
    tree = build_tree(n, 1000, tf)
    alpha = args.alpha if args.alpha else (0.1 if args.network == "trrust" else ALPHA_SEARCHED)
    gc = gc / alpha
    x0 = 2.0 * np.random.rand(n) - 1.0
    vals, infos = [], None
    for i in range(args.warmup + args.steps):
        info = cpu_baseline(gc, tree[0].proj, tree[0].const, x0, max(2.0, args.cpu_seconds / 2))
        if i >= args.warmup:
            vals.append(info["value"])
            infos = info
    value = float(np.mean(vals))
    infos["value"] = value
    line = {"impl": "reference", "metric": "cell_steps_per_sec", "value": value, "unit": "cell-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n, int(np.count_nonzero(gc)), alpha, note="CPU sample of the same network"),
            "cpu_baseline": infos,
            "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, n, nnz, alpha, note=None):
    net = "TRRUST n=%d" % n if args.network == "trrust" else "SimulatedPowerLaw n=%d powerLawExponent=%.1f" % (n, EXPONENT)
    cfg = {"workload": "c5: MatriSeq %s, %d cells per GPU, 7-type/4-leaf lineage tree, count sampling" % (net, args.cells),
           "genes": n, "nnz": nnz, "alpha": alpha, "cells_per_gpu": args.cells, "tree_parents": PARENTS,
           "l2": "inputs exceed L2 (state %.1f GB per GPU)" % (args.cells * (-(-n // 128) * 128) * 4 / 1e9)}
    if note:
        cfg["note"] = note
    return cfg


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from scrutiny_b200 import engine as eng_mod
    from scrutiny_b200.simulate import Comm, LineageSimulation

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm(device=dev)

    per_gpu = args.cells
    cells = per_gpu * world
    gc, tf = build_network(args.genes, args.network)
    n = gc.shape[0]
    nnz = int(np.count_nonzero(gc))
    tree = build_tree(n, cells, tf)
    x0 = 2.0 * np.random.rand(n) - 1.0
    alpha = args.alpha if args.alpha else search_alpha(gc, tree, x0, n)
    gcs = gc / alpha
    from scrutiny_b200.network import dense_to_csr
    csr = dense_to_csr(gcs)

    lo, hi = rank * per_gpu, (rank + 1) * per_gpu
    eng = eng_mod.Engine(local, "f32")
    eng.set_network(csr=csr, n=n)
    eng.alloc_cells(per_gpu, lo)
    sim = LineageSimulation(eng, cells, lo, hi, comm=comm, seed=1337)
    sim.prepare(tree)
    counts_dev = torch.empty((per_gpu, n), dtype=torch.int32, device=dev)
    gathered = [torch.empty((per_gpu, n), dtype=torch.int32, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    layout = eng.layout()

    # the count matrix is sampled in row blocks; each finished block is handed to NCCL at once (asynchronous gather to
    # rank 0), so the transfer of block i runs under the sampling of block i+1
    COUNT_BLOCKS = 8 if world > 1 else 1
    pending = []

    def gather_block(r0, r1):
        if world > 1:
            pending.append(dist.gather(counts_dev[r0:r1], [g[r0:r1] for g in gathered] if rank == 0 else None,
                                       dst=0, async_op=True))

    def gather_counts():
        for w in pending:
            w.wait()
        pending.clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for w in range(args.warmup):
        eng.set_states_broadcast(x0)
        sim.run_all(tree, counts_dev_ptr=counts_dev.data_ptr(), seed=1337 + w, counts_blocks=COUNT_BLOCKS,
                    on_counts_block=gather_block)
        gather_counts()
    barrier()

    # ---- timed: resident inputs ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    launches0 = eng.launch_count()
    eng.step_kernel_stats(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_steps = 0
    local_steps = 0
    elapsed_ms = 0.0
    for k in range(args.steps):
        eng.set_states_broadcast(x0)                            # reset, outside the timed region
        barrier()
        ev0.record()
        t0 = time.perf_counter()
        res = sim.run_all(tree, counts_dev_ptr=counts_dev.data_ptr(), seed=2000 + k, counts_blocks=COUNT_BLOCKS,
                          on_counts_block=gather_block)
        gather_counts()
        barrier()
        ev1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        elapsed_ms += max(ev0.elapsed_time(ev1), wall_ms)
        total_steps += res["global_cell_steps"]
        local_steps += res["local_cell_steps"]
    clocks = sampler.stop()
    launches = eng.launch_count() - launches0 - args.steps      # minus the untimed reset kernels
    kernel_ms, kernel_launches = eng.step_kernel_stats()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = total_steps / (elapsed_ms / 1e3)

    # ---- timed: end to end with host buffers ----
    e2e = None
    if not args.no_e2e:
        try:
            host_counts = torch.empty((per_gpu, n), dtype=torch.int32, pin_memory=True).numpy()
        except RuntimeError:
            host_counts = np.empty((per_gpu, n), dtype=np.int32)
        e_ms, e_steps, h2d = 0.0, 0, 0
        for k in range(args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            eng.set_network(csr=csr, n=n)
            eng.set_states_broadcast(x0)
            res = sim.run_all(tree, counts_out=host_counts, seed=3000 + k)
            barrier()
            e_ms += (time.perf_counter() - t0) * 1e3
            e_steps += res["global_cell_steps"]
            h2d = int(csr[0].nbytes + csr[1].nbytes + csr[2].nbytes + x0.nbytes + res["h2d_bytes"])
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": e_steps / (float(t.item()) / 1e3), "unit": "cell-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(host_counts.nbytes + per_gpu * 8), "steps": args.e2e_steps,
               "api": "scr_set_network_csr + scr_set_states_broadcast + scr_probe/scr_run_phase per node + "
                      "scr_gene_sums + scr_counts(host int32 out)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
        n_pad = layout["n_pad"]
        achieved = (local_steps * 8.0 * n) / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else None
        roofline = {"bound": "hbm", "kernel": "scr::step_kernel<float> (persistent stepper)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_cell_step": 8 * n, "cell_steps_per_launch": local_steps / max(kernel_launches, 1),
                    "kernel_ms_per_launch": kernel_ms / max(kernel_launches, 1), "kernel_launches": int(kernel_launches),
                    "kernel_share_of_step": kernel_ms / elapsed_ms if elapsed_ms else None,
                    "kernel_cell_steps_per_s": local_steps / (kernel_ms / 1e3) if kernel_ms > 0 else None,
                    "traffic": _ncu_traffic(), "traffic_source": "profiles/step_kernel_traffic.json (ncu capture of one "
                    "80e6-cell-step launch: 9.87 GB DRAM vs 1.92 TB algorithmic -- the state is smem-resident)",
                    "layout": layout}
        cpu = None
        if not args.no_cpu:
            cpu = cpu_baseline(gcs, tree[0].proj, tree[0].const, x0, args.cpu_seconds)
        cfg = workload_config(args, n, nnz, alpha)
        cfg["node_iterations"] = sim.node_iterations
        line = {"metric": "cell_steps_per_sec", "value": value, "unit": "cell-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _ncu_traffic():
    """dram bytes per launch of the stepper from the committed ncu capture, if summarised."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


if __name__ == "__main__":
    main()
