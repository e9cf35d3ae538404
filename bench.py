#!/usr/bin/env python
"""Contract benchmark of the MatriSeq simulation hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (BASELINE.json config 5): SimulatedPowerLaw network, n = 3000 genes, powerLawExponent 2.5 (nnz ~4.6k), divided by
alpha = 0.4 (what the reference's own search rule, MatriSeq.py:368-500, returns for this seeded network: re-derived on the
GPU in `extra.alpha_search`), the 4-leaf / 7-type lineage tree of config 2 (parents [-1,0,0,1,1,2,2]), 1,000,000 cells PER
GPU (weak scaling), default dt / noise / thresholds of MatriSeq.py:546-554, then the transformData count read-out
(MatriSeq.py:816-891).  One "step" = one full pass: convergence probe per tree node, ragged stepping of members and
descendants, per-gene mean all-reduce, Poisson counts into this rank's shard of the count matrix (unsigned bytes + overflow
list; `--gather` additionally collects the shards on rank 0 with NCCL).  metric = cell-steps/s = sum of totalCellIterations
over all cells / time.

`value`: inputs resident in HBM.  `e2e`: the same pass through the C ABI with HOST buffers (network, initial state and
schedules copied in, this rank's count shard copied out to pinned memory).  `extra`: the other paths north_star names,
measured after the headline in the same process -- TRRUST and the reference's default exponent 2.05 on the sparse stepper,
the float64 parity context, config c4 on the tcgen05 split-precision stepper (with a cuBLAS matmul of its operand type timed in-run as its denominator),
the count kernel alone, the alpha search, and (N > 1) the strong-scaling point: 10^6 cells in TOTAL split over the ranks.
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
# network divisors: what findOptimalAlpha returned for the seeded networks in round 1 (profiles/r1_final_configs_c1_c4.jsonl;
# powerlaw is re-derived in extra.alpha_search); powerlaw205 simply reuses the contract value (a rate measurement)
ALPHA = {"powerlaw": 0.4, "trrust": 0.1, "powerlaw205": 0.4, "dense_c4": 2.21}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=1000000, help="cells per GPU")
    ap.add_argument("--genes", type=int, default=GENES)
    ap.add_argument("--network", default="powerlaw", choices=["powerlaw", "trrust", "powerlaw205"],
                    help="powerlaw = the contract workload; trrust / powerlaw205 = the same pass on those networks")
    ap.add_argument("--alpha", type=float, default=None, help="override the network divisor")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--extra", default="trrust,powerlaw205,f64,dense_c4,counts,alpha_search,module_api,strong",
                    help="comma-separated extra blocks to measure after the headline")
    ap.add_argument("--gather", action="store_true", help="also gather the count shards on rank 0 with NCCL (N > 1)")
    ap.add_argument("--counts-dtype", default="uint8", choices=["uint8", "uint16", "int32"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workload construction (host, deterministic, identical on every rank and for every N)
# ---------------------------------------------------------------------------------------------
def build_network(n, kind="powerlaw"):
    """kind "powerlaw" (the contract workload), "powerlaw205" (the reference's default exponent), "trrust" (n is then
    TRRUST's own 2895, MS:99-100) or "dense_c4" (Random, networkSparsity 0.5)."""
    from scrutiny_b200 import network
    np.random.seed(1337)
    random.seed(1337)
    if kind == "trrust":
        return network.generate_gene_correlation_matrix(0, "TRRUST")
    if kind == "dense_c4":
        return network.generate_gene_correlation_matrix(n, "Random", networkSparsity=0.5)
    return network.generate_gene_correlation_matrix(n, "SimulatedPowerLaw", powerLawExponent=EXPONENT if kind == "powerlaw" else 2.05)


def build_tree(n, cells, tf, parents=PARENTS, shares=SHARES, const_props=CONST_PROPS):
    """formCellTypeTree (MatriSeq.py:175-297) through the module's scalable assignment (one permutation of the cells split
    by SHARES instead of a Python set difference per type): every cell gets its final type up front, proj/const are
    inherited down the tree."""
    import scrutiny_b200.MatriSeq as ms
    counts = [int(round(s * cells)) for s in shares]
    counts[-1] = cells - sum(counts[:-1])
    tree, _, _ = ms.formCellTypeTree(n, cells, parents, counts, const_props, tf, [0] * len(parents), scalable=True)
    return tree


def build_workload(n, kind, cells):
    """(gc, tf, x0, tree): x0 is drawn BEFORE the cell permutation, so network, initial state and node vectors do not
    depend on the number of cells / ranks."""
    gc, tf = build_network(n, kind)
    x0 = 2.0 * np.random.rand(gc.shape[0]) - 1.0
    single = kind == "dense_c4"
    tree = build_tree(gc.shape[0], cells, tf, *(([-1], [1.0], [0.0]) if single else (PARENTS, SHARES, CONST_PROPS)))
    return gc, tf, x0, tree


def search_alpha(gc, tree, x0, n, device):
    """findOptimalAlpha (MatriSeq.py:368-500) on the GPU over 50 sampled copies of the shared initial state."""
    import scrutiny_b200.MatriSeq as ms
    st_np, st_py = np.random.get_state(), random.getstate()
    ms.init(outputDir=None, seed=1337, verbose=False, device=device)
    states = np.repeat(x0[:, None], 64, axis=1)
    t0 = time.perf_counter()
    alpha = ms.findOptimalAlpha(n, gc, tree, states, numToSample=50)
    dt = time.perf_counter() - t0
    ms.init(outputDir=None, seed=1337, verbose=False, device=device)      # drops the cached context
    np.random.set_state(st_np)
    random.setstate(st_py)
    return alpha, dt


def workload_config(args, n, nnz, alpha):
    """The `config` object: identical in both arms (the driver compares them)."""
    net = {"trrust": "TRRUST n=%d" % n, "powerlaw": "SimulatedPowerLaw n=%d powerLawExponent=%.1f" % (n, EXPONENT),
           "powerlaw205": "SimulatedPowerLaw n=%d powerLawExponent=2.05" % n}[args.network]
    return {"workload": "c5: MatriSeq %s, %d cells per GPU, 7-type/4-leaf lineage tree, count sampling" % (net, args.cells),
            "genes": n, "nnz": nnz, "alpha": round(float(alpha), 4), "cells_per_gpu": args.cells, "tree_parents": PARENTS,
            "l2": "inputs exceed L2 (state %.1f GB per GPU)" % (args.cells * (-(-n // 128) * 128) * 4 / 1e9)}


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
        return self

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
    return {"value": steps / wall, "unit": "cell-steps/s", "cores": workers, "kind": "port", "wall_s": wall,
            "sample": "%d cells x %d iterations of findCellNextState (dense float64 gemv, n=%d) on %d processes, "
                      "%.1f s wall; single-process rate %.0f cell-steps/s" % (workers * cells_each, iters, gc.shape[0],
                                                                          workers, wall, 1.0 / per_step)}


def run_reference(args, rank):
    if rank != 0:
        return
    gc, tf, x0, tree = build_workload(args.genes, args.network, 1000)
    n = gc.shape[0]
    alpha = args.alpha if args.alpha else ALPHA[args.network]
    nnz = int(np.count_nonzero(gc))
    gc = gc / alpha
    vals, walls, infos = [], [], None
    for i in range(args.warmup + args.steps):
        info = cpu_baseline(gc, tree[0].proj, tree[0].const, x0, max(2.0, args.cpu_seconds / 2))
        if i >= args.warmup:
            vals.append(info["value"])
            walls.append(info["wall_s"])
            infos = info
    value = float(np.mean(vals))
    infos["value"] = value
    line = {"impl": "reference", "metric": "cell_steps_per_sec", "value": value, "unit": "cell-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(walls)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n, nnz, alpha),
            "note": "each step is a bounded CPU sample of the same network (cpu_baseline.sample); rank 0 only",
            "cpu_baseline": infos,
            "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Rig:
    """One rank's engine + scheduler for a workload, with the timing loop shared by the headline and the extra blocks."""

    def __init__(self, env, kind, cells_per_gpu, prec="f32", alpha=None, n=GENES):
        from scrutiny_b200 import engine as eng_mod
        from scrutiny_b200.network import dense_to_csr
        from scrutiny_b200.simulate import LineageSimulation
        self.env = env
        world, rank = env["world"], env["rank"]
        self.kind, self.per_gpu, self.cells = kind, cells_per_gpu, cells_per_gpu * world
        self.gc, self.tf, self.x0, self.tree = build_workload(n, kind, self.cells)
        self.n = self.gc.shape[0]
        self.nnz = int(np.count_nonzero(self.gc))
        self.alpha = alpha if alpha else ALPHA[kind]
        self.csr = dense_to_csr(self.gc / self.alpha)
        lo = rank * cells_per_gpu
        self.eng = eng_mod.Engine(env["local"], prec)
        self.eng.set_network(csr=self.csr, n=self.n)
        self.eng.alloc_cells(cells_per_gpu, lo)
        self.sim = LineageSimulation(self.eng, self.cells, lo, lo + cells_per_gpu, comm=env["comm"], seed=1337)
        self.sim.prepare(self.tree)
        self.layout = self.eng.layout()
        self.prec = prec

    def close(self):
        self.eng.close()

    def timed(self, steps, warmup, seed0, clocks=True, **run_kw):
        """W untimed + K timed passes (states reset outside the timed region, barrier + synchronize on both sides, max over
        ranks).  Returns dict(value, ms_per_step, kernel stats ...)."""
        import torch
        env, eng, sim = self.env, self.eng, self.sim
        for w in range(warmup):
            eng.set_states_broadcast(self.x0)
            sim.run_all(self.tree, seed=seed0 + w, **run_kw)
            env["after_pass"]()
        env["barrier"]()
        sampler = ClockSampler(env["local"]).start() if clocks else None
        if clocks:
            time.sleep(0.3)
        launches0 = eng.launch_count()
        eng.step_kernel_stats(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total_steps = local_steps = 0
        elapsed_ms = 0.0
        res = None
        for k in range(steps):
            eng.set_states_broadcast(self.x0)                            # reset, outside the timed region
            env["barrier"]()
            ev0.record()
            t0 = time.perf_counter()
            res = sim.run_all(self.tree, seed=seed0 + 1000 + k, **run_kw)
            env["after_pass"]()
            env["barrier"]()
            ev1.record()
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
            elapsed_ms += max(ev0.elapsed_time(ev1), wall_ms)
            total_steps += res["global_cell_steps"]
            local_steps += res["local_cell_steps"]
        out = {"clocks": sampler.stop() if sampler else None}
        out["launches"] = eng.launch_count() - launches0 - steps        # minus the untimed reset kernels
        kernel_ms, kernel_launches = eng.step_kernel_stats()
        elapsed_ms = env["max_over_ranks"](elapsed_ms)
        out.update(value=total_steps / (elapsed_ms / 1e3), ms_per_step=elapsed_ms / steps, elapsed_ms=elapsed_ms,
                   total_steps=total_steps, local_steps=local_steps, kernel_ms=kernel_ms, kernel_launches=int(kernel_launches),
                   last=res, node_iterations=dict(sim.node_iterations))
        return out

    def sparse_roofline(self, t, peak, peak_source):
        achieved = (t["local_steps"] * 8.0 * self.n) / (t["kernel_ms"] / 1e3) / 1e9 if t["kernel_ms"] > 0 else None
        esz = 8 if self.prec == "f64" else 4
        if achieved and esz == 8:
            achieved *= 2
        return {"bound": "hbm", "kernel": "scr::step_kernel<%s> (persistent stepper)" % ("double" if esz == 8 else "float"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_source, "algorithmic_bytes_per_cell_step": 2 * esz * self.n,
                "cell_steps_per_launch": t["local_steps"] / max(t["kernel_launches"], 1),
                "kernel_ms_per_launch": t["kernel_ms"] / max(t["kernel_launches"], 1), "kernel_launches": t["kernel_launches"],
                "kernel_share_of_step": t["kernel_ms"] / t["elapsed_ms"] if t["elapsed_ms"] else None,
                "kernel_cell_steps_per_s": t["local_steps"] / (t["kernel_ms"] / 1e3) if t["kernel_ms"] > 0 else None}


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
    from scrutiny_b200.simulate import Comm

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm(device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    env = dict(world=world, rank=rank, local=local, comm=comm, barrier=barrier, max_over_ranks=max_over_ranks,
               after_pass=lambda: None)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"

    per_gpu = args.cells
    rig = Rig(env, args.network, per_gpu, alpha=args.alpha, n=args.genes)
    n = rig.n
    cdt = args.counts_dtype
    torch_dt = {"uint8": torch.uint8, "uint16": torch.uint16, "int32": torch.int32}[cdt]
    esz = {"uint8": 1, "uint16": 2, "int32": 4}[cdt]
    counts_dev = torch.empty((per_gpu, n), dtype=torch_dt, device=dev)       # this rank's shard of the count matrix
    gathered = None
    pending = []
    COUNT_BLOCKS = 1
    on_block = None
    if args.gather and world > 1:
        # optional: the shards are collected on rank 0 (north_star's "NCCL only to gather the final count matrix"); each
        # finished row block is handed to NCCL at once, so its transfer runs under the sampling of the next block
        COUNT_BLOCKS = 8
        gathered = [torch.empty((per_gpu, n), dtype=torch_dt, device=dev) for _ in range(world)] if rank == 0 else None

        def on_block(r0, r1):
            pending.append(dist.gather(counts_dev[r0:r1], [g[r0:r1] for g in gathered] if rank == 0 else None,
                                       dst=0, async_op=True))

        def after_pass():
            for w in pending:
                w.wait()
            pending.clear()
        env["after_pass"] = after_pass

    run_kw = dict(counts_dev_ptr=counts_dev.data_ptr(), counts_blocks=COUNT_BLOCKS, on_counts_block=on_block, counts_dtype=cdt)
    head = rig.timed(args.steps, args.warmup, seed0=1337, **run_kw)
    env["after_pass"] = lambda: None
    overflow_entries = 0 if head["last"]["overflow"] is None else int(len(head["last"]["overflow"]))

    # ---- timed: end to end with host buffers ----
    e2e = None
    if not args.no_e2e:
        np_dt = {"uint8": np.uint8, "uint16": np.uint16, "int32": np.int32}[cdt]
        try:
            host_counts = torch.empty((per_gpu, n), dtype=torch_dt, pin_memory=True).numpy()
        except RuntimeError:
            host_counts = np.empty((per_gpu, n), dtype=np_dt)
        e_ms, e_steps, h2d, d2h = 0.0, 0, 0, 0
        for k in range(args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            rig.eng.set_network(csr=rig.csr, n=n)
            rig.eng.set_states_broadcast(rig.x0)
            res = rig.sim.run_all(rig.tree, counts_out=host_counts, seed=3000 + k, counts_dtype=cdt)
            barrier()
            e_ms += (time.perf_counter() - t0) * 1e3
            e_steps += res["global_cell_steps"]
            h2d = int(sum(a.nbytes for a in rig.csr) + rig.x0.nbytes + res["h2d_bytes"])
            d2h = int(host_counts.nbytes + per_gpu * 8 + (0 if res["overflow"] is None else res["overflow"].shape[0] * 12))
        e_ms = max_over_ranks(e_ms)
        e2e = {"value": e_steps / (e_ms / 1e3), "unit": "cell-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "counts_dtype": cdt,
               "api": "scr_set_network_csr + scr_set_states_broadcast + scr_probe/scr_run_phase per node + "
                      "scr_gene_sums + scr_counts_compact(host %s out, per-rank shard)" % cdt}
        del host_counts

    # ---- the other north_star paths, same process ----
    extra = {}
    if not args.no_extra:
        want = [x for x in args.extra.split(",") if x]
        extra = run_extras(args, env, rig, counts_dev, want, peak, peak_source, dev)

    if rank == 0:
        roofline = rig.sparse_roofline(head, peak, peak_source)
        roofline.update(traffic=_ncu_traffic(), traffic_source="profiles/step_kernel_traffic.json (ncu capture of one "
                        "80e6-cell-step launch: DRAM bytes vs 1.92 TB algorithmic -- the state is smem-resident)",
                        layout=rig.layout)
        cpu = None
        if not args.no_cpu:
            cpu = cpu_baseline(rig.gc / rig.alpha, rig.tree[0].proj, rig.tree[0].const, rig.x0, args.cpu_seconds)
        line = {"metric": "cell_steps_per_sec", "value": head["value"], "unit": "cell-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": workload_config(args, n, rig.nnz, rig.alpha), "clocks": head["clocks"],
                "e2e": e2e, "gpu_launches": int(head["launches"]), "roofline": roofline, "cpu_baseline": cpu,
                "details": {"node_iterations": head["node_iterations"], "counts_dtype": cdt,
                            "count_matrix": "per-rank shards (cells x genes, %s + overflow list)%s" % (
                                cdt, ", gathered on rank 0" if args.gather and world > 1 else ""),
                            "overflow_entries_last_pass": overflow_entries,
                            "noise_stream": rig.eng.noise_stream_version()},
                "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, env, rig, counts_dev, want, peak, peak_source, dev):
    """Measured after the headline, in the same process; rank 0 keeps the numbers.  Every block: its own warm-up, CUDA-event
    / barrier timing like the headline, its own clocks sample."""
    import torch
    world, rank = env["world"], env["rank"]
    out = {}
    n = rig.n

    def block(name, fn):
        if name not in want:
            return
        try:
            t0 = time.perf_counter()
            res = fn()
            if res is not None:
                res["block_wall_s"] = time.perf_counter() - t0
                out[name] = res
        except Exception as exc:                                         # an extra block never takes the headline down
            out[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # -- the count kernel alone on the headline's final states -----------------------------------------------------
    def counts_block():
        eng = rig.eng
        sampler = ClockSampler(env["local"]).start()
        shift = rig.sim.gene_shift()
        size = rig.sim._size_factors(0.14)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.counts_compact(shift, size, seed=5 + rep, dtype="uint8", out_device_ptr=counts_dev.data_ptr())
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3
            best = ms if best is None or ms < best else best
        entries = float(rig.per_gpu) * n
        rate = entries / (best / 1e3)
        return {"metric": "entries_per_sec (counts_kernel, uint8 out, one GPU)", "value": rate, "ms_per_call": best,
                "entries": entries, "timing": "host clock around the synchronous scr_counts_compact call (one kernel launch + "
                "4 small copies), best of 4",
                "roofline": {"bound": "hbm", "algorithmic_bytes_per_entry": 5, "achieved": rate * 5 / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": rate * 5 / 1e9 / peak, "peak_source": peak_source},
                "clocks": sampler.stop()}
    block("counts", counts_block)

    # -- alpha search of the contract network (setup of the headline, re-derived here) -----------------------------
    def alpha_block():
        if rank != 0:
            return None
        small = build_tree(n, 1000, rig.tf)
        a, dt = search_alpha(rig.gc, small, rig.x0, n, env["local"])
        return {"alpha_searched": round(float(a), 4), "alpha_used": rig.alpha, "seconds": dt,
                "what": "findOptimalAlpha (MS:368-500) on the GPU, 50 sampled cells, batched ladder"}
    block("alpha_search", alpha_block)
    env["barrier"]()

    def sparse_block(kind, prec, cells, steps, label):
        def fn():
            r = Rig(env, kind, cells, prec=prec)
            try:
                dt = torch.uint8
                buf = torch.empty((cells, r.n), dtype=dt, device=dev)
                t = r.timed(steps, 1, seed0=7000, counts_dev_ptr=buf.data_ptr(), counts_dtype="uint8")
                res = {"metric": "cell_steps_per_sec", "workload": label, "value": t["value"], "ms_per_step": t["ms_per_step"],
                       "n_gpus": world, "cells_per_gpu": cells, "genes": r.n, "nnz": r.nnz, "alpha": r.alpha, "dtype": prec,
                       "steps": steps, "warmup": 1, "node_iterations": t["node_iterations"],
                       "roofline": r.sparse_roofline(t, peak, peak_source), "clocks": t["clocks"],
                       "layout": {k: r.layout[k] for k in ("groups_per_cta", "operator_prefix_bytes", "operator_bytes", "chunks",
                                                           "gather_wavefronts_model", "gather_wavefronts_floor")}}
                del buf
                return res
            finally:
                r.close()
        return fn

    block("trrust", sparse_block("trrust", "f32", args.cells, 2, "c5 pass on TRRUST n=2895 (alpha 0.1), %d cells per GPU" % args.cells))
    block("powerlaw205", sparse_block("powerlaw205", "f32", args.cells, 2,
                                      "c5 pass on SimulatedPowerLaw n=3000 powerLawExponent=2.05 (the reference's default), %d cells per GPU" % args.cells))
    f64_cells = min(args.cells, 250000)
    block("f64", sparse_block(args.network, "f64", f64_cells,
                              2, "c5 pass in the float64 parity context, %d cells per GPU" % f64_cells))
    if world > 1:
        strong_cells = max(4, (1000000 // world) // 4 * 4)
        def strong():
            res = sparse_block(args.network, "f32", strong_cells, 3,
                               "STRONG scaling point: 10^6 cells in total, %d per GPU, contract network" % strong_cells)()
            res["scaling"] = "strong"
            return res
        block("strong", strong)

    # -- the drop-in MODULE surface end to end: the reference's own call sequence with its float64
    #    (genes, cells) host matrices -- upload, lineage run with probes, download, read-out -- on one GPU
    def module_block():
        if rank != 0:
            return None
        import scrutiny_b200.MatriSeq as ms
        cells = min(100000, args.cells)
        ms.init(outputDir=None, seed=1337, verbose=False, device=env["local"])
        import scipy.sparse as sp
        gc = sp.csr_matrix((rig.csr[2], rig.csr[1], rig.csr[0]), shape=(n, n))
        tree = build_tree(n, cells, rig.tf)
        sampler = ClockSampler(env["local"]).start()
        best = None
        for rep in range(3):
            states = np.repeat(rig.x0[:, None], cells, axis=1)
            types = np.zeros(cells, dtype="int64")
            iters = np.zeros(cells, dtype="int64")
            t0 = time.perf_counter()
            ms.findAllNextStates(gc, tree.head(), tree, states, types, iters)
            t1 = time.perf_counter()
            counts, size = ms.transformData(n, cells, states, compact=True)
            t2 = time.perf_counter()
            if rep and (best is None or t2 - t0 < best[0]):
                best = (t2 - t0, t1 - t0, t2 - t1, float(iters.sum()), counts.nbytes_compact())
        ms.init(outputDir=None, seed=1337, verbose=False, device=env["local"])     # drops the cached context
        return {"metric": "cell_steps_per_sec", "workload": "scrutiny_b200.MatriSeq.findAllNextStates + transformData(compact=True) on "
                "%d cells, contract network: float64 (genes, cells) host matrix in, the same matrix and the counts out" % cells,
                "value": best[3] / best[0], "seconds": best[0], "findAllNextStates_s": best[1], "transformData_s": best[2],
                "cell_steps": best[3], "h2d_bytes": int(2 * 8 * n * cells), "d2h_bytes": int(8 * n * cells + best[4]),
                "clocks": sampler.stop()}
    block("module_api", module_block)
    env["barrier"]()

    # -- config c4 on the tcgen05 split-precision stepper ------------------------------------------------------------
    def dense_block():
        cells = min(200000, args.cells)
        r = Rig(env, "dense_c4", cells)
        try:
            assert r.layout["dense_stepper"] == 1
            f16 = r.layout.get("dense_encoding", 2) == 1
            # denominator: a plain cuBLAS matmul in the stepper's operand type (fp16 / tf32), timed here, on this GPU, now,
            # back to back for about a second (the sustained figure under the power cap, like the stepper's own launches)
            torch.backends.cuda.matmul.allow_tf32 = True
            mdt = torch.float16 if f16 else torch.float32
            a = torch.randn(8192, 8192, device=dev).to(mdt)
            b = torch.randn(8192, 8192, device=dev).to(mdt)
            for _ in range(3):
                torch.matmul(a, b)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 1200 if f16 else 600
            ev0.record()
            for _ in range(reps):
                torch.matmul(a, b)
            ev1.record()
            torch.cuda.synchronize()
            mm_peak = 2.0 * 8192 ** 3 / (ev0.elapsed_time(ev1) / reps / 1e3) / 1e12
            del a, b
            t = r.timed(1, 1, seed0=9000, do_counts=False)
            ks = t["local_steps"] / (t["kernel_ms"] / 1e3)
            issued = ks * 6.0 * r.n * r.n / 1e12
            enc = "3 x fp16 split (kind::f16)" if f16 else "3xTF32 (kind::tf32)"
            measured = {}
            try:
                mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                measured = {"bf16_tflops_sustained": mp.get("bf16_tflops_sustained"), "bf16_tflops_burst": mp.get("bf16_tflops")}
            except Exception:
                pass
            return {"metric": "cell_steps_per_sec", "workload": "c4: Random n=3000 networkSparsity=0.5 (alpha 2.21), %d cells per GPU, "
                    "single type, stepping only" % cells, "value": t["value"], "ms_per_step": t["ms_per_step"], "n_gpus": world,
                    "genes": r.n, "nnz": r.nnz, "dtype": "f32 via " + enc, "node_iterations": t["node_iterations"],
                    "kernel_launches": t["kernel_launches"], "kernel_cell_steps_per_s": ks,
                    "roofline": {"bound": "tensor", "kernel": "scr::dense::dense_pair_kernel (tcgen05 cta_group::2, %s, fused update; "
                                 "the phase's last steps: dense_step_kernel, steps inside the launch)" % enc,
                                 "achieved": issued, "useful": issued / 3.0, "peak": mm_peak, "unit": "TFLOP/s",
                                 "frac": issued / mm_peak,
                                 "peak_source": "torch.matmul %s 8192^3 timed in this run, %d back to back (sustained)" % (
                                     "fp16" if f16 else "TF32", reps),
                                 "measured_peaks_json": measured,
                                 "frac_of_measured_bf16_sustained": (issued / measured["bf16_tflops_sustained"]
                                                                     if f16 and measured.get("bf16_tflops_sustained") else None),
                                 "flops_per_cell_step_issued": 6.0 * r.n * r.n},
                    "clocks": t["clocks"]}
        finally:
            r.close()
    block("dense_c4", dense_block)
    return out


def _ncu_traffic():
    """dram bytes per launch of the stepper from the committed ncu capture, if summarised."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        return None


if __name__ == "__main__":
    main()
