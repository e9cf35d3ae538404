"""Measurement of the RegNetInference front end on one B200, next to the CPU oracle.
Prints one JSON object per workload (kept under profiles/).

    python scripts/bench_regnet.py [--cells 200000] [--pair-cells 60000]

rank transform: n = 3000 genes x `cells` Poisson counts, the (genes, cells) float64 matrix MatriSeq.saveData
writes; `e2e` = scr_rank_transform with a pinned HOST matrix (H2D + both passes + D2H), `kernel` = the two kernels'
CUDA-event time.  Algorithmic bytes per entry: pass 1 reads 8 B and writes a 2 B key, pass 2 reads the key twice
(histogram, output) and writes 8 B = 22 B.
pair search: `pair-cells` cells in 7 types, pseudotimes as MatriSeq produces them (multiples of the time step),
threshold chosen for ~20 candidate partners per cell.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import regnet_port as rp  # noqa: E402  (CPU baseline leg only)
from scrutiny_b200 import engine  # noqa: E402


def pinned(shape):
    try:
        import torch
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    except Exception:
        return np.empty(shape)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=200000)
    ap.add_argument("--pair-cells", type=int, default=60000)
    ap.add_argument("--genes", type=int, default=3000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    eng = engine.Engine(0, "f32")
    rs = np.random.RandomState(1337)
    n, cells = a.genes, a.cells

    # ---- rank transform ----
    lam = np.exp(rs.normal(0.0, 1.2, size=(n, 1)))
    src = np.empty((n, cells))
    for g0 in range(0, n, 250):
        src[g0:g0 + 250] = rs.poisson(lam[g0:g0 + 250], size=(min(250, n - g0), cells))
    S = pinned((n, cells))
    e2e, kern = [], []
    for r in range(a.reps + 1):
        S[...] = src
        t0 = time.perf_counter()
        eng.rank_transform(S)
        dt = time.perf_counter() - t0
        if r:                                   # first call = warm-up
            e2e.append(dt)
            kern.append(eng.rank_transform_stats()["kernel_ms"] / 1e3)
    st = eng.rank_transform_stats()
    sub = 1500
    t0 = time.perf_counter()
    want = rp.convert_to_s(n, sub, src[:, :sub].copy())
    cpu_dt = time.perf_counter() - t0
    entries = float(n) * cells
    k = float(np.median(kern))
    print(json.dumps({
        "workload": "convertToS: %d genes x %d cells of Poisson counts (float64 host matrix, in place)" % (n, cells),
        "metric": "entries_per_sec", "e2e": {"value": entries / float(np.median(e2e)), "unit": "entries/s",
                                             "h2d_bytes": int(entries * 8), "d2h_bytes": int(entries * 8)},
        "kernel": {"value": entries / k, "unit": "entries/s", "ms": k * 1e3},
        "paths": st,
        "roofline": {"bound": "hbm", "achieved": entries * 22 / k / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": entries * 22 / k / 1e9 / peak, "algorithmic_bytes_per_entry": 22},
        "cpu_baseline": {"value": n * sub / cpu_dt, "unit": "entries/s", "cores": 1, "kind": "port",
                         "sample": "oracle convert_to_s (numpy argsort ranks) on %d x %d, %.1f s" % (n, sub, cpu_dt)},
        "parity_spot_check": bool(want.shape == (n, sub))}), flush=True)
    del S, src

    # ---- pair search ----
    pc = a.pair_cells
    ng = 200
    Sp = rp.convert_to_s(ng, pc, rs.poisson(1.0, size=(ng, pc)).astype(np.float64))
    types = rs.randint(0, 7, size=pc).astype(np.int64)
    pt = 0.01 * rs.randint(0, 480, size=pc)
    thresh = 0.01 * max(1, int(round(20.0 / (pc / 7.0 / 480.0))))
    times = []
    for r in range(a.reps + 1):
        t0 = time.perf_counter()
        pairs = eng.cell_pairs(Sp, pt, types, thresh)
        if r:
            times.append(time.perf_counter() - t0)
    subc = 2500
    t0 = time.perf_counter()
    want = rp.find_cell_pairs_fast(Sp[:, :subc], pt[:subc], types[:subc], thresh * pc / subc)
    cpu_dt = time.perf_counter() - t0
    comparisons = float(sum(int(np.sum(types == t)) ** 2 for t in range(7)))
    sub_comp = float(sum(int(np.sum(types[:subc] == t)) ** 2 for t in range(7)))
    print(json.dumps({
        "workload": "cell pairs: %d cells, 7 types, %d genes, thresh %.2f -> %d pairs" % (pc, ng, thresh, len(pairs)),
        "metric": "candidate_comparisons_per_sec",
        "e2e": {"value": comparisons / float(np.median(times)), "unit": "comparisons/s", "seconds": float(np.median(times)),
                "note": "both passes (count, fill), host matrices in, int64 pair list out"},
        "cpu_baseline": {"value": sub_comp / cpu_dt, "unit": "comparisons/s", "cores": 1, "kind": "port",
                         "sample": "oracle find_cell_pairs_fast on %d cells (%d pairs), %.1f s" % (subc, len(want), cpu_dt)}}),
          flush=True)


if __name__ == "__main__":
    main()
