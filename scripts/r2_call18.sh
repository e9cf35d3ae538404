#!/bin/bash
out=gpurun_out/r2c18; mkdir -p $out
timeout 600 python -m pytest tests -q -m gpu -k "count or readout or compact or extension or transform" 2>&1 | tail -2
for d in 1 0; do echo "== descending=$d"; SCR_COUNTS_DESCENDING=$d timeout 200 python scripts/counts_bench.py --cells 400000 --reps 4 2>&1 | tail -3; done | tee $out/counts_order.txt
