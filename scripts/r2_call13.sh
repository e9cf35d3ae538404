#!/bin/bash
# round 2, GPU call 13: count kernel with prefetched entry loads (vs the ct128 build without), GPU read-out tests
out=gpurun_out/r2c13; mkdir -p $out
timeout 600 python -m pytest tests -q -m gpu -k "count or readout or compact or extension or transform" 2>&1 | tail -3
for v in "" ct128; do
  if [ -n "$v" ]; then export SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so; else unset SCRUTINY_B200_LIB; fi
  echo "== counts variant ${v:-prefetch}"
  timeout 200 python scripts/counts_bench.py --cells 400000 --reps 4 2>&1 | tail -3
done | tee $out/counts_prefetch.txt
