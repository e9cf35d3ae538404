#!/bin/bash
# round 2, GPU call 9: count kernel CTA-size A/B, GPU suite on the current build
out=gpurun_out/r2c9; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -6 $out/pytest_gpu.log
for v in "" ct128 ct64; do
  echo "== counts variant: ${v:-default256}"
  if [ -n "$v" ]; then export SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so; else unset SCRUTINY_B200_LIB; fi
  timeout 200 python scripts/counts_bench.py --cells 400000 --reps 4 2>&1 | tail -3
done | tee $out/counts_cta_size.txt
unset SCRUTINY_B200_LIB
