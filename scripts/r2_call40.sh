#!/bin/bash
# hub-first row / entry order + bank-aware entry exchange: stepper-only A/B on the three networks, then the GPU suite
out=gpurun_out/r2c40; mkdir -p $out
for net in powerlaw trrust powerlaw205; do
  al=0.4; [ $net = trrust ] && al=0.1
  for rs in 0 1; do
    echo -n "$net SCR_ROW_SORT=$rs: "
    SCR_ROW_SORT=$rs python scripts/quick_bench.py --net $net --alpha $al --cells 200000 --steps 200 --reps 3 2>&1 | grep "rep [12]" | awk '{printf "%s ms %s M/s | ", $4, $9} END {print ""}'
  done
done 2>&1 | tee $out/ab.txt
( timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
