#!/bin/bash
# stepper rate against the size of the operator prefix staged in shared memory (SCR_W_PREFIX_KB; -1 = all that fits)
for net in powerlaw trrust powerlaw205; do
for kb in 0 8 16 24 32 -1; do
  echo -n "$net prefix_kb=$kb: "
  SCR_W_PREFIX_KB=$kb timeout 90 python scripts/quick_bench.py --net $net --cells 200000 --steps 200 --reps 3 2>&1 | grep -E "rep [12]" | awk '{printf "%s M/s | ", $9} END {print ""}'
done; done
