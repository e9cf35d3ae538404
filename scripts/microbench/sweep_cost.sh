#!/bin/bash
for net in powerlaw trrust powerlaw205; do
for cfg in "60 0" "60 4" "60 7" "60 14" "120 7" "30 7" "120 14"; do
  set -- $cfg
  echo -n "$net chunk_cost=$1 ovf_cost=$2: "
  SCR_CHUNK_COST=$1 SCR_OVF_COST=$2 timeout 60 python scripts/quick_bench.py --net $net --cells 100000 --steps 200 --reps 3 2>&1 | grep -E "rep [12]" | awk '{printf "%s M/s | ", $9} END {print ""}'
done; done
