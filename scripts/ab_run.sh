#!/bin/bash
# Developer A/B timing on the GPU box: scripts/ab_run.sh VARIANT... (built by scripts/ab_build.sh)
# prints the stepper-only rate of each variant on the bench-shaped power-law network (and TRRUST with NETS="powerlaw trrust")
for v in "$@"; do
  for net in ${NETS:-powerlaw}; do
    echo -n "$v $net: "
    SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so python scripts/quick_bench.py --net $net --cells ${CELLS:-200000} --steps ${STEPS:-200} --reps 3 2>&1 | grep "rep [12]" | awk '{printf "%s ms %s M/s | ", $4, $9} END {print ""}'
  done
done
