#!/bin/bash
# round 2, GPU call 14: annealed bank assignment of the gene order (stepper-only rates)
out=gpurun_out/r2c14; mkdir -p $out
for cfg in "0 0" "200 0" "1000 0" "4000 0" "1000 1"; do
  set -- $cfg
  for net in powerlaw trrust powerlaw205; do
    echo -n "anneal=$1k opt2=$2 $net: "
    SCR_BANK_ANNEAL=$1 SCR_BANK_OPT2=$2 timeout 200 python scripts/quick_bench.py --net $net --cells 200000 --steps 200 --reps 3 --alpha 0.4 2>&1 | grep -E "rep [12]|layout" | sed -E "s/.*gather_wavefronts_model': ([0-9]+).*/model \1 |/" | awk '{if ($1=="model") printf "%s %s ", $1, $2; else printf "%s M/s | ", $9} END {print ""}'
  done
done | tee $out/anneal.txt
