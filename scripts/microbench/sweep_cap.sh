#!/bin/bash
for net in trrust powerlaw205 powerlaw; do
for cfg in "8 8" "8 4" "8 16" "12 8" "16 8" "16 16" "6 6" "4 4" "24 8"; do
  set -- $cfg
  echo -n "$net cap=$1 lv=$2: "
  SCR_ROW_CAP=$1 SCR_CHUNK_LEN=$2 timeout 60 python scripts/quick_bench.py --net $net --cells 100000 --steps 200 --reps 2 2>&1 | grep -E "rep 1|layout" | sed -E 's/.*tile_slots.: ([0-9]+), .chunks.: ([0-9]+).*/slots \1 chunks \2/' | tr '\n' ' ' | awk '{print $1,$2,$3,$4, $(NF-11), $(NF-10)}'
done; done
