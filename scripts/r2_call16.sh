#!/bin/bash
out=gpurun_out/r2c16; mkdir -p $out
for c in 12500 100000; do echo "== $c cells, device counts"; timeout 200 python scripts/time_breakdown.py $c 2>&1 | tail -3; echo "== $c cells, host counts"; timeout 200 python scripts/time_breakdown.py $c host 2>&1 | tail -2; done | tee $out/breakdown.txt
