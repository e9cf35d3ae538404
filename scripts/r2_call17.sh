#!/bin/bash
# round 2, GPU call 17: 16-warp probe groups -- probe / surface tests, small-config breakdown, c1-c3 configs
out=gpurun_out/r2c17; mkdir -p $out
( timeout 900 python -m pytest tests -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
for c in 12500; do timeout 200 python scripts/time_breakdown.py $c 2>&1 | tail -2; SCR_PROBE_WIDE=0 timeout 200 python scripts/time_breakdown.py $c 2>&1 | tail -1 | sed 's/^/probe_wide=0 /'; done | tee $out/breakdown.txt
timeout 300 python scripts/bench_configs.py c2 c3 2> $out/configs.err | tee $out/configs.jsonl | cut -c1-700
