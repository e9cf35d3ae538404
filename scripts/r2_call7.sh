#!/bin/bash
# round 2, GPU call 7: whole GPU suite (no -x), smoke, small configs
out=gpurun_out/r2c7; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py c1 c2 c3 c4 2> $out/configs.err | tee $out/configs.jsonl | cut -c1-500
