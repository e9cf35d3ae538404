#!/bin/bash
# round 2, GPU call 3: GPU suite, count kernel A/B (float32 PTRS screen on / off, checksums must agree), small configs
out=gpurun_out/r2c3; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -12 $out/pytest_gpu.log
timeout 200 python scripts/counts_bench.py --cells 400000 --reps 4 2>&1 | tail -5 | tee $out/counts_new.txt
SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/noptrs.so timeout 200 python scripts/counts_bench.py --cells 400000 --reps 3 2>&1 | tail -4 | tee $out/counts_noptrs.txt
SCR_COUNTS_EXACT=1 timeout 200 python scripts/counts_bench.py --cells 400000 --reps 2 2>&1 | tail -3 | tee $out/counts_exact.txt
timeout 600 python scripts/bench_configs.py c1 c2 c3 c4 2> $out/configs.err | tee $out/configs.jsonl | cut -c1-1200
tail -3 $out/configs.err
