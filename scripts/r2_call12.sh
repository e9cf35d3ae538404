#!/bin/bash
# round 2, GPU call 12: GPU suite against the bounds-checked build; dense stepper A/B (16-float K blocks, deeper ring)
out=gpurun_out/r2c12; mkdir -p $out
( SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/checks.so timeout 1200 python -m pytest tests -q -m gpu ) > $out/pytest_gpu_debug_checks.log 2>&1; tail -5 $out/pytest_gpu_debug_checks.log
SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/bk16.so timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -k "dense or c4" 2>&1 | tail -3 | tee $out/bk16_tests.txt
for v in "" bk16; do
  if [ -n "$v" ]; then export SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so; else unset SCRUTINY_B200_LIB; fi
  echo "== dense variant ${v:-bk32}"
  timeout 300 python scripts/quick_bench.py --net dense3000 --cells 100000 --steps 40 --reps 3 --alpha 2.21 2>&1 | grep "rep [12]"
  timeout 100 python scripts/quick_bench.py --net dense500 --cells 1000 --steps 400 --reps 3 --alpha 2.69 2>&1 | grep "rep [12]"
done | tee $out/dense_ab.txt
unset SCRUTINY_B200_LIB
