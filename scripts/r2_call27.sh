#!/bin/bash
out=gpurun_out/r2c27; mkdir -p $out
( timeout 300 python -m pytest tests/test_gpu_dense.py -q -m gpu -x -k "cta-pairs" ) > $out/pytest_pairs.log 2>&1; tail -5 $out/pytest_pairs.log
( timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -m gpu -x -k "dense or c4" ) > $out/pytest_dense.log 2>&1; tail -5 $out/pytest_dense.log
for pair in 1 0; do
  echo "== f16 c4-shaped: 100k cells x 40 steps, SCR_DENSE_PAIR=$pair"
  SCR_DENSE_PAIR=$pair timeout 300 python scripts/quick_bench.py --net dense3000 --cells 100000 --steps 40 --alpha 2.21 --reps 3 2>&1 | grep "rep\|finite\|rror" | tee -a $out/quick_pair$pair.log
done
