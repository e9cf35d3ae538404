#!/bin/bash
out=gpurun_out/r2c31; mkdir -p $out
( timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -m gpu -x -k "dense or c4" ) > $out/pytest_dense.log 2>&1; tail -5 $out/pytest_dense.log
