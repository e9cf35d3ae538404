#!/bin/bash
out=gpurun_out/r2c34; mkdir -p $out
( timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
for i in 1 2 3 4 5 6; do timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k philox 2>&1 | tail -1; done
