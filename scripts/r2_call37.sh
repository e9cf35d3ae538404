#!/bin/bash
out=gpurun_out/r2c37; mkdir -p $out
( timeout 900 python -m pytest tests/test_gpu_fuzz.py -q -m gpu -k dense ) > $out/pytest_fuzz.log 2>&1; tail -15 $out/pytest_fuzz.log
