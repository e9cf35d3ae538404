#!/bin/bash
# new reference-fixture lineage tests (errors printed), then the whole GPU suite
out=gpurun_out/r2s3last; mkdir -p $out
( timeout 60 python -m pytest tests/test_gpu_golden_lineages.py -q -m gpu -s ) > $out/pytest_new.log 2>&1
grep -a "lineage replay errors\|passed\|failed\|Error\|assert" $out/pytest_new.log | head -20
( timeout 110 python -m pytest tests -q -m gpu -p no:cacheprovider ) > $out/pytest_gpu.log 2>&1
tail -5 $out/pytest_gpu.log
