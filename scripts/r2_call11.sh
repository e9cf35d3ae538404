#!/bin/bash
# round 2, GPU call 11: compute-sanitizer memcheck of the every-kernel workload (one tool per call, B200_PROFILING.md)
out=gpurun_out/r2c11; mkdir -p $out
timeout 120 python scripts/sanitize_workload.py sparse readout > $out/plain.log 2>&1; echo "plain rc=$?"; tail -3 $out/plain.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python scripts/sanitize_workload.py sparse readout > $out/memcheck_sparse_readout.log 2>&1
echo "memcheck rc=$?"; tail -12 $out/memcheck_sparse_readout.log
