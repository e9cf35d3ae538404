#!/bin/bash
out=gpurun_out/r2c42; mkdir -p $out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $out/pytest_gpu.log 2>&1; grep "passed\|failed" $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee $out/smoke.log
