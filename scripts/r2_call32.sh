#!/bin/bash
out=gpurun_out/r2c32; mkdir -p $out
( timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
python scripts/time_breakdown.py 12500 2>&1 | grep "rep [123]" | tee $out/breakdown_c3.txt
SCR_PROBE_MULTI=0 python scripts/time_breakdown.py 12500 2>&1 | grep "rep [123]" | tee $out/breakdown_c3_seq.txt
timeout 600 python scripts/bench_configs.py c2 c3 > $out/configs.jsonl 2> $out/configs.err; tail -3 $out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2c32/configs.jsonl'):
    d=json.loads(l); print(d['config'], 'whole %.2f M  kernel %.2f M  ratio %.3f probe %.2f ms/%d  wall %.1f ms alpha %s iters %s' % (d['cell_steps_per_s_whole_job']/1e6, d['cell_steps_per_s_kernel']/1e6, d['whole_job_over_kernel'], d['probe_kernel_ms'], d['probe_launches'], d['wall_s']*1e3, d['alpha'], d['node_iterations']))
PY
