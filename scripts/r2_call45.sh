#!/bin/bash
# level-order phases (one probe round per tree level): A/B on c2 / c3, then the GPU tests that run run_all
out=gpurun_out/r2c45; mkdir -p $out
for lo in 1 0; do
  SCR_LEVEL_ORDER=$lo timeout 300 python scripts/bench_configs.py c2 c3 > $out/configs_lo$lo.jsonl 2> $out/configs_lo$lo.err
  python - <<PY
import json
for l in open('$out/configs_lo$lo.jsonl'):
    d=json.loads(l); print('level_order=$lo', d['config'], 'whole %.2f M  kernel %.2f M  ratio %.3f probe %.2f ms/%d  wall %.2f ms alpha %s iters %s' % (d['cell_steps_per_s_whole_job']/1e6, d['cell_steps_per_s_kernel']/1e6, d['whole_job_over_kernel'], d['probe_kernel_ms'], d['probe_launches'], d['wall_s']*1e3, d['alpha'], d['node_iterations']))
PY
done
( timeout 600 python -m pytest tests/test_gpu_bench_configs.py tests/test_gpu_round2.py tests/test_gpu_surface.py tests/test_gpu_parity.py -q -m gpu -x ) > $out/pytest.log 2>&1; grep "passed\|failed" $out/pytest.log
