#!/bin/bash
# round 2, GPU call 10: dense tests with the narrow tile, c1 / c4 configs, bench with all extras
out=gpurun_out/r2c10; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -6 $out/pytest_gpu.log
SCR_DENSE_NARROW=0 timeout 300 python -m pytest tests/test_gpu_dense.py -q 2>&1 | tail -2
SCR_DENSE_NARROW=1 timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -k "dense or c4" 2>&1 | tail -2
timeout 600 python scripts/bench_configs.py c1 c4 2> $out/configs.err | tee $out/configs.jsonl | cut -c1-600
SCR_DENSE_NARROW=0 timeout 300 python scripts/bench_configs.py c1 2>> $out/configs.err | tee $out/configs_c1_wide.jsonl | cut -c1-400
timeout 900 python bench.py --steps 3 --warmup 3 > $out/bench.json 2> $out/bench.err; cut -c1-300 $out/bench.json; tail -3 $out/bench.err
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2c10/bench.json') if l.startswith('{')][0]
    for k,v in d.get('extra',{}).items(): print(k, json.dumps({kk:vv for kk,vv in v.items() if kk not in ('clocks','layout','workload','node_iterations','roofline')})[:400], v.get('roofline',{}).get('frac'))
    print('e2e', d['e2e']['value']); print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_share_of_step','kernel_cell_steps_per_s')})
except Exception as e: print('parse failed', e)
PY
