#!/bin/bash
out=gpurun_out/r2c26; mkdir -p $out
( timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
timeout 600 python scripts/bench_configs.py c1 c4 > $out/configs.jsonl 2> $out/configs.err; tail -3 $out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2c26/configs.jsonl'):
    d=json.loads(l); print(d['config'], d['path'], 'whole %.2f M  kernel %.2f M  launches %d  wall %.1f ms alpha %s issued %.0f TF' % (d['cell_steps_per_s_whole_job']/1e6, d['cell_steps_per_s_kernel']/1e6, d['kernel_launches'], d['wall_s']*1e3, d['alpha'], d['tflops_split_issued']))
PY
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --cells 200000 --extra dense_c4 > $out/bench_dense.json 2> $out/bench_dense.err; tail -2 $out/bench_dense.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2c26/bench_dense.json') if l.startswith('{')][0]
print(json.dumps(d['extra'],indent=1)[:2500])
PY
