#!/bin/bash
# the contract pass on the TRRUST network (north_star's other named network), full line incl. e2e
out=gpurun_out/r2c43; mkdir -p $out
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --network trrust --no-extra > $out/bench_trrust.json 2> $out/bench_trrust.err; echo "rc=$?"; tail -2 $out/bench_trrust.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2c43/bench_trrust.json') if l.startswith('{')][0]
print(d['config']['workload']); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'kernel', d['roofline']['kernel_cell_steps_per_s'], 'frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
