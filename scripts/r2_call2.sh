#!/bin/bash
# round 2, GPU call 2: full GPU suite, smoke, bench with the extra blocks
out=gpurun_out/r2c2; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -25 $out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 2 > $out/bench.json 2> $out/bench.err; cut -c1-1500 $out/bench.json; tail -5 $out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c2/bench.json'))
    for k,v in d.get('extra',{}).items():
        print(k, json.dumps(v)[:700])
    print('e2e', d['e2e'])
    print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_share_of_step','kernel_cell_steps_per_s')})
except Exception as e: print('parse failed', e)
PY
