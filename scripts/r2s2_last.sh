#!/bin/bash
# last check of the round: GPU suite, smoke, a short bench (every extra block), on the code as committed
out=gpurun_out/r2s2last; mkdir -p $out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee $out/smoke.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"; tail -2 $out/bench_default.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2s2last/bench_default.json') if l.startswith('{')][0]
print('value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'frac', d['roofline']['frac'], d['clocks'])
for k,v in d.get('extra',{}).items():
    print('  ',k, {kk:vv for kk,vv in v.items() if kk in ('value','error','seconds')}, v.get('roofline',{}).get('frac'))
PY
