#!/bin/bash
# round 2, final evidence run on one B200: GPU suite, smoke, the bench command with the driver's flags, reference arm
out=gpurun_out/${OUT:-r2final}; mkdir -p $out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/smoke.log
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-250 $out/bench.json; tail -2 $out/bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; cut -c1-250 $out/bench_ref.json
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/'+__import__("os").environ.get("OUT","r2final")+'/bench.json') if l.startswith('{')][0]
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'share', d['roofline']['kernel_share_of_step'], 'kernel', d['roofline']['kernel_cell_steps_per_s'], 'frac', d['roofline']['frac'])
    for k,v in d.get('extra',{}).items():
        print('  ',k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','error','alpha_searched','seconds','cells_per_gpu','scaling')}, v.get('roofline',{}).get('frac'))
    r=[json.loads(l) for l in open('gpurun_out/'+__import__("os").environ.get("OUT","r2final")+'/bench_ref.json') if l.startswith('{')][0]
    print('same config:', r['config']==d['config'], 'ref value', r['value'], 'ref ms_per_step', r['ms_per_step'])
except Exception as e: print('parse failed', e)
PY
