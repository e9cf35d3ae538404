#!/bin/bash
# round 2, GPU call 15 (8 GPUs): the driver's N = 8 command line, short
out=gpurun_out/r2c15; mkdir -p $out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 2 > $out/bench_n8.json 2> $out/bench_n8.err
echo "rc=$?"; tail -4 $out/bench_n8.err
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r2c15/bench_n8.json') if l.startswith('{')][0]
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'share', d['roofline']['kernel_share_of_step'], 'kernel', d['roofline']['kernel_cell_steps_per_s'], 'clocks', d['clocks'])
    for k,v in d.get('extra',{}).items():
        print('  ',k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','error','alpha_searched','seconds','cells_per_gpu','scaling')}, v.get('roofline',{}).get('frac'))
except Exception as e: print('parse failed', e)
PY
