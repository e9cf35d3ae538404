#!/bin/bash
# 8 GPUs: the driver's torchrun command line
out=gpurun_out/r2c39; mkdir -p $out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 2 > $out/bench_n4.json 2> $out/bench_n4.err; echo "rc=$?"; tail -3 $out/bench_n4.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2c39/bench_n4.json') if l.startswith('{')][0]
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'kernel', d['roofline']['kernel_cell_steps_per_s'], 'frac', d['roofline']['frac'], 'share', d['roofline']['kernel_share_of_step'], d['clocks'])
for k,v in d.get('extra',{}).items():
    print('  ',k, {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','error','alpha_searched','seconds','cells_per_gpu','scaling')}, v.get('roofline',{}).get('frac'))
PY
