#!/bin/bash
# round 2, GPU call 8 (2 GPUs): the bench command as the driver launches it for N = 2, reference arm included
out=gpurun_out/r2c8; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
echo "rc=$?"; cut -c1-400 $out/bench_n2.json; tail -5 $out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c8/bench_n2.json'))
    for k,v in d.get('extra',{}).items(): print(k, json.dumps(v)[:400])
    print('value', d['value'], 'e2e', d['e2e']); print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_share_of_step','kernel_cell_steps_per_s')})
except Exception as e: print('parse failed', e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 1 --gather --no-extra --no-cpu > $out/bench_n2_gather.json 2> $out/bench_n2_gather.err
echo "rc=$?"; cut -c1-300 $out/bench_n2_gather.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 6 > $out/ref_n2.json 2> $out/ref_n2.err
echo "rc=$?"; cut -c1-300 $out/ref_n2.json
