#!/bin/bash
# N = 2 contract line of the final code (level-order run_all over NCCL), short: 2 timed steps
out=gpurun_out/r2s3n2; mkdir -p $out
timeout 190 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 2 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$out/bench_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['clocks'])
for k,v in d.get('extra',{}).items():
    print('  ', k, {kk: v[kk] for kk in ('value','seconds') if isinstance(v, dict) and kk in v})
PY
