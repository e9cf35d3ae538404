#!/bin/bash
# 4 GPUs: does binding each rank to its GPU's CPUs (NVML) change the end-to-end rate?  A/B in one call, headline + e2e only
out=gpurun_out/r2c44; mkdir -p $out
for b in 1 0; do
  SCR_BENCH_BIND=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$b bench.py --gpus 4 --steps 3 --warmup 2 --no-extra --no-cpu --e2e-steps 3 > $out/bench_bind$b.json 2> $out/bench_bind$b.err; echo "bind=$b rc=$?"
  python - <<PY
import json
d=[json.loads(l) for l in open('$out/bench_bind$b.json') if l.startswith('{')][0]
print('bind=$b value', d['value'], 'e2e', d['e2e']['value'], d['details'].get('host_affinity'))
PY
done
nvidia-smi topo -m 2>/dev/null | head -12; python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))"
