#!/bin/bash
# dense stepper: fp16-split encoding + steps-inside schedule -- parity first, then rates
out=gpurun_out/r2c22; mkdir -p $out
( timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -m gpu -x -k "dense or c4" ) > $out/pytest_dense.log 2>&1; tail -5 $out/pytest_dense.log
for kind in f16 tf32; do
  echo "== $kind c4-shaped: 100k cells x 40 steps"
  SCR_DENSE_KIND=$kind timeout 300 python scripts/quick_bench.py --net dense3000 --cells 100000 --steps 40 --alpha 2.21 --reps 3 2>&1 | grep "rep\|finite\|rror" | tee -a $out/quick_$kind.log
  echo "== $kind c1-shaped: 1000 cells x 400 steps"
  SCR_DENSE_KIND=$kind timeout 300 python scripts/quick_bench.py --net dense500 --cells 1000 --steps 400 --alpha 1.0 --reps 3 2>&1 | grep "rep\|finite\|rror" | tee -a $out/quick_$kind.log
done
echo "== f16 c1-shaped, one launch per step"
SCR_DENSE_INSIDE=0 timeout 300 python scripts/quick_bench.py --net dense500 --cells 1000 --steps 400 --alpha 1.0 --reps 3 2>&1 | grep "rep\|finite\|rror" | tee -a $out/quick_f16_perstep.log
timeout 600 python scripts/bench_configs.py c1 c4 > $out/configs.jsonl 2> $out/configs.err; tail -3 $out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2c22/configs.jsonl'):
    d=json.loads(l); print(d['config'], 'whole %.2f M  kernel %.2f M  launches %d  probe %.2f ms/%d  wall %.1f ms alpha %s' % (d['cell_steps_per_s_whole_job']/1e6, d['cell_steps_per_s_kernel']/1e6, d['kernel_launches'], d['probe_kernel_ms'], d['probe_launches'], d['wall_s']*1e3, d['alpha']), d.get('tflops_3xtf32_issued'))
PY
