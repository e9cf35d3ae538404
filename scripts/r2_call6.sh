#!/bin/bash
# round 2, GPU call 6: GPU suite after the shared-memory budget fix, stepper A/B (streaming state accesses, prefix sweep), bench
out=gpurun_out/r2c6; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -8 $out/pytest_gpu.log
NETS="powerlaw trrust powerlaw205" STEPS=200 CELLS=200000 timeout 400 bash scripts/ab_run.sh new streamstate 2>&1 | tee $out/rates.txt
for kb in 0 16 30 34 36; do for net in powerlaw powerlaw205; do for v in new streamstate; do echo -n "prefix_kb=$kb $net $v: "; SCR_W_PREFIX_KB=$kb SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so timeout 100 python scripts/quick_bench.py --net $net --cells 200000 --steps 200 --reps 3 2>&1 | grep "rep [12]" | awk '{printf "%s M/s | ", $9} END {print ""}'; done; done; done | tee $out/prefix_sweep.txt
timeout 600 python bench.py --steps 3 --warmup 3 --extra counts --no-cpu > $out/bench.json 2> $out/bench.err; cut -c1-300 $out/bench.json; tail -3 $out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c6/bench.json'))
    for k,v in d.get('extra',{}).items(): print(k, json.dumps(v)[:300])
    print('e2e', d['e2e']['value']); print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_share_of_step','kernel_cell_steps_per_s')})
except Exception as e: print('parse failed', e)
PY
timeout 300 python scripts/bench_configs.py c2 c3 2> $out/configs.err | tee $out/configs.jsonl | cut -c1-600
