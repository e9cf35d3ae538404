#!/bin/bash
# round 2, GPU call 5: GPU suite (large-n fallback, persistent count kernel), count kernel A/B, bench
out=gpurun_out/r2c5; mkdir -p $out
( time timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -12 $out/pytest_gpu.log
timeout 200 python scripts/counts_bench.py --cells 400000 --reps 4 2>&1 | tail -5 | tee $out/counts_stream.txt
SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/c3.so timeout 200 python scripts/counts_bench.py --cells 400000 --reps 3 2>&1 | tail -4 | tee $out/counts_stream_3ctas.txt
SCR_COUNTS_STREAM=0 timeout 200 python scripts/counts_bench.py --cells 400000 --reps 3 2>&1 | tail -4 | tee $out/counts_percell.txt
timeout 600 python bench.py --steps 3 --warmup 3 --extra counts,trrust > $out/bench.json 2> $out/bench.err; cut -c1-400 $out/bench.json; tail -3 $out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c5/bench.json'))
    for k,v in d.get('extra',{}).items(): print(k, json.dumps(v)[:500])
    print('e2e', d['e2e']); print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_share_of_step','kernel_cell_steps_per_s')})
except Exception as e: print('parse failed', e)
PY
timeout 300 python scripts/bench_configs.py c2 c3 2> $out/configs.err | tee $out/configs.jsonl | cut -c1-700
