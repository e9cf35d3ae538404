#!/bin/bash
out=gpurun_out/r2c21; mkdir -p $out
( timeout 900 python -m pytest tests -q -m gpu -x ) > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu --extra module_api > $out/bench.json 2> $out/bench.err; tail -2 $out/bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2c21/bench.json') if l.startswith('{')][0]
print('value', d['value'], 'e2e', d['e2e']['value']); print(json.dumps(d['extra'])[:900])
PY
