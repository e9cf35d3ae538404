#!/bin/bash
# round 2, second session: final evidence run on one B200 (GPU suite, smoke, the driver's bench command, reference arm, the other
# configurations, launch list of the bench command under ncu -- a number printed under ncu is never a bench value)
export OUT=r2s2final
out=gpurun_out/$OUT; mkdir -p $out
bash scripts/r2_final.sh
timeout 900 python scripts/bench_configs.py c1 c2 c3 c4 > $out/configs.jsonl 2> $out/configs.err; tail -2 $out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/r2s2final/configs.jsonl'):
    d=json.loads(l); print(d['config'], d['path'], 'whole %.2f M  kernel %.2f M  ratio %.3f launches %d probe %.2f ms/%d  wall %.1f ms alpha %s' % (d['cell_steps_per_s_whole_job']/1e6, d['cell_steps_per_s_kernel']/1e6, d['whole_job_over_kernel'], d['kernel_launches'], d['probe_kernel_ms'], d['probe_launches'], d['wall_s']*1e3, d['alpha']))
PY
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --extra counts"
$B > $out/bench_short.json 2> $out/bench_short.err && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv $B > $out/ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python scripts/launch_list.py $out/launches.csv 2>/dev/null | tail -25
