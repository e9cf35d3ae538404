#!/bin/bash
# Round-end evidence run on the GPU box (one B200): GPU test-suite, the contract bench line, the reference arm,
# the ncu launch list of the bench command and one ncu --set full capture of the stepper.  Outputs -> gpurun_out/final/.
# Numbers printed by the runs under ncu are never bench values.
out=gpurun_out/final
mkdir -p $out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > $out/pytest_gpu.log 2>&1
tail -3 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err && cat $out/bench.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 1 --launch-count 1 \
  -o $out/prof_step -f python scripts/quick_bench.py --net powerlaw --cells 60000 --steps 100 --reps 2 > $out/ncu_step.log 2>&1
ls -la $out
