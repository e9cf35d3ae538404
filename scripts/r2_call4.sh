#!/bin/bash
# round 2, GPU call 4 (ncu only): full captures of the count kernel and the stepper, launch list of the bench command
out=gpurun_out/r2c4; mkdir -p $out
C1="python scripts/counts_bench.py --cells 40000 --reps 1"
$C1 > $out/counts_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:counts_kernel -c 1 -o $out/prof_counts -f $C1 > $out/ncu_counts.log 2>&1
tail -2 $out/counts_plain.log; tail -2 $out/ncu_counts.log
C2="python scripts/quick_bench.py --net powerlaw --cells 60000 --steps 100 --reps 2 --alpha 0.4"
$C2 > $out/step_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 1 --launch-count 1 -o $out/prof_step -f $C2 > $out/ncu_step.log 2>&1
tail -2 $out/step_plain.log; tail -2 $out/ncu_step.log
C3="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra"
$C3 > $out/bench_plain.json 2> $out/bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv $C3 > $out/ncu_bench.log 2>&1
cut -c1-300 $out/bench_plain.json; tail -2 $out/ncu_bench.log; wc -l $out/launches.csv
ls -la $out
