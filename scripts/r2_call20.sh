#!/bin/bash
# round 2, GPU call 20 (ncu only): final-code captures -- stepper, count kernel, dense narrow tile, launch list of the bench command
out=gpurun_out/r2c20; mkdir -p $out
C2="python scripts/quick_bench.py --net powerlaw --cells 60000 --steps 100 --reps 2 --alpha 0.4"
$C2 > $out/step_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 1 --launch-count 1 -o $out/prof_step -f $C2 > $out/ncu_step.log 2>&1
tail -2 $out/step_plain.log; tail -1 $out/ncu_step.log
C1="python scripts/counts_bench.py --cells 40000 --reps 1"
$C1 > $out/counts_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:counts_kernel -c 1 -o $out/prof_counts -f $C1 > $out/ncu_counts.log 2>&1
tail -2 $out/counts_plain.log; tail -1 $out/ncu_counts.log
C4="python scripts/quick_bench.py --net dense3000 --cells 40000 --steps 6 --reps 1 --alpha 2.21"
$C4 > $out/dense_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_step_kernel --launch-skip 3 --launch-count 1 -o $out/prof_dense -f $C4 > $out/ncu_dense.log 2>&1
tail -2 $out/dense_plain.log; tail -1 $out/ncu_dense.log
C3="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra"
$C3 > $out/bench_plain.json 2> $out/bench_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv $C3 > $out/ncu_bench.log 2>&1
cut -c1-200 $out/bench_plain.json; wc -l $out/launches.csv
