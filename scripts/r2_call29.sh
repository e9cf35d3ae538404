#!/bin/bash
# ncu evidence of the new dense kernels + a look at the sparse probe (numbers printed under ncu are never bench values)
out=gpurun_out/r2c29; mkdir -p $out
python scripts/probe_bench.py --net powerlaw 2>&1 | grep rep | tee $out/probe_plain.log
Q4="python scripts/quick_bench.py --net dense3000 --cells 40000 --steps 4 --alpha 2.21 --reps 1"
$Q4 > $out/q4_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_pair_kernel --launch-skip 2 --launch-count 1 -o $out/prof_dense_pair -f $Q4 > $out/ncu_q4.log 2>&1; tail -2 $out/ncu_q4.log
Q1="python scripts/quick_bench.py --net dense500 --cells 1000 --steps 100 --alpha 1.0 --reps 1"
$Q1 > $out/q1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dense_step_kernel --launch-count 1 -o $out/prof_dense_res -f $Q1 > $out/ncu_q1.log 2>&1; tail -2 $out/ncu_q1.log
P="python scripts/probe_bench.py --net powerlaw --reps 1"
$P > $out/p_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-count 1 -o $out/prof_probe -f $P > $out/ncu_probe.log 2>&1; tail -2 $out/ncu_probe.log
ls -la $out
