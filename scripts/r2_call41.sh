#!/bin/bash
# ncu summaries of the stepper with the hub-first operator order (numbers printed under ncu are never bench values)
out=gpurun_out/r2c41; mkdir -p $out
for net in powerlaw trrust; do
  al=0.4; [ $net = trrust ] && al=0.1
  Q="python scripts/quick_bench.py --net $net --alpha $al --cells 60000 --steps 100 --reps 2"
  $Q > $out/plain_$net.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 1 --launch-count 1 -o $out/prof_step_$net -f $Q > $out/ncu_$net.log 2>&1; echo "$net ncu rc=$?"; grep "rep 1" $out/plain_$net.log
done
