#!/bin/bash
# Round-end evidence run on the GPU box (one B200), most important first; every step has its own time limit and is
# skipped when the call's budget (BUDGET seconds, default 570) would be exceeded.  Outputs -> gpurun_out/final/.
# Numbers printed by the runs under ncu are never bench values.
out=gpurun_out/final
mkdir -p $out
t_start=$(date +%s)
budget=${BUDGET:-570}
left() { echo $(( budget - ( $(date +%s) - t_start ) )); }
step() {   # step NAME NEED_SECONDS LIMIT_SECONDS command...
  local name=$1 need=$2 limit=$3; shift 3
  local l=$(left)
  if [ "$l" -lt "$need" ]; then echo "[skip] $name: $l s left, needs $need" | tee -a $out/steps.log; return 0; fi
  if [ "$limit" -gt "$l" ]; then limit=$l; fi
  local t0=$(date +%s)
  timeout $limit "$@"
  echo "[done] $name: exit $? in $(( $(date +%s) - t0 )) s" | tee -a $out/steps.log
}
: > $out/steps.log
step pytest_gpu 60 200 bash -c "( time python -m pytest tests -x -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log"
step bench 60 170 bash -c "python bench.py > $out/bench.json 2> $out/bench.err; cut -c1-400 $out/bench.json"
if [ -n "$WITH_COUNTS" ]; then
step counts_ab 30 120 bash -c "
  python scripts/counts_bench.py > $out/counts_screened.log 2>&1; tail -3 $out/counts_screened.log
  SCR_COUNTS_EXACT=1 python scripts/counts_bench.py --reps 2 > $out/counts_exact.log 2>&1; tail -2 $out/counts_exact.log
  for v in c3 c2; do SCRUTINY_B200_LIB=\$PWD/scrutiny_b200/_ab/\$v.so python scripts/counts_bench.py --reps 3 > $out/counts_\$v.log 2>&1; tail -2 $out/counts_\$v.log; done"
fi
step ncu_launch_list 60 110 bash -c "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/ncu_bench.log 2>&1"
step ncu_step_full 50 90 bash -c "ncu --set full --clock-control none --import-source on -k regex:step_kernel --launch-skip 1 --launch-count 1 \
  -o $out/prof_step -f python scripts/quick_bench.py --net powerlaw --cells 60000 --steps 100 --reps 2 > $out/ncu_step.log 2>&1"
if [ -n "$WITH_COUNTS" ]; then
step ncu_counts_full 40 70 bash -c "ncu --set full --clock-control none --import-source on -k regex:counts_kernel --launch-count 1 \
  -o $out/prof_counts -f python scripts/counts_bench.py --cells 40000 --reps 1 > $out/ncu_counts.log 2>&1"
fi
step bench_ref 40 100 bash -c "python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err"
step bench_trrust 70 150 bash -c "python bench.py --network trrust > $out/bench_trrust.json 2> $out/bench_trrust.err"
ls -la $out
cat $out/steps.log
