#!/bin/bash
out=gpurun_out/r2c30; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee $out/smoke.log
( timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_bench_configs.py -q -m gpu -x -k "dense or c4" ) > $out/pytest_dense.log 2>&1; tail -3 $out/pytest_dense.log
Q4="python scripts/quick_bench.py --net dense3000 --cells 40000 --steps 4 --alpha 2.21 --reps 1"
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__ops_path_tensor_src_fp16_dst_fp32.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,smsp__inst_executed.sum,launch__grid_size,launch__cluster_size
$Q4 > $out/q4_plain.log 2>&1 && timeout 200 ncu --metrics $M --clock-control none -k regex:dense_pair_kernel --launch-skip 2 --launch-count 1 --csv --log-file $out/ncu_pair_metrics.csv $Q4 > $out/ncu_q4.log 2>&1; echo "ncu pair rc=$?"; tail -3 $out/ncu_q4.log; cat $out/ncu_pair_metrics.csv | tail -14
SCR_DENSE_PAIR=0 $Q4 > $out/q4s_plain.log 2>&1 && SCR_DENSE_PAIR=0 timeout 200 ncu --metrics $M --clock-control none -k regex:dense_step_kernel --launch-skip 2 --launch-count 1 --csv --log-file $out/ncu_single_metrics.csv $Q4 > $out/ncu_q4s.log 2>&1; echo "ncu single rc=$?"; cat $out/ncu_single_metrics.csv | tail -14
