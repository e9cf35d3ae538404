#!/bin/bash
out=gpurun_out/r2c36; mkdir -p $out
for v in defer0 defer4 default defer10; do
  if [ $v = default ]; then lib=$PWD/scrutiny_b200/libscrutiny_b200.so; else lib=$PWD/scrutiny_b200/_ab/$v.so; fi
  echo "== $v"; SCRUTINY_B200_LIB=$lib python scripts/counts_bench.py --cells 400000 --reps 4 2>&1 | grep "rep [123]\|checksum" | tee -a $out/counts_$v.log
done
( timeout 600 python -m pytest tests -q -m gpu -x -k "count or compact or readout or transform or extension" ) > $out/pytest_counts.log 2>&1; tail -3 $out/pytest_counts.log
