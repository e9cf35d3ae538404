#!/bin/bash
# round 2, GPU call 1: full GPU suite (with the new bench-config parity tests), stepper A/B (noise stream v2, barrier variants)
out=gpurun_out/r2c1; mkdir -p $out
( time timeout 600 python -m pytest tests -x -q -m gpu ) > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log
cp scrutiny_b200/libscrutiny_b200.so scrutiny_b200/_ab/new.so
for v in new last early lastearly v1; do echo -n "$v: "; SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so timeout 120 python scripts/ab_hash.py 2>&1 | tr '\n' ' '; echo; done | tee $out/hash.txt
NETS="powerlaw trrust powerlaw205" STEPS=200 CELLS=200000 timeout 600 bash scripts/ab_run.sh new v1 last early lastearly 2>&1 | tee $out/rates.txt
for net in powerlaw trrust powerlaw205; do echo -n "bankopt $net: "; SCR_BANK_OPT=1 timeout 120 python scripts/quick_bench.py --net $net --cells 200000 --steps 200 --reps 3 2>&1 | grep "rep [12]" | awk '{printf "%s ms %s M/s | ", $4, $9} END {print ""}'; done | tee $out/bankopt.txt
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > $out/bench.json 2> $out/bench.err; cut -c1-600 $out/bench.json; tail -3 $out/bench.err
