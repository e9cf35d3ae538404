#!/bin/bash
# One A/B call on the GPU box: bit-identity hashes and stepper-only rates of the listed variants (scrutiny_b200/_ab/*.so)
out=gpurun_out/ab; mkdir -p $out
for v in "$@"; do echo -n "$v: "; SCRUTINY_B200_LIB=$PWD/scrutiny_b200/_ab/$v.so python scripts/ab_hash.py 2>&1 | tr '\n' ' '; echo; done | tee $out/hash.txt
NETS="powerlaw trrust" STEPS=${STEPS:-200} CELLS=${CELLS:-200000} bash scripts/ab_run.sh "$@" 2>&1 | tee $out/rates.txt
