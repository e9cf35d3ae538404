#!/bin/bash
# Developer A/B builds of the library: scripts/ab_build.sh NAME [-DFLAG=VALUE ...] -> scrutiny_b200/_ab/NAME.so
# (git-ignored, travels to the GPU box; select with SCRUTINY_B200_LIB=scrutiny_b200/_ab/NAME.so)
set -e
cd "$(dirname "$0")/../scrutiny_b200/csrc"
name=$1; shift
mkdir -p ../_ab
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
  "$@" -shared -o ../_ab/$name.so scr_api.cu
echo built scrutiny_b200/_ab/$name.so
