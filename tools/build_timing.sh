#!/bin/sh
# Debug build with per-CTA phase time stamps (MGMC_TILE_TIMING, fused.cuh TSTAMP); used as
#   MGMC_LIB=build/libmgmc_timing.so MGMC_TIMING_FILE=gpurun_out/tim MGMC_TIMING_LEVEL=0 python bench.py --steps 10 --warmup 5 --no-cpu-baseline
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DMGMC_TILE_TIMING \
  -o build/libmgmc_timing.so multigridmc_b200/csrc/mgmc_b200.cu
