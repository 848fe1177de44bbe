#!/bin/bash
# usage: scripts/build_variant.sh <name> [extra nvcc flags...]   -> racon_gpu_b200/variants/libb200poa_<name>.so
# Experimental builds (phase timers, occupancy sweeps); select one at run time with B200POA_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
name="$1"; shift
mkdir -p racon_gpu_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3,-pthread -shared \
  --expt-relaxed-constexpr -Xptxas -v "$@" -x cu -I include -I racon_gpu_b200/csrc \
  -o racon_gpu_b200/variants/libb200poa_$name.so racon_gpu_b200/csrc/b200poa.cu racon_gpu_b200/csrc/b200aln.cu \
  racon_gpu_b200/csrc/host/cuda_batch.cpp racon_gpu_b200/csrc/host/cuda_polisher.cpp 2>&1 | grep -E "error|registers" | sort | uniq -c
