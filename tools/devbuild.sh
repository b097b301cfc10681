#!/bin/bash
# Developer helper: build a tuning variant of the CUDA library that only holds the headline
# instantiation (nx=128, 'std').  usage: tools/devbuild.sh <out.so> [extra nvcc flags...]
out=$1; shift
exec nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC \
  -DTRPL_DEV_HEADLINE_ONLY -Xptxas -v --threads 3 "$@" -o "$out" "$(dirname "$0")/../metrotrpl_b200/csrc/trpl_kernels.cu" \
  "$(dirname "$0")/../metrotrpl_b200/csrc/team_kernels.cu" "$(dirname "$0")/../metrotrpl_b200/csrc/team4_kernels.cu"
