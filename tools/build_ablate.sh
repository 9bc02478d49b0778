#!/bin/bash
# Diagnosis builds of the library with parts of the specialised GEMM epilogue compiled out (-DWXF_ABLATE=n; results are
# WRONG by construction): tools/ablate/lib_a{n}.so, timed by tools/gemm_shape_bench.py to see which part bounds a launch.
set -euo pipefail
cd "$(dirname "$0")/../miles_credit_b200/csrc"
bash build.sh
mkdir -p ../../tools/ablate
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include -DWXF_ABLATE=$n -c wxf_gemm_tc.cu -o /tmp/wxf_gemm_tc_a$n.o
  objs=$(ls wxf_*.o | grep -v wxf_gemm_tc.o)
  nvcc -shared -o ../../tools/ablate/lib_a$n.so $objs /tmp/wxf_gemm_tc_a$n.o -lcudart_static -lpthread -ldl -lrt
done
ls -la ../../tools/ablate
