#!/bin/bash
# TMA row-throughput micro-benchmark (tools/tma_bench.cu, built here with
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../include -o tma_bench tma_bench.cu).
mkdir -p gpurun_out
timeout 60 tools/tma_bench > gpurun_out/tma_bench.log 2>&1
echo "exit $?" >> gpurun_out/tma_bench.log
cat gpurun_out/tma_bench.log
