#!/bin/bash
# bench line with the per-launch table, then source-level ncu captures given as "name:regex:skip" arguments
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
cut -c1-300 gpurun_out/bench.log
bash tools/gpu_r2_src.sh "$@"
