#!/bin/bash
# Round 2: the FuXi CUDA path on hardware (1 GPU): kernel + model parity, bench lines, and the conv phase-order change.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fuxi.py -q -m gpu --timeout 300 --durations=5 -x 2>&1 | tail -30 > gpurun_out/pytest_fuxi.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_fuxi.log
timeout 300 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_forward.py -q -m gpu --timeout 200 -x -k "conv or golden or wide" 2>&1 | tail -8 > gpurun_out/pytest_conv.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_conv.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload fuxi_6h_025deg --steps 5 --warmup 3 --profile-out gpurun_out/bench_fuxi_profile.json \
    > gpurun_out/bench_fuxi.log 2> gpurun_out/bench_fuxi.err; echo "bench exit $?" >> gpurun_out/bench_fuxi.err
timeout 200 python bench.py --no-cpu-baseline --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -25 gpurun_out/pytest_fuxi.log; tail -4 gpurun_out/pytest_conv.log; tail -4 gpurun_out/smoke.log
cut -c1-1500 gpurun_out/bench_fuxi.log; tail -5 gpurun_out/bench_fuxi.err; cut -c1-300 gpurun_out/bench.log; tail -2 gpurun_out/bench.err
