#!/bin/bash
# Round 2 (1 GPU): fused pre/post-block kernels, FuXi after the un-patchify change + its reference parity, and the proxy for
# the per-rank work at N = 8 (the same architecture on the 181x360 grid = 12 % of the pixels): default vs WXF_PDL=1.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fuxi.py -q -m gpu --timeout 200 -x 2>&1 | tail -15 > gpurun_out/pytest_pipeline.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_pipeline.log
timeout 600 python bench.py --workload fuxi_6h_025deg --steps 5 --warmup 3 --profile-out gpurun_out/bench_fuxi_profile.json \
    > gpurun_out/bench_fuxi.log 2> gpurun_out/bench_fuxi.err; echo "bench exit $?" >> gpurun_out/bench_fuxi.err
timeout 300 python tools/ab_bench.py --workload wxformer_6h_1deg --b WXF_PDL=1 --rounds 2 --steps 20 > gpurun_out/ab_pdl_1deg.log 2>&1
tail -8 gpurun_out/pytest_pipeline.log; cut -c1-300 gpurun_out/bench_fuxi.log; tail -3 gpurun_out/bench_fuxi.err; cat gpurun_out/ab_pdl_1deg.log | cut -c1-200
