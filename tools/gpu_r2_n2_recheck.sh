#!/bin/bash
# Round 2, N = 2 re-check after the last GroupNorm / attention kernel changes: decomposition parity tests at n = 2 (WXFormer both
# decoder variants + FuXi bands), then the driver-style WXFormer bench line at N = 2.  Run with `gpurun --gpus 2`.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_domain.py -q -m gpu --timeout 300 -x -k "gpus[2]" --durations=3 2>&1 | tail -15 > gpurun_out/recheck_pytest_domain_n2.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/recheck_pytest_domain_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/recheck_bench_n2.log 2> gpurun_out/recheck_bench_n2.err
echo "bench exit $?" >> gpurun_out/recheck_bench_n2.err
tail -8 gpurun_out/recheck_pytest_domain_n2.log
cut -c1-900 gpurun_out/recheck_bench_n2.log; tail -3 gpurun_out/recheck_bench_n2.err
