#!/bin/bash
# Round 2, N = 4 re-check with the final kernels (strictly bounded: a 4-GPU minute costs 4 budget minutes): the WXFormer
# decomposition parity test at n = 4, then the driver-style bench line.  Run with `gpurun --gpus 4`.
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_domain.py -q -m gpu --timeout 90 -x -k "test_domain_decomposition_on_n_gpus and gpus[4]" 2>&1 | tail -6 > gpurun_out/recheck_pytest_domain_n4.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/recheck_pytest_domain_n4.log
timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/recheck_bench_n4.log 2> gpurun_out/recheck_bench_n4.err
echo "bench exit $?" >> gpurun_out/recheck_bench_n4.err
tail -4 gpurun_out/recheck_pytest_domain_n4.log
cut -c1-400 gpurun_out/recheck_bench_n4.log; tail -2 gpurun_out/recheck_bench_n4.err
