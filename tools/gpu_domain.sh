#!/bin/bash
# One `gpurun --gpus N` call: decomposition parity on N GPUs + the N-GPU bench line (domain + replicas).
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_domain.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu_domain.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_domain.py -q -m gpu -s --timeout 800 2>&1 | tail -40 > gpurun_out/pytest_domain.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_domain.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --profile-out gpurun_out/bench_domain_profile_n$N.json \
  > gpurun_out/bench_domain_n$N.log 2> gpurun_out/bench_domain_n$N.err
echo "bench exit $?" >> gpurun_out/bench_domain_n$N.err
tail -15 gpurun_out/pytest_domain.log; cat gpurun_out/bench_domain_n$N.log; tail -5 gpurun_out/bench_domain_n$N.err
