#!/bin/bash
# Round 2, N = 2: the NVLink peer-memory exchanges of the domain decomposition (WXF_DOMAIN_COMM=peer, default) against the
# NCCL path (WXF_DOMAIN_COMM=nccl): parity tests, then one bench line each.  Run with `gpurun --gpus 2`.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_domain.py -q -m gpu --timeout 300 -x -k "single_rank or n_gpus[2]" 2>&1 | tail -25 > gpurun_out/pytest_domain_peer.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_domain_peer.log
for mode in peer nccl; do
  WXF_DOMAIN_COMM=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_n2_${mode}_profile.json \
      > gpurun_out/bench_n2_$mode.log 2> gpurun_out/bench_n2_$mode.err
  echo "bench $mode exit $?" >> gpurun_out/bench_n2_$mode.err
done
tail -20 gpurun_out/pytest_domain_peer.log
for mode in peer nccl; do cut -c1-400 gpurun_out/bench_n2_$mode.log; tail -3 gpurun_out/bench_n2_$mode.err; done
