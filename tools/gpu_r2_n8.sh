#!/bin/bash
# Round 2, 8 GPUs (charged 8x: every command has its own tight timeout): decomposition parity at n = 4 and 8 for both decoder
# variants (vs the single-GPU forward AND the CPU oracle), the 240-step rollout of BASELINE config #5 at n = 8, and the
# bench lines at N = 8 (peer-memory exchanges, then NCCL) and N = 4.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_domain.py -q -m gpu --timeout 400 -x -k "n_gpus[4] or n_gpus[8] or rollout_decomposed_vs_single_gpu[8]" --durations=5 2>&1 \
    | tail -30 > gpurun_out/pytest_domain_n8.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_domain_n8.log
run_bench() {  # $1 = N, $2 = comm mode, extra args
  WXF_DOMAIN_COMM=$2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $1 --steps 20 --warmup 5 $3 --profile-out gpurun_out/bench_n$1_$2_profile.json \
      > gpurun_out/bench_n$1_$2.log 2> gpurun_out/bench_n$1_$2.err
  echo "bench N=$1 $2 exit $?" >> gpurun_out/bench_n$1_$2.err
}
run_bench 8 peer ""
run_bench 8 nccl "--no-cpu-baseline"
run_bench 4 peer "--no-cpu-baseline"
tail -12 gpurun_out/pytest_domain_n8.log
for f in n8_peer n8_nccl n4_peer; do cut -c1-330 gpurun_out/bench_$f.log; tail -2 gpurun_out/bench_$f.err; done
