#!/bin/bash
# Round 2, N = 2: the latitude-band decomposition of FuXi (fuxi_domain.py) on hardware: parity tests at n = 1 and 2, the
# WXFormer decomposition again (GroupNorm sums now put before the wait), then one bench line each.  Run with `gpurun --gpus 2`.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_domain.py -q -m gpu --timeout 500 -x -k "fuxi_band and (gpus[1] or gpus[2]) or n_gpus[2]" --durations=5 2>&1 | tail -25 > gpurun_out/pytest_fuxi_domain.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_fuxi_domain.log
run_bench() {  # $1 = tag, rest = args
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 2 --steps 10 --warmup 3 "$@" --profile-out gpurun_out/bench_n2_${tag}_profile.json \
      > gpurun_out/bench_n2_$tag.log 2> gpurun_out/bench_n2_$tag.err
  echo "bench $tag exit $?" >> gpurun_out/bench_n2_$tag.err
}
run_bench fuxi --workload fuxi_6h_025deg --no-cpu-baseline
run_bench wxf --no-cpu-baseline
tail -20 gpurun_out/pytest_fuxi_domain.log
for t in fuxi wxf; do cut -c1-600 gpurun_out/bench_n2_$t.log; tail -3 gpurun_out/bench_n2_$t.err; done
