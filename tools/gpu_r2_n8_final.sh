#!/bin/bash
# Round 2, final 8-GPU evidence (charged 8x: tight timeouts): WXFormer decomposition parity at n = 8 with the current kernels,
# FuXi band decomposition at n = 8 (full 0.25 deg grid) and n = 4, then the bench lines: WXFormer N = 8, 4 and FuXi N = 8.
mkdir -p gpurun_out
timeout 330 python -m pytest tests/test_gpu_domain.py -q -m gpu --timeout 300 -x -k "n_gpus[8] and not 240" --durations=5 2>&1 \
    | tail -14 > gpurun_out/pytest_domain_n8_final.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_domain_n8_final.log
run_bench() {  # $1 = N, $2 = tag, extra args
  n=$1; tag=$2; shift 2
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline "$@" --profile-out gpurun_out/bench_n${n}_${tag}_profile.json \
      > gpurun_out/bench_n${n}_$tag.log 2> gpurun_out/bench_n${n}_$tag.err
  echo "bench N=$n $tag exit $?" >> gpurun_out/bench_n${n}_$tag.err
}
run_bench 8 wxf
run_bench 8 fuxi --workload fuxi_6h_025deg
run_bench 4 wxf
tail -9 gpurun_out/pytest_domain_n8_final.log
for f in n8_wxf n8_fuxi n4_wxf; do cut -c1-330 gpurun_out/bench_$f.log; tail -2 gpurun_out/bench_$f.err; done
