#!/bin/bash
# Round 2: window attention with multiply-high tile / window decodes (no division sequences on the softmax warps' critical
# path) + window-major even-L groups: parity tests, isolated launches previous library vs this one, then one bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_forward.py tests/test_gpu_ops.py -q -m gpu --timeout 200 -x 2>&1 | tail -8 > gpurun_out/attn_fastdiv_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/attn_fastdiv_pytest.log
cat gpurun_out/attn_fastdiv_pytest.log
for rep in 1 2; do
  echo "--- previous"; timeout 120 python tools/attn_time.py tools/ablate/lib_prev.so
  echo "--- new"; timeout 120 python tools/attn_time.py
done 2>&1 | tee gpurun_out/attn_fastdiv_times.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/attn_fastdiv_bench_launches.json > gpurun_out/attn_fastdiv_bench.log 2> gpurun_out/attn_fastdiv_bench.err
echo "bench exit $?"; cut -c1-330 gpurun_out/attn_fastdiv_bench.log
