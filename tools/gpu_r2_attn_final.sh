#!/bin/bash
# Round 2: final window-attention kernel (multiply-high decodes, P over its own S half, row maximum over every loaded column,
# late epilogue for every tile shape, window-major even-L groups): parity tests, isolated launches, one bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_forward.py tests/test_gpu_ops.py -q -m gpu --timeout 200 -x 2>&1 | tail -4 | tee gpurun_out/attn_final_pytest.log
( echo "--- previous (round-2 kernel before this work)"; timeout 120 python tools/attn_time.py tools/ablate/lib_prev.so; echo "--- final"; timeout 120 python tools/attn_time.py ) 2>&1 | tee gpurun_out/attn_final_times.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/attn_final_bench_launches.json > gpurun_out/attn_final_bench.log 2> gpurun_out/attn_final_bench.err
echo "bench exit $?"; cut -c1-330 gpurun_out/attn_final_bench.log
