#!/bin/bash
# Round 2: uniform-tile softmax (straight-line, NH active halves) and window-major packing of the even-L dilated groups:
# parity tests, then the isolated attention launches of the headline workload, previous library vs this one.  One GPU.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_forward.py -q -m gpu --timeout 200 -x 2>&1 | tail -8 > gpurun_out/attn_uniform_pytest.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/attn_uniform_pytest.log
cat gpurun_out/attn_uniform_pytest.log
for rep in 1 2; do
  echo "--- previous"; timeout 120 python tools/attn_time.py tools/ablate/lib_prev.so
  echo "--- new"; timeout 120 python tools/attn_time.py
  echo "--- new, uniform path off"; WXF_ATTN_UNIFORM=0 timeout 120 python tools/attn_time.py
done 2>&1 | tee gpurun_out/attn_uniform_times.log
