#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu -x --timeout 300 -s 2>&1 | tail -25 > gpurun_out/pytest_attn.log
cat gpurun_out/pytest_attn.log
for cfg in "WXF_ATTN_V2=0" "WXF_ATTN_V2=1"; do
  echo "== $cfg"; env $cfg timeout 120 python tools/attn_time.py 2>&1 | tail -8
done > gpurun_out/attn_exp.log 2>&1
cat gpurun_out/attn_exp.log
