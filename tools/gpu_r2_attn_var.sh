#!/bin/bash
# Round 2: window attention with P written over its own S half (no deferred lo words, 161 -> 128 registers) and two
# experiment switches (WXF_ATTN_VAR bit 0: rolled loop over the halves, bit 1: row maximum over every loaded column).
mkdir -p gpurun_out
for v in 0 1 4 5; do
  echo "=== WXF_ATTN_VAR=$v"
  WXF_ATTN_VAR=$v timeout 200 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu --timeout 100 -x 2>&1 | tail -2
  WXF_ATTN_VAR=$v timeout 120 python tools/attn_time.py
done 2>&1 | tee gpurun_out/attn_var_times.log
echo "=== skip previous"

