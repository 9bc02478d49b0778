#!/bin/bash
mkdir -p gpurun_out
for cfg in "WXF_ATTN_V2=0" "WXF_ATTN_V2=1" "WXF_ATTN_DEBUG=1" "WXF_ATTN_DEBUG=2" "WXF_ATTN_DEBUG=4" "WXF_ATTN_DEBUG=3" "WXF_ATTN_DEBUG=7"; do
  echo "== $cfg"; env $cfg timeout 120 python tools/attn_time.py 2>&1 | tail -8
done > gpurun_out/attn_exp.log 2>&1
cat gpurun_out/attn_exp.log
