#!/bin/bash
# Round 2: persistent Toeplitz kernel for the short-K cross-embed branches: parity (also with k = 16 on it), then interleaved A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_forward.py -q -m gpu -x -s --timeout 500 -k "toeplitz or golden or 1deg or full_grid" 2>&1 | grep -E "toeplitz|passed|failed|rror|rel" | tail -16 > gpurun_out/pytest_toep.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_toep.log
cat gpurun_out/pytest_toep.log
WXF_TOEP_PERSISTENT_MAXK=2048 timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_forward.py -q -m gpu -x -s --timeout 500 -k "toeplitz or 1deg or full_grid" 2>&1 | grep -E "toeplitz|passed|failed|rror|rel" | tail -12 > gpurun_out/pytest_toep2048.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_toep2048.log
cat gpurun_out/pytest_toep2048.log
timeout 500 python tools/ab_bench.py --a WXF_TOEP_PERSISTENT_MAXK=0 --b WXF_TOEP_PERSISTENT_MAXK=1024 --rounds 2 --steps 5 > gpurun_out/ab_toep.log 2>&1
cut -c1-150 gpurun_out/ab_toep.log | grep -v "^  [a-df-z]"
timeout 300 python tools/ab_bench.py --a WXF_TOEP_PERSISTENT_MAXK=1024 --b WXF_TOEP_PERSISTENT_MAXK=2048 --rounds 1 --steps 5 > gpurun_out/ab_toep2.log 2>&1
cut -c1-150 gpurun_out/ab_toep2.log | grep -v "^  [a-df-z]"
