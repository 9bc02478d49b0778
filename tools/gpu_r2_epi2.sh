#!/bin/bash
# Round 2: resident-W small-K GEMM with the specialised epilogues, and the K threshold of the 16-epilogue-warp shape.
mkdir -p gpurun_out
WXF_GEMM_RESIDENT_W=1 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_forward.py -q -m gpu -x --timeout 500 -k "not full_grid and not 240" 2>&1 | tail -8 > gpurun_out/pytest_epi2.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_epi2.log
tail -4 gpurun_out/pytest_epi2.log
timeout 500 python tools/ab_bench.py --b WXF_GEMM_RESIDENT_W=1 --rounds 2 --steps 5 > gpurun_out/ab_rw.log 2>&1
cut -c1-120 gpurun_out/ab_rw.log
timeout 500 python tools/ab_bench.py --a WXF_GEMM_EW16_MAXK=128 --b WXF_GEMM_EW16_MAXK=128 WXF_GEMM_RESIDENT_W=1 --rounds 2 --steps 5 > gpurun_out/ab_maxk.log 2>&1
cut -c1-140 gpurun_out/ab_maxk.log
