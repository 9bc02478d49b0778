#!/bin/bash
# Round 2: the specialised GEMM epilogues on hardware: parity tests, then interleaved A/B against the generic epilogue
# (WXF_GEMM_EPI=0) and one arm with parked mbarrier waits.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_forward.py -q -m gpu -x --timeout 500 -k "not full_grid and not 240" 2>&1 | tail -15 > gpurun_out/pytest_epi.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_epi.log
tail -6 gpurun_out/pytest_epi.log
timeout 500 python tools/ab_bench.py --a WXF_GEMM_EPI=0 --b WXF_GEMM_EPI=1 --rounds 2 --steps 5 > gpurun_out/ab_epi.log 2>&1
cat gpurun_out/ab_epi.log | cut -c1-200
timeout 300 python tools/ab_bench.py --a WXF_MBAR_PARK_NS=1000 --b WXF_MBAR_PARK_NS=20000 --rounds 1 --steps 5 > gpurun_out/ab_park.log 2>&1
cat gpurun_out/ab_park.log | cut -c1-200
