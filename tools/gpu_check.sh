#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python -m pytest tests/test_gpu_attention_tc.py -q -m gpu -s --timeout 200 2>&1 | tail -40 > gpurun_out/pytest_attn_tc.log
if [ "${PIPESTATUS[0]}" != "0" ]; then echo "attention_tc FAILED: rest of the run uses WXF_ATTN_TC=0" >> gpurun_out/pytest_attn_tc.log; export WXF_ATTN_TC=0; fi
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_conv_tc.py -q -m gpu -s --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_gemm_tc.log
echo "pytest gemm_tc exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gemm_tc.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_gemm_tc.py --deselect tests/test_gpu_conv_tc.py --deselect tests/test_gpu_attention_tc.py -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --profile-out gpurun_out/bench_profile.json ${BENCH_EXTRA:-} > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
if [ -n "${AB:-}" ]; then
  WXF_TC_PERSISTENT=0 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_conv_tc.py -q -m gpu --timeout 300 2>&1 | tail -5 > gpurun_out/pytest_tc_nonpersistent.log
  WXF_TC_PERSISTENT=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile_np.json > gpurun_out/bench_np.log 2>> gpurun_out/bench.err
  tail -3 gpurun_out/pytest_tc_nonpersistent.log; python -c "import json;d=json.loads(open('gpurun_out/bench_np.log').read());print('non-persistent ms/step',d['ms_per_step'],{k:v['ms'] for k,v in d['kernel_families'].items()})"
fi
grep -E 'rel-max|passed|failed|FAILED|rror' gpurun_out/pytest_attn_tc.log | head -30; tail -12 gpurun_out/pytest_gemm_tc.log; tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
