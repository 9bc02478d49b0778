#!/bin/bash
# Round 2, first GPU call (1 GPU): parity suite, the new bench line (reference + eager-GPU baselines), ncu launch list with
# DRAM bytes, then the never-run gated candidates (parity first, interleaved A/B if green).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -q -m gpu --timeout 400 --durations=8 -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/launches_dram.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list exit $?" >> gpurun_out/profile_step.log
# gated candidates, parity first
WXF_GEMM_CLUSTER=1 timeout 150 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_forward.py -q -m gpu -x --timeout 100 \
    -k "gemm or golden or 1deg" 2>&1 | tail -15 > gpurun_out/cand_cluster.log; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/cand_cluster.log
WXF_FF_FUSED=1 timeout 150 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_forward.py -q -m gpu -x --timeout 100 \
    -k "ff_fused or golden or 1deg" 2>&1 | tail -15 > gpurun_out/cand_ff_fused.log; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/cand_ff_fused.log
timeout 200 python tools/ab_bench.py --b WXF_PDL=1 --rounds 2 > gpurun_out/ab_pdl.log 2>&1
timeout 200 python tools/ab_bench.py --b WXF_ATTN_SIMT_SMALL=1 --rounds 2 > gpurun_out/ab_attn_small.log 2>&1
if grep -q "exit 0" gpurun_out/cand_cluster.log; then timeout 200 python tools/ab_bench.py --b WXF_GEMM_CLUSTER=1 --rounds 2 > gpurun_out/ab_cluster.log 2>&1; fi
if grep -q "exit 0" gpurun_out/cand_ff_fused.log; then timeout 200 python tools/ab_bench.py --b WXF_FF_FUSED=1 --rounds 2 > gpurun_out/ab_ff_fused.log 2>&1; fi
tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
tail -2 gpurun_out/profile_step.log; tail -4 gpurun_out/cand_cluster.log gpurun_out/cand_ff_fused.log; tail -3 gpurun_out/ab_*.log
