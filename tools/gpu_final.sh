#!/bin/bash
# Round-end evidence in ONE bounded gpurun call (1 GPU): parity tests, smoke, the bench line, the ncu launch list of one
# forecast step and a --set full capture of the dominant kernels (exported to CSV on the box).
mkdir -p gpurun_out /tmp/prof
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 150 python -m pytest tests -q -m gpu --timeout 120 --durations=5 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 120 python bench.py --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list exit $?" >> gpurun_out/profile_step.log
# full sets: stage-0 block (pad, toeplitz x4, LN, qkv, attention, out_proj, LN, ff1, ff2) + the un-pad kernel
timeout 120 ncu --set full --clock-control none --profile-from-start off -c 13 -o /tmp/prof/full_s0 \
    python tools/profile_step.py > gpurun_out/profile_full.log 2>&1
echo "full exit $?" >> gpurun_out/profile_full.log
ncu -i /tmp/prof/full_s0.ncu-rep --page raw --csv > gpurun_out/prof_full_s0_raw.csv 2>/dev/null
timeout 60 ncu --set full --clock-control none --profile-from-start off -k regex:"unpad_resize" -c 1 -o /tmp/prof/full_unpad \
    python tools/profile_step.py >> gpurun_out/profile_full.log 2>&1
ncu -i /tmp/prof/full_unpad.ncu-rep --page raw --csv > gpurun_out/prof_full_unpad_raw.csv 2>/dev/null
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.log | cut -c1-700; tail -2 gpurun_out/bench.err
tail -2 gpurun_out/profile_step.log; tail -2 gpurun_out/profile_full.log; ls -la gpurun_out | tail -12
