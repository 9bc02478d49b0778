#!/bin/bash
# ncu evidence: (1) launch list with per-launch durations of one forecast step, (2) --set full on the first kernels of
# every family.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list exit $?" >> gpurun_out/profile_step.log
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"tc_persistent|tc_contract|window_attention_tc|layernorm_vec|gn_silu_vec|pad_to_pixel|unpad_resize" -c ${NCU_COUNT:-40} \
    -o gpurun_out/prof_full python tools/profile_step.py > gpurun_out/profile_full.log 2>&1
echo "full exit $?" >> gpurun_out/profile_full.log
ls -la gpurun_out/ | tail -8; tail -3 gpurun_out/profile_step.log gpurun_out/profile_full.log
