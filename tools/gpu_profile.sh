#!/bin/bash
# ncu evidence: (1) launch list with per-launch durations of one forecast step, (2) --set full on the first kernels of
# every family, exported to CSV on the box (the .ncu-rep of 40 kernels exceeds the 64 MiB return limit), (3) small
# .ncu-rep files with source for the attention and GEMM kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out /tmp/prof
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list exit $?" >> gpurun_out/profile_step.log
timeout 1200 ncu --set full --clock-control none --profile-from-start off \
    -k regex:"tc_persistent|tc_contract|window_attention_tc|layernorm_vec|gn_silu_vec|pad_to_pixel|unpad_resize" -c ${NCU_COUNT:-40} \
    -o /tmp/prof/prof_full python tools/profile_step.py > gpurun_out/profile_full.log 2>&1
echo "full exit $?" >> gpurun_out/profile_full.log
ncu -i /tmp/prof/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"window_attention_tc" -c 2 -o gpurun_out/prof_attention python tools/profile_step.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"tc_persistent" -s 4 -c 2 -o gpurun_out/prof_gemm python tools/profile_step.py > /dev/null 2>&1
ls -la gpurun_out/ | tail -9; cat gpurun_out/profile_step.log | tail -2
