#!/bin/bash
# ncu --set full with source for ONE launch selected by kernel regex ($1), skip count ($2) and output name ($3)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$1" -s ${2:-0} -c 1 -o gpurun_out/${3:-prof_one} python tools/profile_step.py > gpurun_out/profile_one.log 2>&1
echo "exit $?" >> gpurun_out/profile_one.log; tail -2 gpurun_out/profile_one.log; ls -la gpurun_out/*.ncu-rep
