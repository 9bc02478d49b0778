#!/bin/bash
# Source-level ncu capture of chosen launches of one FuXi step: $@ = "name:regex:skip" ...
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; regex=${rest%%:*}; skip=${rest##*:}
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k regex:"$regex" -s $skip -c 1 -f -o gpurun_out/src_$name python tools/profile_step_fuxi.py > gpurun_out/src_$name.log 2>&1
  echo "$name exit $?" | tee -a gpurun_out/src_$name.log
done
ls -la gpurun_out/*.ncu-rep
