#!/bin/bash
# Round-2 final evidence on ONE GPU: whole -m gpu suite, smoke(), the two full bench lines (CPU reference baseline, eager-GPU
# competitor, full-grid parity), the ncu launch lists (duration + DRAM bytes per launch) of one WXFormer and one FuXi step,
# and an ncu --set full raw page of the first launches of a WXFormer step.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out /tmp/prof
timeout 900 python -m pytest tests -q -m gpu --timeout 600 --durations=5 2>&1 | tail -14 > gpurun_out/final_pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/final_bench_launches.json > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err
echo "bench exit $?" >> gpurun_out/final_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload fuxi_6h_025deg --profile-out gpurun_out/final_bench_fuxi_launches.json > gpurun_out/final_bench_fuxi.log 2> gpurun_out/final_bench_fuxi.err
echo "bench fuxi exit $?" >> gpurun_out/final_bench_fuxi.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_dram.csv \
    python tools/profile_step.py > gpurun_out/final_profile_step.log 2>&1
cp gpurun_out/step_tags.json gpurun_out/final_step_tags.json
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_dram_fuxi.csv \
    python tools/profile_step_fuxi.py > gpurun_out/final_profile_step_fuxi.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -c 16 -f -o /tmp/prof/final_full \
    python tools/profile_step.py > gpurun_out/final_profile_full.log 2>&1
ncu -i /tmp/prof/final_full.ncu-rep --page raw --csv > gpurun_out/final_ncu_full_first16_raw.csv 2>/dev/null
tail -4 gpurun_out/final_pytest_gpu.log; tail -3 gpurun_out/final_smoke.log
for f in final_bench final_bench_fuxi; do cut -c1-260 gpurun_out/$f.log; tail -1 gpurun_out/$f.err; done
ls -la gpurun_out/final_*
