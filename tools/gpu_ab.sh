#!/bin/bash
# One gpurun call: GPU parity tests + bench, then the same bench with one environment switch flipped (A/B).
#   AB_VAR=WXF_TC_CONCAT AB_VAL=0 bash tools/gpu_ab.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest ${TESTS:-tests} -q -m gpu -x --timeout 600 --durations=8 -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
if [ -n "${AB_VAR:-}" ]; then
  env ${AB_VAR}=${AB_VAL} timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile_b.json > gpurun_out/bench_b.log 2>> gpurun_out/bench.err
fi
grep -E "rel-max|passed|failed|FAILED|rror|exit" gpurun_out/pytest_gpu.log | tail -40
for f in gpurun_out/bench.log gpurun_out/bench_b.log; do
  [ -s $f ] && python - $f <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k:v['ms'] for k,v in d['kernel_families'].items()})
P
done
tail -3 gpurun_out/bench.err
