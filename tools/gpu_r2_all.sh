#!/bin/bash
# Whole single-GPU test suite, then the headline bench with the per-launch table (optionally an A/B given as $1 / $2).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 --durations=6 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['clocks'], {k:v['ms'] for k,v in d['kernel_families'].items()})
P
tail -2 gpurun_out/bench.err
