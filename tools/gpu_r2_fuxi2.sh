#!/bin/bash
# FuXi: kernel parity tests, then the bench line with the per-launch table
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fuxi.py -q -m gpu -x --timeout 500 2>&1 | tail -5 > gpurun_out/pytest_fuxi2.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_fuxi2.log; cat gpurun_out/pytest_fuxi2.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload fuxi_6h_025deg --profile-out gpurun_out/bench_fuxi_profile.json > gpurun_out/bench_fuxi.log 2> gpurun_out/bench_fuxi.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_fuxi.log').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k:round(v['ms'],3) for k,v in d['kernel_families'].items()})
P
tail -2 gpurun_out/bench_fuxi.err
