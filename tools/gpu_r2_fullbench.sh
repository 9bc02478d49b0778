#!/bin/bash
# The two full bench lines at N = 1 (CPU reference baseline, eager-GPU competitor and full-grid parity included).
mkdir -p gpurun_out
timeout 700 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/bench_full_profile.json > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 700 python bench.py --steps 10 --warmup 3 --workload fuxi_6h_025deg --profile-out gpurun_out/bench_fuxi_full_profile.json > gpurun_out/bench_fuxi_full.log 2> gpurun_out/bench_fuxi_full.err
echo "bench fuxi exit $?" >> gpurun_out/bench_fuxi_full.err
for f in bench_full bench_fuxi_full; do python - $f <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/{sys.argv[1]}.log').read().strip().splitlines()[-1])
print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'parity', d.get('parity'), 'cpu', d.get('cpu_baseline'), 'eager', d.get('gpu_eager_baseline'), 'roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','traffic')})
P
tail -2 gpurun_out/$f.err; done
