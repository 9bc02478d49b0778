#!/bin/bash
# One `gpurun --gpus N` call: decomposition parity on N GPUs + the N-GPU bench line (graph replay and eager launches).
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_domain.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpu_domain.txt 2>&1
if [ -z "${SKIP_TESTS:-}" ]; then
timeout 900 python -m pytest tests/test_gpu_domain.py -q -m gpu -s --timeout 800 2>&1 | tail -40 > gpurun_out/pytest_domain.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_domain.log
tail -15 gpurun_out/pytest_domain.log
fi
for G in ${GRAPHS:-1 0}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$G \
  bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --graph $G --profile-out gpurun_out/bench_domain_profile_n$N.json \
  > gpurun_out/bench_domain_n${N}_g$G.log 2> gpurun_out/bench_domain_n${N}_g$G.err
echo "bench exit $?" >> gpurun_out/bench_domain_n${N}_g$G.err
python - gpurun_out/bench_domain_n${N}_g$G.log <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['config'].get('launch'), 'launches', d['gpu_launches'], 'replicas', d.get('replicas',{}).get('value'))
    print({k:v['ms'] for k,v in d['kernel_families'].items()})
except Exception as e:
    print('no bench line', e)
P
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_domain_n${N}_g$G.err | tail -6
done
