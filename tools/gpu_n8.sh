#!/bin/bash
# One `gpurun --gpus N` call, strictly bounded: the decomposed bench line at N GPUs (graph replay).
# ENVS="WXF_PDL=1" adds environment switches (e.g. programmatic dependent launch, which should matter most here: the
# rank's kernels last ~15 us at N=8).
N=${N:-8}
mkdir -p gpurun_out
SECONDS=0
env ${ENVS:-WXF_NONE=0} timeout ${TMO:-80} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --graph ${G:-1} --profile-out gpurun_out/bench_domain_profile_n$N.json \
  > gpurun_out/bench_domain_n${N}.log 2> gpurun_out/bench_domain_n${N}.err
echo "bench exit $? after ${SECONDS}s" >> gpurun_out/bench_domain_n${N}.err
python - gpurun_out/bench_domain_n${N}.log <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('ms/step', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), d['config'].get('launch'), 'launches', d['gpu_launches'], 'replicas', d.get('replicas',{}).get('value'))
    print({k:v['ms'] for k,v in d['kernel_families'].items()})
except Exception as e:
    print('no bench line', e)
P
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_domain_n${N}.err | tail -8 | cut -c1-400
