#!/bin/bash
# Last bounded GPU call of round 1: default-path parity + bench, then the two gated candidates
# (WXF_GEMM_RESIDENT_W=1, WXF_PDL=1): parity tests and one short bench line each.  Results are appended as they come.
mkdir -p gpurun_out
L=gpurun_out/lastshot.log
: > $L
T="tests/test_gpu_gemm_tc.py tests/test_gpu_forward.py"
K="gemm or golden or graph"
run_tests() { echo "== tests $1" >> $L; env $1 timeout 30 python -m pytest $T -q -m gpu -x -k "$K" --timeout 25 2>&1 | tail -3 >> $L; }
run_bench() { echo "== bench $1" >> $L; env $1 timeout 30 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>$L | python -c "
import json,sys
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: continue
    print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), {k:v['ms'] for k,v in d['kernel_families'].items()})
" >> $L; }
run_tests "WXF_NONE=0"
run_bench "WXF_NONE=0"
run_tests "WXF_GEMM_RESIDENT_W=1"
run_bench "WXF_GEMM_RESIDENT_W=1"
run_tests "WXF_PDL=1"
run_bench "WXF_PDL=1"
echo "== bench both" >> $L
env WXF_PDL=1 WXF_GEMM_RESIDENT_W=1 timeout 30 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>$L | python -c "
import json,sys
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: continue
    print('ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))
" >> $L
cat $L
