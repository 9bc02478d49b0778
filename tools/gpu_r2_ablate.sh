#!/bin/bash
# Which part of a GEMM launch bounds it: the shapes of the forecast step timed alone, default library vs diagnosis builds.
mkdir -p gpurun_out
S="320000,512,128,gelu 320000,384,128,planes 320000,128,128,red 320000,128,512,red 80000,1024,256,gelu 20000,2048,512,gelu 20000,512,2048,red"
{
python tools/gemm_shape_bench.py $S
WXF_GEMM_RESIDENT_W=1 python tools/gemm_shape_bench.py 320000,512,128,gelu 320000,384,128,planes 320000,128,128,red | sed 's/default /resident_w/'
WXF_GEMM_EW16_MAXK=256 python tools/gemm_shape_bench.py 80000,1024,256,gelu | sed 's/default /ew16@256 /'
for n in 1 2 4; do python tools/gemm_shape_bench.py --lib tools/ablate/lib_a$n.so $S; done
} 2>&1 | tee gpurun_out/ablate.log
