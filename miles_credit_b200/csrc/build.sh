#!/bin/bash
# Builds libwxformer_b200.so (sm_100a only) next to the sources.  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include ${WXF_NVCC_EXTRA:-}"
OBJS=""
PIDS=""
for f in wxf_*.cu; do
  o="${f%.cu}.o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ wxf_common.cuh -nt "$o" ] || [ wxf_tc_ptx.cuh -nt "$o" ] || [ wxf_tc_host.cuh -nt "$o" ] || [ wxf_fastdiv.h -nt "$o" ] || [ ../../include/wxformer_b200.h -nt "$o" ]; then
    echo "nvcc $f" >&2
    ( $NVCC $FLAGS -c "$f" -o "$o" || { rm -f "$o"; exit 1; } ) &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for p in $PIDS; do wait $p || { echo "nvcc failed" >&2; exit 1; }; done
$NVCC -shared -o libwxformer_b200.so $OBJS -lcudart_static -lpthread -ldl -lrt
echo "built $(pwd)/libwxformer_b200.so" >&2
