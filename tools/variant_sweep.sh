#!/bin/bash
# Run tools/sweep.py for every library variant under gpurun_variants/ (kernel experiments).
# usage: tools/variant_sweep.sh <workload> <warps-per-sm list>
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in reveal_graph_embedding_b200/libarcte_cuda.so gpurun_variants/*.so; do
  echo "== $lib"
  ARCTE_CUDA_LIB="$PWD/$lib" timeout 300 python tools/sweep.py "$1" "$2" 2>&1 | grep -v "^set_graph\|^assemble"
done
