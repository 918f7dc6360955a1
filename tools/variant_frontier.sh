#!/bin/bash
# Run tools/frontier_sweep.py for the shipped library and every variant under gpurun_variants/.
cd "$(dirname "$0")/.."
for lib in reveal_graph_embedding_b200/libarcte_cuda.so gpurun_variants/*.so; do
  echo "== $lib"
  ARCTE_CUDA_LIB="$PWD/$lib" timeout 600 python tools/frontier_sweep.py "$1" "$2" 2>&1
done
