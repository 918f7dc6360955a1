#!/bin/bash
# A/B of library variants of the compact push kernel on the bench shape (kernel time, counters and checksums per
# line; the checksums must agree across variants).  usage: tools/ab_variants.sh <out.jsonl>
#   default build                       compact:48
#   gpurun_variants/libarcte_mb8.so     -DARCTE_COMPACT_MIN_BLOCKS=8 (32 registers)      compact:64,compact:56
#   gpurun_variants/libarcte_hintg.so   -DARCTE_HINT_GRAPH=1 (graph arrays L2 evict_last) compact:48
#   gpurun_variants/libarcte_hintg_mb8.so  both                                          compact:64
#   default build, ARCTE_CUDA_L2_PERSIST_MB=32                                           compact:48
cd "$(dirname "$0")/.."
out="${1:-gpurun_out/ab_variants.jsonl}"
run() {  # label, lib, configs, extra env
  echo "{\"variant\": \"$1\"}" >> "$out"
  env ARCTE_CUDA_LIB="$2" $4 timeout 40 python tools/engine_sweep.py youtube "$3" "$out" 2>&1 | grep -v "^set_graph" | cut -c1-200
}
D="$PWD/reveal_graph_embedding_b200/libarcte_cuda.so"
V="$PWD/gpurun_variants"
run default "$D" compact:48 ""
[ -f "$V/libarcte_mb8.so" ] && run mb8 "$V/libarcte_mb8.so" compact:64,compact:56 ""
[ -f "$V/libarcte_hintg.so" ] && run hintg "$V/libarcte_hintg.so" compact:48 ""
[ -f "$V/libarcte_hintg_mb8.so" ] && run hintg_mb8 "$V/libarcte_hintg_mb8.so" compact:64 ""
run l2persist32 "$D" compact:48 "ARCTE_CUDA_L2_PERSIST_MB=32"
