#!/bin/bash
# Second A/B: the two variants that won on the YouTube shape, on the Flickr shape and once more on YouTube, then the
# compact-engine tests and the full-size tests against each variant library.  usage: tools/ab_variants2.sh <deadline s>
cd "$(dirname "$0")/.."
T0=$(date +%s); LIMIT=${1:-150}
left() { echo $(( LIMIT - ( $(date +%s) - T0 ) )); }
out=gpurun_out/final_ab_variants2.jsonl
D="$PWD/reveal_graph_embedding_b200/libarcte_cuda.so"; V="$PWD/gpurun_variants"
run() {  # label lib workload configs
  local l=$(left); [ $l -lt 12 ] && { echo "skip $1 $3"; return; }
  [ $l -gt 40 ] && l=40
  echo "{\"variant\": \"$1\", \"workload\": \"$3\"}" >> "$out"
  ARCTE_CUDA_LIB="$2" timeout $l python tools/engine_sweep.py "$3" "$4" "$out" 2>&1 | grep -v "^set_graph" | cut -c1-150
}
run default "$D" flickr compact:48
run mb8 "$V/libarcte_mb8.so" flickr compact:64
run hintg "$V/libarcte_hintg.so" flickr compact:48
run mb8 "$V/libarcte_mb8.so" youtube compact:64
run default "$D" youtube compact:48
run hintg "$V/libarcte_hintg.so" youtube compact:48
for v in mb8 hintg; do
  l=$(left); [ $l -lt 20 ] && { echo "skip tests $v"; continue; }
  ARCTE_CUDA_LIB="$V/libarcte_$v.so" timeout $l python -m pytest tests/test_gpu_engines.py tests/test_gpu_fullsize.py -m gpu -x -q -k "compact or shape" > gpurun_out/final_tests_$v.log 2>&1
  echo "== tests $v rc=$? ($(left) s left)"; tail -1 gpurun_out/final_tests_$v.log
done
echo "total $(( $(date +%s) - T0 )) s"
