#!/bin/bash
# One GPU-box call that validates the tree and collects the round's last measurements under a hard deadline
# (seconds from start, $1): every step gets what is left.  Outputs under gpurun_out/final_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s); LIMIT=${1:-400}
left() { echo $(( LIMIT - ( $(date +%s) - T0 ) )); }
step() {  # name, max seconds, command...
  local name=$1 max=$2; shift 2
  local l=$(left); [ $l -lt 8 ] && { echo "== $name: skipped (deadline)"; return; }
  [ $max -lt $l ] && l=$max
  local t=$(date +%s)
  timeout $l "$@"
  echo "== $name: rc=$? in $(( $(date +%s) - t )) s ($(left) s left)"
}
step pytest 200 python -m pytest tests -m gpu -x -q > gpurun_out/final_gputests.log 2>&1
tail -3 gpurun_out/final_gputests.log
step smoke 40 python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1
tail -1 gpurun_out/final_smoke.log
step bench 150 sh -c 'python bench.py --no-python-ref > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err'
step ab 120 tools/ab_variants.sh gpurun_out/final_ab_variants.jsonl
step snow 90 sh -c 'python tools/snow_standin.py gpu > gpurun_out/final_snow_gpu.json 2> gpurun_out/final_snow_gpu.err'
step flickr 60 sh -c 'python bench.py --workload flickr --no-cpu --no-cold > gpurun_out/final_bench_flickr_1gpu.json 2> gpurun_out/final_bench_flickr.err'
step latency 30 sh -c 'python tools/small_latency.py politicsuk 20 > gpurun_out/final_small_latency.txt 2>&1'
echo "total $(( $(date +%s) - T0 )) s"
