#!/bin/sh
# engines x workloads (run on the GPU box): tools/engine_sweep_many.sh out.jsonl "fifo:32,dense:16,hash:16" politicsuk ba5000x5 ...
out=$1; cfg=$2; shift 2
for w in "$@"; do timeout 300 python tools/engine_sweep.py $w $cfg $out 2>&1 | grep "^{" | cut -c1-250; done
