"""Engines x walks-per-SM on one workload (run on the GPU box).

    python tools/engine_sweep.py youtube fifo:32,dense:16,hash:8,hash:16,hash:24 [out.jsonl]

Each configuration is extracted twice (the first run sizes pools and page-faults them in), the second is
reported: push-kernel ms (CUDA events inside the library), algorithmic GB/s, counters and a checksum of
the segment table, which must agree across engines.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import EPS, RHO, make_graph  # noqa: E402
from reveal_graph_embedding_b200.engine import Engine  # noqa: E402

workload = sys.argv[1]
configs = [c.split(":") for c in sys.argv[2].split(",")]
out = open(sys.argv[3], "a") if len(sys.argv) > 3 else None
A = make_graph(workload)
eng = Engine(0)
t = time.time()
eng.set_graph(A)
print("set_graph %.3fs n=%d nnz=%d" % (time.time() - t, A.shape[0], A.nnz), flush=True)
for cfg in configs:
    name, wps = cfg[0], cfg[1]
    mem_pct = int(cfg[2]) if len(cfg) > 2 else 0       # engine:warps_per_sm[:mem_percent]
    eng.set_engine(name)
    eng.configure(warps_per_sm=int(wps), mem_percent=mem_pct)
    for rep in range(2):
        t = time.time()
        eng.extract(0, RHO, EPS)
        dt = time.time() - t
    st = eng.stats()
    seg_seed, seg_cnt, seg_off, mem = eng.segments()
    line = {"workload": workload, "engine": name, "wps": int(wps), "mem_percent": mem_pct, "slots": st["n_slots"],
            "ms_push": round(st["ms_push"], 3), "extract_wall_ms": round(dt * 1e3, 1),
            "util": round(st["slot_utilisation"], 3),
            "GBps_alg": round(st["alg_bytes_push"] / st["ms_push"] / 1e6, 1),
            "frac_of_6550.7": round(st["alg_bytes_push"] / st["ms_push"] / 1e6 / 6550.7, 4),
            "pushes": st["pushes"], "edges": st["edge_touches"], "support": st["support"], "touched": st["touched"],
            "members": st["members"], "retries": st["retries"], "maxq": st["max_queue"],
            "seg_checksum": int((seg_cnt.astype(np.int64) * (seg_seed.astype(np.int64) + 1)).sum()),
            "member_sum": int(mem.astype(np.int64).sum())}
    print(json.dumps(line), flush=True)
    if out:
        out.write(json.dumps(line) + "\n")
        out.flush()
