"""Sweep the launch geometry of the frontier schedule (csrc/push_frontier.cu) on a bench workload.

    python tools/frontier_sweep.py <workload> "<permille>:<hthreads>:<hctas>:<lthreads>:<lctas>,..."
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import EPS, RHO, make_graph
from reveal_graph_embedding_b200.engine import Engine

workload = sys.argv[1]
geoms = [tuple(int(x) for x in g.split(":")) for g in sys.argv[2].split(",")] if len(sys.argv) > 2 else [(-1, 0, 0, 0, 0)]
A = make_graph(workload)
eng = Engine(0)
eng.set_graph(A)
if "--exact" in sys.argv:
    for rep in range(2):
        eng.extract(0, RHO, EPS)
    st = eng.stats()
    print(json.dumps({"workload": workload, "schedule": "fifo", "ms_push": round(st["ms_push"], 2),
                      "members": st["members"], "pushes": st["pushes"], "edges": st["edge_touches"]}), flush=True)
for g in geoms:
    eng.set_schedule("frontier", *g)
    for rep in range(2):
        t = time.time()
        eng.extract(0, RHO, EPS)
        wall = time.time() - t
    st = eng.stats()
    print(json.dumps({"workload": workload, "schedule": "frontier", "geometry": g, "slots": st["n_slots"],
                      "ms_push": round(st["ms_push"], 2), "wall_ms": round(wall * 1e3, 1),
                      "seeds_per_s": round(st["n_seeds_shard"] / (st["ms_push"] / 1e3)),
                      "alg_GBps": round(st["alg_bytes_push"] / st["ms_push"] / 1e6, 1),
                      "pushes": st["pushes"], "edges": st["edge_touches"], "rounds": st["rounds"],
                      "support": st["support"], "members": st["members"], "emitted": st["emitted"],
                      "max_frontier": st["max_queue"], "retries": st["retries"]}), flush=True)
