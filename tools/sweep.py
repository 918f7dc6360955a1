"""Sweep walk states per SM for one workload (run on the GPU box)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_graph, RHO, EPS
from reveal_graph_embedding_b200.engine import Engine

workload = sys.argv[1]
wps_list = [int(x) for x in sys.argv[2].split(",")]
A = make_graph(workload)
eng = Engine(0)
t = time.time(); eng.set_graph(A); print("set_graph %.3fs" % (time.time() - t), flush=True)
for wps in wps_list:
    eng.configure(warps_per_sm=wps)
    for rep in range(2):
        t = time.time(); eng.extract(0, RHO, EPS); dt = time.time() - t
    st = eng.stats()
    print(json.dumps({"workload": workload, "wps": wps, "slots": st["n_slots"], "ms_push": round(st["ms_push"], 3),
                      "extract_wall_ms": round(dt * 1e3, 1), "util": round(st["slot_utilisation"], 3),
                      "GBps_alg": round(st["alg_bytes_push"] / st["ms_push"] / 1e6, 1),
                      "pushes": st["pushes"], "edges": st["edge_touches"], "support": st["support"],
                      "members": st["members"], "retries": st["retries"], "maxq": st["max_queue"]}), flush=True)
t = time.time(); eng.assemble(); print("assemble %.3fs nnz=%d" % (time.time() - t, eng.out_nnz), eng.stats()["ms_assemble"])
