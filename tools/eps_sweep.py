"""BASELINE.json config 5: epsilon sweep on an R-MAT graph (frontier size vs achieved rate)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bench import make_graph, RHO
from reveal_graph_embedding_b200.engine import Engine

workload = sys.argv[1] if len(sys.argv) > 1 else "rmat20"
eps_list = [float(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1e-3,1e-4,1e-5,1e-6").split(",")]
max_seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t = time.time(); A = make_graph(workload); print("graph %s n=%d nnz=%d (%.1fs)" % (workload, A.shape[0], A.nnz, time.time() - t), flush=True)
eng = Engine(0)
t = time.time(); eng.set_graph(A); print("set_graph %.3fs, seeds=%d" % (time.time() - t, eng.seeds().size), flush=True)
if max_seeds:
    seeds = eng.seeds()
    idx = np.unique(np.linspace(0, seeds.size - 1, max_seeds).astype(np.int64))  # degree-stratified sample
    eng.set_seeds(seeds[idx])
for eps in eps_list:
    t = time.time(); eng.extract(0, RHO, eps); dt = time.time() - t
    st = eng.stats()
    ns = st["n_seeds_shard"]
    print(json.dumps({"workload": workload, "epsilon": eps, "seeds": ns, "slots": st["n_slots"],
                      "ms_push": round(st["ms_push"], 2), "seeds_per_s": round(ns / (st["ms_push"] / 1e3)),
                      "pushes_per_seed": round(st["pushes"] / ns, 1), "edges_per_seed": round(st["edge_touches"] / ns, 1),
                      "support_per_seed": round(st["support"] / ns, 1), "members_per_seed": round(st["members"] / ns, 1),
                      "max_queue": st["max_queue"], "retries": st["retries"],
                      "alg_GBps": round(st["alg_bytes_push"] / st["ms_push"] / 1e6, 1),
                      "frac_of_6550.7": round(st["alg_bytes_push"] / st["ms_push"] / 1e6 / 6550.7, 4)}), flush=True)
