"""End-to-end arcte() in ONE process driving N GPUs (host threads, in-library NCCL exchange)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_graph, RHO, EPS
from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte

workload, n_gpus = sys.argv[1], int(sys.argv[2])
A = make_graph(workload)
ts = []
for i in range(4):
    t = time.perf_counter(); X = arcte(A, RHO, EPS, n_gpus); ts.append(time.perf_counter() - t)
print(json.dumps({"workload": workload, "gpus": n_gpus, "ms_per_call_first_to_last": [round(1e3 * t, 1) for t in ts],
                  "nnz": int(X.nnz), "shape": list(X.shape)}))
