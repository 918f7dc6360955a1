"""One extraction (one launch of the fused push kernel) for ncu; prints the library's stats."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_graph, RHO, EPS
from reveal_graph_embedding_b200 import graphs
from reveal_graph_embedding_b200.engine import Engine

workload = sys.argv[1]
wps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if workload == "youtube_small":
    A = graphs.chung_lu(300_000, 790_000, gamma=2.2, max_degree=12000, seed=7)
else:
    A = make_graph(workload)
eng = Engine(0)
if wps:
    eng.configure(warps_per_sm=wps)
eng.set_graph(A)
eng.extract(0, RHO, EPS)
st = eng.stats()
print(json.dumps({k: st[k] for k in ("n_seeds_shard", "pushes", "edge_touches", "support", "members", "n_slots",
                                     "ms_push", "alg_bytes_push", "slot_utilisation")}))
