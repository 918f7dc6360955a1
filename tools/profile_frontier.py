"""One extraction under the frontier schedule for ncu (smaller YouTube-like graph by default)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_graph, RHO, EPS
from reveal_graph_embedding_b200 import graphs
from reveal_graph_embedding_b200.engine import Engine
workload = sys.argv[1] if len(sys.argv) > 1 else "youtube_small"
geom = [int(x) for x in sys.argv[2].split(":")] if len(sys.argv) > 2 else [-1, 0, 0, 0, 0]
A = graphs.chung_lu(300_000, 790_000, gamma=2.2, max_degree=12000, seed=7) if workload == "youtube_small" else make_graph(workload)
eng = Engine(0)
eng.set_schedule("frontier", *geom)
eng.set_graph(A)
eng.extract(0, RHO, EPS)
st = eng.stats()
print(json.dumps({k: st[k] for k in ("n_seeds_shard", "pushes", "edge_touches", "support", "members", "rounds", "n_slots", "ms_push", "alg_bytes_push")}))
