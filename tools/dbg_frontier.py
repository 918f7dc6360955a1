import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from helpers import load_golden, RHO, EPS
from oracle import arcte_oracle as O
from reveal_graph_embedding_b200.engine import Engine
A, z = load_golden("ba300")
g = O.Graph(A)
e = Engine(0); e.set_schedule("frontier"); e.set_graph(A)
e.extract(0, RHO, EPS, eps_override=z["eps_eff"])
seg_seed, seg_cnt, seg_off, mem = e.segments()
with O.schedule(1):
    sd, seg, omem, eff, st = O.extract(g, 0, RHO, EPS, z["seeds"], 1, eps_override=z["eps_eff"])
print("gpu stats", {k: e.stats()[k] for k in ("pushes","edge_touches","enqueues","support","members","emitted","max_queue","rounds")})
print("ora stats", st)
print("seed order equal", np.array_equal(seg_seed, z["seeds"]))
o = 0; bad = 0
for i in range(seg.size):
    c = int(seg[i]); a = set(omem[o:o+c].tolist()); o += c
    b = set(mem[seg_off[i]:seg_off[i]+seg_cnt[i]].tolist())
    if a != b:
        bad += 1
        if bad <= 5:
            sdn = int(z["seeds"][i]); eps = float(z["eps_eff"][i])
            with O.schedule(1): s, r, nop, stt = O.push(g, 0, sdn, RHO, eps)
            sg, rg, nopg = e.push(0, sdn, RHO, eps)
            base = np.append(A.indices[A.indptr[sdn]:A.indptr[sdn+1]], sdn)
            q = s / g.d_in; tau = q[base].min()
            print("pos", i, "seed", sdn, "deg", base.size-1, "oracle m", len(a), "gpu m", len(b), "only oracle", sorted(a-b)[:10], "only gpu", sorted(b-a)[:10], "s equal", np.array_equal(s, sg), "tau", tau, "n pass", int((q[s>0] >= tau).sum()))
print("mismatching seeds", bad, "of", seg.size)
