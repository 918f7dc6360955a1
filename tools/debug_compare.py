"""Debug helper: per-seed comparison of GPU segments with the oracle (run on the GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import *
from oracle import arcte_oracle as O
from reveal_graph_embedding_b200.engine import Engine

rng = np.random.default_rng(9)
A, _ = load_golden("ba2000")
A = A.copy()
A.data = rng.uniform(0.2, 4.0, A.nnz)
g = O.Graph(A)
eng = Engine(0)
eng.set_graph(A)
seeds = eng.seeds()
eps_dev = eng.epsilon_effective(EPS, seeds)
for rule in (0, 1, 2):
    rho = RHO if rule != 2 else (RHO * 0.5) / (1 - 0.5 * RHO)
    sd, seg, mem, eff, st = O.extract(g, rule, rho, EPS, seeds, 8, eps_override=eps_dev)
    eng.extract(rule, rho, EPS)
    gs = eng.stats()
    print("rule", rule, "oracle", st)
    print("        gpu", {k: gs[k] for k in ("pushes", "edge_touches", "enqueues", "max_queue", "support", "members", "emitted", "retries", "n_slots")})
    s_seed, s_cnt, s_off, s_mem = eng.segments()
    assert np.array_equal(s_seed, seeds.astype(np.int32))
    bad = np.nonzero(s_cnt != seg)[0]
    print("   mismatching segments:", bad.size, bad[:10])
    for k in bad[:3]:
        seed = int(seeds[k])
        s, r, nop = eng.push(rule, seed, rho, float(eps_dev[k]))
        so, ro, nopo, sto = O.push(g, rule, seed, rho, float(eps_dev[k]))
        print("   seed", seed, "deg", A.indptr[seed+1]-A.indptr[seed], "gpu cnt", s_cnt[k], "oracle cnt", seg[k], "nop", nop, nopo,
              "s equal", np.array_equal(s, so), "r equal", np.array_equal(r, ro), sto)
