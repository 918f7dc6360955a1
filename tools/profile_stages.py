"""One pass over every stage around the push kernel for ncu (K1 transition build, K2 seeds + epsilon-effective,
K5 assembly, K6 column normalisation / chi2 + peak-SNR weights), on the bench shape with a small walk-state pool so
that ncu's save/restore of device memory between replay passes stays cheap.

    ncu --set full --clock-control none --import-source on -k regex:'<kernels>' -c 80 -o gpurun_out/r2_stages \
        python tools/profile_stages.py youtube
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import scipy.sparse as sparse  # noqa: E402

from bench import EPS, RHO, make_graph  # noqa: E402
from reveal_graph_embedding_b200.engine import Engine  # noqa: E402

A = make_graph(sys.argv[1] if len(sys.argv) > 1 else "youtube")
eng = Engine(0)
eng.set_engine("compact")
eng.configure(warps_per_sm=8, mem_percent=20)
eng.set_graph(A)                       # K1 + K2a
eng.extract(0, RHO, EPS)               # K2b + K3/K4 (the push kernel is not in ncu's filter)
eng.assemble()                         # K5
st = eng.stats()
rng = np.random.default_rng(0)
n = A.shape[0]
train = np.sort(rng.choice(n, size=3170, replace=False))
test = np.setdiff1d(np.arange(n), train)[:20000]
Y = sparse.csr_matrix((np.ones(train.size), (np.arange(train.size), rng.integers(0, 47, size=train.size))), shape=(train.size, 47))
eng.normalize_features()               # K6: column histogram + scale
eng.store_features()                   # keep the normalised matrix resident
eng.weighted_fold(train, test, Y)      # K6: row gather, chi2, peak SNR, weighting of both blocks
print(json.dumps({k: st[k] for k in ("ms_transition", "ms_seeds", "ms_push", "ms_assemble", "members", "n_slots")}))
