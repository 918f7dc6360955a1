"""BASELINE.json config 5: epsilon sweep on an R-MAT graph over N GPUs driven by ONE process (host thread per GPU,
in-library NCCL exchange): push-queue size vs achieved HBM rate.

    python tools/eps_sweep_multi.py rmat22 8 1e-3,1e-4,1e-5,1e-6 [max_seeds] [engine]

Every epsilon is one complete extraction over ALL seeds (or an evenly spaced, degree-stratified sample of max_seeds
of them -- stated in every line): per-GPU shard walk, all-to-all of the communities, per-GPU row-block assembly.
A line per epsilon: wall time of the whole step (max over GPUs), the push kernel's time and algorithmic GB/s per
GPU (sum of the algorithmic bytes / slowest kernel / N), queue statistics, and the 64-bit content hash of the result.
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import RHO, make_graph  # noqa: E402
from reveal_graph_embedding_b200.engine import Engine, device_count  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "rmat20"
n_gpus = int(sys.argv[2]) if len(sys.argv) > 2 else device_count()
eps_list = [float(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1e-3,1e-4,1e-5,1e-6").split(",")]
max_seeds = int(sys.argv[4]) if len(sys.argv) > 4 else 0
engine = sys.argv[5] if len(sys.argv) > 5 else "auto"

t = time.time()
A = make_graph(workload)
print("graph %s n=%d nnz=%d (%.1fs)" % (workload, A.shape[0], A.nnz, time.time() - t), flush=True)
engines = [Engine(d) for d in range(n_gpus)]


def parallel(fn):
    errs = []

    def call(r):
        try:
            fn(r)
        except BaseException as e:   # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=call, args=(r,)) for r in range(n_gpus)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if errs:
        raise errs[0]


def setup(r):
    engines[r].set_engine(engine)
    engines[r].set_graph(A, canonical=True)


t = time.time()
parallel(setup)
seeds = engines[0].seeds()
sample = None
if max_seeds and max_seeds < seeds.size:
    sample = seeds[np.unique(np.linspace(0, seeds.size - 1, max_seeds).astype(np.int64))]
    parallel(lambda r: engines[r].set_seeds(sample))
n_seeds = int(sample.size if sample is not None else seeds.size)
print("set_graph on %d GPUs %.2fs, seeds=%d of %d" % (n_gpus, time.time() - t, n_seeds, seeds.size), flush=True)
if n_gpus > 1:
    Engine.comm_init_all(engines)

for eps in eps_list:
    nnz = [0] * n_gpus

    def step(r):
        e = engines[r]
        e.extract(0, RHO, eps, shard_rank=r, shard_count=n_gpus)
        nnz[r] = e.exchange_assemble() if n_gpus > 1 else e.assemble()
    for rep in range(2):                      # the first pass sizes pools and rings
        t = time.perf_counter()
        parallel(step)
        wall = time.perf_counter() - t
    sts = [e.stats() for e in engines]
    off = np.concatenate([[0], np.cumsum(nnz)])
    h = sum(engines[r].features_hash(int(off[r])) for r in range(n_gpus)) & ((1 << 64) - 1)
    ms_push = max(s["ms_push"] for s in sts)
    alg = sum(s["alg_bytes_push"] for s in sts)
    tot = {k: sum(s[k] for s in sts) for k in ("pushes", "edge_touches", "support", "members", "enqueues", "retries")}
    print(json.dumps({"workload": workload, "n_gpus": n_gpus, "epsilon": eps, "seeds": n_seeds,
                      "seeds_are": "all" if sample is None else "evenly spaced sample of the degree-sorted list",
                      "engine": sts[0]["engine"], "slots_per_gpu": sts[0]["n_slots"],
                      "step_wall_ms": round(wall * 1e3, 1), "seeds_per_s": round(n_seeds / wall),
                      "ms_push_max": round(ms_push, 2), "ms_exchange_max": round(max(s["ms_exchange"] for s in sts), 2),
                      "ms_assemble_max": round(max(s["ms_assemble"] for s in sts), 2),
                      "alg_GBps_per_gpu": round(alg / ms_push / 1e6 / n_gpus, 1),
                      "frac_of_6550.7": round(alg / ms_push / 1e6 / n_gpus / 6550.7, 4),
                      "pushes_per_seed": round(tot["pushes"] / n_seeds, 1), "edges_per_seed": round(tot["edge_touches"] / n_seeds, 1),
                      "support_per_seed": round(tot["support"] / n_seeds, 1), "members_per_seed": round(tot["members"] / n_seeds, 1),
                      "enqueues_per_seed": round(tot["enqueues"] / n_seeds, 1), "max_queue": max(s["max_queue"] for s in sts),
                      "retries": tot["retries"], "features_nnz": int(off[-1]), "result_hash": "%016x" % h}), flush=True)
