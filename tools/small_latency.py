"""Latency of one arcte() call on a small graph (BASELINE.json config 2: PoliticsUK shape), with
the host-side wall time of every stage, next to the CPU oracle port on the same host."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from bench import EPS, RHO, make_graph
from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
from reveal_graph_embedding_b200.engine import canonical_csr, get_engine

workload = sys.argv[1] if len(sys.argv) > 1 else "politicsuk"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
A = make_graph(workload)
eng = get_engine(0)
for _ in range(3):
    X = arcte(A, RHO, EPS, 1)
names = ("canonical_csr", "set_graph", "extract", "assemble", "features")
acc = np.zeros((reps, len(names)))
for r in range(reps):
    t = [time.perf_counter()]
    Ac = canonical_csr(A); t.append(time.perf_counter())
    eng.set_graph(Ac, canonical=True); t.append(time.perf_counter())
    eng.extract(0, RHO, EPS); t.append(time.perf_counter())
    eng.assemble(); t.append(time.perf_counter())
    X = eng.features(); t.append(time.perf_counter())
    acc[r] = np.diff(t) * 1e3
calls = []
for r in range(reps):
    t0 = time.perf_counter(); X = arcte(A, RHO, EPS, 1); calls.append((time.perf_counter() - t0) * 1e3)
steps = []
for r in range(reps):
    eng.flush_l2()
    eng.timer_start(); eng.build_transition(); eng.extract(0, RHO, EPS); eng.assemble(); steps.append(eng.timer_stop())
st = eng.stats()
print("%s [schedule %s]: n=%d nnz=%d seeds=%d  features nnz=%d"
      % (workload, os.environ.get("ARCTE_CUDA_SCHEDULE", "fifo"), A.shape[0], A.nnz, st["n_seeds_shard"], X.nnz))
print("stage wall ms (median of %d): " % reps + ", ".join("%s=%.3f" % (n, v) for n, v in zip(names, np.median(acc, axis=0))))
print("arcte() wall ms: median %.3f  min %.3f  max %.3f" % (np.median(calls), np.min(calls), np.max(calls)))
print("device step ms (K1..K5, resident): median %.3f min %.3f max %.3f; kernels: transition %.3f seeds %.3f push %.3f assemble %.3f; launches/extract+assemble %d"
      % (np.median(steps), np.min(steps), np.max(steps), st["ms_transition"], st["ms_seeds"], st["ms_push"], st["ms_assemble"], st["launches"]))
from oracle import arcte_oracle as O
t0 = time.perf_counter(); Y = O.arcte(A, RHO, EPS, os.cpu_count() or 1); t_cpu = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter(); Y = O.arcte(A, RHO, EPS, 1); t_cpu1 = (time.perf_counter() - t0) * 1e3
print("CPU oracle port (reference FIFO order), whole arcte(): %.3f ms on %d threads, %.3f ms on 1 thread; entries that differ from it: %d of %d"
      % (t_cpu, os.cpu_count() or 1, t_cpu1, int((X != Y).nnz), int(Y.nnz)))
