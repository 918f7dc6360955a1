"""Where the host side of one arcte() call goes (run with ARCTE_CUDA_DEBUG=1): python tools/e2e_breakdown.py youtube"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ARCTE_CUDA_DEBUG"] = "1"
from bench import make_graph, RHO, EPS
from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
A = make_graph(sys.argv[1] if len(sys.argv) > 1 else "youtube")
for i in range(3):
    t = time.perf_counter(); X = arcte(A, RHO, EPS, 1); print("call %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t)), file=sys.stderr)
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), open("/sys/kernel/mm/transparent_hugepage/shmem_enabled").read().strip(), file=sys.stderr)
