"""torchrun target: where one distributed arcte() call spends its host time on rank 0 (ARCTE_CUDA_DEBUG=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ARCTE_CUDA_DEBUG"] = "1"
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from bench import make_graph, RHO, EPS
from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
A = make_graph(sys.argv[1] if len(sys.argv) > 1 else "youtube")
for i in range(4):
    dist.barrier(); t = time.perf_counter(); X = arcte(A, RHO, EPS); dist.barrier()
    if dist.get_rank() == 0:
        print("call %d: %.1f ms" % (i, 1e3 * (time.perf_counter() - t)), file=sys.stderr)
dist.destroy_process_group()
