#!/usr/bin/env python
"""ARCTE extraction benchmark (BASELINE.json metric: seeds/sec + extraction wall time).

    python bench.py --gpus N --steps K --warmup W [--workload youtube|flickr|politicsuk|rmatS|baNxM]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU arm (oracle port, all host threads)

A "step" is one complete pass of the hot path over the synthetic graph with the adjacency CSR already
resident in HBM: K1 transition build, K2 seed selection + epsilon-effective, K3+K4 fused push/threshold
kernel over ALL seeds, (N>1: the in-library NCCL all-to-all of the communities split by row block), K5
assembly of the n x 2n CSR (N>1: each rank its own row block).  value = seeds / step time (whole job), timed
with CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.

`e2e` is the same metric through the public call arcte(A, rho, eps) with an ordinary (pageable) scipy matrix in
and an ordinary scipy matrix out: host-to-device copy of the graph and device-to-host copy of the feature
matrix inside the timed region, steady state (the process has made the call before).  `e2e_cold` is the FIRST
such call of a fresh process (CUDA context, library load, pool allocation included) -- what the reference's
console script and experiments/demo.py pay (one call per process).  Total work is fixed as N grows ("strong"
scaling): the graph and its seed set do not depend on N.

`result_hash` is a 64-bit content hash of the feature matrix (row starts, column indices, non-unit values;
additive over row blocks, see include/arcte_cuda.h): equal hashes at N = 1, 2, 4, 8 mean identical matrices.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RHO, EPS = 0.1, 1e-5  # reference defaults (entry_points/arcte.py:37,40)


def make_graph(workload):
    from reveal_graph_embedding_b200 import graphs
    if workload.startswith("rmat"):
        return graphs.rmat(scale=int(workload[4:] or 22))
    if workload.startswith("ba"):
        n, m = workload[2:].split("x")
        return graphs.barabasi_albert(int(n), int(m), seed=1)
    return graphs.WORKLOADS[workload]()


def describe(workload, A, n_seeds, eps=EPS):
    names = {"youtube": "synthetic ASU-YouTube-shaped graph (Chung-Lu power law, seed 1138499)",
             "flickr": "synthetic ASU-Flickr-shaped graph (Chung-Lu power law, seed 80513)",
             "politicsuk": "synthetic PoliticsUK-shaped planted-partition graph (seed 419)"}
    return {"workload": names.get(workload, workload), "nodes": int(A.shape[0]), "nnz": int(A.nnz),
            "seeds": int(n_seeds), "rho": RHO, "epsilon": eps, "rule": "absorbing (arcte)",
            "l2": "explicit 512 MB flush between timed steps; per-step working set >> 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs, one sample per 500 ms: started before
    the warm-up so that NVML's start-up cost is not inside the timed steps."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "500"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples from here on belong to the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        lines = self.lines[getattr(self, "first", 0):] or self.lines[-2:]
        sm, mx, reasons = [], [], set()
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- CPU arms
def cpu_reference_run(A, n_threads, target_seconds, sample_hint=None):
    """Oracle port (oracle/arcte_oracle.c; bit-identical to the Python reference, see
    tests/test_oracle_golden.py) on n_threads host threads: transition build over the whole graph, walks +
    thresholds over a degree-stratified sample of the seed list (seeds dealt round-robin to the workers like
    arcte.py:651), assembly of the sample's communities.  Seeds/s is for the whole job, the per-seed stages
    extrapolated linearly from the sample."""
    from oracle import arcte_oracle as O
    t0 = time.perf_counter()
    g = O.Graph(A)                      # transition.py:43-68 restated: degrees + row normalisation
    seeds = g.seeds()
    t_graph = time.perf_counter() - t0
    k = min(seeds.size, sample_hint or 2000)
    while True:
        idx = np.unique(np.linspace(0, seeds.size - 1, k).astype(np.int64))  # stratified over the degree order
        t0 = time.perf_counter()
        sd, seg, mem, eff, st = O.extract(g, 0, RHO, EPS, seeds[idx], n_threads)
        dt = time.perf_counter() - t0
        if dt >= 0.4 * target_seconds or idx.size >= seeds.size:
            break
        # aim the walks at 0.6 of the target: the assembly of the sample's communities takes about half as long again
        k = int(min(seeds.size, max(k * 2, k * 0.6 * target_seconds / max(dt, 1e-3))))
    t0 = time.perf_counter()
    O.assemble(g, sd, seg, mem)
    t_asm = time.perf_counter() - t0
    frac = idx.size / seeds.size
    total = t_graph + dt / frac + t_asm     # assembly: the sample's only, not extrapolated (favours the CPU side)
    return {"value": seeds.size / total, "unit": "seeds/s", "cores": n_threads, "kind": "port",
            "sample": "%d of %d seeds (evenly spaced over the degree-sorted seed list): transition build %.2f s (whole "
                      "graph, 1 thread) + walks and thresholds %.1f s on the sample, extrapolated linearly in the seed count, + "
                      "assembly of the sample %.2f s (not extrapolated)" % (idx.size, seeds.size, t_graph, dt, t_asm),
            "walk_only_seeds_per_s": idx.size / dt, "sample_seeds": int(idx.size)}, seeds.size


def _python_reference_worker(args):
    """One process of the unmodified reference (baseline/_ref): arcte_worker (arcte.py:279) on its share of the
    sample, exactly as arcte() hands work to its pool (arcte.py:650-668)."""
    ref_dir, seeds, indices, indptr, data, d_out, d_in = args
    sys.path.insert(0, ref_dir)
    from reveal_graph_embedding.embedding.arcte.arcte import arcte_worker
    t0 = time.perf_counter()
    arcte_worker(seeds, indices, indptr, data, d_out, d_in, RHO, EPS)
    return time.perf_counter() - t0


def python_reference_run(A, n_procs, seeds_per_proc):
    """The UNMODIFIED Python reference on the box's host cores: arcte_worker on a degree-stratified sample of
    seeds_per_proc seeds per process, all processes at once; seeds/s extrapolated to the whole seed list.  The
    transition matrix is handed over ready-made (the reference builds it with a Python loop over all rows,
    transition.py:61-63: 9.6 s at this size on one core)."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "reveal_graph_embedding")):
        return {"unavailable": "baseline/_ref is absent (run baseline/make_ref.sh where /root/reference exists)"}
    import multiprocessing as mp
    from oracle import arcte_oracle as O
    g = O.Graph(A)
    seeds = g.seeds()
    k = min(seeds.size, n_procs * seeds_per_proc)
    idx = np.unique(np.linspace(0, seeds.size - 1, k).astype(np.int64))
    sample = seeds[idx]
    w_data = g.w
    indices = np.asarray(A.indices, dtype=np.int64)
    indptr = np.asarray(A.indptr, dtype=np.int64)
    chunks = [sample[r::n_procs] for r in range(n_procs)]           # roundrobin_chunks, arcte.py:19-23
    chunks = [c for c in chunks if c.size]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(len(chunks)) as pool:
        per = pool.map(_python_reference_worker, [(ref_dir, c, indices, indptr, w_data, g.d_out, g.d_in) for c in chunks])
    wall = time.perf_counter() - t0
    return {"value": sample.size / wall, "unit": "seeds/s", "cores": len(chunks), "kind": "reference",
            "sample": "%d of %d seeds (evenly spaced over the degree-sorted list), %d per process, arcte_worker of the "
                      "unmodified reference in %d processes: %.1f s wall (slowest worker %.1f s); extrapolated linearly "
                      "in the seed count; transition build and final hstack not charged"
                      % (sample.size, seeds.size, seeds_per_proc, len(chunks), wall, max(per)),
            "extrapolated_full_run_s": seeds.size / (sample.size / wall)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    A = make_graph(args.workload)
    cores = os.cpu_count() or 1
    n_seeds = 0
    vals = []
    # every step walks a bounded sample sized so that the WHOLE run (warm-up + steps) ends within a few minutes
    # whatever --steps/--warmup the caller passes (ARCTE_BENCH_REF_BUDGET_S, default 240 s of CPU arm in total)
    budget = float(os.environ.get("ARCTE_BENCH_REF_BUDGET_S", "240"))
    per_step = max(2.0, min(args.cpu_seconds, budget / max(1, args.warmup + args.steps)))
    hint = None
    for i in range(args.warmup + args.steps):
        res, n_seeds = cpu_reference_run(A, cores, per_step, hint)
        hint = res["sample_seeds"]   # the next step starts from the sample size this one settled on
        if i >= args.warmup:
            vals.append(res)
    v = float(np.mean([r["value"] for r in vals]))
    best = vals[-1]
    best["value"] = v
    line = {"impl": "reference", "metric": "arcte_seeds_per_sec", "value": v, "unit": "seeds/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * n_seeds / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": describe(args.workload, A, n_seeds), "cpu_baseline": best,
            "e2e": {"value": v, "unit": "seeds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference arm = C port of the reference's Python path (bit-identical to it on the fixtures) on all "
                    "host threads, same stages as the GPU arm (transition build, walks + thresholds, assembly); each "
                    "step walks a bounded degree-stratified SAMPLE of the seeds and ms_per_step is the linear "
                    "extrapolation to all seeds, so it is longer than the run took"}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- cold call
COLD_CHILD = r"""
import json, os, sys, time
t_start = time.perf_counter()
sys.path.insert(0, {root!r})
import numpy as np, scipy.sparse as sparse
z = np.load({path!r})
A = sparse.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=tuple(z["shape"]))
t_loaded = time.perf_counter()
from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
from reveal_graph_embedding_b200.engine import csr_hash
t0 = time.perf_counter()
X = arcte(A, {rho!r}, {eps!r}, 1)
t1 = time.perf_counter()
X2 = arcte(A, {rho!r}, {eps!r}, 1)
t2 = time.perf_counter()
print(json.dumps({{"first_call_ms": 1e3 * (t1 - t0), "second_call_ms": 1e3 * (t2 - t1), "nnz": int(X.nnz),
                  "import_ms": 1e3 * (t0 - t_loaded)}}))
"""


def cold_call(A):
    """First arcte() of a fresh process on a pageable scipy matrix (CUDA context creation, library load and
    pool allocation included), then the second call of the same process."""
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(d, "arcte_bench_graph_%d.npz" % os.getpid())
    np.savez(path, data=A.data, indices=A.indices, indptr=A.indptr, shape=np.array(A.shape))
    try:
        env = dict(os.environ)
        for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        out = subprocess.run([sys.executable, "-c", COLD_CHILD.format(root=ROOT, path=path, rho=RHO, eps=EPS)],
                             capture_output=True, text=True, timeout=900, env=env)
        if out.returncode != 0:
            return {"error": out.stderr[-400:]}
        return json.loads(out.stdout.strip().splitlines()[-1])
    finally:
        try:
            os.unlink(path)
        except OSError:
            pass


# --------------------------------------------------------------------------------------- GPU arm
ENGINE_NAMES = {-2: "frontier", 0: "fifo (one queue entry per warp iteration, dense state)",
                1: "batched, direct-mapped 32-byte state", 2: "batched, per-walk hash tables",
                3: "fifo (one queue entry per warp iteration), compact first-touch-ordered state behind an "
                   "epoch-tagged index map, walk labels"}
KERNEL_NAMES = {-2: "k_push_frontier", 0: "k_push_threshold<absorbing>", 1: "k_walk_batched<direct>",
                2: "k_walk_batched<hash>", 3: "k_push_compact<absorbing, 6 or 8 CTAs per SM>"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ARCTE_BENCH_WORKLOAD", "youtube"))
    ap.add_argument("--epsilon", type=float, default=EPS)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work per cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-python-ref", action="store_true", help="skip the unmodified-Python-reference leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cold", action="store_true")
    ap.add_argument("--warps-per-sm", type=int, default=0)
    ap.add_argument("--engine", default="auto")
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries ONE JSON line: whatever NCCL logs (NCCL_DEBUG=VERSION/INFO set by the caller) goes to stderr.
        # NCCL opens the file with "w": not when stderr is a regular file, which that would truncate.
        import stat
        try:
            if not stat.S_ISREG(os.fstat(2).st_mode):
                os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        except OSError:
            pass
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from reveal_graph_embedding_b200 import distributed as ardist
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.engine import RULE_ABSORBING, csr_hash, get_engine

    eps = args.epsilon
    A = make_graph(args.workload)     # ordinary scipy CSR in pageable memory, as a caller would have it
    eng = get_engine(local_rank)
    eng.set_engine(args.engine)
    if args.warps_per_sm:
        eng.configure(warps_per_sm=args.warps_per_sm)
    eng.set_graph(A)
    n_seeds = int(eng.seeds().size)
    if world > 1:
        ardist.ensure_communicator(eng)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """K1..K5 with the adjacency resident in HBM; returns the stored entries of this rank's rows."""
        eng.build_transition()
        if world > 1:
            eng.extract(RULE_ABSORBING, RHO, eps, shard_rank=rank, shard_count=world)
            return eng.exchange_assemble()
        eng.extract(RULE_ABSORBING, RHO, eps)
        return eng.assemble()

    sampler = ClockSampler(local_rank)
    sampler.start()  # NVML start-up happens during the warm-up, not inside the timed steps
    for _ in range(args.warmup):
        step()
    launches0 = eng.stats()["launches"]
    step_ms, push_ms, alg_bytes, stage_ms = [], [], [], []
    nnz_block = 0
    barrier()
    sampler.mark()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.flush_l2()
        barrier()
        eng.timer_start()
        nnz_block = step()
        ms = eng.timer_stop()
        st = eng.stats()
        step_ms.append(ms)
        push_ms.append(st["ms_push"])
        alg_bytes.append(st["alg_bytes_push"])
        stage_ms.append((st["ms_transition"], st["ms_seeds"], st["ms_push"], st["ms_exchange"] if world > 1 else 0.0,
                         st["ms_assemble"]))
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    st = eng.stats()
    launches = st["launches"] - launches0

    # ---- content hash of the result of the last timed step: additive over the ranks' row blocks ----
    if world > 1:
        nnz_all = ardist.all_gather_int64([nnz_block])[:, 0]
        nnz_lo = int(nnz_all[:rank].sum())
        features_nnz = int(nnz_all.sum())
    else:
        nnz_lo, features_nnz = 0, int(nnz_block)
    h_dev = eng.features_hash(nnz_lo)

    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms, float(np.sum(push_ms)), float(np.sum(alg_bytes)),
                          float(np.mean([s[3] for s in stage_ms]))], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0])
        push_ms_job = float(tmax[1]) / args.steps
        alg_job = float(tsum[2]) / args.steps
        exchange_ms = float(tmax[3])
        hs = torch.tensor([np.int64(np.uint64(h_dev))], dtype=torch.int64, device="cuda")   # sums wrap: mod 2^64
        dist.all_reduce(hs, op=dist.ReduceOp.SUM)
        h_dev = int(np.uint64(np.int64(int(hs[0]))))
    else:
        push_ms_job = float(np.mean(push_ms))
        alg_job = float(np.mean(alg_bytes))
        exchange_ms = 0.0
    ms_per_step = total_ms / args.steps
    value = n_seeds / (ms_per_step / 1e3)

    # ---- end to end through the public API: pageable scipy in, scipy out ----
    e2e = None
    if not args.no_e2e:
        X = None
        n_calls = max(args.warmup, 3)
        for _ in range(n_calls):   # steady state: the process has made the call before
            X = arcte(A, RHO, eps, args.gpus if world == 1 else None)
        barrier()
        calls = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tc = time.perf_counter()
            X = arcte(A, RHO, eps, args.gpus if world == 1 else None)
            calls.append(1e3 * (time.perf_counter() - tc))
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        if rank == 0:
            print("e2e calls ms: " + ", ".join("%.1f" % c for c in calls), file=sys.stderr)
        h2d = int(A.data.nbytes + A.indices.size * 4 + (A.shape[0] + 1) * 8) * world
        e2e = {"value": n_seeds / dt, "unit": "seeds/s", "ms_per_call": dt * 1e3, "h2d_bytes_per_step": h2d,
               "api": "reveal_graph_embedding_b200.embedding.arcte.arcte.arcte(A, 0.1, %g)" % eps,
               "host_memory": "pageable scipy/numpy arrays in and out; the library streams through a fixed 128 MB ring "
                              "of pinned slots (csrc/hostcopy.cu)"}
        if X is not None:
            # the 1.0 values are written by the host copy threads, not copied (arcte.py:379-381): bytes over PCIe
            e2e["d2h_bytes_per_step"] = int(X.indices.size * 4 + (A.shape[0] + 1) * 8)
            e2e["result_nnz"] = int(X.indptr[-1])
            e2e["values"] = "not copied: every stored value is 1.0 except self-loop diagonals (patched on the host)"
            t0 = time.perf_counter()
            e2e["result_hash"] = "%016x" % csr_hash(X)
            e2e["hash_matches_device"] = e2e["result_hash"] == "%016x" % h_dev
            print("host hash %.1f s" % (time.perf_counter() - t0), file=sys.stderr)
        X = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    engine_id = int(st.get("engine", 0))
    # DRAM bytes of one launch of the same kernel on the same workload from the committed ncu capture
    # (per launch, like `achieved`); null for any other configuration.
    traffic = traffic_note = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r2_push_youtube_traffic.json")))
        if args.workload == cap["workload"] and world == cap["n_gpus"] and engine_id == cap["engine"]:
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
            traffic_note = cap.get("note")
    except (OSError, KeyError, ValueError):
        pass
    achieved = alg_job / (push_ms_job / 1e3) / 1e9 / world  # per GPU: each GPU ran alg_job/world bytes
    roofline = {"bound": "hbm", "kernel": KERNEL_NAMES.get(engine_id, "?"), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_job / world, "kernel_ms": push_ms_job,
                "pushes": st["pushes"], "edge_touches": st["edge_touches"], "support": st["support"],
                "note": "achieved = SURVEY 8(d) algorithmic bytes of this GPU's seeds / push-kernel time "
                        "(CUDA events on the launching stream)"}
    if traffic_note:
        roofline["traffic_note"] = traffic_note
    cfg = describe(args.workload, A, n_seeds, eps)
    cfg["engine"] = ENGINE_NAMES.get(engine_id, str(engine_id))
    mean_stage = [float(x) for x in np.mean(np.array(stage_ms), axis=0)]
    if world > 1:
        mean_stage[3] = exchange_ms
    line = {"metric": "arcte_seeds_per_sec", "value": value, "unit": "seeds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "extraction_wall_ms": ms_per_step,
            "stage_ms": dict(zip(("transition", "seeds_eps", "push_threshold", "exchange", "assemble"), mean_stage)),
            "features_nnz": features_nnz, "result_hash": "%016x" % h_dev,
            "n_slots": st["n_slots"], "retries": st["retries"], "slot_utilisation": st["slot_utilisation"],
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "timed_region_wall_s": wall}
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cold and not args.no_e2e:
        eng.close()   # a fresh process must find the GPU as a first caller would: this process's pools are released
        time.sleep(1.0)   # ... and the driver has finished taking back the 100 GB they held
        cold = cold_call(A)
        if "first_call_ms" in cold:
            line["e2e_cold"] = {"value": n_seeds / (cold["first_call_ms"] / 1e3), "unit": "seeds/s",
                                "ms_first_call": cold["first_call_ms"], "ms_second_call": cold["second_call_ms"],
                                "what": "first arcte() of a fresh Python process on a pageable scipy matrix: CUDA "
                                        "context creation, library load and pool allocation inside the timed call"}
        else:
            line["e2e_cold"] = cold
    if not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        cb, _ = cpu_reference_run(A, cores, args.cpu_seconds)
        line["cpu_baseline"] = cb
        if not args.no_python_ref:
            try:
                line["cpu_baseline_python"] = python_reference_run(A, cores, 2000 if A.shape[0] > 200000 else 200)
            except Exception as exc:   # the baseline is a courtesy figure: never lose the bench line over it
                line["cpu_baseline_python"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
