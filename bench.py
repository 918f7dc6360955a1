#!/usr/bin/env python
"""ARCTE extraction benchmark (BASELINE.json metric: seeds/sec + extraction wall time).

    python bench.py --gpus N --steps K --warmup W [--workload youtube|flickr|politicsuk|rmatS]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU arm (oracle port, all host threads)

A "step" is one complete pass of the hot path over the synthetic graph with the adjacency
CSR already resident in HBM: K1 transition build, K2 seed selection + epsilon-effective,
K3+K4 fused push/threshold kernel over ALL seeds, (N>1: NCCL all-gather of the per-GPU
segments), K5 assembly of the n x 2n CSR.  value = seeds / step time (whole job).
`e2e` is the same metric through the public call arcte(A, rho, eps) with HOST scipy input
and HOST scipy output (pinned-host H2D of the graph and D2H of the feature matrix inside
the timed region).  Total work is fixed as N grows ("strong" scaling): the graph and its
seed set do not depend on N.
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


def describe(workload, A, n_seeds):
    names = {"youtube": "synthetic ASU-YouTube-shaped graph (Chung-Lu power law, seed 1138499)",
             "flickr": "synthetic ASU-Flickr-shaped graph (Chung-Lu power law, seed 80513)",
             "politicsuk": "synthetic PoliticsUK-shaped planted-partition graph (seed 419)"}
    return {"workload": names.get(workload, workload), "nodes": int(A.shape[0]), "nnz": int(A.nnz),
            "seeds": int(n_seeds), "rho": RHO, "epsilon": EPS, "rule": "absorbing (arcte)",
            "l2": "explicit 512 MB flush between timed steps; per-step working set >> 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
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


def pinned_csr(A):
    """The same matrix with its three arrays in page-locked host memory (dtype as the ABI wants)."""
    import scipy.sparse as sparse
    import torch

    def pin(a, dtype):
        t = torch.empty(a.size, dtype=dtype).pin_memory()
        v = t.numpy()
        v[:] = a
        return v, t
    data, k0 = pin(A.data, torch.float64)
    indices, k1 = pin(A.indices, torch.int32)
    indptr, k2 = pin(A.indptr, torch.int64)
    B = sparse.csr_matrix((data, indices, indptr), shape=A.shape, copy=False)
    B.has_sorted_indices = True
    B.has_canonical_format = True
    return B, (k0, k1, k2)


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(A, n_threads, target_seconds, sample_hint=None):
    """Oracle port (oracle/arcte_oracle.c; bit-identical to the Python reference, see
    tests/test_oracle_golden.py) over a degree-stratified sample of the seed list, seeds
    dealt round-robin to n_threads workers like arcte.py:651."""
    from oracle import arcte_oracle as O
    g = O.Graph(A)
    seeds = g.seeds()
    k = min(seeds.size, sample_hint or 2000)
    while True:
        idx = np.unique(np.linspace(0, seeds.size - 1, k).astype(np.int64))  # stratified over the degree order
        t0 = time.perf_counter()
        O.extract(g, 0, RHO, EPS, seeds[idx], n_threads)
        dt = time.perf_counter() - t0
        if dt >= 0.5 * target_seconds or idx.size >= seeds.size:
            break
        k = int(min(seeds.size, max(k * 2, k * target_seconds / max(dt, 1e-3))))
    return {"value": idx.size / dt, "unit": "seeds/s", "cores": n_threads, "kind": "port",
            "sample": "%d of %d seeds (evenly spaced over the degree-sorted seed list), %.1f s; push + threshold only"
                      % (idx.size, seeds.size, dt)}, seeds.size


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    A = make_graph(args.workload)
    cores = os.cpu_count() or 1
    best, n_seeds = None, 0
    vals = []
    for i in range(args.warmup + args.steps):
        res, n_seeds = cpu_reference_run(A, cores, args.cpu_seconds)
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
            "note": "reference arm = C port of the reference's Python path on all host threads; "
                    "ms_per_step extrapolated linearly from the sample to all seeds"}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ARCTE_BENCH_WORKLOAD", "youtube"))
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work per cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--warps-per-sm", type=int, default=0)
    args = ap.parse_args()

    if args.impl == "reference":
        return run_reference_arm(args)

    os.environ["NCCL_DEBUG"] = "WARN"  # NCCL's version banner goes to stdout otherwise; one JSON line only
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from reveal_graph_embedding_b200 import distributed as ardist
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte
    from reveal_graph_embedding_b200.engine import RULE_ABSORBING, get_engine

    A = make_graph(args.workload)
    A_pinned, _keep = pinned_csr(A)
    eng = get_engine(local_rank)
    if args.warps_per_sm:
        eng.configure(warps_per_sm=args.warps_per_sm)
    eng.set_graph(A_pinned, canonical=True)
    n_seeds = int(eng.seeds().size)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        """K1..K5 with the adjacency resident in HBM."""
        eng.build_transition()
        if world > 1:
            # walk own seed shard, all-gather segments, row-sharded assembly, concatenate on rank 0
            out = ardist.extract_and_concatenate(eng, RULE_ABSORBING, RHO, EPS)
            return None if out is None else out[3]
        eng.extract(RULE_ABSORBING, RHO, EPS)
        eng.assemble()
        return eng.out_nnz

    sampler = ClockSampler(local_rank)
    sampler.start()  # NVML start-up happens during the warm-up, not inside the timed steps
    for _ in range(args.warmup):
        step()
    launches0 = eng.stats()["launches"]
    step_ms, push_ms, alg_bytes, stage_ms = [], [], [], []
    nnz_out = 0
    barrier()
    sampler.lines.clear()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        eng.flush_l2()
        barrier()
        eng.timer_start()
        nnz_step = step()
        ms = eng.timer_stop()
        if nnz_step is not None:
            nnz_out = nnz_step
        st = eng.stats()
        step_ms.append(ms)
        push_ms.append(st["ms_push"])
        alg_bytes.append(st["alg_bytes_push"])
        stage_ms.append((st["ms_transition"], st["ms_seeds"], st["ms_push"], st["ms_assemble"]))
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    st = eng.stats()
    launches = st["launches"] - launches0

    total_ms = float(np.sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms, float(np.sum(push_ms)), float(np.sum(alg_bytes))], dtype=torch.float64,
                         device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms = float(tmax[0])
        push_ms_job = float(tmax[1]) / args.steps
        alg_job = float(tsum[2]) / args.steps
    else:
        push_ms_job = float(np.mean(push_ms))
        alg_job = float(np.mean(alg_bytes))
    ms_per_step = total_ms / args.steps
    value = n_seeds / (ms_per_step / 1e3)

    # ---- end to end through the public API: host scipy in, host scipy out ----
    e2e = None
    if not args.no_e2e:
        X = None
        checksum = 0
        if world == 1:  # where the end-to-end call spends its host time (stderr, not part of the JSON line)
            from reveal_graph_embedding_b200.engine import canonical_csr
            t = [time.perf_counter()]
            Ac = canonical_csr(A_pinned); t.append(time.perf_counter())
            eng.set_graph(Ac, canonical=True); t.append(time.perf_counter())
            eng.extract(RULE_ABSORBING, RHO, EPS); t.append(time.perf_counter())
            eng.assemble(); t.append(time.perf_counter())
            X = eng.features(); t.append(time.perf_counter())
            X = None
            names = ("canonical_csr", "set_graph(H2D+K1+K2a)", "extract", "assemble", "features(D2H)")
            print("e2e breakdown ms: " + ", ".join("%s=%.1f" % (n_, 1e3 * (b - a)) for n_, a, b in
                                                   zip(names, t[:-1], t[1:])), file=sys.stderr)
        from reveal_graph_embedding_b200 import hostmem
        for _ in range(max(args.warmup, 3)):
            # same rebinding pattern as the timed loop (the previous result is alive during the call,
            # so two sets of result buffers rotate); buffers are page-locked in the background
            X = arcte(A_pinned, RHO, EPS, args.gpus if world == 1 else None)
            hostmem.wait_idle()
        barrier()
        ones_hits0 = hostmem.counters["ones_hits"]
        t0 = time.perf_counter()
        for _ in range(args.steps):
            tc = time.perf_counter()
            X = arcte(A_pinned, RHO, EPS, args.gpus if world == 1 else None)
            if rank == 0:
                print("e2e call %.1f ms" % (1e3 * (time.perf_counter() - tc)), file=sys.stderr)
            if X is not None:  # under torchrun the matrix is returned on rank 0
                checksum = int(X.indptr[-1])  # the result is in host memory
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        h2d = int(A_pinned.data.nbytes + A_pinned.indices.nbytes + (A.shape[0] + 1) * 8) * world
        # the value array (all ones except self-loop diagonals) is not copied when a pre-filled
        # page-locked block was ready (hostmem.ones): count only the bytes that crossed PCIe
        values_copied = (hostmem.counters["ones_hits"] - ones_hits0) < args.steps
        d2h = int((X.data.nbytes if values_copied else 0) + X.indices.size * 4 + (A.shape[0] + 1) * 8) if X is not None else 0
        checksum = checksum if X is not None else 0
        e2e = {"value": n_seeds / dt, "unit": "seeds/s", "ms_per_call": dt * 1e3,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "result_nnz": checksum,
               "values": "copied from the device" if values_copied else
                         "not copied: pre-filled ones on the host, self-loop diagonals patched (indices and indptr copied)",
               "api": "reveal_graph_embedding_b200.embedding.arcte.arcte.arcte(A, 0.1, 1e-5)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    # DRAM bytes of one launch of the same kernel on the same workload from the committed
    # `ncu --set full` capture (per launch, like `achieved`); null for any other configuration.
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "r1_push_youtube_traffic.json")))
        if args.workload == cap["workload"] and world == cap["n_gpus"]:
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    achieved = alg_job / (push_ms_job / 1e3) / 1e9 / world  # per GPU: each GPU ran alg_job/world bytes
    roofline = {"bound": "hbm", "kernel": "k_push_threshold<absorbing>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_job / world, "kernel_ms": push_ms_job,
                "pushes": st["pushes"], "edge_touches": st["edge_touches"], "support": st["support"],
                "note": "achieved = SURVEY 8(d) algorithmic bytes of this GPU's seeds / push-kernel time "
                        "(CUDA events on the launching stream)"}
    line = {"metric": "arcte_seeds_per_sec", "value": value, "unit": "seeds/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": describe(args.workload, A, n_seeds),
            "extraction_wall_ms": ms_per_step,
            "stage_ms": dict(zip(("transition", "seeds_eps", "push_threshold", "assemble"),
                                 [float(x) for x in np.mean(np.array(stage_ms), axis=0)])),
            "features_nnz": int(nnz_out), "n_slots": st["n_slots"], "retries": st["retries"],
            "slot_utilisation": st["slot_utilisation"],
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "timed_region_wall_s": wall}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu and world == 1:
        cb, _ = cpu_reference_run(A, os.cpu_count() or 1, args.cpu_seconds)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
