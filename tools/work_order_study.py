"""How well does 1/epsilon-effective predict the cost of a walk?  (CPU only: oracle walks on the bench shape.)

The push engines hand the seeds of a shard to persistent warps through one atomic counter; the launch ends when the
last walk ends, so the order of the list decides how long the tail is.  This script walks a degree-stratified sample
of the YouTube-shaped bench graph with the oracle (test infrastructure, never the product path), records the edge
touches of every walk, and compares three orders of the sample on a list-scheduling simulation: the seed list's own
order (count-descending, arcte.py:610-617), ascending epsilon-effective (what extract_shard does now), and the
unknowable optimum (descending true cost).

    python tools/work_order_study.py [n_sample] > profiles/r2_work_order.json
"""
import heapq
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_graph, RHO, EPS  # noqa: E402
from oracle import arcte_oracle as O  # noqa: E402


def spearman(a, b):
    ra = np.argsort(np.argsort(a, kind="stable"), kind="stable").astype(np.float64)
    rb = np.argsort(np.argsort(b, kind="stable"), kind="stable").astype(np.float64)
    return float(np.corrcoef(ra, rb)[0, 1])


def makespan(cost, order, n_slots):
    """Persistent workers pulling from one counter: the next item goes to the worker that frees first."""
    free = [0.0] * n_slots
    heapq.heapify(free)
    for k in order:
        t = heapq.heappop(free)
        heapq.heappush(free, t + cost[k])
    return max(free)


def main():
    n_sample = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    A = make_graph("youtube")
    g = O.Graph(A)
    seeds = g.seeds()
    pick = np.linspace(0, seeds.size - 1, n_sample).astype(np.int64)
    sample = seeds[pick]
    eff = np.array([O.epsilon_effective(g, EPS, int(s)) for s in sample])
    edges = np.empty(n_sample, dtype=np.int64)
    pushes = np.empty(n_sample, dtype=np.int64)
    support = np.empty(n_sample, dtype=np.int64)
    for k, s in enumerate(sample):
        _, _, nop, st = O.push(g, O.RULE_ABSORBING, int(s), RHO, float(eff[k]))
        edges[k], pushes[k], support[k] = st["edges"], nop, st["support"]
    cost = (edges + 8 * pushes + 2 * support).astype(np.float64)  # dependent accesses of a walk, roughly
    counts = np.diff(g.indptr)[sample]
    out = {
        "graph": "YouTube shape (bench.py make_graph('youtube'))", "sample": n_sample,
        "spearman_cost_vs_inv_eps": spearman(cost, 1.0 / eff),
        "spearman_cost_vs_count": spearman(cost, counts.astype(np.float64)),
        "share_of_cost_in_top_1pct_by_true_cost": float(np.sort(cost)[::-1][: n_sample // 100].sum() / cost.sum()),
        "share_of_cost_in_first_1pct_by_eps": float(cost[np.argsort(eff, kind="stable")][: n_sample // 100].sum() / cost.sum()),
        "share_of_cost_in_first_1pct_by_count": float(cost[: n_sample // 100].sum() / cost.sum()),
        "makespan_over_ideal": {},
    }
    # the sample is 1/300 of the list: scale the slot count the same way (7104 walks in flight -> 24; an 8-GPU shard
    # has 1/8 of the seeds on the same 7104 slots -> 190)
    for label, slots in (("1 GPU (7104 slots : 894174 seeds)", max(1, round(7104 * n_sample / seeds.size))),
                         ("8 GPUs (7104 slots : 111772 seeds)", max(1, round(7104 * n_sample * 8 / seeds.size)))):
        ideal = cost.sum() / slots
        out["makespan_over_ideal"][label] = {
            "slots_in_simulation": slots,
            "seed_list_order": makespan(cost, np.arange(n_sample), slots) / ideal,
            "ascending_epsilon": makespan(cost, np.argsort(eff, kind="stable"), slots) / ideal,
            "descending_true_cost": makespan(cost, np.argsort(-cost, kind="stable"), slots) / ideal,
            "longest_walk_over_ideal": float(cost.max() / ideal),
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
