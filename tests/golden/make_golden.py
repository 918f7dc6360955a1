"""Generate the golden fixtures in this directory from the UNMODIFIED Python
reference (MKLab-ITI/reveal-graph-embedding mounted at /root/reference).

Run (only in the build container; the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each fixture <name>.npz holds the input adjacency CSR and everything the
reference computes on it along the ARCTE path:

  A_indptr/A_indices/A_data         input (canonical CSR, float64)
  W_data, d_out, d_in               get_natural_random_walk_matrix   (transition.py:43)
  seeds, eps_eff                    seed list (arcte.py:614-617) and calculate_epsilon_effective
                                    per seed exactly as arcte_worker calls it (arcte.py:340)
  X{0,1,2}_indptr/_indices/_data    arcte / arcte_with_pagerank / arcte_with_lazy_pagerank output
  probe_seeds, probe_rule{r}_s/_r/_nop   dense s, r and push count of the three
                                    similarity.py drivers for a few seeds
  pairwise_*                        np.mean inputs/outputs pinning numpy's summation order
"""
import os
import sys

import numpy as np
import scipy.sparse as sparse

sys.path.insert(0, "/root/reference")

from reveal_graph_embedding.embedding.arcte import arcte as ref_arcte  # noqa: E402
from reveal_graph_embedding.eps_randomwalk import similarity as ref_sim  # noqa: E402
from reveal_graph_embedding.eps_randomwalk.transition import get_natural_random_walk_matrix  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RHO, EPS = 0.1, 1e-5  # reference defaults: entry_points/arcte.py:37,40; experiments/demo.py:20-21


def sym(rows, cols, vals, n):
    A = sparse.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def graph_ba(n, m, seed):
    import networkx as nx
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    e = np.array(G.edges(), dtype=np.int64)
    r = np.concatenate([e[:, 0], e[:, 1]])
    c = np.concatenate([e[:, 1], e[:, 0]])
    return sym(r, c, np.ones(r.size), n)


def graph_weighted(n, p, seed):
    rng = np.random.default_rng(seed)
    U = sparse.random(n, n, density=p, random_state=rng, format="coo",
                      data_rvs=lambda k: rng.uniform(0.1, 3.0, size=k))
    U = sparse.triu(U, k=1)
    A = (U + U.T).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def graph_planted(n, groups, p_in, p_out, seed):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, groups, size=n)
    P = np.where(lab[:, None] == lab[None, :], p_in, p_out)
    M = np.triu(rng.random((n, n)) < P, k=1)
    r, c = np.nonzero(M)
    return sym(np.concatenate([r, c]), np.concatenate([c, r]), np.ones(2 * r.size), n)


def graph_edgecases(seed):
    """Self loops, an isolated node, degree-1 leaves, a hub, halved weights like the
    CLI's (A + A.T)/2 (entry_points/arcte.py:70-71)."""
    rng = np.random.default_rng(seed)
    n = 160
    rows, cols = [], []
    for v in range(1, 60):            # hub 0 with 59 spokes
        rows.append(0); cols.append(v)
    for _ in range(420):              # random directed edges among 0..149
        a, b = rng.integers(0, 150, size=2)
        rows.append(int(a)); cols.append(int(b))     # may create self loops
    for v in range(150, 158):         # leaves hanging off random nodes
        rows.append(v); cols.append(int(rng.integers(0, 150)))
    # node 158 has a self loop and two neighbours; node 159 is isolated
    rows += [158, 158, 158]; cols += [158, 3, 7]
    D = sparse.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)).tocsr()
    A = (D + D.transpose()) / 2
    A = sparse.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A


def run_reference(A, probe_count, with_variants=True):
    out = {}
    A = sparse.csr_matrix(A, dtype=np.float64)
    n = A.shape[0]
    out["A_indptr"] = A.indptr.astype(np.int64)
    out["A_indices"] = A.indices.astype(np.int32)
    out["A_data"] = A.data.astype(np.float64)

    W, d_out, d_in = get_natural_random_walk_matrix(A, make_shared=False)
    assert np.array_equal(W.indices, A.indices) and np.array_equal(W.indptr, A.indptr)
    out["W_data"], out["d_out"], out["d_in"] = W.data.copy(), d_out.copy(), d_in.copy()

    # seed list exactly as arcte.py:610-617
    a = A.copy()
    a.data = np.ones_like(a.data)
    cnt = np.squeeze(np.asarray(a.sum(axis=0), dtype=np.int64))
    it = np.where(cnt != 0)[0]
    it = it[np.argsort(cnt[it])][::-1]
    it = it[np.where(cnt[it] > 1.0)[0]]
    out["seeds"] = it.astype(np.int64)

    adj = [W.indices[W.indptr[i]:W.indptr[i + 1]] for i in range(n)]
    wts = [W.data[W.indptr[i]:W.indptr[i + 1]] for i in range(n)]
    a_i = np.ndarray(n, dtype=np.ndarray)
    w_i = np.ndarray(n, dtype=np.ndarray)
    for i in range(n):
        a_i[i], w_i[i] = adj[i], wts[i]
    mean_degree = np.mean(d_out)
    out["eps_eff"] = np.array([
        ref_arcte.calculate_epsilon_effective(RHO, EPS, d_out[s], d_out[adj[s]], mean_degree)
        for s in it], dtype=np.float64)

    fns = [ref_arcte.arcte, ref_arcte.arcte_with_pagerank, ref_arcte.arcte_with_lazy_pagerank]
    for rule, fn in enumerate(fns):
        if rule > 0 and not with_variants:
            continue
        X = sparse.csr_matrix(fn(A.copy(), RHO, EPS, 1))
        X.sort_indices()
        assert X.shape == (n, 2 * n)
        out["X%d_indptr" % rule] = X.indptr.astype(np.int64)
        out["X%d_indices" % rule] = X.indices.astype(np.int64)
        out["X%d_data" % rule] = X.data.astype(np.float64)

    rng = np.random.default_rng(12345)
    k = min(probe_count, it.size)
    probe_idx = np.sort(rng.choice(it.size, size=k, replace=False))
    probe_idx[0] = 0  # always include the highest-degree seed
    out["probe_seeds"] = it[probe_idx].astype(np.int64)
    out["probe_eps"] = out["eps_eff"][probe_idx]
    drivers = [
        lambda s, r, sd, e: ref_sim.fast_approximate_cumulative_pagerank_difference(
            s, r, w_i, a_i, d_out, d_in, sd, RHO, e),
        lambda s, r, sd, e: ref_sim.fast_approximate_personalized_pagerank(
            s, r, w_i, a_i, d_out, d_in, sd, RHO, e),
        lambda s, r, sd, e: ref_sim.lazy_approximate_personalized_pagerank(
            s, r, w_i, a_i, d_out, d_in, sd, (RHO * 0.5) / (1 - 0.5 * RHO), e),
    ]
    for rule, drv in enumerate(drivers):
        S, R, NOP = [], [], []
        for sd, e in zip(out["probe_seeds"], out["probe_eps"]):
            s = np.zeros(n); r = np.zeros(n)
            NOP.append(drv(s, r, int(sd), float(e)))
            S.append(s); R.append(r)
        out["probe_rule%d_s" % rule] = np.array(S)
        out["probe_rule%d_r" % rule] = np.array(R)
        out["probe_rule%d_nop" % rule] = np.array(NOP, dtype=np.int64)
    return out


def pairwise_fixture():
    """Pins numpy's float64 add.reduce order behind neighbor_degrees.mean() (arcte.py:32)."""
    rng = np.random.default_rng(7)
    lens = [1, 2, 7, 8, 9, 15, 16, 17, 63, 64, 127, 128, 129, 130, 136, 255, 256, 257, 300, 511,
            1000, 1023, 1025, 4097, 33333]
    vals, means = [], []
    for L in lens:
        v = rng.uniform(0.5, 5000.0, size=L) * rng.choice([1.0, 1e-3, 1e3], size=L)
        vals.append(v)
        means.append(v.mean())
    return {"pairwise_lens": np.array(lens, dtype=np.int64),
            "pairwise_vals": np.concatenate(vals),
            "pairwise_means": np.array(means, dtype=np.float64)}


def main():
    graphs = {
        "ba300": (graph_ba(300, 3, 1), 6, True),
        "weighted200": (graph_weighted(200, 0.04, 3), 6, True),
        "planted419": (graph_planted(419, 5, 0.35, 0.05, 419), 4, False),
        "edgecases160": (graph_edgecases(5), 8, True),
        "ba2000": (graph_ba(2000, 4, 11), 4, False),
        # SURVEY.md section 4 known answer: nnz 233,123, local block 178,173, 533,063 pushes
        "ba5000": (graph_ba(5000, 5, 7), 3, False),
    }
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    for name, (A, probes, variants) in graphs.items():
        if only and name not in only:
            continue
        out = run_reference(A, probes, variants)
        if name == "ba300":
            out.update(pairwise_fixture())
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "n=%d nnz=%d seeds=%d X0.nnz=%d  %.1f KB" % (
            A.shape[0], A.nnz, out["seeds"].size, out["X0_data"].size,
            os.path.getsize(path) / 1024))




def cli_fixture():
    """Edge list -> reference CLI pipeline (entry_points/arcte.py:62-84) -> feature file."""
    import scipy.sparse as spsp
    from reveal_graph_embedding.datautil.datarw import read_adjacency_matrix, write_features
    rng = np.random.default_rng(77)
    ids = rng.choice(np.arange(1000, 5000), size=120, replace=False)
    lines = ["# a comment line", "#another\tcomment"]
    for _ in range(420):
        a, b = rng.choice(ids, size=2)
        lines.append("%d\t%d\t%s" % (a, b, repr(float(rng.integers(1, 4)))))
    edge_path = os.path.join(HERE, "cli_edges.tsv")
    with open(edge_path, "w") as f:
        f.write("\n".join(lines) + "\n")
    A, node_to_id = read_adjacency_matrix(file_path=edge_path, separator="\t", undirected=False)
    A = spsp.csr_matrix(A)
    coo = spsp.coo_matrix(A)
    np.savez_compressed(os.path.join(HERE, "cli_adjacency.npz"), row=coo.row, col=coo.col, data=coo.data,
                        n=A.shape[0], node_ids=np.array([node_to_id[i] for i in range(A.shape[0])]))
    A = (A + A.transpose()) / 2
    X = ref_arcte.arcte(adjacency_matrix=A, rho=RHO, epsilon=EPS, number_of_threads=1)
    X = spsp.csr_matrix(X)
    write_features(file_path=os.path.join(HERE, "cli_features.tsv"), features=X, separator="\t",
                   node_to_id=node_to_id)
    print("cli fixture: n=%d nnz=%d features nnz=%d" % (A.shape[0], A.nnz, X.nnz))


if __name__ == "__main__":
    if "--cli-only" not in sys.argv:
        main()
    if not [a for a in sys.argv[1:] if not a.startswith("--")]:
        cli_fixture()
