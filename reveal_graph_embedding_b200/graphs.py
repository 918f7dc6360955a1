"""Synthetic graphs of the shapes BASELINE.json names (no datasets can be fetched).

All generators are deterministic (numpy default_rng with a fixed seed), return a
canonical symmetric scipy CSR with unit weights and cost seconds on the host.
"""
import numpy as np
import scipy.sparse as sparse


def _symmetric_unit_csr(n, u, v):
    m = u != v
    u, v = u[m], v[m]
    rows = np.concatenate([u, v])
    cols = np.concatenate([v, u])
    A = sparse.coo_matrix((np.ones(rows.size, dtype=np.float64), (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.data[:] = 1.0
    A.sort_indices()
    return A


def planted_partition(n=419, groups=5, p_in=0.35, p_out=0.05, seed=419):
    """PoliticsUK-shape: small dense mention/follow graph with planted groups (config 2)."""
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, groups, size=n)
    P = np.where(lab[:, None] == lab[None, :], p_in, p_out)
    M = np.triu(rng.random((n, n)) < P, k=1)
    r, c = np.nonzero(M)
    return _symmetric_unit_csr(n, r.astype(np.int64), c.astype(np.int64))


def chung_lu(n, n_edges, gamma, max_degree, seed, min_degree_one=True):
    """Power-law graph: expected degrees w_i ~ (i + i0)^(-1/(gamma-1)), endpoints drawn
    proportionally to w (Chung-Lu).  With min_degree_one every node also gets one
    preferential edge, so no node is isolated (like the ASU social graphs)."""
    rng = np.random.default_rng(seed)
    alpha = 1.0 / (gamma - 1.0)
    target_sum = 2.0 * n_edges
    # choose i0 so that the largest expected degree is about max_degree
    lo, hi = 0.0, float(n)
    for _ in range(60):
        i0 = 0.5 * (lo + hi)
        w = (np.arange(n, dtype=np.float64) + 1.0 + i0) ** (-alpha)
        if w[0] / w.sum() * target_sum > max_degree:
            lo = i0
        else:
            hi = i0
    w = (np.arange(n, dtype=np.float64) + 1.0 + i0) ** (-alpha)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = rng.permutation(n)  # hubs scattered over the id range

    def draw(k):
        return perm[np.minimum(np.searchsorted(cdf, rng.random(k)), n - 1)]

    us, vs = [], []
    remaining = n_edges
    if min_degree_one:
        us.append(np.arange(n, dtype=np.int64))
        vs.append(draw(n))
        remaining -= n
    if remaining > 0:
        k = int(remaining * 1.03)  # head-room for duplicates / self loops
        us.append(draw(k))
        vs.append(draw(k))
    return _symmetric_unit_csr(n, np.concatenate(us), np.concatenate(vs))


def flickr_like(seed=80513):
    """ASU-Flickr shape: 80,513 nodes, ~5.9 M undirected edges, power-law degrees (config 3)."""
    return chung_lu(80513, 5_900_000, gamma=2.3, max_degree=5700, seed=seed)


def youtube_like(seed=1138499):
    """ASU-YouTube shape: 1,138,499 nodes, ~3 M undirected edges, hubs in the tens of
    thousands (config 4)."""
    return chung_lu(1_138_499, 2_990_000, gamma=2.2, max_degree=28000, seed=seed)


def rmat(scale=22, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=22):
    """Graph500-style R-MAT, symmetrised and de-duplicated (config 5)."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = edge_factor * n
    u = np.zeros(m, dtype=np.int64)
    v = np.zeros(m, dtype=np.int64)
    ab, abc = a + b, a + b + c
    for bit in range(scale):
        r = rng.random(m)
        u_bit = r >= ab
        v_bit = ((r >= a) & (r < ab)) | (r >= abc)
        u |= u_bit.astype(np.int64) << bit
        v |= v_bit.astype(np.int64) << bit
    return _symmetric_unit_csr(n, u, v)


def barabasi_albert(n, m, seed):
    """Preferential attachment without networkx (tests and small benchmarks)."""
    rng = np.random.default_rng(seed)
    targets = np.empty(2 * n * m, dtype=np.int64)
    us = np.empty(n * m, dtype=np.int64)
    vs = np.empty(n * m, dtype=np.int64)
    k = 0
    t = 0
    for i in range(m, n):
        if t == 0:
            chosen = np.arange(m)
        else:
            chosen = np.unique(targets[rng.integers(0, t, size=m)])
        for c in chosen:
            us[k], vs[k] = i, c
            k += 1
            targets[t] = i
            targets[t + 1] = c
            t += 2
    return _symmetric_unit_csr(n, us[:k], vs[:k])


WORKLOADS = {
    "politicsuk": planted_partition,
    "flickr": flickr_like,
    "youtube": youtube_like,
}
