"""The library's second walk schedule (csrc/push_frontier.cu: synchronous frontier rounds on
fixed-point state, opt-in) against the reference.

What is promised and tested here (BASELINE.json north_star; SURVEY.md section 8a, error-bound note):
  * per entry:  0 <= (G_seed - s)[x]/d[x] < eps (1-rho)/rho  and  r[x]/d[x] < eps at the end,
    hence |s_frontier - s_reference|[x]/d[x] < eps (1-rho)/rho with eps the per-seed
    epsilon-effective;
  * the thresholded support equals the reference's except for DOCUMENTED TIES: an entry
    (x, seed) present in only one of the two results must have |q[x] - tau| below that bound in
    at least one of them (q = s/d_in, tau = the seed's threshold, arcte.py:355-367); when a seed
    emits a community in only one of them (arcte.py:370), every member outside the base
    community must be such a tie;
  * the schedule is deterministic: the GPU result is bit-identical to its CPU restatement
    (oracle_push_frontier) and independent of launch geometry and seed sharding.
The CPU tests pin the schedule itself against the reference fixtures; the GPU tests pin the CUDA
implementation against the CPU restatement bit for bit.
"""
import numpy as np
import pytest
import scipy.sparse as sparse
import scipy.sparse.linalg as spla

from helpers import EPS, RHO, assert_csr_identical, golden_features, load_golden

BOUND = (1.0 - RHO) / RHO
NAMES = ["ba300", "weighted200", "edgecases160", "ba2000"]


def classify_against_reference(oracle, A, z, X_frontier, s_frontier_of):
    """Every difference between the reference's local block and the frontier schedule's must be a
    documented tie.  s_frontier_of(seed, eps) -> dense s of the frontier schedule.
    Returns (differing entries, union size, seeds that differ)."""
    n = A.shape[0]
    g = oracle.Graph(A)
    ref = golden_features(z, 0, n)
    assert (ref[:, :n] != X_frontier[:, :n]).nnz == 0          # base block: always identical
    Lr, Lf = ref[:, n:].tocsc(), sparse.csr_matrix(X_frontier)[:, n:].tocsc()
    eps_of = dict(zip(z["seeds"].tolist(), z["eps_eff"].tolist()))
    differing = union = seeds_diff = 0
    for seed in z["seeds"].tolist():
        a = set(Lr.indices[Lr.indptr[seed]:Lr.indptr[seed + 1]].tolist())
        b = set(Lf.indices[Lf.indptr[seed]:Lf.indptr[seed + 1]].tolist())
        union += len(a | b)
        if a == b:
            continue
        seeds_diff += 1
        eps = eps_of[seed]
        band = eps * BOUND
        s_ref, _, _, _ = oracle.push(g, 0, seed, RHO, eps)       # bit-identical to the reference (golden-pinned)
        s_fr = s_frontier_of(seed, eps)
        base = np.append(A.indices[A.indptr[seed]:A.indptr[seed + 1]], seed)
        q_ref, q_fr = s_ref / g.d_in, s_fr / g.d_in
        tau_ref, tau_fr = q_ref[base].min(), q_fr[base].min()

        def tie(x):
            return abs(q_ref[x] - tau_ref) < band or abs(q_fr[x] - tau_fr) < band
        if bool(a) != bool(b):                                   # emitted by one side only (arcte.py:370)
            extra = (a | b) - set(base.tolist())
            assert extra and all(tie(x) for x in extra), "seed %d: emission differs outside the band" % seed
        else:
            assert all(tie(x) for x in a ^ b), "seed %d: support differs outside the band" % seed
        differing += len(a ^ b)
    return differing, union, seeds_diff


# ------------------------------------------------------------------ CPU: the schedule itself
@pytest.mark.parametrize("name", NAMES)
def test_schedule_error_bound_and_termination(oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    n = A.shape[0]
    W = sparse.csr_matrix((g.w, A.indices, A.indptr), shape=(n, n))
    M = (sparse.identity(n, format="csc") - (1.0 - RHO) * W.T.tocsc()).tocsc()
    lu = spla.splu(M)
    for seed, eps, s_ref in zip(z["probe_seeds"], z["probe_eps"], z["probe_rule0_s"]):
        with oracle.schedule(oracle.SCHEDULE_FRONTIER):
            s, r, nop, st = oracle.push(g, 0, int(seed), RHO, float(eps))
        e = np.zeros(n)
        e[int(seed)] = 1.0
        G = lu.solve(e)
        d = np.where(g.d_in > 0, g.d_in, 1.0)
        gap = (G - s) / d
        assert gap.min() > -1e-12 and gap.max() < eps * BOUND      # SURVEY 8a error bound
        assert np.all(r / d < eps)                                 # similarity.py:204 stopping rule
        assert (np.abs(s - s_ref) / d).max() < eps * BOUND         # against the reference's own s
        assert nop == st["pushes"] > 0


@pytest.mark.parametrize("name", NAMES)
def test_schedule_support_equals_reference_up_to_documented_ties(oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    with oracle.schedule(oracle.SCHEDULE_FRONTIER):
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, z["seeds"], 4, eps_override=z["eps_eff"])
        X = oracle.assemble(g, sd, seg, mem)

    def s_frontier_of(seed, eps):
        with oracle.schedule(oracle.SCHEDULE_FRONTIER):
            return oracle.push(g, 0, seed, RHO, eps)[0]
    differing, union, seeds_diff = classify_against_reference(oracle, A, z, X, s_frontier_of)
    assert differing <= 0.05 * max(union, 1)                       # and they are few


# ------------------------------------------------------------------ GPU: the CUDA implementation
@pytest.fixture(scope="module")
def feng():
    from reveal_graph_embedding_b200.engine import Engine
    e = Engine(0)
    e.set_schedule("frontier")
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_push_bit_identical_to_restatement(feng, oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    feng.set_graph(A)
    for seed, eps in zip(z["probe_seeds"], z["probe_eps"]):
        s, r, nop = feng.push(0, int(seed), RHO, float(eps))
        with oracle.schedule(oracle.SCHEDULE_FRONTIER):
            so, ro, nopo, st = oracle.push(g, 0, int(seed), RHO, float(eps))
        assert np.array_equal(s, so) and np.array_equal(r, ro) and nop == nopo


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES + ["planted419"])
def test_gpu_features_bit_identical_to_restatement_and_in_band_of_reference(feng, oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    feng.set_graph(A)
    feng.set_seeds(z["seeds"])     # the fixture's order among equal counts (eps_override is positional)
    feng.extract(0, RHO, EPS, eps_override=z["eps_eff"])
    feng.assemble()
    X = feng.features()
    with oracle.schedule(oracle.SCHEDULE_FRONTIER):
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, z["seeds"], 4, eps_override=z["eps_eff"])
        Xo = oracle.assemble(g, sd, seg, mem)
    assert_csr_identical(X, Xo)
    gs = feng.stats()
    for k in ("pushes", "enqueues", "support", "members", "emitted", "max_queue", "seed_degree"):
        assert gs[k] == st[k], k
    assert gs["edge_touches"] == st["edges"] and gs["rounds"] > 0
    classify_against_reference(oracle, A, z, X, lambda seed, eps: feng.push(0, seed, RHO, eps)[0])


@pytest.mark.gpu
def test_gpu_frontier_medium_graph_and_geometry_independence(oracle):
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    A = graphs.barabasi_albert(20000, 3, seed=2)
    g = oracle.Graph(A)
    e0 = Engine(0)
    e0.set_graph(A)
    seeds = e0.seeds()
    e0.close()
    with oracle.schedule(oracle.SCHEDULE_FRONTIER):
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8)
        Xo = oracle.assemble(g, sd, seg, mem)
    results = []
    for geom in (dict(), dict(heavy_permille=0, light_threads=128, light_ctas_per_sm=8),
                 dict(heavy_permille=1000, heavy_threads=1024, heavy_ctas_per_sm=1),
                 dict(heavy_permille=500, heavy_threads=256, heavy_ctas_per_sm=3, light_threads=512, light_ctas_per_sm=2)):
        e = Engine(0)
        e.set_schedule("frontier", **geom)
        e.set_graph(A)
        e.extract(0, RHO, EPS, eps_override=eff)
        e.assemble()
        results.append(e.features())
        gs = e.stats()
        assert gs["pushes"] == st["pushes"] and gs["members"] == st["members"] and gs["edge_touches"] == st["edges"]
        e.close()
    for X in results:
        assert_csr_identical(X, Xo)
    # seed sharding (3 parts, as 3 GPUs would): same matrix
    engines, parts = [], []
    for r in range(3):
        e = Engine(0)
        e.set_schedule("frontier")
        e.set_graph(A)
        ns, nm = e.extract(0, RHO, EPS, shard_rank=r, shard_count=3, eps_override=eff)
        parts.append((ns, nm) + e.segments_device())
        engines.append(e)
    engines[0].assemble(parts)
    assert_csr_identical(engines[0].features(), Xo)
    for e in engines:
        e.close()


@pytest.mark.gpu
def test_gpu_frontier_member_buffer_overflow_is_retried(oracle):
    from reveal_graph_embedding_b200.engine import Engine
    A, z = load_golden("ba2000")
    g = oracle.Graph(A)
    with oracle.schedule(oracle.SCHEDULE_FRONTIER):
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, z["seeds"], 4, eps_override=z["eps_eff"])
        Xo = oracle.assemble(g, sd, seg, mem)
    e = Engine(0)
    e.configure(member_capacity=1 << 10)          # far too small: forces retry passes
    e.set_schedule("frontier")
    e.set_graph(A)
    e.set_seeds(z["seeds"])
    e.extract(0, RHO, EPS, eps_override=z["eps_eff"])
    assert e.stats()["retries"] > 0
    e.assemble()
    assert_csr_identical(e.features(), Xo)
    e.close()


@pytest.mark.gpu
def test_gpu_frontier_rejects_other_rules(feng):
    from reveal_graph_embedding_b200.engine import ArcteCudaError
    A, z = load_golden("ba300")
    feng.set_graph(A)
    with pytest.raises(ArcteCudaError, match="absorbing"):
        feng.extract(1, RHO, EPS)


@pytest.mark.gpu
def test_gpu_frontier_full_size_sample(oracle):
    """YouTube shape (n = 1,138,499): 1200 degree-stratified seeds on the full graph, bit for bit."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    A = graphs.youtube_like()
    g = oracle.Graph(A)
    e = Engine(0)
    e.set_schedule("frontier")
    e.set_graph(A)
    seeds = e.seeds()
    sample = seeds[np.unique(np.linspace(0, seeds.size - 1, 1200).astype(np.int64))]
    eps_dev = e.epsilon_effective(EPS, sample)
    e.set_seeds(sample)
    e.extract(0, RHO, EPS)
    seg_seed, seg_cnt, seg_off, mem = e.segments()
    with oracle.schedule(oracle.SCHEDULE_FRONTIER):
        sd, seg, omem, eff, st = oracle.extract(g, 0, RHO, EPS, sample, 8, eps_override=eps_dev)
    assert np.array_equal(seg_cnt, seg)
    o = 0
    for i in range(sample.size):
        c = int(seg[i])
        assert np.array_equal(np.sort(mem[seg_off[i]:seg_off[i] + c]), np.sort(omem[o:o + c])), "seed %d" % sample[i]
        o += c
    gs = e.stats()
    assert gs["pushes"] == st["pushes"] and gs["edge_touches"] == st["edges"] and gs["members"] == st["members"]
    e.close()


@pytest.mark.gpu
def test_auto_schedule_picks_by_seed_count_and_rule(oracle):
    """ARCTE_CUDA_SCHEDULE=auto / set_default_schedule("auto"): frontier for few seeds and the
    absorbing rule, FIFO otherwise; the FIFO default is untouched afterwards."""
    from reveal_graph_embedding_b200 import engine, graphs
    from reveal_graph_embedding_b200.embedding.arcte.arcte import arcte, arcte_with_pagerank
    A, z = load_golden("ba2000")
    try:
        engine.set_default_schedule("auto")
        X = arcte(A, RHO, EPS, 1)
        eng = engine.get_engine(0)
        assert eng.stats()["rounds"] > 0                       # walked by the frontier schedule
        g = oracle.Graph(A)
        with oracle.schedule(oracle.SCHEDULE_FRONTIER):
            sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, eng.seeds(), 4)
            assert_csr_identical(X, oracle.assemble(g, sd, seg, mem))
        Xp = arcte_with_pagerank(A, RHO, EPS, 1)               # PageRank rule: FIFO under "auto"
        assert eng.stats()["rounds"] == 0
        big = graphs.barabasi_albert(50000, 2, seed=3)         # > 40,000 seeds: FIFO
        arcte(big, RHO, EPS, 1)
        assert eng.stats()["rounds"] == 0 and eng.stats()["n_seeds_shard"] > engine.AUTO_FRONTIER_MAX_SEEDS
    finally:
        engine.set_default_schedule(None)
    X0 = arcte(A, RHO, EPS, 1)
    assert_csr_identical(X0, golden_features(z, 0, A.shape[0]))  # back to the exact default


def _classify_sample_against_fifo_oracle(oracle, A, engine, k):
    """Frontier schedule on the GPU against the reference's FIFO order (oracle, golden-pinned) on a degree-
    stratified sample of k seeds walked on the whole graph: every entry present in only one of the two
    communities must be a documented tie (|q - tau| < eps_eff (1-rho)/rho in at least one of them), with the
    ORACLE's epsilon-effective.  Returns (differing entries, union size, differing seeds, seeds)."""
    g = oracle.Graph(A)
    engine.set_graph(A)
    seeds = engine.seeds()
    sample = seeds[np.unique(np.linspace(0, seeds.size - 1, k).astype(np.int64))]
    eps_ora = np.array([oracle.epsilon_effective(g, EPS, int(s)) for s in sample])
    engine.set_seeds(sample)
    all_eps = np.zeros(sample.size)
    all_eps[:] = eps_ora
    engine.extract(0, RHO, EPS, eps_override=all_eps)
    seg_seed, seg_cnt, seg_off, mem = engine.segments()
    assert np.array_equal(seg_seed, sample)
    sd, seg, omem, eff, st = oracle.extract(g, 0, RHO, EPS, sample, 8, eps_override=eps_ora)   # FIFO, the reference's order
    differing = union = seeds_diff = 0
    o = 0
    for i, seed in enumerate(sample.tolist()):
        a = set(omem[o:o + int(seg[i])].tolist())
        o += int(seg[i])
        b = set(mem[seg_off[i]:seg_off[i] + max(int(seg_cnt[i]), 0)].tolist())
        union += len(a | b)
        if a == b:
            continue
        seeds_diff += 1
        eps = float(eps_ora[i])
        band = eps * BOUND
        s_ref = oracle.push(g, 0, seed, RHO, eps)[0]
        with oracle.schedule(oracle.SCHEDULE_FRONTIER):
            s_fr = oracle.push(g, 0, seed, RHO, eps)[0]         # bit-identical to the GPU's frontier walk (tested above)
        base = np.append(A.indices[A.indptr[seed]:A.indptr[seed + 1]], seed)
        nz = np.union1d(np.flatnonzero(s_ref), np.flatnonzero(s_fr))
        q_ref = dict(zip(nz.tolist(), (s_ref[nz] / g.d_in[nz]).tolist()))
        q_fr = dict(zip(nz.tolist(), (s_fr[nz] / g.d_in[nz]).tolist()))
        tau_ref = min(q_ref.get(int(x), 0.0) for x in base)
        tau_fr = min(q_fr.get(int(x), 0.0) for x in base)

        def tie(x):
            return abs(q_ref.get(x, 0.0) - tau_ref) < band or abs(q_fr.get(x, 0.0) - tau_fr) < band
        if bool(a) != bool(b):                                   # emitted by one side only (arcte.py:370)
            extra = (a | b) - set(base.tolist())
            assert extra and all(tie(x) for x in extra), "seed %d: emission differs outside the band" % seed
        else:
            assert all(tie(x) for x in a ^ b), "seed %d: support differs outside the band" % seed
        differing += len(a ^ b)
    return differing, union, seeds_diff, sample.size


@pytest.mark.gpu
@pytest.mark.parametrize("shape,k", [("youtube", 1000), ("flickr", 150)])
def test_gpu_frontier_full_size_differences_are_documented_ties(oracle, shape, k):
    """BASELINE.json configs 3 and 4 at full size: the tolerance-parity schedule against the reference's own
    order (the full-size frontier test compared the GPU with its own restatement only)."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    A = graphs.youtube_like() if shape == "youtube" else graphs.flickr_like()
    e = Engine(0)
    try:
        e.set_schedule("frontier")
        differing, union, seeds_diff, n_seeds = _classify_sample_against_fifo_oracle(oracle, A, e, k)
    finally:
        e.close()
    assert differing <= 0.02 * max(union, 1)     # measured: about half a percent of the entries
    assert seeds_diff <= 0.25 * n_seeds
