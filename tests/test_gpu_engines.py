"""The four engines of the exact FIFO schedule (include/arcte_cuda.h: ARCTE_ENGINE_*) against the
reference fixtures and the oracle.  All of them must be BIT-EXACT: the batched engines (shared-memory
staging of several queue entries per warp iteration, dense or hashed walk state) change how the state
is laid out and fetched, not the arithmetic nor its order (similarity.py:149-222, push.py:41-64,
arcte.py:328-376).
"""
import numpy as np
import pytest
import scipy.sparse as sparse

from helpers import EPS, GOLDEN_NAMES, RHO, assert_csr_identical, golden_features, load_golden

pytestmark = pytest.mark.gpu

ENGINES = ["fifo", "dense", "hash", "compact"]
COUNTERS = ("pushes", "enqueues", "support", "members", "emitted", "max_queue", "seed_degree")


@pytest.fixture(scope="module")
def eng():
    from reveal_graph_embedding_b200.engine import Engine
    e = Engine(0)
    yield e
    e.set_engine("auto")
    e.close()


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_push_vectors_bit_exact_vs_reference(eng, name, engine):
    """Dense s, r and the push count of the absorbing driver, probe seeds of every fixture."""
    A, z = load_golden(name)
    eng.set_engine(engine)
    eng.set_graph(A)
    for k, (seed, eps) in enumerate(zip(z["probe_seeds"], z["probe_eps"])):
        s, r, nop = eng.push(0, int(seed), RHO, float(eps))
        assert nop == z["probe_rule0_nop"][k]
        assert np.array_equal(s, z["probe_rule0_s"][k])
        assert np.array_equal(r, z["probe_rule0_r"][k])


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_features_identical_to_reference(eng, name, engine):
    A, z = load_golden(name)
    n = A.shape[0]
    eng.set_engine(engine)
    eng.set_graph(A)
    seeds = eng.seeds()
    order = {int(s): i for i, s in enumerate(z["seeds"])}
    ov = np.array([z["eps_eff"][order[int(s)]] for s in seeds])
    eng.extract(0, RHO, EPS, eps_override=ov)
    assert eng.stats()["engine"] == {"fifo": 0, "dense": 1, "hash": 2, "compact": 3}[engine]
    eng.assemble()
    assert_csr_identical(eng.features(), golden_features(z, 0, n))


@pytest.mark.parametrize("engine", ["dense", "hash", "compact"])
@pytest.mark.parametrize("eps", [1e-3, 1e-5, 1e-7])
def test_raw_epsilon_long_walks_vs_oracle(eng, oracle, engine, eps):
    """Raw epsilon: long walks, repeated pushes of the same node, duplicates in the queue, deep FIFO."""
    from reveal_graph_embedding_b200 import graphs
    A = graphs.barabasi_albert(4000, 4, seed=2)
    g = oracle.Graph(A)
    eng.set_engine(engine)
    eng.set_graph(A)
    for seed in (0, 17, 3999):
        s, r, nop = eng.push(0, seed, RHO, eps)
        so, ro, nopo, st = oracle.push(g, 0, seed, RHO, eps)
        assert nop == nopo
        assert np.array_equal(s, so) and np.array_equal(r, ro)


@pytest.mark.parametrize("engine", ["dense", "hash", "compact"])
def test_counters_and_matrix_vs_oracle(eng, oracle, engine):
    from reveal_graph_embedding_b200 import graphs
    A = graphs.barabasi_albert(20000, 3, 2)
    g = oracle.Graph(A)
    eng.set_engine(engine)
    eng.set_graph(A)
    seeds = eng.seeds()
    eps_dev = eng.epsilon_effective(EPS, seeds)
    sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8, eps_override=eps_dev)
    eng.extract(0, RHO, EPS)
    eng.assemble()
    assert_csr_identical(eng.features(), oracle.assemble(g, sd, seg, mem))
    gs = eng.stats()
    for k in COUNTERS:
        assert gs[k] == st[k], k
    assert gs["edge_touches"] == st["edges"]


@pytest.mark.parametrize("engine", ["dense", "hash", "compact"])
def test_hub_rows_self_loops_and_real_weights(eng, oracle, engine):
    """Rows longer than a batch (the cp.async staged path), a hub with a self loop, asymmetric real
    weights (the per-entry weight path instead of the uniform-row one), duplicates inside a batch."""
    from reveal_graph_embedding_b200 import graphs
    rng = np.random.default_rng(11)
    A = graphs.chung_lu(6000, 40000, 2.1, 2500, seed=5).tolil()
    A[0, 0] = 1.0
    A[7, 7] = 1.0
    A = A.tocsr()
    for weighted in (False, True):
        B = A.copy()
        if weighted:
            B.data = rng.uniform(0.2, 4.0, B.nnz)
        g = oracle.Graph(B)
        eng.set_engine(engine)
        eng.set_graph(B)
        seeds = eng.seeds()
        assert np.diff(B.indptr).max() > 600
        eps_dev = eng.epsilon_effective(EPS, seeds)
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8, eps_override=eps_dev)
        eng.extract(0, RHO, EPS)
        eng.assemble()
        assert_csr_identical(eng.features(), oracle.assemble(g, sd, seg, mem))
        gs = eng.stats()
        for k in COUNTERS:
            assert gs[k] == st[k], (k, weighted)
        for seed in (int(seeds[0]), int(seeds[seeds.size // 2])):
            s, r, nop = eng.push(0, seed, RHO, 1e-6)
            so, ro, nopo, _ = oracle.push(g, 0, seed, RHO, 1e-6)
            assert nop == nopo and np.array_equal(s, so) and np.array_equal(r, ro)


def test_table_growth_and_region_overflow_fall_back(oracle):
    """A tiny table region: walks outgrow it, are undone and re-run by the dense FIFO engine; a tiny ring
    does the same through the ring-overflow path.  Results identical, slots left clean (second run equal)."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    A = graphs.barabasi_albert(5000, 5, 7)
    g = oracle.Graph(A)
    e = Engine(0)
    try:
        e.set_engine("hash", table_capacity=1024)
        e.set_graph(A)
        seeds = e.seeds()
        eps_dev = e.epsilon_effective(EPS, seeds)
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8, eps_override=eps_dev)
        want = oracle.assemble(g, sd, seg, mem)
        for _ in range(2):
            e.extract(0, RHO, EPS)
            assert e.stats()["retries"] > 0
            e.assemble()
            assert_csr_identical(e.features(), want)
        for engine in ("hash", "dense", "compact"):
            e.set_engine(engine)
            e.configure(queue_capacity=64)
            for _ in range(2):
                e.extract(0, RHO, EPS)
                assert e.stats()["retries"] > 0
                e.assemble()
                assert_csr_identical(e.features(), want)
            e.configure()
            e.extract(0, RHO, EPS)
            assert e.stats()["retries"] == 0
            e.assemble()
            assert_csr_identical(e.features(), want)
    finally:
        e.close()


@pytest.mark.parametrize("engine", ["dense", "hash", "compact"])
def test_rmat_sample_vs_oracle(eng, oracle, engine):
    """BASELINE.json config 5's graph family (R-MAT a,b,c = 0.57,0.19,0.19, edge factor 16, symmetrised):
    extreme hubs, long FIFOs.  Degree-stratified seed sample walked on the whole graph by both sides."""
    from reveal_graph_embedding_b200 import graphs
    A = graphs.rmat(18, 16, seed=22)
    g = oracle.Graph(A)
    eng.set_engine(engine)
    eng.set_graph(A)
    seeds = eng.seeds()
    idx = np.unique(np.linspace(0, seeds.size - 1, 400).astype(np.int64))
    sample = seeds[idx]
    eps_dev = eng.epsilon_effective(EPS, sample)
    eng.set_seeds(sample)
    eng.extract(0, RHO, EPS)
    seg_seed, seg_cnt, seg_off, mem = eng.segments()
    sd, seg, omem, eff, st = oracle.extract(g, 0, RHO, EPS, sample, 8, eps_override=eps_dev)
    assert np.array_equal(seg_seed, sample) and np.array_equal(seg_cnt, seg)
    o = 0
    for i in range(sample.size):
        c = int(seg[i])
        assert np.array_equal(np.sort(mem[seg_off[i]:seg_off[i] + c]), np.sort(omem[o:o + c])), "seed %d" % sample[i]
        o += c
    gs = eng.stats()
    for k in COUNTERS:
        assert gs[k] == st[k], k
    assert gs["edge_touches"] == st["edges"]


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_compact_engine_pagerank_and_lazy_rules(eng, name):
    """The compact-state engine implements all three push rules: PageRank (similarity.py:11-63) and lazy PageRank
    (:66-146) vectors and the guarded feature matrices of arcte_with_pagerank / arcte_with_lazy_pagerank against
    the reference fixtures, bit for bit."""
    A, z = load_golden(name)
    n = A.shape[0]
    eng.set_engine("compact")
    eng.set_graph(A)
    for rule in (1, 2):
        rho = RHO if rule != 2 else (RHO * 0.5) / (1.0 - 0.5 * RHO)   # arcte.py:109
        for k, (seed, eps) in enumerate(zip(z["probe_seeds"], z["probe_eps"])):
            s, r, nop = eng.push(rule, int(seed), rho, float(eps))
            assert nop == z["probe_rule%d_nop" % rule][k]
            assert np.array_equal(s, z["probe_rule%d_s" % rule][k])
            assert np.array_equal(r, z["probe_rule%d_r" % rule][k])
    seeds = eng.seeds()
    order = {int(s): i for i, s in enumerate(z["seeds"])}
    ov = np.array([z["eps_eff"][order[int(s)]] for s in seeds])
    for rule in (1, 2):
        if ("X%d_indptr" % rule) not in z.files:
            continue
        eng.extract(rule, RHO, EPS, eps_override=ov)
        assert eng.stats()["engine"] == 3
        eng.assemble()
        assert_csr_identical(eng.features(), golden_features(z, rule, n))


def test_compact_engine_epoch_wrap(oracle):
    """Few slots, many walks per slot, three extractions on the same pool: the slots' epochs carry over from walk to
    walk and from call to call, and nothing is ever reset in between (the wrap of the epoch field, 2^19 walks of
    one slot on this graph, is covered by ARCTE_CUDA_COMPACT_EPOCH_BITS=3 in test_compact_engine_forced_epoch_wrap)."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    A = graphs.barabasi_albert(5000, 5, 7)
    g = oracle.Graph(A)
    e = Engine(0)
    try:
        e.set_engine("compact")
        e.configure(warps_per_sm=1)
        e.set_graph(A)
        seeds = e.seeds()
        eps_dev = e.epsilon_effective(EPS, seeds)
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8, eps_override=eps_dev)
        want = oracle.assemble(g, sd, seg, mem)
        for _ in range(3):   # the slots' epochs carry over from one extraction to the next
            e.extract(0, RHO, EPS)
            e.assemble()
            assert_csr_identical(e.features(), want)
    finally:
        e.close()


def test_compact_engine_forced_epoch_wrap(oracle, monkeypatch):
    """Three epoch bits: every slot clears its index map after seven walks.  Same matrix as the oracle's."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    monkeypatch.setenv("ARCTE_CUDA_COMPACT_EPOCH_BITS", "3")
    A = graphs.barabasi_albert(3000, 4, 3)
    g = oracle.Graph(A)
    e = Engine(0)
    try:
        e.set_engine("compact")
        e.configure(warps_per_sm=1)
        e.set_graph(A)
        seeds = e.seeds()
        eps_dev = e.epsilon_effective(EPS, seeds)
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8, eps_override=eps_dev)
        want = oracle.assemble(g, sd, seg, mem)
        for _ in range(2):
            e.extract(0, RHO, EPS)
            e.assemble()
            assert_csr_identical(e.features(), want)
        s, r, nop = e.push(0, int(seeds[0]), RHO, 1e-6)
        so, ro, nopo, _ = oracle.push(g, 0, int(seeds[0]), RHO, 1e-6)
        assert nop == nopo and np.array_equal(s, so) and np.array_equal(r, ro)
    finally:
        e.close()


def test_compact_engine_capacity_overflow_retry(oracle, monkeypatch):
    """64 pairs per slot: almost every walk outgrows its compact arrays, is abandoned and re-run in the retry pass
    with room for all n nodes (extract) / at once (push).  Same matrix, vectors and counters as the oracle's."""
    from reveal_graph_embedding_b200 import graphs
    from reveal_graph_embedding_b200.engine import Engine
    monkeypatch.setenv("ARCTE_CUDA_COMPACT_CAP", "64")
    A = graphs.barabasi_albert(3000, 4, 3)
    g = oracle.Graph(A)
    e = Engine(0)
    try:
        e.set_engine("compact")
        e.set_graph(A)
        seeds = e.seeds()
        eps_dev = e.epsilon_effective(EPS, seeds)
        sd, seg, mem, eff, st = oracle.extract(g, 0, RHO, EPS, seeds, 8, eps_override=eps_dev)
        want = oracle.assemble(g, sd, seg, mem)
        for _ in range(2):
            e.extract(0, RHO, EPS)
            gs = e.stats()
            assert gs["retries"] > 0
            for k in COUNTERS:
                assert gs[k] == st[k], k
            e.assemble()
            assert_csr_identical(e.features(), want)
        s, r, nop = e.push(0, int(seeds[0]), RHO, 1e-6)
        so, ro, nopo, _ = oracle.push(g, 0, int(seeds[0]), RHO, 1e-6)
        assert nop == nopo and np.array_equal(s, so) and np.array_equal(r, ro)
    finally:
        e.close()
