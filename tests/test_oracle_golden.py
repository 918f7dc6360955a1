"""Pins the CPU oracle (oracle/arcte_oracle.c) against fixtures produced by the
unmodified Python reference (tests/golden/make_golden.py).  Everything is
bit-exact except epsilon-effective, whose two logarithms go through numpy's
SIMD log on the generating host (<= 2 ulp allowed)."""
import numpy as np
import pytest

from helpers import EPS, GOLDEN_NAMES, RHO, assert_csr_identical, golden_features, load_golden, ulp_diff


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_transition_bit_exact(oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    assert np.array_equal(g.w, z["W_data"])
    assert np.array_equal(g.d_out, z["d_out"])
    assert np.array_equal(g.d_in, z["d_in"])


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_seed_set(oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    seeds = g.seeds()
    assert np.array_equal(np.sort(seeds), np.sort(z["seeds"]))
    cnt = g.column_counts()
    assert np.all(np.diff(cnt[seeds]) <= 0)  # degree-descending like arcte.py:614-617


def test_pairwise_sum_matches_numpy_mean(oracle):
    _, z = load_golden("ba300")
    off = 0
    for L, m in zip(z["pairwise_lens"], z["pairwise_means"]):
        v = z["pairwise_vals"][off:off + L]
        off += L
        assert oracle.pairwise_sum(v) / float(L) == m
        assert oracle.pairwise_sum(v) / float(L) == v.mean()  # and the numpy installed here


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_epsilon_effective(oracle, name):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    got = np.array([oracle.epsilon_effective(g, EPS, int(s)) for s in z["seeds"]])
    assert ulp_diff(got, z["eps_eff"]).max() <= 2


@pytest.mark.parametrize("rule", [0, 1, 2])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_push_drivers_bit_exact(oracle, name, rule):
    A, z = load_golden(name)
    g = oracle.Graph(A)
    rho = RHO if rule != 2 else (RHO * 0.5) / (1 - 0.5 * RHO)
    for k, (seed, eps) in enumerate(zip(z["probe_seeds"], z["probe_eps"])):
        s, r, nop, st = oracle.push(g, rule, int(seed), rho, float(eps))
        assert nop == z["probe_rule%d_nop" % rule][k]
        assert np.array_equal(s, z["probe_rule%d_s" % rule][k])
        assert np.array_equal(r, z["probe_rule%d_r" % rule][k])


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_arcte_features_identical(oracle, name, threads):
    A, z = load_golden(name)
    n = A.shape[0]
    # use the reference's own epsilon-effective so the comparison is exact by construction
    g = oracle.Graph(A)
    order = {int(s): i for i, s in enumerate(z["seeds"])}
    seeds = g.seeds()
    ov = np.array([z["eps_eff"][order[int(s)]] for s in seeds])
    for rule in (0, 1, 2):
        if "X%d_data" % rule not in z:
            continue
        sd, seg, mem, eff, st = oracle.extract(g, rule, RHO, EPS, seeds, threads, eps_override=ov)
        X = oracle.assemble(g, sd, seg, mem)
        assert_csr_identical(X, golden_features(z, rule, n))
        assert X.dtype == np.float64 and X.has_sorted_indices


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_arcte_end_to_end_own_epsilon(oracle, name):
    """With the oracle's own epsilon-effective (libm log) the features still match."""
    A, z = load_golden(name)
    X = oracle.arcte(A, RHO, EPS, number_of_threads=2)
    assert_csr_identical(X, golden_features(z, 0, A.shape[0]))
