"""GPU form of reveal_graph_embedding/eps_randomwalk/similarity.py -- the operator seam.

Same signatures as the reference: `s` and `r` are dense float64 vectors that must be
zero on entry and are filled in place; the return value is the number of pushes.
`w_i` / `a_i` are the reference's arrays-of-arrays (arcte.py:296-300); they are flattened
to CSR and uploaded once per (w_i, a_i) pair.
"""
import numpy as np

from ..engine import RULE_ABSORBING, RULE_LAZY, RULE_PAGERANK, get_engine

_cache = {"key": None}


def _ensure_graph(w_i, a_i, out_degree, in_degree):
    key = (id(w_i), id(a_i), id(out_degree), id(in_degree))
    eng = get_engine(0)
    if _cache["key"] != key or eng.n != len(a_i):
        lens = np.fromiter((len(a) for a in a_i), dtype=np.int64, count=len(a_i))
        indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        indices = np.concatenate([np.asarray(a) for a in a_i]).astype(np.int32) if indptr[-1] else np.zeros(0, np.int32)
        w = np.concatenate([np.asarray(x, dtype=np.float64) for x in w_i]) if indptr[-1] else np.zeros(0)
        eng.set_transition(indptr, indices, w, out_degree, in_degree)
        _cache["key"] = key
    return eng


def _run(rule, s, r, w_i, a_i, out_degree, in_degree, seed_node, rho, epsilon):
    eng = _ensure_graph(w_i, a_i, out_degree, in_degree)
    s_new, r_new, nop = eng.push(rule, seed_node, rho, epsilon)
    s[:] = s_new
    r[:] = r_new
    return nop


def fast_approximate_cumulative_pagerank_difference(s, r, w_i, a_i, out_degree, in_degree, seed_node,
                                                    rho=0.2, epsilon=0.00001):
    """similarity.py:149-222."""
    return _run(RULE_ABSORBING, s, r, w_i, a_i, out_degree, in_degree, seed_node, rho, epsilon)


def fast_approximate_personalized_pagerank(s, r, w_i, a_i, out_degree, in_degree, seed_node,
                                           rho=0.2, epsilon=0.00001):
    """similarity.py:11-63."""
    return _run(RULE_PAGERANK, s, r, w_i, a_i, out_degree, in_degree, seed_node, rho, epsilon)


def lazy_approximate_personalized_pagerank(s, r, w_i, a_i, out_degree, in_degree, seed_node,
                                           rho=0.2, epsilon=0.00001, laziness_factor=0.5):
    """similarity.py:66-146.  Only the reference's own laziness factor (0.5) is built in."""
    if laziness_factor != 0.5:
        raise ValueError("laziness_factor other than the reference's 0.5 is not supported")
    return _run(RULE_LAZY, s, r, w_i, a_i, out_degree, in_degree, seed_node, rho, epsilon)
