"""GPU form of reveal_graph_embedding/eps_randomwalk/transition.py."""
import numpy as np
import scipy.sparse as sparse

from ..engine import canonical_csr, get_engine


def get_natural_random_walk_matrix(adjacency_matrix, make_shared=False):
    """transition.py:43-99: returns (W as CSR, out_degree, in_degree).  `make_shared`
    asked for multiprocessing shared-memory copies (transition.py:70-97); there are no
    worker processes here, so it is accepted and ignored."""
    A = canonical_csr(adjacency_matrix)
    eng = get_engine(0)
    eng.set_graph(A, canonical=True)
    w, d_out, d_in = eng.transition()
    W = sparse.csr_matrix((w.copy(), A.indices.copy(), A.indptr.copy()), shape=A.shape)
    return W, d_out, d_in
