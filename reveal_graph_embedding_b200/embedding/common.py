"""Column normalisation of community features on the GPU -- drop-in for
reveal_graph_embedding/embedding/common.py:49-67 of the reference (the step every experiment
applies right after arcte(), experiments/utility.py:66-69).

    normalize_columns(features) -> scipy.sparse.csr_matrix

The reference loops over all columns in Python (2n `getcol` calls for an n x 2n ARCTE
matrix); here it is one histogram and one scaling pass over the stored entries on the
device.  There is no CPU path.
"""
from ..engine import get_engine


def normalize_columns(features):
    """Divide every column that has more than one stored entry by sqrt(log(stored entries))
    (common.py:59-63).  Returns a new canonical CSR; the argument is not modified."""
    return get_engine(0).normalize_columns(features)
