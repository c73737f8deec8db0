"""numpy/scipy restatement of ``sklearn.metrics.pairwise.cosine_similarity``.  Test infrastructure.

The reference's whole path is calls to this third-party function
(ml/similarity_computer.py:41,58,86; scripts/populate_database.py:180-186).  Its published
algorithm (scikit-learn 1.7.2, sklearn/metrics/pairwise.py ``cosine_similarity``):

1. ``check_pairwise_arrays``: both operands become float32 only if BOTH are float32, else
   float64 (``_return_float_dtype``) -- int64 / bool / float64 inputs are all promoted to float64.
2. ``normalize(X, norm="l2", copy=True)``: dense -> ``X / sqrt(einsum('ij,ij->i', X, X))`` with
   zero norms replaced by 1; CSR -> per-row ``sqrt(sum(x*x))``, rows with zero norm left alone
   (``inplace_csr_row_normalize_l2``).
3. ``safe_sparse_dot(Xn, Yn.T, dense_output=True)``.

Consequences pinned by tests/test_oracle.py: zero rows give similarity 0 everywhere, diagonal
included, never NaN; non-negative inputs give scores in [0, 1 + eps].
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _float_dtype(X, Y):
    xd = getattr(X, "dtype", None)
    yd = getattr(Y, "dtype", None)
    if xd == np.float32 and (Y is None or yd == np.float32):
        return np.float32
    return np.float64


def normalize_rows(X, dtype=np.float64):
    """Step 2: row-wise L2 normalisation, returns a new array / CSR matrix."""
    if sp.issparse(X):
        Xc = sp.csr_matrix(X, dtype=dtype, copy=True)
        sq = np.asarray(Xc.multiply(Xc).sum(axis=1)).ravel()
        norms = np.sqrt(sq)
        norms[norms == 0.0] = 1.0
        # divide each stored value by its row norm (what inplace_csr_row_normalize_l2 does)
        Xc.data /= np.repeat(norms, np.diff(Xc.indptr))
        return Xc
    Xa = np.array(X, dtype=dtype, copy=True, ndmin=2)
    norms = np.sqrt(np.einsum("ij,ij->i", Xa, Xa))
    norms[norms == 0.0] = 1.0
    Xa /= norms[:, None]
    return Xa


def cosine_similarity(X, Y=None) -> np.ndarray:
    """Dense [n_x, n_y] cosine similarity; float64 unless every input is float32."""
    dtype = _float_dtype(X, Y)
    Xn = normalize_rows(X, dtype)
    Yn = Xn if Y is None or Y is X else normalize_rows(Y, dtype)
    out = Xn @ Yn.T
    if sp.issparse(out):
        out = out.toarray()
    return np.asarray(out)
