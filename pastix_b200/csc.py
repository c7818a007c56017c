"""Internal (permuted) CSC of the numeric phase.

Mirror of what CscOrdistrib leaves in SopalinParam.cscmtx / .transcsc
(src/sopalin/src/csc_intern_build.c:352-560): the matrix in the NEW ordering,
0-based, rows sorted inside each column; symmetric/hermitian inputs (lower
triangle given) are expanded to both triangles; for unsymmetric inputs
`tvalues` holds the values of A^T on the same pattern (the pattern must be
symmetric, as the reference requires for LU).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def internal_csc(A: sp.spmatrix, permtab: np.ndarray, sym: str, dtype=None):
    """sym in {'yes','her','no'}; permtab[old] = new (0-based).
    Returns dict(colptr, rows, values, tvalues|None)."""
    A = sp.coo_matrix(A)
    dtype = np.dtype(dtype or A.dtype)
    r = permtab[A.row]; c = permtab[A.col]; v = A.data.astype(dtype)
    n = A.shape[0]
    if sym in ("yes", "her"):
        off = A.row != A.col
        vo = np.conj(v[off]) if sym == "her" else v[off]
        r, c, v = np.concatenate([r, c[off]]), np.concatenate([c, r[off]]), np.concatenate([v, vo])
    # sort by (col, row) without summing duplicates away silently
    order = np.lexsort((r, c))
    r, c, v = r[order], c[order], v[order]
    if r.size > 1 and np.any((r[1:] == r[:-1]) & (c[1:] == c[:-1])):
        raise ValueError("duplicate entries in the matrix")
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, c + 1, 1)
    colptr = np.cumsum(colptr)
    out = dict(colptr=colptr, rows=r.astype(np.int64), values=np.ascontiguousarray(v), tvalues=None)
    if sym == "no":
        # value of A^T at (row r, col c) = A[c, r]: look (c, r) up in the sorted (col,row) list
        key = c.astype(np.int64) * n + r
        tkey = r.astype(np.int64) * n + c
        pos = np.searchsorted(key, tkey)
        if np.any(pos >= key.size) or np.any(key[np.minimum(pos, key.size - 1)] != tkey):
            raise ValueError("LU needs a structurally symmetric pattern")
        out["tvalues"] = np.ascontiguousarray(v[pos])
    return out


def permute_rhs(b: np.ndarray, permtab: np.ndarray) -> np.ndarray:
    """User ordering -> permuted, column-major (CscUpdownRhs, csc_intern_updown.c:86-130)."""
    b2 = b.reshape(b.shape[0], -1)
    x = np.empty(b2.shape, dtype=b.dtype, order="F")
    x[permtab, :] = b2
    return x if b.ndim == 2 else x[:, 0]


def unpermute_solution(x: np.ndarray, permtab: np.ndarray) -> np.ndarray:
    """Permuted -> user ordering (CscRhsUpdown, csc_intern_updown.c:184-280)."""
    return np.asfortranarray(x[permtab] if x.ndim == 1 else x[permtab, :])
