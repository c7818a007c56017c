"""Synthetic benchmark matrices and the geometric nested-dissection ordering.

These are the inputs SURVEY.md §8(d) fixes for BASELINE.json's configs (the
reference itself only ships a 1-D generator, matrix_drivers/src/laplacian.c:92):
grid index i = (z*N + y)*N + x, CSC with 1-based indices when handed to the
reference-style API, lower triangle only for symmetric matrices.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def laplacian_1d(n: int, dtype=np.float64) -> sp.csc_matrix:
    """Values of genlaplacian (matrix_drivers/src/laplacian.c:150-190): diag 2,
    sub-diagonal -1, lower triangle only."""
    d = np.full(n, 2.0, dtype=dtype)
    o = np.full(n - 1, -1.0, dtype=dtype)
    return sp.diags([d, o], [0, -1], format="csc", dtype=dtype)


def _grid_offsets(stencil: int):
    offs = []
    if stencil == 7:
        offs = [(0, 0, 1), (0, 1, 0), (1, 0, 0)]  # (dz, dy, dx), lower triangle neighbours
    elif stencil == 27:
        for dz in (0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if dz == 0 and (dy < 0 or (dy == 0 and dx <= 0)):
                        continue
                    offs.append((dz, dy, dx))
    else:
        raise ValueError("stencil must be 7 or 27")
    return offs


def laplacian_3d(N: int, stencil: int = 7, dtype=np.float64, lower: bool = True) -> sp.csc_matrix:
    """3-D Laplacian on an N^3 grid. 7-pt: diag 6, off -1. 27-pt: diag 26, off -1.
    Returns the lower triangle (rows >= cols) as CSC with sorted indices."""
    n = N ** 3
    z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    z = z.ravel(); y = y.ravel(); x = x.ravel()
    idx = (z * N + y) * N + x
    rows = [idx]
    cols = [idx]
    vals = [np.full(n, 26.0 if stencil == 27 else 6.0)]
    for dz, dy, dx in _grid_offsets(stencil):
        zz, yy, xx = z + dz, y + dy, x + dx
        ok = (zz >= 0) & (zz < N) & (yy >= 0) & (yy < N) & (xx >= 0) & (xx < N)
        r = ((zz * N + yy) * N + xx)[ok]
        rows.append(r); cols.append(idx[ok]); vals.append(np.full(r.size, -1.0))
    rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals).astype(dtype)
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()
    if not lower:
        A = (A + sp.tril(A, -1).T).tocsc()
    A.sort_indices()
    return A


def convection_diffusion_3d(N: int, dtype=np.complex128) -> sp.csc_matrix:
    """Config 4: 7-pt pattern, both triangles, nonsymmetric values:
    diag 6+0.5i (complex) / 6 (real), neighbours -1 -/+ c with c = (0.3,0.2,0.1)
    for the x/y/z axes: entry A[i, i+e_axis] = -1 + c, A[i+e_axis, i] = -1 - c."""
    n = N ** 3
    z, y, x = np.meshgrid(np.arange(N), np.arange(N), np.arange(N), indexing="ij")
    z = z.ravel(); y = y.ravel(); x = x.ravel()
    idx = (z * N + y) * N + x
    cplx = np.issubdtype(np.dtype(dtype), np.complexfloating)
    rows = [idx]; cols = [idx]; vals = [np.full(n, (6.0 + 0.5j) if cplx else 6.0, dtype=dtype)]
    for (dz, dy, dx), c in (((0, 0, 1), 0.3), ((0, 1, 0), 0.2), ((1, 0, 0), 0.1)):
        zz, yy, xx = z + dz, y + dy, x + dx
        ok = (zz < N) & (yy < N) & (xx < N)
        j = ((zz * N + yy) * N + xx)[ok]
        i = idx[ok]
        rows += [j, i]; cols += [i, j]
        vals += [np.full(i.size, -1.0 - c, dtype=dtype), np.full(i.size, -1.0 + c, dtype=dtype)]
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsc()
    A.sort_indices()
    return A


def rhs_vector(n: int, nrhs: int = 1, dtype=np.float64) -> np.ndarray:
    """b_i = 1 + (i mod 7)/4 (+0.1i if complex) (+0.01*rhs_index); column-major n x nrhs."""
    i = np.arange(n)
    b = np.empty((n, nrhs), dtype=dtype, order="F")
    for k in range(nrhs):
        col = 1.0 + (i % 7) * 0.25 + 0.01 * k
        if np.issubdtype(np.dtype(dtype), np.complexfloating):
            col = col + 0.1j
        b[:, k] = col
    return b


def nested_dissection_perm(N: int, leaf: int = 8) -> np.ndarray:
    """Recursive coordinate bisection on the N^3 grid: split the longest axis
    (x before y before z on ties) at its midpoint, number both halves first and
    the separator plane last; boxes of <= `leaf` nodes are numbered in natural
    order. Returns perm (0-based): perm[old] = new."""
    perm = np.empty((N, N, N), dtype=np.int64)  # indexed [z, y, x]
    cnt = 0
    # explicit stack of (kind, box); kind 0 = recurse, 1 = number this slab now
    stack = [(0, (0, N, 0, N, 0, N))]
    while stack:
        kind, (x0, x1, y0, y1, z0, z1) = stack.pop()
        dx, dy, dz = x1 - x0, y1 - y0, z1 - z0
        if dx <= 0 or dy <= 0 or dz <= 0:
            continue
        if kind == 1 or dx * dy * dz <= leaf:
            m = dx * dy * dz
            perm[z0:z1, y0:y1, x0:x1] = (cnt + np.arange(m)).reshape(dz, dy, dx)
            cnt += m
            continue
        if dx >= dy and dx >= dz:
            mid = x0 + dx // 2
            a = (x0, mid, y0, y1, z0, z1); b = (mid + 1, x1, y0, y1, z0, z1); s = (mid, mid + 1, y0, y1, z0, z1)
        elif dy >= dz:
            mid = y0 + dy // 2
            a = (x0, x1, y0, mid, z0, z1); b = (x0, x1, mid + 1, y1, z0, z1); s = (x0, x1, mid, mid + 1, z0, z1)
        else:
            mid = z0 + dz // 2
            a = (x0, x1, y0, y1, z0, mid); b = (x0, x1, y0, y1, mid + 1, z1); s = (x0, x1, y0, y1, mid, mid + 1)
        # LIFO: push separator first so that a, then b, then s are numbered in this order
        stack.append((1, s)); stack.append((0, b)); stack.append((0, a))
    assert cnt == N ** 3
    return perm.ravel()


def nested_dissection_perm_1d(n: int, leaf: int = 4) -> np.ndarray:
    """1-D recursive bisection (halves first, the separating vertex last)."""
    perm = np.empty(n, dtype=np.int64)
    cnt = 0
    stack = [(0, 0, n)]
    while stack:
        kind, lo, hi = stack.pop()
        if hi <= lo:
            continue
        if kind == 1 or hi - lo <= leaf:
            perm[lo:hi] = cnt + np.arange(hi - lo)
            cnt += hi - lo
            continue
        mid = lo + (hi - lo) // 2
        stack.append((1, mid, mid + 1)); stack.append((0, mid + 1, hi)); stack.append((0, lo, mid))
    return perm


def permute_symmetric(A: sp.spmatrix, perm: np.ndarray, full: bool = True) -> sp.csc_matrix:
    """P A P^T in CSC with sorted rows (0-based). If `A` holds only the lower
    triangle of a symmetric matrix and full=True, both triangles are produced —
    the layout of the reference's internal CSC (csc_intern_build.c:352-560)."""
    A = A.tocoo()
    r = perm[A.row]; c = perm[A.col]
    v = A.data
    if full:
        off = A.row != A.col
        r, c, v = np.concatenate([r, c[off]]), np.concatenate([c, r[off]]), np.concatenate([v, v[off]])
    B = sp.coo_matrix((v, (r, c)), shape=A.shape).tocsc()
    B.sort_indices()
    return B
