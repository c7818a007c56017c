"""Generates the golden fixtures tests/golden/*.npz by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref, built by oracle/build_ref.sh from /root/reference).
Run in the build container:  python tests/golden/make_golden.py
Each fixture holds: the SolverMatrix arrays and final permutation produced by the
reference's order/kass/blend analysis, the internal CSC the reference built
(CscOrdistrib), the pivot threshold, the reference's factor panels
(coeftab/ucoeftab), IPARM_STATIC_PIVOTING / IPARM_INERTIA, a right-hand side and
the reference's solution."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refpastix import RefPastix  # noqa: E402
from pastix_b200 import generators as G  # noqa: E402
from pastix_b200.csc import internal_csc  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def case_matrix(kind, N, dt):
    if kind == "lap1d":
        return G.laplacian_1d(N, dt), G.nested_dissection_perm_1d(N)
    if kind == "lap7":
        return G.laplacian_3d(N, 7, dt), G.nested_dissection_perm(N)
    if kind == "lap27":
        return G.laplacian_3d(N, 27, dt), G.nested_dissection_perm(N)
    if kind == "cd":
        return G.convection_diffusion_3d(N, dt), G.nested_dissection_perm(N)
    if kind == "lap7shift":   # complex symmetric: Laplacian + i*0.5 on the diagonal
        A = G.laplacian_3d(N, 7, dt).tolil()
        A.setdiag(A.diagonal() + 0.5j)
        return A.tocsc(), G.nested_dissection_perm(N)
    if kind == "lap7her":     # hermitian positive definite: off-diagonals -1 +/- 0.3i
        A = G.laplacian_3d(N, 7, dt).tocoo()
        v = A.data.copy(); off = A.row != A.col
        v[off] = v[off] + 0.3j * np.where((A.row[off] - A.col[off]) % 2 == 0, 1, -1)
        return sp.csc_matrix((v, (A.row, A.col)), shape=A.shape), G.nested_dissection_perm(N)
    if kind == "diag":        # no coupling at all: every cblk is a root without off-diagonal bloks
        return sp.diags(np.arange(1, N + 1).astype(dt)).tocsc(), np.arange(N)[::-1].copy()
    if kind == "lap7sing":    # zero pivots: forces the static-pivoting rule
        A = G.laplacian_3d(N, 7, dt).tolil()
        n = N ** 3
        for i in (0, n // 3, n - 1):
            A[i, i] = 0.0
        return A.tocsc(), G.nested_dissection_perm(N)
    raise ValueError(kind)


CASES = [
    # name, kind, N, prec, facto, iparm overrides, nrhs
    ("lap1d100_llt_d", "lap1d", 100, "d", "llt", {}, 1),      # BASELINE config 1 (simple -lap 100)
    ("lap1d100_ldlt_d", "lap1d", 100, "d", "ldlt", {}, 1),
    ("lap7_8_llt_d", "lap7", 8, "d", "llt", {}, 3),
    ("lap27_6_ldlt_d", "lap27", 6, "d", "ldlt", {}, 2),
    ("cd_8_lu_d", "cd", 8, "d", "lu", {}, 2),
    ("cd_6_lu_z", "cd", 6, "z", "lu", {}, 2),
    ("lap7shift_6_ldlt_z", "lap7shift", 6, "z", "ldlt", {}, 1),
    ("lap7her_6_ldlh_z", "lap7her", 6, "z", "ldlh", {}, 1),
    ("lap7_6_llt_s", "lap7", 6, "s", "llt", {}, 2),
    ("cd_6_lu_c", "cd", 6, "c", "lu", {}, 1),
    ("lap7_8_ilu2_llt_d", "lap7", 8, "d", "llt", {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 2}, 1),
    ("lap7sing_6_ldlt_d", "lap7sing", 6, "d", "ldlt", {}, 1),
    ("lap7_10_llt_d_bs16", "lap7", 10, "d", "llt", {"IPARM_MIN_BLOCKSIZE": 8, "IPARM_MAX_BLOCKSIZE": 16}, 1),
    # edge cases: one unknown, one cblk, no off-diagonal blok anywhere
    ("lap1d1_llt_d", "lap1d", 1, "d", "llt", {}, 1),
    ("lap1d3_ldlt_d", "lap1d", 3, "d", "ldlt", {}, 2),
    ("diag20_lu_d", "diag", 20, "d", "lu", {}, 1),
    # remaining precision x factorization combinations of the generic (SIMT) path
    ("lap7_6_ldlt_s", "lap7", 6, "s", "ldlt", {}, 1),
    ("cd_6_lu_s", "cd", 6, "s", "lu", {}, 2),
    ("lap7her_6_ldlh_c", "lap7her", 6, "c", "ldlh", {}, 1),
    # IPARM_SCHUR: the last cblk (top separator, numbered last by the nested dissection) is left unfactored = Schur
    # complement; the solve is the interior solve (sopalin_compute.c:767-772, updo.c:425-428)
    ("lap7_8_llt_d_schur", "lap7", 8, "d", "llt", {"IPARM_SCHUR": 1}, 2),
    ("lap27_6_ldlt_d_schur", "lap27", 6, "d", "ldlt", {"IPARM_SCHUR": 1}, 1),
    ("cd_8_lu_d_schur", "cd", 8, "d", "lu", {"IPARM_SCHUR": 1}, 2),
    ("cd_6_lu_z_schur", "cd", 6, "z", "lu", {"IPARM_SCHUR": 1}, 1),
    ("lap7_12_llt_d_schur_bs", "lap7", 12, "d", "llt", {"IPARM_SCHUR": 1, "IPARM_MIN_BLOCKSIZE": 20, "IPARM_MAX_BLOCKSIZE": 40}, 1),
    # complex LLt with a cblk wider than MAXSIZEOFBLOCKS = 64: the reference's unblocked kernel is symmetric (csqrt + geru,
    # compute_diag.c:140) but its blocked trailing update is zherk (sopalin_compute.h:178-179) — whatever that computes
    # on a complex SYMMETRIC matrix is the reference's result, block boundaries at multiples of 64 included
    ("lap7shift_8_llt_z_wide", "lap7shift", 8, "z", "llt", {"IPARM_MIN_BLOCKSIZE": 100, "IPARM_MAX_BLOCKSIZE": 200}, 1),
]


def check_schur(name, A, sym, peritab, S, x, b, w):
    """The reference's Schur mode against dense linear algebra: S = A_SS - A_SI A_II^-1 A_IS on the last w unknowns
    of the final ordering, and the solve = interior solve with the Schur unknowns left at their right-hand side."""
    Af = A if sym == "no" else (A + (sp.tril(A, -1).T.conj() if sym == "her" else sp.tril(A, -1).T))
    P = sp.csc_matrix(Af)[peritab][:, peritab].toarray()
    n = P.shape[0]; k = n - w
    St = P[k:, k:] - P[k:, :k] @ np.linalg.solve(P[:k, :k], P[:k, k:])
    e = np.abs((S if sym == "no" else np.tril(S)) - (St if sym == "no" else np.tril(St))).max() / np.abs(St).max()
    bp = b.reshape(n, -1)[peritab]
    xp = bp.copy(); xp[:k] = np.linalg.solve(P[:k, :k], bp[:k])
    ex = np.abs(x.reshape(n, -1)[peritab] - xp).max() / np.abs(xp).max()
    tolv = 1e-12 if P.dtype in (np.float64, np.complex128) else 1e-4
    assert e < tolv and ex < tolv, (name, e, ex)
    print(f"  schur: width {w}, |S - dense Schur| {e:.1e}, |x - interior solve| {ex:.1e}")


def make(name, kind, N, prec, facto, over, nrhs):
    dt = DT[prec]
    A, perm0 = case_matrix(kind, N, dt)
    n = A.shape[0]
    sym = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}[facto]
    r = RefPastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze()
    s = r.solver()
    permtab, peritab = r.order()
    r.numfact()
    csc = r.csc()
    mine = internal_csc(A, permtab, sym, dt)
    assert np.array_equal(mine["colptr"], csc["colptr"]) and np.array_equal(mine["rows"], csc["rows"]), name
    assert np.array_equal(mine["values"], csc["vals"]), name
    L, U = r.coef()
    out = r.out()
    b = G.rhs_vector(n, nrhs, dt)
    x = r.solve(b)
    eps = out["epsilon_magn_ctrl"]
    crit = r.norm1() * np.sqrt(eps)
    d = dict(cblknbr=s["cblknbr"], bloknbr=s["bloknbr"], fcol=s["fcol"], lcol=s["lcol"], bloknum=s["bloknum"],
             stride=s["stride"], frow=s["frow"], lrow=s["lrow"], fcblk=s["fcblk"], coefind=s["coefind"],
             permtab=permtab, colptr=csc["colptr"], rows=csc["rows"], values=csc["vals"],
             critere=crit, norm1=r.norm1(), L=L, nbpivot=out["static_pivoting"], inertia=out["inertia"],
             nnzeros=out["nnzeros"], fact_flops=out["fact_flops"], b=b, x=x,
             prec=prec, facto=facto, sym=sym, kind=kind, N=N, schur=int(over.get("IPARM_SCHUR", 0)))
    if d["schur"]:
        cb = s["cblknbr"]; w = int(s["lcol"][cb - 1] - s["fcol"][cb - 1] + 1)
        S = r.get_schur(w)                       # pastix_getSchur = the last cblk's coeftab
        assert np.array_equal(S.ravel(order="F"), L[-w * w:]), name
        check_schur(name, A, sym, peritab, S, x, b, w)
    if mine["tvalues"] is not None:
        d["tvalues"] = mine["tvalues"]
    if U is not None:
        d["U"] = U
    # compact the index arrays
    for k in ("fcol", "lcol", "bloknum", "stride", "frow", "lrow", "fcblk", "coefind", "permtab", "colptr", "rows"):
        d[k] = d[k].astype(np.int32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: n={n} cblk={s['cblknbr']} blok={s['bloknbr']} coefnbr={s['coefnbr']} nbpivot={out['static_pivoting']} "
          f"size={os.path.getsize(os.path.join(OUT, name + '.npz')) / 1024:.0f} KiB")
    # r.clean() is skipped: the reference frees with its own allocator bookkeeping and the process exits anyway


if __name__ == "__main__":
    only = sys.argv[1:]
    for c in CASES:
        if only and c[0] not in only:
            continue
        make(*c)
