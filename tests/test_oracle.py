"""CPU tests (no GPU): the C oracle (oracle/sopalin_oracle.c) is PINNED against
(a) the committed golden dumps of the unmodified reference (tests/golden/*.npz,
made by tests/golden/make_golden.py) and (b) — when oracle/_ref was built in this
container — the reference itself, run live on fresh cases."""
import os

import numpy as np
import pytest

from conftest import golden_names, load_golden, lower_mask, relerr, tol
from oracle.oracle import Oracle


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g = load_golden(name)
    o = Oracle(g, g["prec"])
    lu = g["facto"] == "lu"
    # CscNorm1 + threshold of init_struct_sopalin
    n1 = o.norm1(g["colptr"], g["values"])
    assert abs(n1 - g["norm1"]) <= 1e-13 * g["norm1"]
    L, U = o.assemble(g["colptr"], g["rows"], g["values"], g["tvalues"], herm=False, lu=lu)
    nb = o.factorize(g["facto"], L, U, g["critere"], schur=g["schur"])
    assert nb == g["nbpivot"]
    m = lower_mask(g) if not lu else slice(None)
    t = tol(g["prec"])
    assert relerr(L[m], g["L"][m]) <= t
    if lu:
        assert relerr(U, g["U"]) <= t
    if g["facto"] == "ldlt" and g["prec"] in ("s", "d"):
        assert o.inertia(L) == g["inertia"]
    from pastix_b200.csc import permute_rhs, unpermute_solution
    x = permute_rhs(g["b"], g["permtab"])
    o.solve(g["facto"], L, U, x, schur=g["schur"])
    # a replaced pivot is ~1e-15: the solution of that (numerically singular) system is not a parity quantity
    st = 50 * t if g["nbpivot"] == 0 else 1e-1
    assert relerr(unpermute_solution(x, g["permtab"]), g["x"]) <= st


@pytest.mark.parametrize("name", ["lap7_8_llt_d", "cd_8_lu_d", "lap7her_6_ldlh_z"])
def test_oracle_solve_with_reference_factors(name):
    """up_down restatement alone, fed the reference's own factor panels."""
    from pastix_b200.csc import permute_rhs, unpermute_solution
    g = load_golden(name)
    o = Oracle(g, g["prec"])
    x = permute_rhs(g["b"], g["permtab"])
    o.solve(g["facto"], g["L"], g["U"], x)
    assert relerr(unpermute_solution(x, g["permtab"]), g["x"]) <= 50 * tol(g["prec"])


def test_schur_goldens_hold_the_dense_schur_complement():
    """IPARM_SCHUR fixtures: the last cblk of the reference's coeftab is A_SS - A_SI A_II^-1 A_IS of the internal CSC
    (dense linear algebra, independent of both the reference and the oracle), and x_S == b_S in its solution."""
    import scipy.sparse as sp
    names = [n for n in golden_names() if load_golden(n)["schur"]]
    assert names, "no Schur fixtures"
    for name in names:
        g = load_golden(name)
        n = len(g["colptr"]) - 1
        A = sp.csc_matrix((g["values"], g["rows"], g["colptr"]), shape=(n, n)).toarray()   # permuted ordering
        if g["sym"] != "no":                       # the internal CSC of a symmetric matrix holds the lower triangle
            A = np.tril(A) + np.tril(A, -1).T
        cb = g["cblknbr"]; w = int(g["lcol"][cb - 1] - g["fcol"][cb - 1] + 1); k = n - w
        S = g["L"][-w * w:].reshape(w, w, order="F")
        St = A[k:, k:] - A[k:, :k] @ np.linalg.solve(A[:k, :k], A[:k, k:])
        if g["sym"] != "no":
            S, St = np.tril(S), np.tril(St)
        assert relerr(S, St) <= 1e-12, name
        xp = g["x"].reshape(n, -1)[np.argsort(g["permtab"])]      # permuted solution
        bp = g["b"].reshape(n, -1)[np.argsort(g["permtab"])]
        assert np.array_equal(xp[k:], bp[k:]), name
        assert relerr(xp[:k], np.linalg.solve(A[:k, :k], bp[:k])) <= 1e-12, name


def test_golden_structures_are_consistent():
    """Invariants the CUDA engine relies on (checked again in pb200_create)."""
    for name in golden_names():
        g = load_golden(name)
        cb = g["cblknbr"]
        assert g["bloknum"][cb] == g["bloknbr"]
        assert g["fcol"][0] == 0
        for c in range(cb):
            b0, b1 = g["bloknum"][c], g["bloknum"][c + 1]
            assert g["frow"][b0] == g["fcol"][c] and g["lrow"][b0] == g["lcol"][c]
            rows = (g["lrow"][b0:b1] - g["frow"][b0:b1] + 1)
            assert rows.sum() == g["stride"][c]
            assert np.array_equal(np.concatenate([[0], np.cumsum(rows)[:-1]]), g["coefind"][b0:b1])
            assert np.all(g["fcblk"][b0 + 1:b1] > c)


def _ref_available(prec):
    try:
        from oracle import refpastix
        return refpastix.available(prec)
    except Exception:
        return False


LIVE = [
    ("lap7", 7, "d", "llt", {}),
    ("lap27", 5, "d", "ldlt", {}),
    ("cd", 7, "d", "lu", {}),
    ("cd", 5, "z", "lu", {}),
    ("lap7", 9, "d", "ldlt", {"IPARM_MIN_BLOCKSIZE": 4, "IPARM_MAX_BLOCKSIZE": 8}),
    ("lap7", 7, "d", "llt", {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 1}),
]


@pytest.mark.parametrize("kind,N,prec,facto,over", LIVE)
def test_oracle_matches_live_reference(kind, N, prec, facto, over):
    """Fresh (non-golden) cases: run the unmodified reference here and compare.
    Skipped where oracle/_ref has not been built (it needs /root/reference)."""
    if not _ref_available(prec):
        pytest.skip("oracle/_ref not built")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix
    from pastix_b200.csc import internal_csc, permute_rhs, unpermute_solution
    from pastix_b200 import generators as G
    A, perm0 = case_matrix(kind, N, DT[prec])
    sym = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}[facto]
    r = RefPastix(prec, threads=2).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze()
    s = r.solver(); permtab, _ = r.order()
    r.numfact()
    Lr, Ur = r.coef()
    csc = internal_csc(A, permtab, sym, DT[prec])
    o = Oracle(s, prec)
    crit = o.norm1(csc["colptr"], csc["values"]) * np.sqrt(r.out()["epsilon_magn_ctrl"])
    assert abs(o.norm1(csc["colptr"], csc["values"]) - r.norm1()) <= 1e-13 * r.norm1()
    L, U = o.assemble(csc["colptr"], csc["rows"], csc["values"], csc["tvalues"], lu=(facto == "lu"))
    nb = o.factorize(facto, L, U, crit)
    assert nb == r.out()["static_pivoting"]
    m = lower_mask(s) if facto != "lu" else slice(None)
    assert relerr(L[m], Lr[m]) <= tol(prec)
    if Ur is not None:
        assert relerr(U, Ur) <= tol(prec)
    b = G.rhs_vector(A.shape[0], 2, DT[prec])
    xr = r.solve(b)
    x = permute_rhs(b, permtab)
    o.solve(facto, L, U, x)
    assert relerr(unpermute_solution(x, permtab), xr) <= 50 * tol(prec)
