"""GPU parity through the reference-facing boundary: the SAME pastix() calls (iparm/dparm,
API_TASK_* sequence) served by (a) the drop-in library = reference host code + B200 numeric phase
and (b) the unmodified reference built in oracle/_ref (CPU sopalin).  Reads like the reference's
own examples (src/example/src/simple.c:60-256 + CHECK_SOL utils.h:74-137), with asserts."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from conftest import relerr, tol  # noqa: E402

CASES = [
    # kind, N, prec, facto, iparm overrides, nrhs
    ("lap1d", 100, "d", "llt", {}, 1),          # BASELINE config 1: simple -lap 100
    ("lap1d", 100, "d", "ldlt", {}, 1),
    ("lap7", 12, "d", "llt", {}, 3),
    ("lap27", 14, "d", "ldlt", {}, 2),       # (the unmodified reference itself corrupts its heap on 27-pt LDLt for N in 7..12)
    ("cd", 10, "d", "lu", {}, 2),
    ("cd", 8, "z", "lu", {}, 1),
    ("lap7shift", 8, "z", "ldlt", {}, 1),
    ("lap7her", 8, "z", "ldlh", {}, 1),
    ("lap7", 8, "s", "llt", {}, 1),
    ("cd", 6, "c", "lu", {}, 1),
    ("lap7", 20, "d", "llt", {"IPARM_MIN_BLOCKSIZE": 160, "IPARM_MAX_BLOCKSIZE": 320}, 1),   # wide cblks: sub-panel path
    ("lap7", 10, "d", "llt", {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 2}, 1),        # ILU(2): generic path
]


def full_matrix(A, sym):
    if sym == "no":
        return A
    lo = sp.tril(A, -1)
    return (A + (lo.conj().T if sym == "her" else lo.T)).tocsc()


@pytest.mark.parametrize("kind,N,prec,facto,over,nrhs", CASES)
def test_pastix_dropin_matches_reference(kind, N, prec, facto, over, nrhs):
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, N, DT[prec])
    sym = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}[facto]
    b = G.rhs_vector(A.shape[0], nrhs, DT[prec])
    ref = RefPastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze().numfact()
    xr = ref.solve(b)
    gpu = Pastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze().numfact()
    xg = gpu.solve(b)
    og, orf = gpu.out(), ref.out()
    # analysis is the same unchanged host code: identical structure-derived outputs
    assert og["nnzeros"] == orf["nnzeros"] and og["fact_flops"] == orf["fact_flops"]
    assert og["static_pivoting"] == orf["static_pivoting"]
    if facto == "ldlt" and prec in ("s", "d"):
        assert og["inertia"] == orf["inertia"]
    assert og["fact_time"] > 0 and og["solv_time"] > 0
    incomplete = bool(over.get("IPARM_INCOMPLETE"))
    t = tol(prec)
    assert relerr(xg, xr) <= 50 * t, "solution differs from the reference's"
    if not incomplete:
        Af = full_matrix(A, sym)
        res = np.linalg.norm(Af @ xg - b) / np.linalg.norm(b)
        assert res <= (1e-12 if prec in ("d", "z") else 1e-4), res     # north_star: ||b-Ax||/||b|| <= 1e-12 in double
    gpu.release()


@pytest.mark.parametrize("kind,N,facto,nrhs", [("lap7", 12, "llt", 2), ("cd", 10, "lu", 1), ("lap27", 14, "ldlt", 1)])
def test_int32_dropin_matches_reference(kind, N, facto, nrhs):
    """The drop-in built with the reference's default 32-bit PASTIX_INT (libpastix_dropin_d_i32.so; the shim widens the
    SolverMatrix into the int64 C ABI, the internal CSC takes the reference's host CscOrdistrib) against the 64-bit
    reference: same pastix() calls, same outputs."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix, dropin_path
    from pastix_b200 import generators as G
    if not available("d") or not os.path.exists(dropin_path("d", 32)):
        pytest.skip("oracle/_ref or the int32 drop-in not built")
    A, perm0 = case_matrix(kind, N, DT["d"])
    sym = {"llt": "yes", "ldlt": "yes", "lu": "no"}[facto]
    b = G.rhs_vector(A.shape[0], nrhs, DT["d"])
    ref = RefPastix("d", threads=1).setup(A, perm0, facto, sym=sym).analyze().numfact()
    xr = ref.solve(b)
    gpu = Pastix("d", threads=1, int_bits=32).setup(A, perm0, facto, sym=sym).analyze().numfact()
    xg = gpu.solve(b)
    og, orf = gpu.out(), ref.out()
    assert og["nnzeros"] == orf["nnzeros"] and og["fact_flops"] == orf["fact_flops"]
    assert og["static_pivoting"] == orf["static_pivoting"]
    if facto == "ldlt":
        assert og["inertia"] == orf["inertia"]
    assert relerr(xg, xr) <= 50 * tol("d")
    Af = full_matrix(A, sym)
    assert np.linalg.norm(Af @ xg - b) / np.linalg.norm(b) <= 1e-12
    gpu.release()


def test_two_live_instances_and_refactorization_with_new_values():
    """Step-by-step use of pastix() (src/example/src/step-by-step.c, reentrant.c): two pastix_data alive at once, their
    NUMFACT / SOLVE calls interleaved (one device handle per SolverMatrix in the shim's side table), then a second
    NUMFACT on the SAME analysis with new matrix values — the internal CSC is rebuilt from the user's avals on every
    NUMFACT (pastix.c:3486-3489) and the resident factors must follow."""
    from make_golden import case_matrix, DT
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    A1, p1 = case_matrix("lap7", 10, DT["d"])
    A2, p2 = case_matrix("cd", 8, DT["d"])
    b1 = G.rhs_vector(A1.shape[0], 1, DT["d"])[:, 0].copy()
    b2 = G.rhs_vector(A2.shape[0], 2, DT["d"])
    g1 = Pastix("d", threads=1).setup(A1, p1, "ldlt").analyze()
    g2 = Pastix("d", threads=1).setup(A2, p2, "lu", sym="no").analyze()
    g1.numfact(); g2.numfact()
    assert g1.handle() != g2.handle() and g1.handle() and g2.handle()
    x2 = g2.solve(b2); x1 = g1.solve(b1)
    F1 = full_matrix(A1, "yes")
    assert np.linalg.norm(F1 @ x1 - b1) / np.linalg.norm(b1) <= 1e-12
    assert np.linalg.norm(A2 @ x2 - b2) / np.linalg.norm(b2) <= 1e-12
    h1 = g1.handle()
    g1.vals *= 2.0                                   # same pattern, new values: A1 <- 2 A1
    g1.numfact()
    assert g1.handle() == h1, "the analysis did not change: the device handle (schedule, slabs) must be reused"
    y1 = g1.solve(b1)
    assert np.linalg.norm(2.0 * (F1 @ y1) - b1) / np.linalg.norm(b1) <= 1e-12
    assert relerr(2.0 * y1, x1) <= 50 * tol("d")
    assert relerr(g2.solve(b2), x2) <= 1e-14         # the other instance's factors are untouched
    g1.release(); g2.release()


def test_new_analysis_with_another_structure_on_the_same_pastix_data():
    """The shim keys its device state by &pastix_data->solvmatr, an address that survives a second API_TASK_ANALYSE on
    the same pastix_data (pastix_task_blend only calls CoefMatrix_Free, pastix.c:2714-2718).  A new ordering (hence a
    different SolverMatrix at the same address, same n) and new values must get a NEW schedule / slab — the structural
    fingerprint in the side table — not the stale one.  (A new ordering is the re-analysis the unmodified reference
    itself survives; it corrupts its heap when the PATTERN changes on a live pastix_data.)"""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    A1, p1 = case_matrix("lap7", 10, DT["d"])
    n = A1.shape[0]
    b = G.rhs_vector(n, 1, DT["d"])[:, 0].copy()
    g = Pastix("d", threads=1).setup(A1, p1, "ldlt").analyze().numfact()
    x1 = g.solve(b)
    F1 = full_matrix(A1, "yes")
    assert np.linalg.norm(F1 @ x1 - b) / np.linalg.norm(b) <= 1e-12
    c1 = g.sopalin().coefnbr
    p2 = (n - 1 - p1).astype(p1.dtype)                 # the mirrored elimination order: another symbol structure
    g.set_perm(p2)
    g.vals *= 3.0
    g.analyze().numfact()
    assert g.sopalin().coefnbr != c1, "the handle was not rebuilt for the new structure"
    x2 = g.solve(b)
    assert np.linalg.norm(3.0 * (F1 @ x2) - b) / np.linalg.norm(b) <= 1e-12
    ref = RefPastix("d", threads=1).setup(3.0 * A1, p2, "ldlt").analyze().numfact()
    assert relerr(x2, ref.solve(b)) <= 50 * tol("d")
    assert g.live_entries() == 1
    g.clean()
    assert g.live_entries() == 0


def test_clean_releases_the_device_and_a_recycled_address_gets_a_fresh_handle():
    """API_TASK_CLEAN (pastix_task_clean -> solverExit, pastix.c:4539) must free the HBM held for that pastix_data
    without any non-reference call, and a later pastix_data that malloc places at the same address must not inherit
    anything: several init..clean cycles with different matrices, no explicit release."""
    import torch
    from make_golden import case_matrix, DT
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    cases = [("lap7", 10, "llt", "yes"), ("cd", 8, "lu", "no"), ("lap27", 8, "ldlt", "yes"), ("lap7", 9, "ldlt", "yes")]
    free0 = None
    for it, (kind, N, facto, sym) in enumerate(cases * 2):
        A, p0 = case_matrix(kind, N, DT["d"])
        b = G.rhs_vector(A.shape[0], 1, DT["d"])[:, 0].copy()
        g = Pastix("d", threads=1).setup(A, p0, facto, sym=sym).analyze().numfact()
        x = g.solve(b)
        assert np.linalg.norm(full_matrix(A, sym) @ x - b) / np.linalg.norm(b) <= 1e-12, (it, kind)
        assert g.live_entries() == 1
        g.clean()
        assert g.live_entries() == 0
        torch.cuda.synchronize()
        free = torch.cuda.mem_get_info()[0]
        if free0 is None:
            free0 = free                               # after the first cycle: context, module and pools are loaded
        assert free >= free0 - (8 << 20), f"cycle {it}: {free0 - free} bytes of HBM not returned by API_TASK_CLEAN"


@pytest.mark.parametrize("prec,sym,kind", [("d", "yes", "lap7"), ("z", "yes", "lap7shift"), ("z", "her", "lap7her"), ("c", "her", "lap7her")])
def test_lu_on_a_symmetric_or_hermitian_typed_matrix(prec, sym, kind):
    """IPARM_FACTORIZATION = LU with IPARM_SYM = YES / HER: the reference builds the internal CSC of type 'S' / 'H' with
    forcetrans (pastix.c:3294-3310) and fills ucoeftab with the transposed values — CONJUGATED for 'H'
    (csc_intern_solve.c:110-116).  Same calls on both libraries."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, 8, DT[prec])
    b = G.rhs_vector(A.shape[0], 2, DT[prec])
    ref = RefPastix(prec, threads=1).setup(A, perm0, "lu", sym=sym).analyze().numfact()
    xr = ref.solve(b)
    for host_csc in (False, True):                     # device-built internal CSC and the reference's host CscOrdistrib
        if host_csc:
            os.environ["PB200_HOST_CSC"] = "1"
        try:
            gpu = Pastix(prec, threads=1).setup(A, perm0, "lu", sym=sym).analyze().numfact()
            xg = gpu.solve(b)
        finally:
            os.environ.pop("PB200_HOST_CSC", None)
        res = np.linalg.norm(full_matrix(A, sym) @ xg - b) / np.linalg.norm(b)
        assert res <= (1e-12 if prec in ("d", "z") else 1e-4), (res, host_csc)
        assert relerr(xg, xr) <= 50 * tol(prec), host_csc
        gpu.clean()


def test_analysis_with_several_blend_threads():
    """IPARM_THREAD_NBR > 1 only shapes blend's task vectors (ttsktab); the numeric phase on the GPU ignores them and
    must give the same answer as the reference run with the same iparm."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    if not available("d"):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix("lap7", 12, DT["d"])
    b = G.rhs_vector(A.shape[0], 1, DT["d"])[:, 0].copy()
    xr = RefPastix("d", threads=4).setup(A, perm0, "llt").analyze().numfact().solve(b)
    gpu = Pastix("d", threads=4).setup(A, perm0, "llt").analyze().numfact()
    xg = gpu.solve(b)
    assert relerr(xg, xr) <= 50 * tol("d")
    assert np.linalg.norm(full_matrix(A, "yes") @ xg - b) / np.linalg.norm(b) <= 1e-12
    gpu.release()


def test_factors_match_reference_through_handle():
    """coeftab read back from HBM through the handle the shim keeps == the reference's panels."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from conftest import lower_mask
    if not available("d"):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix("lap27", 14, DT["d"])
    ref = RefPastix("d", threads=1).setup(A, perm0, "ldlt").analyze().numfact()
    Lr, _ = ref.coef()
    gpu = Pastix("d", threads=1).setup(A, perm0, "ldlt").analyze().numfact()
    Lg, _ = gpu.sopalin().get_coeftab()
    m = lower_mask(ref.solver())
    assert relerr(Lg[m], Lr[m]) <= tol("d")
    assert abs(gpu.critere() - ref.norm1() * np.sqrt(ref.out()["epsilon_magn_ctrl"])) <= 1e-15 * ref.norm1()
    gpu.release()


def test_refinement_runs_on_top_of_gpu_updown():
    """API_TASK_REFINE: the reference's host GMRES loop preconditioned by the GPU up_down (ILU(1) factor)."""
    from make_golden import case_matrix, DT
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    A, perm0 = case_matrix("lap7", 10, DT["d"])
    over = {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 1, "IPARM_REFINEMENT": None}
    gpu = Pastix("d", threads=1)
    over["IPARM_REFINEMENT"] = gpu.E["API_RAF_GMRES"]
    gpu.setup(A, perm0, "llt", iparm_over=over, dparm_over={"DPARM_EPSILON_REFINEMENT": 1e-10}).analyze().numfact()
    b = G.rhs_vector(A.shape[0], 1, DT["d"])[:, 0].copy()
    x = gpu.solve(b)
    x = gpu.refine(b, x)
    Af = full_matrix(A, "yes")
    res = np.linalg.norm(Af @ x - b) / np.linalg.norm(b)
    assert res <= 1e-8, res
    assert gpu.out()["nbiter"] >= 1
    gpu.release()


@pytest.mark.parametrize("prec", ["d", "z"])
def test_transpose_solve_lu(prec):
    """IPARM_TRANSPOSE_SOLVE on an LU factorization (updo.c:165-260, 1553-1600): A^T x = b, same answer as the reference."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix("cd", 9, DT[prec])
    b = G.rhs_vector(A.shape[0], 1, DT[prec])
    ref = RefPastix(prec, threads=1)
    over = {"IPARM_TRANSPOSE_SOLVE": ref.E["API_YES"]}
    ref.setup(A, perm0, "lu", sym="no", iparm_over=over).analyze().numfact()
    xr = ref.solve(b)
    gpu = Pastix(prec, threads=1).setup(A, perm0, "lu", sym="no", iparm_over=over).analyze().numfact()
    xg = gpu.solve(b)
    At = sp.csc_matrix(A).T
    assert np.linalg.norm(At @ xg - b) / np.linalg.norm(b) <= 1e-12
    assert np.linalg.norm(A @ xg - b) / np.linalg.norm(b) > 1e-3          # really the transposed system
    assert relerr(xg, xr) <= 50 * tol(prec)
    # and back to the plain system on the same factors
    gpu.iparm[gpu.E["IPARM_TRANSPOSE_SOLVE"]] = gpu.E["API_NO"]
    x2 = gpu.solve(b)
    assert np.linalg.norm(A @ x2 - b) / np.linalg.norm(b) <= 1e-12
    gpu.release()


SCHUR_CASES = [
    # kind, N, prec, facto, iparm overrides, user-provided Schur array (pastix_setSchurArray)
    ("lap7", 10, "d", "llt", {}, False),
    ("lap7", 10, "d", "ldlt", {}, True),
    ("cd", 8, "d", "lu", {}, False),
    ("cd", 6, "z", "lu", {}, True),
    ("lap7", 14, "d", "llt", {"IPARM_MIN_BLOCKSIZE": 20, "IPARM_MAX_BLOCKSIZE": 40}, False),   # Schur cblk wider than a sub-panel
    ("lap7", 8, "s", "llt", {}, False),                                                          # generic (SIMT) factorization path
    ("lap1d", 100, "d", "llt", {}, False),                                                       # every cblk small: the solve must still take the Schur-aware sweeps
]


@pytest.mark.parametrize("kind,N,prec,facto,over,user_array", SCHUR_CASES)
def test_schur_complement_matches_reference(kind, N, prec, facto, over, user_array):
    """IPARM_SCHUR = API_YES through pastix(): the last column block is never factored (sopalin_compute.c:767-772),
    pastix_getSchur (pastix.c:6434-6475) returns the same Schur complement from the drop-in as from the reference —
    and the one dense linear algebra gives —, API_TASK_SOLVE is the interior solve that leaves the Schur unknowns at
    their right-hand side (updo.c:425-428, 1154-1180)."""
    import ctypes as C
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, N, DT[prec])
    n = A.shape[0]
    sym = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}[facto]
    over = dict(over, IPARM_SCHUR=1)
    b = G.rhs_vector(n, 2, DT[prec])
    ref = RefPastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze()
    sr = ref.solver()
    w = int(sr["lcol"][sr["cblknbr"] - 1] - sr["fcol"][sr["cblknbr"] - 1] + 1)
    ref.numfact()
    Sr = ref.get_schur(w)
    xr = ref.solve(b)
    gpu = Pastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze()
    mine = None
    if user_array:                                   # the Schur complement lands in user memory (pastix.c:3400-3412)
        mine = np.zeros(w * w, dtype=DT[prec])
        gpu.lib.pastix_setSchurArray.argtypes = [C.c_void_p, C.c_void_p]
        gpu.lib.pastix_setSchurArray(gpu.pd, mine.ctypes.data)
    gpu.numfact()
    Sg = gpu.get_schur(w)
    if mine is not None:
        assert np.array_equal(mine.reshape(w, w, order="F"), Sg)
    xg = gpu.solve(b)
    t = tol(prec)
    lo = (lambda M: M) if facto == "lu" else np.tril
    assert relerr(lo(Sg), lo(Sr)) <= t, "Schur complement differs from the reference's"
    # dense check, independent of both
    _, peritab = gpu.order()
    P = full_matrix(A, sym)[peritab][:, peritab].toarray()
    k = n - w
    St = P[k:, k:] - P[k:, :k] @ np.linalg.solve(P[:k, :k], P[:k, k:])
    assert relerr(lo(Sg), lo(St)) <= (1e-12 if prec in ("d", "z") else 1e-4)
    assert relerr(xg, xr) <= 50 * t
    assert np.array_equal(xg[peritab][k:], b[peritab][k:]), "Schur unknowns must keep their right-hand side"
    # a second NUMFACT + SOLVE on the same analysis gives the same answer (handle reuse, re-assembly)
    gpu.numfact()
    assert relerr(lo(gpu.get_schur(w)), lo(Sg)) <= t
    assert relerr(gpu.solve(b), xg) <= 50 * t
    gpu.release()


@pytest.mark.parametrize("prec,kind,facto,sym", [("z", "cd", "lu", "no"), ("c", "cd", "lu", "no"), ("s", "lap7", "llt", "yes")])
def test_static_pivot_threshold_from_device_norm(prec, kind, facto, sym):
    """critere = ||A||_1 * sqrt(DPARM_EPSILON_MAGN_CTRL) (sopalin3d.c:586-606) with the 1-norm taken on the internal
    CSC in HBM: identical sums for real types, within an ulp of cabs() per entry for complex ones."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, 8, DT[prec])
    ref = RefPastix(prec, threads=1).setup(A, perm0, facto, sym=sym).analyze().numfact()
    gpu = Pastix(prec, threads=1).setup(A, perm0, facto, sym=sym).analyze().numfact()
    want = ref.norm1() * np.sqrt(ref.out()["epsilon_magn_ctrl"])
    assert abs(gpu.critere() - want) <= 4e-16 * want * (1 if prec in ("s", "d") else 8)
    gpu.release()


@pytest.mark.parametrize("eps", [-1e-6, -0.25])
def test_absolute_static_pivot_threshold(eps):
    """DPARM_EPSILON_MAGN_CTRL < 0 is an ABSOLUTE static-pivot threshold (sopalin3d.c:586-590).  On a matrix with zero
    diagonal entries the replaced-pivot count (IPARM_STATIC_PIVOTING) and the inertia must equal the reference's."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    if not available("d"):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix("lap7sing", 8, DT["d"])
    dp = {"DPARM_EPSILON_MAGN_CTRL": eps}
    ref = RefPastix("d", threads=1).setup(A, perm0, "ldlt", dparm_over=dp).analyze().numfact()
    gpu = Pastix("d", threads=1).setup(A, perm0, "ldlt", dparm_over=dp).analyze().numfact()
    assert gpu.critere() == -eps
    assert gpu.out()["static_pivoting"] == ref.out()["static_pivoting"] >= 1
    assert gpu.out()["inertia"] == ref.out()["inertia"]
    gpu.release()


REFINE_CASES = [
    # kind, N, prec, facto, sym, refinement, incomplete level (None: complete factorization)
    ("lap7", 10, "d", "llt", "yes", "API_RAF_GMRES", 1),
    ("lap7", 10, "d", "llt", "yes", "API_RAF_GRAD", 1),
    ("lap7", 10, "d", "ldlt", "yes", "API_RAF_GRAD", 1),
    ("cd", 9, "d", "lu", "no", "API_RAF_BICGSTAB", 1),
    ("cd", 9, "d", "lu", "no", "API_RAF_GMRES", 1),
    ("cd", 8, "z", "lu", "no", "API_RAF_GMRES", 1),
    ("lap7her", 8, "z", "ldlh", "her", "API_RAF_GMRES", 1),
    ("lap7", 8, "s", "llt", "yes", "API_RAF_GRAD", 1),
    ("cd", 9, "d", "lu", "no", "API_RAF_PIVOT", None),      # static-pivot refinement: host vectors around the GPU up_down
]


@pytest.mark.parametrize("kind,N,prec,facto,sym,raf,ilu", REFINE_CASES)
def test_refinement_on_device_vectors_matches_reference(kind, N, prec, facto, sym, raf, ilu):
    """API_TASK_REFINE: the reference's unchanged drivers (raff_gmres.c, raff_grad.c, raff_bicgstab.c) over the device
    vector back end (shim_raff.c + kernels_raff.cuh) against the same drivers over the reference's host back end
    (raff_functions.c): same iteration count (+-1: the dot products are summed in a different order), same solution."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, N, DT[prec])
    b = G.rhs_vector(A.shape[0], 1, DT[prec])[:, 0].copy()
    eps = 1e-10 if prec in ("d", "z") else 1e-5
    out = []
    for cls in (RefPastix, Pastix):
        p = cls(prec, threads=1)
        over = {"IPARM_REFINEMENT": p.E[raf], "IPARM_ITERMAX": 60, "IPARM_GMRES_IM": 25}
        if ilu is not None:
            over.update({"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": ilu})
        p.setup(A, perm0, facto, sym=sym, iparm_over=over, dparm_over={"DPARM_EPSILON_REFINEMENT": eps}).analyze().numfact()
        x = p.solve(b)
        x = p.refine(b, x)
        out.append((x, p.out()["nbiter"], p.out()["relative_error"]))
        if cls is Pastix:
            p.release()
    (xr, itr, errr), (xg, itg, errg) = out
    Af = full_matrix(A, sym)
    res = np.linalg.norm(Af @ xg - b) / np.linalg.norm(b)
    assert res <= 50 * eps, res
    assert abs(itg - itr) <= 1 and (ilu is None or itg >= 1), (itg, itr)
    assert relerr(xg, xr) <= 1e3 * eps


@pytest.mark.parametrize("prec,kind,facto,sym", [("d", "lap7", "llt", "yes"), ("z", "cd", "lu", "no")])
def test_generated_rhs_reads_the_host_csc_on_demand(prec, kind, facto, sym, monkeypatch):
    """IPARM_RHS_MAKING = API_RHS_1: pastix.c:716 builds b = A * 1 from the HOST CscMatrix (Csc2updown,
    csc_intern_updown.c:339) — the one reader of the internal CSC values inside the unchanged pastix.c.  The drop-in
    leaves rows / values in HBM after CscOrdistrib and fills the host arrays when such a reader shows up
    (shim_csc.c, pb200_shim_csc_host): the solution must be the vector of ones, like the reference's, and equal to what
    the eager copy (PB200_EAGER_CSC=1) gives."""
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, 10, DT[prec])
    n = A.shape[0]
    E = Pastix(prec).E
    over = {"IPARM_RHS_MAKING": E["API_RHS_1"]}
    ref = RefPastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze().numfact()
    xr = ref.solve(np.zeros(n, dtype=DT[prec]))
    ref.clean()
    out = []
    for eager in (False, True):
        if eager:
            monkeypatch.setenv("PB200_EAGER_CSC", "1")
        gpu = Pastix(prec, threads=1).setup(A, perm0, facto, sym=sym, iparm_over=over).analyze().numfact()
        out.append(gpu.solve(np.zeros(n, dtype=DT[prec])))
        gpu.clean()
    assert np.abs(xr - 1.0).max() <= 1e-8, "the reference itself does not return the vector of ones"
    assert relerr(out[0], xr) <= 50 * tol(prec)
    assert relerr(out[0], out[1]) <= 50 * tol(prec)
