"""Device-side CscOrdistrib (pastix_b200/csrc/csc_build.cu behind pastix_b200/shim/shim_csc.c) against the
reference's own CscOrdistrib (src/sopalin/src/csc_intern_build.c:352-570): the internal CSC that a
pastix(API_TASK_NUMFACT) call leaves in pastix_data must be IDENTICAL — column pointers, row indices, values and
the transposed values bit for bit — to the one the unmodified reference (oracle/_ref) builds from the same user
CSC and ordering.  Covers every branch of the reference routine: 'S' (mirror), 'H' (conjugated mirror), 'U' with
transcsc, 'S' with forcetrans (LU on a symmetric matrix) and all four precisions."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

CASES = [
    # kind, N, prec, facto, sym
    ("lap1d", 100, "d", "llt", "yes"),
    ("lap7", 12, "d", "llt", "yes"),
    ("lap27", 14, "d", "ldlt", "yes"),
    ("lap7", 9, "d", "lu", "yes"),          # LU on a symmetric matrix: forcetrans, transcsc aliases the values
    ("cd", 10, "d", "lu", "no"),            # 'U' + transposed values
    ("cd", 8, "z", "lu", "no"),
    ("lap7shift", 8, "z", "ldlt", "yes"),
    ("lap7her", 8, "z", "ldlh", "her"),     # 'H': mirror entries conjugated
    ("lap7", 8, "s", "llt", "yes"),
    ("cd", 6, "c", "lu", "no"),
]


def _same(a, b):
    return a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("kind,N,prec,facto,sym", CASES)
def test_device_csc_identical_to_reference(kind, N, prec, facto, sym):
    from make_golden import case_matrix, DT
    from oracle.refpastix import RefPastix, available
    from pastix_b200.pastix_api import Pastix
    if not available(prec):
        pytest.skip("oracle/_ref not built")
    A, perm0 = case_matrix(kind, N, DT[prec])
    ref = RefPastix(prec, threads=1).setup(A, perm0, facto, sym=sym).analyze().numfact()
    cr = ref.csc()
    gpu = Pastix(prec, threads=1).setup(A, perm0, facto, sym=sym).analyze().numfact()
    cg = gpu.csc()
    assert cg["type"] == cr["type"]
    assert _same(cg["colptr"], cr["colptr"])
    assert _same(cg["rows"], cr["rows"])
    assert _same(cg["vals"], cr["vals"]), "values differ bitwise"
    # CoefMatrix_Init frees the transposed values after assembly (coefinit.c:327-341) and so does the shim
    assert cg["tvals"] is None and cr["tvals"] is None
    # what the assembly kernel actually read in HBM, incl. the transposed values (A^T on the pattern of A;
    # alias of the values for LU on a symmetric matrix)
    lu = facto == "lu"
    cd = gpu.csc_device(want_t=lu)
    assert _same(cd["colptr"], cr["colptr"]) and _same(cd["rows"], cr["rows"]) and _same(cd["vals"], cr["vals"])
    if lu:
        from pastix_b200.csc import internal_csc
        permtab, _ = gpu.order()
        ic = internal_csc(A, permtab, sym, DT[prec])
        want_t = ic["tvalues"] if sym == "no" else ic["values"]
        assert _same(ic["rows"], cr["rows"])
        assert _same(cd["tvals"], np.ascontiguousarray(want_t)), "transposed values differ bitwise"
    # a second NUMFACT on the same analysis (new values, same pattern) goes through the same path
    gpu.vals *= 2
    gpu.numfact()
    c2 = gpu.csc()
    assert _same(c2["rows"], cr["rows"]) and _same(c2["vals"], 2 * cr["vals"])
    gpu.release()


def test_host_csc_switch_gives_the_same_factorization():
    """PB200_HOST_CSC=1 keeps the reference's host CscOrdistrib (the path multi-dof matrices take): same solution."""
    from make_golden import case_matrix, DT
    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    A, perm0 = case_matrix("cd", 10, DT["d"])
    b = G.rhs_vector(A.shape[0], 1, DT["d"])
    xs = []
    for host in (False, True):
        if host:
            os.environ["PB200_HOST_CSC"] = "1"
        try:
            gpu = Pastix("d", threads=1).setup(A, perm0, "lu", sym="no").analyze().numfact()
            xs.append(gpu.solve(b))
            gpu.release()
        finally:
            os.environ.pop("PB200_HOST_CSC", None)
    # same panels assembled, but the update order on the device is not reproducible bit for bit (L2 reductions)
    assert np.max(np.abs(xs[0] - xs[1])) <= 1e-12 * np.max(np.abs(xs[1]))
