"""TEST INFRASTRUCTURE — not part of the product path.

ctypes front-end to the UNMODIFIED reference built by oracle/build_ref.sh
(oracle/_ref/libpastix_ref_<p>.so).  It drives the reference through its own
public entry point pastix() (src/sopalin/src/pastix.h:219-244, task state
machine pastix.c:4734-5098) and reads internal structures through
oracle/ref_driver.c.  Used by tests/ (checker), by tools that generate golden
fixtures / workload structures, and by bench.py's cpu_baseline / reference arm.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

_DTYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def available(prec: str = "d") -> bool:
    return os.path.exists(os.path.join(_REF, f"libpastix_ref_{prec}.so")) and \
        os.path.exists(os.path.join(_REF, "api_enums.json"))


def enums() -> dict:
    return json.load(open(os.path.join(_REF, "api_enums.json")))


class RefPastix:
    """One reference pastix_data_t instance (one matrix)."""

    def __init__(self, prec: str = "d", threads: int = 1, verbose: int = 0):
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")  # PaStiX threads itself
        self.prec = prec
        self.dtype = np.dtype(_DTYPES[prec])
        self.lib = C.CDLL(os.path.join(_REF, f"libpastix_ref_{prec}.so"), mode=C.RTLD_LOCAL)
        self.E = enums()
        assert self.lib.refdrv_int_size() == 8
        self.lib.refdrv_norm1.restype = C.c_double
        self.pd = C.c_void_p(None)
        self.iparm = np.zeros(self.E["IPARM_SIZE"], dtype=np.int64)
        self.dparm = np.zeros(self.E["DPARM_SIZE"], dtype=np.float64)
        self.threads = threads
        self.verbose = verbose
        self._init_done = False

    # -- raw call ---------------------------------------------------------
    def _call(self, start: int, end: int, b=None, nrhs: int = 1):
        E = self.E
        self.iparm[E["IPARM_START_TASK"]] = start
        self.iparm[E["IPARM_END_TASK"]] = end
        bp = b.ctypes.data_as(C.c_void_p) if b is not None else None
        self.lib.pastix(C.byref(self.pd), C.c_int(0), C.c_int64(self.n),
                        self.colptr.ctypes.data_as(C.c_void_p), self.rows.ctypes.data_as(C.c_void_p),
                        self.vals.ctypes.data_as(C.c_void_p), self.perm.ctypes.data_as(C.c_void_p),
                        self.invp.ctypes.data_as(C.c_void_p), bp, C.c_int64(nrhs),
                        self.iparm.ctypes.data_as(C.c_void_p), self.dparm.ctypes.data_as(C.c_void_p))
        err = int(self.iparm[E["IPARM_ERROR_NUMBER"]])
        if err != 0:
            raise RuntimeError(f"reference pastix() returned IPARM_ERROR_NUMBER={err}")

    # -- setup ------------------------------------------------------------
    def setup(self, A: sp.spmatrix, perm0: np.ndarray, facto: str, sym: str = None, iparm_over: dict = None,
              dparm_over: dict = None):
        """A: CSC, lower triangle for symmetric ('sym'/'her'), full for 'no'.
        perm0: 0-based perm[old]=new.  facto in {'llt','ldlt','lu','ldlh'}."""
        E = self.E
        A = sp.csc_matrix(A)
        A.sort_indices()
        self.n = A.shape[0]
        self.colptr = (A.indptr.astype(np.int64) + 1)
        self.rows = (A.indices.astype(np.int64) + 1)
        self.vals = np.ascontiguousarray(A.data.astype(self.dtype))
        self.perm = perm0.astype(np.int64) + 1
        self.invp = np.empty_like(self.perm)
        self.invp[self.perm - 1] = np.arange(1, self.n + 1)
        # defaults (pastix.c:334-456)
        self.iparm[E["IPARM_MODIFY_PARAMETER"]] = E["API_NO"]
        self._call(E["API_TASK_INIT"], E["API_TASK_INIT"])
        fact = {"llt": "API_FACT_LLT", "ldlt": "API_FACT_LDLT", "lu": "API_FACT_LU", "ldlh": "API_FACT_LDLH"}[facto]
        if sym is None:
            sym = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}[facto]
        self.facto, self.sym = facto, sym
        ip = self.iparm
        ip[E["IPARM_THREAD_NBR"]] = self.threads
        ip[E["IPARM_SYM"]] = {"yes": E["API_SYM_YES"], "no": E["API_SYM_NO"], "her": E["API_SYM_HER"]}[sym]
        ip[E["IPARM_FACTORIZATION"]] = E[fact]
        ip[E["IPARM_VERBOSE"]] = self.verbose
        ip[E["IPARM_ORDERING"]] = E["API_ORDER_PERSONAL"]
        ip[E["IPARM_MATRIX_VERIFICATION"]] = E["API_NO"]
        ip[E["IPARM_LEVEL_OF_FILL"]] = -1
        ip[E["IPARM_RHS_MAKING"]] = E["API_RHS_B"]
        for k, v in (iparm_over or {}).items():
            ip[E[k]] = v
        for k, v in (dparm_over or {}).items():
            self.dparm[E[k]] = v
        return self

    def analyze(self):
        E = self.E
        self._call(E["API_TASK_ORDERING"], E["API_TASK_ANALYSE"])
        return self

    def numfact(self):
        E = self.E
        self._call(E["API_TASK_NUMFACT"], E["API_TASK_NUMFACT"])
        return self

    def solve(self, b: np.ndarray) -> np.ndarray:
        """b: (n,) or (n,nrhs) in the USER ordering; returns x likewise."""
        E = self.E
        x = np.array(b, dtype=self.dtype, order="F", copy=True)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        self._call(E["API_TASK_SOLVE"], E["API_TASK_SOLVE"], b=x, nrhs=nrhs)
        return x

    def clean(self):
        E = self.E
        if self.pd:
            self._call(E["API_TASK_CLEAN"], E["API_TASK_CLEAN"])
            self.pd = C.c_void_p(None)

    # -- outputs ----------------------------------------------------------
    def out(self) -> dict:
        E = self.E
        return {
            "nnzeros": int(self.iparm[E["IPARM_NNZEROS"]]),
            "static_pivoting": int(self.iparm[E["IPARM_STATIC_PIVOTING"]]),
            "inertia": int(self.iparm[E["IPARM_INERTIA"]]),
            "fact_flops": float(self.dparm[E["DPARM_FACT_FLOPS"]]),
            "fact_time": float(self.dparm[E["DPARM_FACT_TIME"]]),
            "solv_time": float(self.dparm[E["DPARM_SOLV_TIME"]]),
            "epsilon_magn_ctrl": float(self.dparm[E["DPARM_EPSILON_MAGN_CTRL"]]),
        }

    def solver(self) -> dict:
        """Flat copy of the SolverMatrix (blend/src/solver.h:94-168)."""
        s = np.zeros(16, dtype=np.int64)
        self.lib.refdrv_solver_sizes(self.pd, s.ctypes.data_as(C.c_void_p))
        cb, bl = int(s[0]), int(s[1])
        a = {k: np.zeros(cb + 1, dtype=np.int64) for k in ("fcol", "lcol", "bloknum", "stride")}
        b = {k: np.zeros(bl, dtype=np.int64) for k in ("frow", "lrow", "fcblk", "levf", "coefind")}
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        self.lib.refdrv_solver_get(self.pd, p(a["fcol"]), p(a["lcol"]), p(a["bloknum"]), p(a["stride"]),
                                   p(b["frow"]), p(b["lrow"]), p(b["fcblk"]), p(b["levf"]), p(b["coefind"]))
        d = dict(cblknbr=cb, bloknbr=bl, nodenbr=int(s[2]), coefnbr=int(s[3]), coefmax=int(s[4]),
                 ftgtnbr=int(s[5]), tasknbr=int(s[6]), indnbr=int(s[7]), thrdnbr=int(s[8]),
                 clustnbr=int(s[9]), clustnum=int(s[10]), procnbr=int(s[11]))
        d.update(a); d.update(b)
        return d

    def tasks(self) -> dict:
        s = self.solver()
        t = {k: np.zeros(s["tasknbr"], dtype=np.int64) for k in
             ("taskid", "prionum", "cblknum", "bloknum", "ftgtcnt", "ctrbcnt", "indnum")}
        ind = np.zeros(max(s["indnbr"], 1), dtype=np.int64)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        self.lib.refdrv_tasks_get(self.pd, p(t["taskid"]), p(t["prionum"]), p(t["cblknum"]), p(t["bloknum"]),
                                  p(t["ftgtcnt"]), p(t["ctrbcnt"]), p(t["indnum"]), p(ind))
        t["indtab"] = ind[: s["indnbr"]]
        return t

    def csc(self) -> dict:
        """Internal CSC after NUMFACT (CscOrdistrib): 0-based, permuted, rows sorted."""
        s = np.zeros(4, dtype=np.int64)
        self.lib.refdrv_csc_sizes(self.pd, s.ctypes.data_as(C.c_void_p))
        if not s[3]:
            raise RuntimeError("internal CSC not filled (run numfact first)")
        n, nnz = int(s[0]), int(s[1])
        colptr = np.zeros(n + 1, dtype=np.int64); rows = np.zeros(nnz, dtype=np.int64)
        vals = np.zeros(nnz, dtype=self.dtype); tv = np.zeros(nnz, dtype=self.dtype) if s[2] else None
        p = lambda x: x.ctypes.data_as(C.c_void_p) if x is not None else None
        self.lib.refdrv_csc_get(self.pd, p(colptr), p(rows), p(vals), p(tv))
        return dict(colptr=colptr, rows=rows, vals=vals, tvals=tv, type=chr(self.lib.refdrv_csc_type(self.pd)))

    def coef(self):
        s = self.solver()
        L = np.zeros(s["coefnbr"], dtype=self.dtype)
        U = np.zeros(s["coefnbr"], dtype=self.dtype) if self.facto == "lu" else None
        rc = self.lib.refdrv_coef_get(self.pd, L.ctypes.data_as(C.c_void_p),
                                      U.ctypes.data_as(C.c_void_p) if U is not None else None)
        if rc:
            raise RuntimeError(f"coeftab not allocated (rc={rc})")
        return L, U

    def order(self):
        """Final permutation kept by the reference after fax/kass (ordemesh, order.h:52-57):
        permtab[old] = new and peritab[new] = old, 0-based.  KASS amalgamation may
        re-number inside supernodes, so this can differ from the PERSONAL input."""
        pt = np.zeros(self.n, dtype=np.int64); pi = np.zeros(self.n, dtype=np.int64)
        self.lib.refdrv_order_get(self.pd, pt.ctypes.data_as(C.c_void_p), pi.ctypes.data_as(C.c_void_p))
        return pt, pi

    def norm1(self) -> float:
        return float(self.lib.refdrv_norm1(self.pd))
