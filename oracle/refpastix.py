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

import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pastix_b200.pastix_api import PastixLib  # noqa: E402  (the same pastix() binding, pointed at the reference build)

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

_DTYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def available(prec: str = "d") -> bool:
    return os.path.exists(os.path.join(_REF, f"libpastix_ref_{prec}.so")) and \
        os.path.exists(os.path.join(_REF, "api_enums.json"))


def enums() -> dict:
    return json.load(open(os.path.join(_REF, "api_enums.json")))


class RefPastix(PastixLib):
    """One reference pastix_data_t instance (one matrix), plus read access to its internals."""

    def __init__(self, prec: str = "d", threads: int = 1, verbose: int = 0):
        super().__init__(prec, os.path.join(_REF, f"libpastix_ref_{prec}.so"), enums(), threads=threads, verbose=verbose)
        assert self.lib.refdrv_int_size() == 8
        self.lib.refdrv_norm1.restype = C.c_double

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
