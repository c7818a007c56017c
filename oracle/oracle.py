"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/libsopalin_oracle.so (the
plain-C restatement in oracle/sopalin_oracle.c).  Only tests/, smoke() and
bench.py's cpu_baseline/reference leg may import this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsopalin_oracle.so")
_DTYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
FACTO = {"llt": 0, "ldlt": 1, "lu": 2, "ldlh": 3}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


class _OSolver(C.Structure):
    _fields_ = [("cblknbr", C.c_int64), ("bloknbr", C.c_int64)] + \
        [(k, C.c_void_p) for k in ("fcol", "lcol", "bloknum", "stride", "frow", "lrow", "fcblk", "coefind", "poff")]


class Oracle:
    def __init__(self, solver: dict, prec: str = "d"):
        if not os.path.exists(_LIB):
            build()
        self.lib = C.CDLL(_LIB)
        self.p = prec
        self.dtype = np.dtype(_DTYPES[prec])
        self.s = {k: np.ascontiguousarray(solver[k], dtype=np.int64) for k in
                  ("fcol", "lcol", "bloknum", "stride", "frow", "lrow", "fcblk", "coefind")}
        self.cblknbr, self.bloknbr = int(solver["cblknbr"]), int(solver["bloknbr"])
        w = self.s["lcol"][:self.cblknbr] - self.s["fcol"][:self.cblknbr] + 1
        self.poff = np.concatenate([[0], np.cumsum(self.s["stride"][:self.cblknbr] * w)]).astype(np.int64)
        self.coefnbr = int(self.poff[-1])
        self.n = int(self.s["lcol"][self.cblknbr - 1]) + 1
        self.os = _OSolver(self.cblknbr, self.bloknbr, *[self.s[k].ctypes.data for k in
                           ("fcol", "lcol", "bloknum", "stride", "frow", "lrow", "fcblk", "coefind")],
                           self.poff.ctypes.data)
        f = lambda name: getattr(self.lib, f"{prec}_oracle_{name}")
        self._norm1 = f("norm1"); self._norm1.restype = C.c_double
        self._assemble = f("assemble"); self._assemble.restype = C.c_int64
        self._factorize = f("factorize"); self._factorize.restype = C.c_int
        self._factorize_schur = f("factorize_schur"); self._factorize_schur.restype = C.c_int
        self._solve_schur = f("solve_schur"); self._solve_schur.restype = None
        self._inertia = f("inertia"); self._inertia.restype = C.c_int64
        self._solve = f("solve"); self._solve.restype = None

    def norm1(self, colptr, vals) -> float:
        colptr = np.ascontiguousarray(colptr, dtype=np.int64); vals = np.ascontiguousarray(vals, dtype=self.dtype)
        return float(self._norm1(C.c_int64(len(colptr) - 1), C.c_void_p(colptr.ctypes.data), C.c_void_p(vals.ctypes.data)))

    def assemble(self, colptr, rows, vals, tvals=None, herm=False, lu=False):
        colptr = np.ascontiguousarray(colptr, dtype=np.int64); rows = np.ascontiguousarray(rows, dtype=np.int64)
        vals = np.ascontiguousarray(vals, dtype=self.dtype)
        L = np.empty(self.coefnbr, dtype=self.dtype)
        U = np.empty(self.coefnbr, dtype=self.dtype) if lu else None
        tv = np.ascontiguousarray(tvals, dtype=self.dtype) if tvals is not None else None
        self.dropped = int(self._assemble(C.byref(self.os), C.c_void_p(colptr.ctypes.data), C.c_void_p(rows.ctypes.data),
                                          C.c_void_p(vals.ctypes.data), C.c_void_p(tv.ctypes.data) if tv is not None else None,
                                          C.c_int(int(herm)), C.c_void_p(L.ctypes.data),
                                          C.c_void_p(U.ctypes.data) if U is not None else None))
        return L, U

    def factorize(self, facto: str, L, U, crit: float, schur: bool = False) -> int:
        """schur: IPARM_SCHUR semantics — the last cblk is left unfactored (it holds the Schur complement)."""
        nb = C.c_int64(0)
        rc = (self._factorize_schur if schur else self._factorize)(C.byref(self.os), C.c_int(FACTO[facto]), C.c_void_p(L.ctypes.data),
                             C.c_void_p(U.ctypes.data) if U is not None else None, C.c_double(crit), C.byref(nb))
        if rc:
            raise RuntimeError("oracle: negative diagonal term")
        return int(nb.value)

    def inertia(self, L) -> int:
        return int(self._inertia(C.byref(self.os), C.c_void_p(L.ctypes.data)))

    def solve(self, facto: str, L, U, x, schur: bool = False):
        """x: (n,) or (n,nrhs) Fortran-ordered, permuted ordering; solved in place.
        schur: the last cblk and the bloks facing it are ignored (interior solve, x_S = b_S)."""
        assert x.dtype == self.dtype and (x.ndim == 1 or x.flags.f_contiguous)
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        (self._solve_schur if schur else self._solve)(C.byref(self.os), C.c_int(FACTO[facto]), C.c_void_p(L.ctypes.data),
                    C.c_void_p(U.ctypes.data) if U is not None else None, C.c_void_p(x.ctypes.data),
                    C.c_int64(x.shape[0]), C.c_int64(nrhs))
        return x
