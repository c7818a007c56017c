"""Python binding of the drop-in host library (pastix_b200/lib/libpastix_dropin_<p>.so):
the reference's own public entry point pastix() (src/sopalin/src/pastix.h:219-244) with its
iparm/dparm arrays and API_TASK_* state machine (pastix.c:4734-5098; api.h:252-261), unchanged.
Ordering, symbolic factorization and blend analysis run as the reference's host code inside that
library; API_TASK_NUMFACT and API_TASK_SOLVE land in pastix_b200/shim/sopalin_b200_shim.c and
run on the GPU.  This is the call a PaStiX user already makes — bench.py's `e2e` goes through it
with host buffers."""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(_HERE, "lib")
DTYPES = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def dropin_path(prec: str, int_bits: int = 64) -> str:
    """int_bits=32: the build with the reference's default 32-bit PASTIX_INT (libpastix_dropin_<p>_i32.so)."""
    return os.path.join(LIBDIR, f"libpastix_dropin_{prec}{'_i32' if int_bits == 32 else ''}.so")


# GPU-aware block sizes for blend (pb200_tune_iparm in shim/shim_hooks.c): same analysis
# code, different parameters.  The reference's defaults are 60 / 120.
TUNED_IPARM = {"IPARM_MIN_BLOCKSIZE": 120, "IPARM_MAX_BLOCKSIZE": 240}


def load_enums(path: str = None) -> dict:
    return json.load(open(path or os.path.join(LIBDIR, "api_enums.json")))


class PastixLib:
    """One pastix_data_t instance driven through pastix(); `libpath` decides which build of
    libpastix serves it."""

    def __init__(self, prec: str, libpath: str, enums: dict, threads: int = 1, verbose: int = 0, int_bits: int = 64):
        """int_bits: width of PASTIX_INT in that build (INTSIZE64 / INTSIZE32, common_pastix.h:331-347)."""
        os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")  # PaStiX threads itself
        self.idt = np.int64 if int_bits == 64 else np.int32
        self.cint = C.c_int64 if int_bits == 64 else C.c_int32
        if not os.path.exists(libpath):
            raise RuntimeError(f"{libpath} is missing: run __graft_entry__.build() where the reference tree is available")
        self.prec = prec
        self.dtype = np.dtype(DTYPES[prec])
        self.lib = C.CDLL(libpath, mode=C.RTLD_LOCAL)
        self.E = enums
        self.pd = C.c_void_p(None)
        self.iparm = np.zeros(self.E["IPARM_SIZE"], dtype=self.idt)
        self.dparm = np.zeros(self.E["DPARM_SIZE"], dtype=np.float64)
        self.threads = threads
        self.verbose = verbose

    # -- raw call: pastix(&pastix_data, comm, n, colptr, row, avals, perm, invp, b, rhs, iparm, dparm)
    def _call(self, start: int, end: int, b=None, nrhs: int = 1):
        E = self.E
        self.iparm[E["IPARM_START_TASK"]] = start
        self.iparm[E["IPARM_END_TASK"]] = end
        bp = b.ctypes.data_as(C.c_void_p) if b is not None else None
        self.lib.pastix(C.byref(self.pd), C.c_int(0), self.cint(self.n),
                        self.colptr.ctypes.data_as(C.c_void_p), self.rows.ctypes.data_as(C.c_void_p),
                        self.vals.ctypes.data_as(C.c_void_p), self.perm.ctypes.data_as(C.c_void_p),
                        self.invp.ctypes.data_as(C.c_void_p), bp, self.cint(nrhs),
                        self.iparm.ctypes.data_as(C.c_void_p), self.dparm.ctypes.data_as(C.c_void_p))
        err = int(self.iparm[E["IPARM_ERROR_NUMBER"]])
        if err != 0:
            raise RuntimeError(f"pastix() returned IPARM_ERROR_NUMBER={err}")

    def setup(self, A: sp.spmatrix, perm0: np.ndarray, facto: str, sym: str = None, iparm_over: dict = None,
              dparm_over: dict = None):
        """A: CSC, lower triangle for symmetric ('yes'/'her'), full for 'no'.
        perm0: 0-based perm[old]=new (handed over as API_ORDER_PERSONAL).  facto in {'llt','ldlt','lu','ldlh'}."""
        E = self.E
        A = sp.csc_matrix(A)
        A.sort_indices()
        self.n = A.shape[0]
        self.colptr = (A.indptr.astype(self.idt) + 1)
        self.rows = (A.indices.astype(self.idt) + 1)
        self.vals = np.ascontiguousarray(A.data.astype(self.dtype))
        self.perm = perm0.astype(self.idt) + 1
        self.invp = np.empty_like(self.perm)
        self.invp[self.perm - 1] = np.arange(1, self.n + 1, dtype=self.idt)
        self.iparm[E["IPARM_MODIFY_PARAMETER"]] = E["API_NO"]      # defaults: pastix.c:334-456
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

    def rebind(self, A: sp.spmatrix):
        """Another user matrix of the same order on the SAME pastix_data (the ordering handed over at setup is kept):
        the caller re-runs analyze() / numfact()."""
        A = sp.csc_matrix(A)
        A.sort_indices()
        assert A.shape[0] == self.n
        self.colptr = (A.indptr.astype(self.idt) + 1)
        self.rows = (A.indices.astype(self.idt) + 1)
        self.vals = np.ascontiguousarray(A.data.astype(self.dtype))
        return self

    def set_perm(self, perm0: np.ndarray):
        """Another API_ORDER_PERSONAL permutation for the next analyze() on the same pastix_data."""
        self.perm = perm0.astype(self.idt) + 1
        self.invp = np.empty_like(self.perm)
        self.invp[self.perm - 1] = np.arange(1, self.n + 1, dtype=self.idt)
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

    def refine(self, b: np.ndarray, x: np.ndarray) -> np.ndarray:
        """API_TASK_REFINE: the reference's host refinement loop around the GPU up_down."""
        E = self.E
        x = np.array(x, dtype=self.dtype, order="F", copy=True)
        self._call(E["API_TASK_REFINE"], E["API_TASK_REFINE"], b=x, nrhs=1)
        return x

    def get_schur(self, w: int) -> np.ndarray:
        """pastix_getSchur (pastix.c:6434-6475) after a NUMFACT with IPARM_SCHUR = API_YES: the w x w panel of the last,
        never-factored column block (w = its width; lower triangle meaningful for symmetric factorizations)."""
        S = np.zeros(w * w, dtype=self.dtype)
        self.lib.pastix_getSchur.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.pastix_getSchur(self.pd, S.ctypes.data)
        return S.reshape(w, w, order="F")

    def clean(self):
        E = self.E
        if self.pd:
            self._call(E["API_TASK_CLEAN"], E["API_TASK_CLEAN"])
            self.pd = C.c_void_p(None)

    def out(self) -> dict:
        E = self.E
        return {
            "nnzeros": int(self.iparm[E["IPARM_NNZEROS"]]),
            "static_pivoting": int(self.iparm[E["IPARM_STATIC_PIVOTING"]]),
            "inertia": int(self.iparm[E["IPARM_INERTIA"]]),
            "nbiter": int(self.iparm[E["IPARM_NBITER"]]),
            "fact_flops": float(self.dparm[E["DPARM_FACT_FLOPS"]]),
            "fact_time": float(self.dparm[E["DPARM_FACT_TIME"]]),
            "solv_time": float(self.dparm[E["DPARM_SOLV_TIME"]]),
            "relative_error": float(self.dparm[E["DPARM_RELATIVE_ERROR"]]),
            "epsilon_magn_ctrl": float(self.dparm[E["DPARM_EPSILON_MAGN_CTRL"]]),
        }


class Pastix(PastixLib):
    """pastix() served by the drop-in library: reference host code + B200 numeric phase."""

    def __init__(self, prec: str = "d", threads: int = 1, verbose: int = 0, int_bits: int = 64):
        super().__init__(prec, dropin_path(prec, int_bits), load_enums(), threads=threads, verbose=verbose, int_bits=int_bits)

    # -- hooks of the drop-in (pastix_b200/shim/shim_hooks.c) ----------------------------------
    def handle(self) -> int:
        """pb200_handle_t* the shim keeps for this pastix_data (0 before the first NUMFACT)."""
        f = self.lib.pb200_shim_get_handle
        f.restype = C.c_void_p
        f.argtypes = [C.c_void_p]
        return int(f(self.pd) or 0)

    def critere(self) -> float:
        f = self.lib.pb200_shim_get_critere
        f.restype = C.c_double
        f.argtypes = [C.c_void_p]
        return float(f(self.pd))

    def order(self):
        """Final permutation kept by the reference (ordemesh): permtab[old]=new, peritab[new]=old, 0-based."""
        pt = np.zeros(self.n, dtype=np.int64); pi = np.zeros(self.n, dtype=np.int64)
        self.lib.pb200_shim_get_order.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.pb200_shim_get_order(self.pd, pt.ctypes.data, pi.ctypes.data)
        return pt, pi

    def solver(self) -> dict:
        """Flat copy of the SolverMatrix produced by the analysis (blend/src/solver.h:94-168), keys as in
        pb200_solver_t — the input of `Sopalin(...)` / pb200_create_dist for the multi-GPU path."""
        sz = np.zeros(2, dtype=np.int64)
        self.lib.pb200_shim_solver_sizes.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.pb200_shim_solver_sizes(self.pd, sz.ctypes.data)
        cb, bl = int(sz[0]), int(sz[1])
        a = {k: np.zeros(cb + 1, dtype=np.int64) for k in ("fcolnum", "lcolnum", "bloknum", "stride")}
        b = {k: np.zeros(bl, dtype=np.int64) for k in ("frownum", "lrownum", "cblknum", "coefind")}
        self.lib.pb200_shim_solver_get.argtypes = [C.c_void_p] * 9
        self.lib.pb200_shim_solver_get(self.pd, *[a[k].ctypes.data for k in ("fcolnum", "lcolnum", "bloknum", "stride")],
                                       *[b[k].ctypes.data for k in ("frownum", "lrownum", "cblknum", "coefind")])
        d = dict(cblknbr=cb, bloknbr=bl); d.update(a); d.update(b)
        return d

    def csc(self) -> dict:
        """Internal CSC after NUMFACT (CscOrdistrib — built on the device by shim_csc.c unless PB200_HOST_CSC=1):
        0-based, new numbering, rows sorted in every column."""
        s = np.zeros(4, dtype=np.int64)
        self.lib.pb200_shim_csc_sizes.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.pb200_shim_csc_sizes(self.pd, s.ctypes.data)
        if not s[3]:
            raise RuntimeError("internal CSC not filled (run numfact first)")
        n, nnz = int(s[0]), int(s[1])
        colptr = np.zeros(n + 1, dtype=np.int64); rows = np.zeros(nnz, dtype=np.int64)
        vals = np.zeros(nnz, dtype=self.dtype); tv = np.zeros(nnz, dtype=self.dtype) if s[2] else None
        self.lib.pb200_shim_csc_get.argtypes = [C.c_void_p] * 5
        self.lib.pb200_shim_csc_get.restype = C.c_int
        t = self.lib.pb200_shim_csc_get(self.pd, colptr.ctypes.data, rows.ctypes.data, vals.ctypes.data,
                                        tv.ctypes.data if tv is not None else None)
        return dict(colptr=colptr, rows=rows, vals=vals, tvals=tv, type=chr(t))

    def csc_device(self, want_t: bool) -> dict:
        """The internal CSC as it sits in HBM after the device-side CscOrdistrib (what the assembly kernel read)."""
        host = self.csc()
        n, nnz = len(host["colptr"]) - 1, len(host["rows"])
        colptr = np.zeros(n + 1, dtype=np.int64); rows = np.zeros(nnz, dtype=np.int64)
        vals = np.zeros(nnz, dtype=self.dtype); tv = np.zeros(nnz, dtype=self.dtype) if want_t else None
        f = self.lib.pb200_shim_csc_device_get
        f.argtypes = [C.c_void_p] * 5
        f.restype = C.c_int64
        got = f(self.pd, colptr.ctypes.data, rows.ctypes.data, vals.ctypes.data, tv.ctypes.data if want_t else None)
        if got != nnz:
            raise RuntimeError(f"device CSC not available / size mismatch ({got} vs {nnz})")
        return dict(colptr=colptr, rows=rows, vals=vals, tvals=tv)

    def sopalin(self):
        """The GPU numeric phase behind this pastix_data as a `Sopalin` (borrowed handle)."""
        from .sopalin import Sopalin, SolverMatrix
        h = self.handle()
        if not h:
            raise RuntimeError("no device handle yet: run numfact() first")
        s = Sopalin.from_handle(h, self.prec, self.facto)
        s.solver = SolverMatrix.from_dict(self.solver())      # panel shapes for get_cblk
        return s

    def live_entries(self) -> int:
        """Number of SolverMatrix entries the shim's side table holds (this precision's library)."""
        self.lib.pb200_shim_live.restype = C.c_int
        return int(self.lib.pb200_shim_live())

    def release(self):
        """Free the HBM held for this pastix_data (before clean())."""
        self.lib.pb200_shim_release_data.argtypes = [C.c_void_p]
        self.lib.pb200_shim_release_data(self.pd)
