"""Host-side mirror of the reference's numeric-phase interface.

`Sopalin` plays the role of the reference's `Sopalin_Data_t` + the
`{po,sy,he,ge}_sopalin_thread` / `*_updo_thread` entry points
(src/sopalin/src/sopalin3d.c:1388,1467; updo.c:67): it is built from the
SolverMatrix the unchanged blend analysis produced, is fed the internal
(permuted) CSC, and runs NUMFACT / SOLVE on the GPU through the C ABI in
include/pastix_b200.h.  Names follow the reference's domain: cblk, blok,
coeftab, ucoeftab, critere, nbpivot, sm2xtab.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

FLTTYPE = {"s": 0, "d": 1, "c": 2, "z": 3}                 # API_REALSINGLE.. (common/src/api.h)
FACTO = {"llt": 0, "ldlt": 1, "lu": 2, "ldlh": 3}          # API_FACT_*      (common/src/api.h:381-384)
DTYPE = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
DPARM_EPSILON_MAGN_CTRL_DEFAULT = 1e-31                    # pastix.c:448


class PastixB200Error(RuntimeError):
    pass


def _check(rc: int):
    if rc != 0:
        raise PastixB200Error(f"pastix_b200 error {rc}: {_lib.lib().pb200_last_error().decode()}")


@dataclass
class SolverMatrix:
    """Flat copy of the fields of the reference's SolverMatrix this path reads
    (blend/src/solver.h:94-168). Arrays are int64; `bloknum` has cblknbr+1 entries."""
    cblknbr: int
    bloknbr: int
    fcolnum: np.ndarray
    lcolnum: np.ndarray
    bloknum: np.ndarray
    stride: np.ndarray
    frownum: np.ndarray
    lrownum: np.ndarray
    cblknum: np.ndarray
    coefind: np.ndarray

    @classmethod
    def from_dict(cls, d: dict) -> "SolverMatrix":
        g = lambda *names: next(np.ascontiguousarray(d[k], dtype=np.int64) for k in names if k in d)
        cb = int(d["cblknbr"])
        return cls(cb, int(d["bloknbr"]), g("fcolnum", "fcol")[:cb], g("lcolnum", "lcol")[:cb],
                   g("bloknum")[:cb + 1], g("stride")[:cb], g("frownum", "frow"), g("lrownum", "lrow"),
                   g("cblknum", "fcblk"), g("coefind"))

    @property
    def n(self) -> int:
        return int(self.lcolnum[-1]) + 1

    def panel_offsets(self) -> np.ndarray:
        w = self.lcolnum - self.fcolnum + 1
        return np.concatenate([[0], np.cumsum(self.stride * w)]).astype(np.int64)


def critere_from_norm(norm1: float, epsilon_magn_ctrl: float = DPARM_EPSILON_MAGN_CTRL_DEFAULT) -> float:
    """Static-pivot threshold of init_struct_sopalin (sopalin3d.c:586-606):
    ||A||_1 * sqrt(eps), or |eps| taken as an absolute threshold when eps < 0."""
    if epsilon_magn_ctrl < 0:
        return -epsilon_magn_ctrl
    return norm1 * float(np.sqrt(epsilon_magn_ctrl))


class Sopalin:
    """GPU numeric phase bound to one SolverMatrix."""

    @classmethod
    def from_handle(cls, handle: int, prec: str, facto: str) -> "Sopalin":
        """Wrap a pb200_handle_t* owned by someone else (the drop-in shim keeps one per SolverMatrix)."""
        self = cls.__new__(cls)
        self.solver, self.prec, self.facto = None, prec, facto
        self.dtype = np.dtype(DTYPE[prec])
        self.L = _lib.lib()
        self.h = C.c_void_p(handle)
        self._borrowed = True
        self._read_info()
        return self

    def _read_info(self):
        info = _lib.Info()
        _check(self.L.pb200_info(self.h, C.byref(info)))
        self.n, self.coefnbr, self.nlevels = int(info.n), int(info.coefnbr), int(info.nlevels)
        self.device_bytes, self.device, self.sm_count = int(info.device_bytes), int(info.device), int(info.sm_count)
        self.cc = (int(info.cc_major), int(info.cc_minor))
        self.nbpivot = 0
        self.fact_time = 0.0
        self.solv_time = 0.0

    def set_profile(self, on: bool):
        _check(self.L.pb200_set_profile(self.h, int(on)))

    def get_profile(self) -> dict:
        ms = (C.c_double * 4)(); n = (C.c_int64 * 4)(); fl = C.c_double(0)
        _check(self.L.pb200_get_profile(self.h, ms, n, C.byref(fl)))
        names = ("diag", "trsm", "gemm_scatter", "inpanel_update")
        return {"ms": dict(zip(names, list(ms))), "launches": dict(zip(names, [int(v) for v in n])), "gemm_flops": float(fl.value)}

    def __init__(self, solver: SolverMatrix | dict, prec: str = "d", facto: str = "llt", device: int = -1,
                 rank: int = 0, nranks: int = 1, schur: bool = False):
        """rank/nranks > 1: one process per GPU; call `attach()` (collective) before assembling.
        schur: IPARM_SCHUR semantics — the last cblk is never factored (after `factorize` its panel is the Schur
        complement, `get_schur()`), and `solve` ignores it (sopalin_compute.c:767-772, updo.c:425-428)."""
        self._borrowed = False
        self.schur = bool(schur)
        self.rank, self.nranks = rank, nranks
        if isinstance(solver, dict):
            solver = SolverMatrix.from_dict(solver)
        self.solver, self.prec, self.facto = solver, prec, facto
        self.dtype = np.dtype(DTYPE[prec])
        self.L = _lib.lib()
        p = lambda a: a.ctypes.data
        self._desc = _lib.SolverDesc(solver.cblknbr, solver.bloknbr, p(solver.fcolnum), p(solver.lcolnum),
                                     p(solver.bloknum), p(solver.stride), p(solver.frownum), p(solver.lrownum),
                                     p(solver.cblknum), p(solver.coefind))
        self.h = C.c_void_p(None)
        opts = _lib.Options(int(self.schur))
        _check(self.L.pb200_create_opts(C.byref(self.h), C.byref(self._desc), FLTTYPE[prec], FACTO[facto], device,
                                        rank, nranks, C.byref(opts)))
        self._read_info()

    # -- multi-GPU plumbing ----------------------------------------------------
    def attach(self, dist=None):
        """Exchange the CUDA IPC blobs of the slabs over torch.distributed (any backend; the blobs are
        pb200_ipc_size() = 256 opaque bytes per rank) and map every peer's slab (pb200_ipc_attach).  Collective."""
        if self.nranks == 1:
            return self
        import torch
        import torch.distributed as td
        dist = dist or td
        nb = int(self.L.pb200_ipc_size())
        blob = (C.c_ubyte * nb)()
        _check(self.L.pb200_ipc_export(self.h, blob))
        mine = torch.tensor(list(bytes(blob)), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            mine = mine.cuda()
        allb = [torch.empty_like(mine) for _ in range(self.nranks)]
        dist.all_gather(allb, mine)
        raw = b"".join(bytes(t.cpu().tolist()) for t in allb)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        _check(self.L.pb200_ipc_attach(self.h, buf))
        return self

    def barrier(self):
        _check(self.L.pb200_dist_barrier(self.h))

    @staticmethod
    def dist_plan(solver, facto: str, nranks: int):
        """The column-block -> GPU mapping alone (host only): (owner, contrib mask, flops per rank)."""
        if isinstance(solver, dict):
            solver = SolverMatrix.from_dict(solver)
        L = _lib.lib()
        p = lambda a: a.ctypes.data
        desc = _lib.SolverDesc(solver.cblknbr, solver.bloknbr, p(solver.fcolnum), p(solver.lcolnum), p(solver.bloknum),
                               p(solver.stride), p(solver.frownum), p(solver.lrownum), p(solver.cblknum), p(solver.coefind))
        owner = np.zeros(solver.cblknbr, dtype=np.int32); contrib = np.zeros(solver.cblknbr, dtype=np.uint32)
        load = np.zeros(nranks, dtype=np.float64)
        _check(L.pb200_dist_plan(C.byref(desc), FACTO[facto], nranks, owner.ctypes.data, contrib.ctypes.data, load.ctypes.data))
        return owner, contrib, load

    def close(self):
        if self.h and not self._borrowed:
            self.L.pb200_destroy(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- assembly (CoefMatrix_Init / Csc2solv_cblk) ---------------------------
    def norm1(self, colptr, values) -> float:
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=self.dtype)
        return float(self.L.pb200_norm1(FLTTYPE[self.prec], len(colptr) - 1, colptr.ctypes.data, values.ctypes.data))

    def assemble(self, colptr, rows, values, tvalues=None):
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=self.dtype)
        tv = np.ascontiguousarray(tvalues, dtype=self.dtype) if tvalues is not None else None
        _check(self.L.pb200_assemble(self.h, colptr.ctypes.data, rows.ctypes.data, values.ctypes.data,
                                     tv.ctypes.data if tv is not None else None))

    def reassemble(self):
        _check(self.L.pb200_reassemble(self.h))

    # -- NUMFACT -------------------------------------------------------------
    def factorize(self, critere: float) -> int:
        nb, sec = C.c_int64(0), C.c_double(0)
        _check(self.L.pb200_factorize(self.h, float(critere), C.byref(nb), C.byref(sec)))
        self.nbpivot, self.fact_time = int(nb.value), float(sec.value)
        return self.nbpivot

    def inertia(self) -> int:
        v = C.c_int64(0)
        _check(self.L.pb200_inertia(self.h, C.byref(v)))
        return int(v.value)

    # -- SOLVE ---------------------------------------------------------------
    def solve(self, x: np.ndarray) -> np.ndarray:
        """x: (n,) or Fortran-ordered (n, nrhs), permuted ordering; overwritten with the solution."""
        if x.dtype != self.dtype or not (x.ndim == 1 or x.flags.f_contiguous):
            raise ValueError("x must be a column-major array of the factorization's dtype")
        nrhs = 1 if x.ndim == 1 else x.shape[1]
        sec = C.c_double(0)
        _check(self.L.pb200_solve(self.h, x.ctypes.data, x.shape[0], nrhs, C.byref(sec)))
        self.solv_time = float(sec.value)
        return x

    def solve_device(self, x_ptr: int, ldx: int, nrhs: int) -> float:
        sec = C.c_double(0)
        _check(self.L.pb200_solve_device(self.h, C.c_void_p(x_ptr), ldx, nrhs, C.byref(sec)))
        self.solv_time = float(sec.value)
        return self.solv_time

    # -- coeftab / ucoeftab ----------------------------------------------------
    def get_coeftab(self):
        Lh = np.empty(self.coefnbr, dtype=self.dtype)
        Uh = np.empty(self.coefnbr, dtype=self.dtype) if self.facto == "lu" else None
        _check(self.L.pb200_get_coeftab(self.h, Lh.ctypes.data, Uh.ctypes.data if Uh is not None else None))
        return Lh, Uh

    def get_cblk(self, c: int, with_u: bool = False):
        """coeftab[c] (and ucoeftab[c]) as stride x width column-major arrays."""
        off = self.solver.panel_offsets()
        w = int(self.solver.lcolnum[c] - self.solver.fcolnum[c] + 1); ld = int(self.solver.stride[c])
        Lh = np.empty(int(off[c + 1] - off[c]), dtype=self.dtype)
        Uh = np.empty_like(Lh) if with_u else None
        _check(self.L.pb200_get_cblk(self.h, c, Lh.ctypes.data, Uh.ctypes.data if with_u else None))
        Lh = Lh.reshape(ld, w, order="F")
        return (Lh, Uh.reshape(ld, w, order="F")) if with_u else Lh

    def get_schur(self) -> np.ndarray:
        """The Schur complement left in the last cblk by a `schur=True` factorization (what pastix_getSchur copies,
        pastix.c:6434-6475): w x w, lower triangle meaningful for the symmetric factorizations, full for LU."""
        return self.get_cblk(self.solver.cblknbr - 1)

    def set_coeftab(self, Lh, Uh=None, factorized: bool = False):
        Lh = np.ascontiguousarray(Lh, dtype=self.dtype)
        Uh = np.ascontiguousarray(Uh, dtype=self.dtype) if Uh is not None else None
        _check(self.L.pb200_set_coeftab(self.h, Lh.ctypes.data, Uh.ctypes.data if Uh is not None else None))
        if factorized:
            _check(self.L.pb200_mark_factorized(self.h))

    def last_launches(self) -> int:
        return int(self.L.pb200_last_launches(self.h))
