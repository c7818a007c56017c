"""ctypes binding of include/pastix_b200.h.  Fails loudly when the CUDA library
is missing: there is no CPU fallback behind this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(_HERE, "lib", "libpastix_b200.so")   # PB200_LIB: tuning variants

# every symbol include/pastix_b200.h declares
SYMBOLS = [
    "pb200_last_error", "pb200_version", "pb200_create", "pb200_destroy", "pb200_info", "pb200_panel_offsets",
    "pb200_norm1", "pb200_assemble", "pb200_reassemble", "pb200_factorize", "pb200_inertia", "pb200_solve",
    "pb200_solve_device", "pb200_set_transpose_solve", "pb200_get_coeftab", "pb200_set_coeftab", "pb200_mark_factorized",
    "pb200_last_launches", "pb200_probe_fp64_gflops", "pb200_set_profile", "pb200_get_profile",
    "pb200_create_dist", "pb200_ipc_size", "pb200_ipc_export", "pb200_ipc_attach", "pb200_dist_barrier", "pb200_dist_plan",
    "pb200_vec_alloc", "pb200_vec_free", "pb200_vec_set", "pb200_vec_get", "pb200_vec_zero", "pb200_vec_copy", "pb200_vec_scal",
    "pb200_vec_axpy", "pb200_vec_dot", "pb200_csc_ax", "pb200_precond",
    "pb200_csc_create", "pb200_csc_destroy", "pb200_csc_build", "pb200_csc_fetch", "pb200_csc_fetch_colptr", "pb200_csc_norm1", "pb200_assemble_csc",
    "pb200_create_opts", "pb200_get_cblk", "pb200_set_hermitian", "pb200_attach_local", "pb200_destroy_group",
]


class SolverDesc(C.Structure):
    """pb200_solver_t"""
    _fields_ = [("cblknbr", C.c_int64), ("bloknbr", C.c_int64)] + \
        [(k, C.c_void_p) for k in ("fcolnum", "lcolnum", "bloknum", "stride", "frownum", "lrownum", "cblknum", "coefind")]


class Info(C.Structure):
    """pb200_info_t"""
    _fields_ = [("n", C.c_int64), ("coefnbr", C.c_int64), ("nlevels", C.c_int64), ("device_bytes", C.c_int64),
                ("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32)]


class Options(C.Structure):
    """pb200_options_t"""
    _fields_ = [("schur", C.c_int32), ("reserved0", C.c_int32), ("owner", C.c_void_p), ("reserved", C.c_int32 * 4)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m pastix_b200.build` "
            "(pastix_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.pb200_last_error.restype = C.c_char_p
    L.pb200_version.restype = C.c_char_p
    L.pb200_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(SolverDesc), C.c_int, C.c_int, C.c_int]
    L.pb200_create_dist.argtypes = [C.POINTER(C.c_void_p), C.POINTER(SolverDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pb200_create_opts.argtypes = [C.POINTER(C.c_void_p), C.POINTER(SolverDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(Options)]
    L.pb200_get_cblk.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.pb200_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
    L.pb200_ipc_attach.argtypes = [C.c_void_p, C.c_void_p]
    L.pb200_dist_barrier.argtypes = [C.c_void_p]
    L.pb200_dist_plan.argtypes = [C.POINTER(SolverDesc), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pb200_destroy.argtypes = [C.c_void_p]
    L.pb200_info.argtypes = [C.c_void_p, C.POINTER(Info)]
    L.pb200_panel_offsets.argtypes = [C.c_void_p, C.c_void_p]
    L.pb200_norm1.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.pb200_norm1.restype = C.c_double
    L.pb200_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pb200_reassemble.argtypes = [C.c_void_p]
    L.pb200_csc_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int]
    L.pb200_csc_destroy.argtypes = [C.c_void_p]
    L.pb200_csc_build.argtypes = [C.c_void_p, C.c_char, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.POINTER(C.c_int64)]
    L.pb200_csc_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pb200_csc_fetch_colptr.argtypes = [C.c_void_p, C.c_void_p]
    L.pb200_assemble_csc.argtypes = [C.c_void_p, C.c_void_p]
    L.pb200_csc_norm1.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.pb200_factorize.argtypes = [C.c_void_p, C.c_double, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.pb200_inertia.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.pb200_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
    L.pb200_solve_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_double)]
    L.pb200_set_transpose_solve.argtypes = [C.c_void_p, C.c_int]
    L.pb200_get_coeftab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pb200_set_coeftab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.pb200_mark_factorized.argtypes = [C.c_void_p]
    L.pb200_last_launches.argtypes = [C.c_void_p]
    L.pb200_last_launches.restype = C.c_int64
    L.pb200_set_profile.argtypes = [C.c_void_p, C.c_int]
    L.pb200_get_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.pb200_probe_fp64_gflops.argtypes = [C.c_int, C.c_int]
    L.pb200_probe_fp64_gflops.restype = C.c_double
    _lib = L
    return L
