"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/pastix_b200.h
declares; the drop-in host library exports the reference's numeric-phase entry points; without a
GPU the product path fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    from pastix_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pastix_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pastix_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in lib.pb200_version()


@pytest.mark.parametrize("prec,P", [("d", "D"), ("z", "Z"), ("s", "S"), ("c", "C")])
def test_dropin_exports_reference_entry_points(prec, P):
    """Names from sopalin3d.h:290-467 with the variant prefixes of sopalin_define.h:453-465; the
    precision prefix of redefine_functions.h:84-101 is empty in a one-precision-per-library build
    (no -DMULTIPLE_TYPE_DEFINE), which is how both the oracle and the drop-in are built."""
    from pastix_b200.pastix_api import dropin_path
    path = dropin_path(prec)
    if not os.path.exists(path):
        pytest.skip("drop-in library not built (needs the reference tree at build time)")
    lib = C.CDLL(path)
    names = ["pastix", "dpastix", "pastix_fortran"]
    for v in ("po", "sy", "he", "ge"):
        names += [f"{v}_sopalin_thread", f"{v}_sopalin_updo_thread", f"{v}_updo_thread", f"{v}_up_down_smp"]
        names += [f"{v}_sopalin_updo_gmres_thread", f"{v}_gmres_thread"]
    names += ["ge_sopalin_updo_pivot_thread", "ge_sopalin_updo_bicgstab_thread", "po_sopalin_updo_grad_thread"]
    names += ["pb200_shim_get_handle", "pb200_shim_get_critere", "pb200_shim_release_data"]
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from conftest import load_golden
    from pastix_b200 import Sopalin, PastixB200Error
    g = load_golden("lap7_6_llt_s")
    with pytest.raises(PastixB200Error, match="no CUDA device|CUDA"):
        Sopalin(g, "s", "llt")


def test_argument_errors_are_reported_before_any_device_work():
    """Bad arguments come back as PB200_ERR_BADARG with a message (no GPU needed to reach these checks)."""
    from conftest import load_golden
    from pastix_b200 import Sopalin, PastixB200Error
    g = load_golden("lap7_8_llt_d_schur")
    with pytest.raises(PastixB200Error, match="Schur mode is single-GPU"):
        Sopalin(g, "d", "llt", rank=0, nranks=2, schur=True)
    with pytest.raises(PastixB200Error, match="bad rank"):
        Sopalin(g, "d", "llt", rank=3, nranks=2)
    with pytest.raises(PastixB200Error, match="bad rank"):
        Sopalin(g, "d", "llt", rank=0, nranks=99)


_REFUSAL = r"""
import sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests/golden")
from make_golden import case_matrix, DT
from pastix_b200.pastix_api import Pastix
A, perm0 = case_matrix("lap7", 6, DT["d"])
p = Pastix("d").setup(A, perm0, "llt", iparm_over={over!r}).analyze()
p.numfact()
print("NUMFACT RETURNED")
"""


@pytest.mark.parametrize("over,msg", [({"IPARM_FILL_MATRIX": 1}, "IPARM_FILL_MATRIX"),
                                      ({"IPARM_DISTRIBUTION_LEVEL": 2}, "2D distribution"),
                                      ({}, "no CUDA device")])
def test_dropin_refuses_loudly(over, msg, tmp_path):
    """What the shim does not handle ends the way the reference's own fatal paths do (errorPrint + EXIT -> abort,
    common/src/errors.h:161-165) with a message naming the reason — never a silent different computation.  The last
    case: pastix(API_TASK_NUMFACT) on a machine without a GPU (skipped where one is present)."""
    import subprocess
    import sys
    import torch
    from pastix_b200.pastix_api import dropin_path
    if not os.path.exists(dropin_path("d")):
        pytest.skip("drop-in library not built")
    if not over and torch.cuda.is_available():
        pytest.skip("a GPU is present")
    script = tmp_path / "refuse.py"
    script.write_text(_REFUSAL.format(root=ROOT, over=over))
    # PB200_HOST_CSC=1: the internal CSC through the reference's host routine, so that on a machine without a GPU the
    # first device contact is the numeric phase itself (otherwise the device-side CscOrdistrib already reports "no CUDA")
    env = dict(os.environ, PB200_HOST_CSC="1") if over else dict(os.environ)
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode != 0 and "NUMFACT RETURNED" not in out.stdout
    assert msg in out.stdout + out.stderr, (out.stdout[-500:], out.stderr[-500:])


def test_flop_model_matches_reference_count():
    """DPARM_FACT_FLOPS of the golden fixtures (the metric's numerator, blend_symbol_cost.c:52-88)
    re-derived from the SolverMatrix arrays with the formulas of flops.h:74-117."""
    from conftest import golden_names, load_golden
    for name in golden_names():
        g = load_golden(name)
        if g["facto"] != "llt" or "ilu" in name:
            continue
        cb = g["cblknbr"]
        tot = 0.0
        fm, fa = (6.0, 2.0) if g["prec"] in ("c", "z") else (1.0, 1.0)     # complex: 6 per multiply, 2 per add (flops.h:211-271)
        for c in range(cb):
            n = int(g["lcol"][c] - g["fcol"][c] + 1); ld = int(g["stride"][c]); m = ld - n
            potrf = fm * (n ** 3 / 6 + n ** 2 / 2 + n / 3) + fa * (n ** 3 / 6 - n / 6)   # FMULS / FADDS_POTRF
            trsm = (fm + fa) * m * n * (n + 1) / 2                   # FADDS_TRSM is defined as FMULS_TRMM (flops.h:99-100)
            gemm = 0.0
            for b in range(int(g["bloknum"][c]) + 1, int(g["bloknum"][c + 1])):
                nk = int(g["lrow"][b] - g["frow"][b] + 1); mk = ld - int(g["coefind"][b])
                gemm += (fm + fa) * mk * nk * n
            tot += potrf + trsm + gemm
        assert abs(tot - g["fact_flops"]) <= 1e-9 * g["fact_flops"], (name, tot, g["fact_flops"])


@pytest.mark.parametrize("name,kind,N", [("lap7_8_llt_d", "lap7", 8), ("cd_8_lu_d", "cd", 8)])
def test_int32_dropin_analysis_gives_the_golden_structures(name, kind, N):
    """The drop-in built with the reference's DEFAULT 32-bit PASTIX_INT (libpastix_dropin_d_i32.so): pastix() runs
    ordering -> blend on the CPU (no GPU needed before NUMFACT) and the SolverMatrix the shim hands to the C ABI,
    widened to int64, is bit-identical to the one the 64-bit reference produced for the golden fixture."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from conftest import load_golden
    from make_golden import case_matrix, DT
    from pastix_b200.pastix_api import Pastix, dropin_path
    if not os.path.exists(dropin_path("d", 32)):
        pytest.skip("int32 drop-in not built (needs the reference tree at build time)")
    g = load_golden(name)
    A, perm0 = case_matrix(kind, N, DT["d"])
    p = Pastix("d", int_bits=32).setup(A, perm0, g["facto"], sym=g["sym"]).analyze()
    s = p.solver(); permtab, _ = p.order()
    assert s["cblknbr"] == g["cblknbr"] and s["bloknbr"] == g["bloknbr"]
    for k, k2 in (("fcolnum", "fcol"), ("lcolnum", "lcol"), ("bloknum", "bloknum"), ("stride", "stride"),
                  ("frownum", "frow"), ("lrownum", "lrow"), ("cblknum", "fcblk"), ("coefind", "coefind")):
        assert np.array_equal(s[k][:len(g[k2])], g[k2]), k
    assert np.array_equal(permtab, g["permtab"])
    assert p.out()["fact_flops"] == g["fact_flops"] and p.out()["nnzeros"] == g["nnzeros"]


def test_options_struct_layout_matches_the_python_binding(tmp_path):
    """pb200_options_t (include/pastix_b200.h) as a C compiler lays it out against the ctypes mirror the harness uses:
    same size, `owner` at the same offset (the struct grew a pointer in round 2 inside its reserved space)."""
    import subprocess
    from pastix_b200 import _lib
    src = tmp_path / "o.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pastix_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu\\n", sizeof(pb200_options_t), offsetof(pb200_options_t, schur), '
                   'offsetof(pb200_options_t, owner));return 0;}\n')
    exe = tmp_path / "o"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, off_schur, off_owner = map(int, subprocess.check_output([str(exe)]).split())
    assert size == C.sizeof(_lib.Options) == 32
    assert off_schur == _lib.Options.schur.offset == 0
    assert off_owner == _lib.Options.owner.offset == 8
