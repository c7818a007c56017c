import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    for k in ("prec", "facto", "sym", "kind"):
        d[k] = str(d[k])
    for k in ("cblknbr", "bloknbr", "nbpivot", "inertia", "nnzeros", "N"):
        d[k] = int(d[k])
    for k in ("critere", "norm1", "fact_flops"):
        d[k] = float(d[k])
    for k in ("fcol", "lcol", "bloknum", "stride", "frow", "lrow", "fcblk", "coefind", "permtab", "colptr", "rows"):
        d[k] = d[k].astype(np.int64)
    d["schur"] = bool(int(d["schur"])) if "schur" in d else False
    d.setdefault("tvalues", None)
    d.setdefault("U", None)
    return d


def lower_mask(sol):
    """Boolean mask over the slab selecting what the reference defines for a
    symmetric factorization: everything except the strict upper triangle of each
    diagonal blok (never referenced: compute_diag.c works on 'L')."""
    cb = sol["cblknbr"]
    w = sol["lcol"][:cb] - sol["fcol"][:cb] + 1
    poff = np.concatenate([[0], np.cumsum(sol["stride"][:cb] * w)])
    mask = np.ones(int(poff[-1]), dtype=bool)
    for c in range(cb):
        ld = int(sol["stride"][c])
        for j in range(1, int(w[c])):
            mask[poff[c] + j * ld: poff[c] + j * ld + j] = False
    return mask


def tol(prec, kind="factor"):
    """Stated parity tolerances (relative to the largest magnitude of the array):
    double 1e-12 (north_star), single 2e-5."""
    return {"d": 1e-12, "z": 1e-12, "s": 2e-5, "c": 2e-5}[prec]


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
