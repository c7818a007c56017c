"""Dev tool: incomplete factorization + Krylov refinement (API_TASK_REFINE) on the drop-in (device vector back end) and
on the unmodified reference (host back end, all cores).  usage: python tools/refine_case.py N [gmres|grad|bicgstab] [level]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refpastix import RefPastix  # noqa: E402
from pastix_b200.pastix_api import Pastix  # noqa: E402
from pastix_b200 import generators as G  # noqa: E402
import scipy.sparse as sp  # noqa: E402

N = int(sys.argv[1]); meth = sys.argv[2] if len(sys.argv) > 2 else "gmres"; lvl = int(sys.argv[3]) if len(sys.argv) > 3 else 2
raf = {"gmres": "API_RAF_GMRES", "grad": "API_RAF_GRAD", "bicgstab": "API_RAF_BICGSTAB"}[meth]
A = G.laplacian_3d(N, 7, np.float64); perm0 = G.nested_dissection_perm(N)
b = G.rhs_vector(A.shape[0], 1, np.float64)[:, 0].copy()
Af = (A + sp.tril(A, -1).T).tocsc()
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
for name, cls, thr in (("drop-in (B200)", Pastix, 1), ("reference (CPU)", RefPastix, os.cpu_count())):
    os.dup2(devnull, 1)
    try:
        p = cls("d", threads=thr)
        over = {"IPARM_REFINEMENT": p.E[raf], "IPARM_ITERMAX": 250, "IPARM_GMRES_IM": 25, "IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": lvl}
        p.setup(A, perm0, "llt", iparm_over=over, dparm_over={"DPARM_EPSILON_REFINEMENT": 1e-10}).analyze().numfact()
        x = p.solve(b)
        t0 = time.perf_counter(); x = p.refine(b, x); t1 = time.perf_counter()
        o = p.out(); rt = float(p.dparm[p.E["DPARM_RAFF_TIME"]])
    finally:
        os.dup2(saved, 1)
    res = np.linalg.norm(Af @ x - b) / np.linalg.norm(b)
    print(f"{name:16s} ILU({lvl}) {N}^3 {meth}: {o['nbiter']} iterations, DPARM_RAFF_TIME {rt * 1e3:.1f} ms ({rt * 1e3 / max(o['nbiter'], 1):.2f} ms/iter), "
          f"refine call {1e3 * (t1 - t0):.1f} ms, fact {o['fact_time'] * 1e3:.1f} ms, ||b-Ax||/||b|| = {res:.2e}", flush=True)
    if cls is Pastix:
        p.release()
