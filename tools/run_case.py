"""Dev tool: run one synthetic case end to end on the GPU (analysis = the reference's unchanged host code inside
the drop-in library, numeric phase by pastix_b200 through the C ABI) and print timings and
the backward error.  usage: python tools/run_case.py N stencil facto prec [nrhs] [--ref] [--reps=K] [--json=path]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pastix_b200.pastix_api import Pastix  # noqa: E402
from pastix_b200 import Sopalin, critere_from_norm, generators as G  # noqa: E402
from pastix_b200.csc import internal_csc, permute_rhs, unpermute_solution  # noqa: E402

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    N = int(args[0]); stencil = args[1]; facto = args[2]; prec = args[3]
    nrhs = int(args[4]) if len(args) > 4 else 1
    dt = DT[prec]
    t0 = time.time()
    if stencil == "cd":
        A = G.convection_diffusion_3d(N, dt)
    else:
        A = G.laplacian_3d(N, int(stencil), dt)
    perm0 = G.nested_dissection_perm(N)
    sym = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}[facto]
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
    try:
        r = Pastix(prec, threads=1).setup(A, perm0, facto, sym=sym).analyze()
    finally:
        os.dup2(saved, 1)
    s = r.solver(); permtab, _ = r.order(); out = r.out()
    cb = s["cblknbr"]
    s["coefnbr"] = int(np.sum(s["stride"][:cb] * (s["lcolnum"][:cb] - s["fcolnum"][:cb] + 1)))
    t1 = time.time()
    print(f"analysis {t1 - t0:.2f}s: n={A.shape[0]} cblk={s['cblknbr']} blok={s['bloknbr']} coefnbr={s['coefnbr']} "
          f"flops={out['fact_flops']:.4g} nnzL={out['nnzeros']}")
    csc = internal_csc(A, permtab, sym, dt)
    eng = Sopalin(s, prec, facto)
    print(f"levels={eng.nlevels} device_bytes={eng.device_bytes / 1e9:.2f} GB")
    eng.assemble(csc["colptr"], csc["rows"], csc["values"], csc["tvalues"])
    crit = critere_from_norm(eng.norm1(csc["colptr"], csc["values"]))
    reps = next((int(a.split("=")[1]) for a in sys.argv if a.startswith("--reps=")), 3)
    rec = {"N": N, "stencil": stencil, "facto": facto, "prec": prec, "n": int(A.shape[0]), "cblknbr": int(s["cblknbr"]),
           "coefnbr": int(s["coefnbr"]), "fact_flops": float(out["fact_flops"]), "analysis_s": t1 - t0,
           "device_bytes": int(eng.device_bytes), "levels": int(eng.nlevels), "fact_ms": [], "solve_ms_per_rhs": []}
    for it in range(reps):
        if it:
            eng.reassemble()
        nb = eng.factorize(crit)
        print(f"factorize: {eng.fact_time * 1e3:.2f} ms  {out['fact_flops'] / eng.fact_time / 1e9:.1f} GFLOP/s "
              f"nbpivot={nb} launches={eng.last_launches()}", flush=True)
        rec["fact_ms"].append(eng.fact_time * 1e3)
    b = G.rhs_vector(A.shape[0], nrhs, dt)
    for it in range(2):
        x = permute_rhs(b, permtab)
        eng.solve(x)
        print(f"solve: {eng.solv_time * 1e3:.2f} ms ({eng.solv_time * 1e3 / nrhs:.3f} ms/rhs) launches={eng.last_launches()}", flush=True)
        rec["solve_ms_per_rhs"].append(eng.solv_time * 1e3 / nrhs)
    xs = unpermute_solution(x, permtab)
    Af = A if sym == "no" else (A + sp.tril(A, -1).T if sym == "yes" else A + sp.tril(A, -1).conj().T)
    res = np.linalg.norm(Af @ xs - b) / np.linalg.norm(b)
    print(f"backward error ||b-Ax||/||b|| = {res:.3e}")
    rec.update(backward_error=float(res), nbpivot=int(nb), gflops=float(out["fact_flops"] / (min(rec["fact_ms"]) * 1e-3) / 1e9))
    jp = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--json=")), None)
    if jp:
        import json
        json.dump(rec, open(jp, "w"))
    if "--ref" in sys.argv:      # comparison leg only: the unmodified reference on the host cores (test infrastructure)
        from oracle.refpastix import RefPastix
        r = RefPastix(prec, threads=os.cpu_count()).setup(A, perm0, facto, sym=sym).analyze()
        t = time.time(); r.numfact(); xr = r.solve(b)
        o = r.out()
        print(f"reference CPU ({os.cpu_count()} threads): fact {o['fact_time']:.3f}s solve {o['solv_time'] * 1e3:.1f} ms; "
              f"max |x-xref|/|xref| = {np.abs(xs - xr).max() / np.abs(xr).max():.3e}")


if __name__ == "__main__":
    main()
