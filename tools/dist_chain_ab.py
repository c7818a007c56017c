"""Dev tool (torchrun, >= 2 GPUs): the two ways of dealing the shared column blocks of the top separators over the
GPUs (csrc/dist_plan.h, PB200_DIST_CHAIN=deal|group) on the same box, same analysis, back to back.
usage: torchrun --nproc-per-node N tools/dist_chain_ab.py [c2] [c3] [--json=path]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import scipy.sparse as sp
import bench
from pastix_b200.pastix_api import Pastix
from pastix_b200 import Sopalin, critere_from_norm, generators as G
from pastix_b200.csc import internal_csc, permute_rhs, unpermute_solution

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
r, w = dist.get_rank(), dist.get_world_size()
os.dup2(2, 1)


def maxr(v):
    t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = []
for wl in [a for a in sys.argv[1:] if not a.startswith("--")] or ["c2"]:
    desc, kind, N, prec, facto, nrhs, over = bench.WORKLOADS[wl]
    dt = bench.DT[prec]
    A, perm0 = bench.case_matrix(kind, N, dt)
    n = A.shape[0]
    an = Pastix(prec, threads=1).setup(A, perm0, facto, sym=bench.SYM[facto], iparm_over=dict(over)).analyze()
    flops = an.out()["fact_flops"]
    sol = an.solver(); permtab, _ = an.order()
    csc = internal_csc(A, permtab, bench.SYM[facto], dt)
    b = G.rhs_vector(n, 1, dt)
    Af = A if bench.SYM[facto] == "no" else (A + (sp.tril(A, -1).conj().T if bench.SYM[facto] == "her" else sp.tril(A, -1).T)).tocsc()
    for mode in ("deal", "group"):
        os.environ["PB200_DIST_CHAIN"] = mode
        s = Sopalin(sol, prec, facto, device=local, rank=r, nranks=w).attach()
        owner, contrib, load = Sopalin.dist_plan(sol, facto, w)
        s.assemble(csc["colptr"], csc["rows"], csc["values"], csc["tvalues"])
        crit = critere_from_norm(s.norm1(csc["colptr"], csc["values"]))
        ts = []
        for it in range(5):
            if it:
                s.reassemble()
            s.factorize(crit)
            ts.append(maxr(s.fact_time))
        x = permute_rhs(b, permtab); s.solve(x)
        berr = float(np.linalg.norm(Af @ unpermute_solution(x, permtab) - b) / np.linalg.norm(b))
        rec = {"workload": wl, "n_gpus": w, "chain": mode, "fact_ms": [t * 1e3 for t in ts], "best_ms": min(ts[2:]) * 1e3,
               "tflops": flops / min(ts[2:]) / 1e12, "load_share": [float(v) for v in load / load.sum()], "backward_error": berr}
        out.append(rec)
        if r == 0:
            print(json.dumps(rec), file=sys.stderr, flush=True)
        s.close()
        dist.barrier()
jp = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--json=")), None)
if r == 0 and jp:
    json.dump(out, open(jp, "w"), indent=1)
dist.barrier(); dist.destroy_process_group()
