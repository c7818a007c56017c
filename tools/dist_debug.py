"""Dev tool (torchrun, >= 2 GPUs): distributed vs single-GPU factors, cblk by cblk."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from pastix_b200.pastix_api import Pastix
from pastix_b200 import Sopalin, critere_from_norm, generators as G
from pastix_b200.csc import internal_csc, permute_rhs
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
r, w = dist.get_rank(), dist.get_world_size()
os.dup2(2, 1)
for kind, N, facto in [("lap7", 20, "llt"), ("lap7", 32, "llt"), ("lap7", 48, "llt"), ("lap7", 64, "llt"), ("lap27", 40, "ldlt")]:
    A, perm0 = bench.case_matrix(kind, N, np.float64)
    an = Pastix("d").setup(A, perm0, facto).analyze()
    sol = an.solver(); permtab, _ = an.order()
    csc = internal_csc(A, permtab, "yes", np.float64)
    one = Sopalin(sol, "d", facto, device=local)
    one.assemble(csc["colptr"], csc["rows"], csc["values"]); crit = critere_from_norm(one.norm1(csc["colptr"], csc["values"]))
    one.factorize(crit); L1, _ = one.get_coeftab(); one.close()
    s = Sopalin(sol, "d", facto, device=local, rank=r, nranks=w).attach()
    owner, contrib, load = Sopalin.dist_plan(sol, facto, w)
    cb = sol["cblknbr"]; wd = sol["lcolnum"][:cb] - sol["fcolnum"][:cb] + 1
    poff = np.concatenate([[0], np.cumsum(sol["stride"][:cb] * wd)])
    level = np.zeros(cb, int)
    for c in range(cb):
        for b in range(sol["bloknum"][c] + 1, sol["bloknum"][c + 1]):
            fc = sol["cblknum"][b]; level[fc] = max(level[fc], level[c] + 1)
    for it in range(3):
        s.assemble(csc["colptr"], csc["rows"], csc["values"])
        s.factorize(crit)
        L2, _ = s.get_coeftab()
        bad = []
        for c in range(cb):
            a, b_ = L1[poff[c]:poff[c + 1]], L2[poff[c]:poff[c + 1]]
            ld = int(sol["stride"][c]); m = np.ones(a.size, bool)
            for j in range(1, int(wd[c])): m[j * ld: j * ld + j] = False
            e = np.max(np.abs(a[m] - b_[m])) / max(np.max(np.abs(a[m])), 1e-300)
            if e > 1e-10: bad.append((c, int(level[c]), int(owner[c]), int(contrib[c]), int(wd[c]), ld, float(e)))
        if r == 0:
            print(f"{kind} N={N} {facto} it={it}: cblk={cb} levels={level.max()+1} bad={len(bad)} first={bad[:6]}", file=sys.stderr, flush=True)
    s.close()
dist.barrier(); dist.destroy_process_group()
