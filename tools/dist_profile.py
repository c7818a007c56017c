"""Dev tool (torchrun, >= 2 GPUs): per-rank, per-level, per-kind serialised step times of a multi-GPU factorization
(PB200_PROFILE_VERBOSE) + the unserialised time, to see where a rank spends its share.
usage: torchrun --nproc-per-node N tools/dist_profile.py c3"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from pastix_b200.pastix_api import Pastix
from pastix_b200 import Sopalin, critere_from_norm
from pastix_b200.csc import internal_csc

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
r, w = dist.get_rank(), dist.get_world_size()
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
log = open(os.path.join(ROOT, "gpurun_out", f"distprof_{wl}_n{w}_rank{r}.txt"), "w")
os.dup2(log.fileno(), 2); os.dup2(log.fileno(), 1)
desc, kind, N, prec, facto, nrhs, over = bench.WORKLOADS[wl]
dt = bench.DT[prec]
A, perm0 = bench.case_matrix(kind, N, dt)
an = Pastix(prec, threads=1).setup(A, perm0, facto, sym=bench.SYM[facto], iparm_over=dict(over)).analyze()
flops = an.out()["fact_flops"]
sol = an.solver(); permtab, _ = an.order()
csc = internal_csc(A, permtab, bench.SYM[facto], dt)
s = Sopalin(sol, prec, facto, device=local, rank=r, nranks=w).attach()
s.assemble(csc["colptr"], csc["rows"], csc["values"], csc["tvalues"])
crit = critere_from_norm(s.norm1(csc["colptr"], csc["values"]))
for it in range(3):
    if it:
        s.reassemble()
    s.factorize(crit)
    print(f"[rank {r}] plain run {it}: {s.fact_time * 1e3:.2f} ms", file=sys.stderr, flush=True)
os.environ["PB200_TIMELINE"] = "1"
s.reassemble(); s.factorize(crit)
print(f"[rank {r}] timeline run: {s.fact_time * 1e3:.2f} ms", file=sys.stderr, flush=True)
del os.environ["PB200_TIMELINE"]
os.environ["PB200_PROFILE"] = "1"; os.environ["PB200_PROFILE_VERBOSE"] = "1"
s.reassemble(); s.factorize(crit)
print(f"[rank {r}] serialised run: {s.fact_time * 1e3:.2f} ms", file=sys.stderr, flush=True)
del os.environ["PB200_PROFILE"], os.environ["PB200_PROFILE_VERBOSE"]
dist.barrier()
s.close()
dist.destroy_process_group()
