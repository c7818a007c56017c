"""Dev tool for ncu: host analysis + warm-up steps outside the capture range, then ONE step of the hot
path (device assembly, numeric factorization, up_down) between cudaProfilerStart/Stop.
usage: ncu --profile-from-start off ... python tools/profile_step.py [workload]   (workloads: bench.py)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from pastix_b200.pastix_api import Pastix  # noqa: E402
from pastix_b200 import generators as G  # noqa: E402
from pastix_b200.csc import permute_rhs  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
desc, kind, N, prec, facto, nrhs, over = bench.WORKLOADS[wl]
os.dup2(2, 1)
A, perm0 = bench.case_matrix(kind, N, bench.DT[prec])
gpu = Pastix(prec).setup(A, perm0, facto, sym=bench.SYM[facto], iparm_over=dict(over)).analyze().numfact()
s = gpu.sopalin(); crit = gpu.critere(); permtab, _ = gpu.order()
n = A.shape[0]
xp = permute_rhs(G.rhs_vector(n, nrhs, bench.DT[prec]), permtab)
x_src = torch.from_numpy(np.ascontiguousarray(xp.T)).cuda(); x_dev = x_src.clone()
for _ in range(2):
    s.reassemble(); s.factorize(crit); x_dev.copy_(x_src); torch.cuda.synchronize(); s.solve_device(x_dev.data_ptr(), n, nrhs)
torch.cuda.synchronize()
torch.cuda.profiler.start()
s.reassemble(); s.factorize(crit); x_dev.copy_(x_src); torch.cuda.synchronize(); s.solve_device(x_dev.data_ptr(), n, nrhs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{wl}: fact {s.fact_time * 1e3:.2f} ms solve {s.solv_time * 1e3:.3f} ms", file=sys.stderr)
gpu.release()
