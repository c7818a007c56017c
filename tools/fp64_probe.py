"""Dev tool: FP64 roof probes on the GPU box (our micro-kernels + cuBLAS DGEMM via torch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pastix_b200 import _lib
L = _lib.lib()
for v, name in enumerate(["DFMA", "DMMA m8n8k4", "DMMA m16n8k8", "DMMA m16n8k16"]):
    print(f"{name}: {L.pb200_probe_fp64_gflops(0, v):.0f} GFLOP/s")
for w in (4, 8, 12, 16, 32):
    print(f"DMMA m16n8k8, {w} warps/SM: {L.pb200_probe_fp64_gflops(0, 100 + w):.0f} GFLOP/s")
for c in (1, 2, 3, 4, 6, 8):
    print(f"gemm main loop (smem fragments, 32x32 warp tiles), {c} CTAs x 4 warps per SM: no barrier {L.pb200_probe_fp64_gflops(0, 200 + c):.0f}, "
          f"barrier per 16-k chunk {L.pb200_probe_fp64_gflops(0, 300 + c):.0f} GFLOP/s")
for dt, nm in ((torch.float64, "DGEMM"), (torch.float32, "SGEMM"), (torch.complex128, "ZGEMM")):
    n = 8192 if dt != torch.complex128 else 4096
    torch.backends.cuda.matmul.allow_tf32 = False
    a = torch.randn(n, n, device="cuda", dtype=dt); b = torch.randn(n, n, device="cuda", dtype=dt)
    for _ in range(2): c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    f = 2 * n ** 3 * (4 if dt == torch.complex128 else 1)
    print(f"cuBLAS {nm} {n}^3: {f / best / 1e6:.0f} GFLOP/s ({best:.2f} ms)")
