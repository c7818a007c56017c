#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on the B200: numeric factorization GFLOP/s (PaStiX's own flop
count, DPARM_FACT_FLOPS) and % of measured FP64 peak on the 3-D Laplacian, plus solve ms/RHS.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|...] [--impl reference]

One "step" = one pass of the sopalin numeric phase over the synthetic matrix: device-side assembly
of the panels from the CSC resident in HBM, numeric factorization, one up_down solve.  `value`
follows SURVEY.md §8(d): DPARM_FACT_FLOPS / the DPARM_FACT_TIME interval (factorization only,
device-timed with CUDA events on the launching stream; assembly and solve are reported beside it and
all three are inside `ms_per_step`).  `e2e` is the same metric through the reference-facing call a
PaStiX user makes — pastix(API_TASK_NUMFACT) then pastix(API_TASK_SOLVE) on the drop-in library with
HOST buffers (host CSC/RHS -> HBM and the solution back inside the timed region).
`roofline` is for the dominant kernel (k_gemm_scatter, the fused GEMM + scatter-add of the
supernodal updates): its algorithmic flops (PaStiX's GEMM term) / its summed CUDA-event duration.
`cpu_baseline` / `--impl reference` time the UNMODIFIED reference's threaded CPU sopalin
(oracle/_ref, built from /root/reference in the build container) on the box's host cores.

Multi-GPU (--gpus N>1, launched by torch.distributed.run): see DESIGN.md "Multi-GPU".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# name -> (description, stencil kind, N, precision, factorization, nrhs, iparm overrides)
WORKLOADS = {
    "c1": ("1-D Laplacian n=100 LLt double (example/bin/simple -lap 100)", "lap1d", 100, "d", "llt", 1, {}),
    "c2": ("3-D 7-point Laplacian 64^3 (n=262144) LLt double, nested dissection", "lap7", 64, "d", "llt", 1, {}),
    "c3": ("3-D 27-point Laplacian 100^3 (n=1000000) LDLt double, nested dissection", "lap27", 100, "d", "ldlt", 1, {}),
    "c2s": ("3-D 7-point Laplacian 64^3 (n=262144) LLt SINGLE precision, nested dissection", "lap7", 64, "s", "llt", 1, {}),
    "c3s": ("3-D 27-point Laplacian 100^3 (n=1000000) LDLt SINGLE precision, nested dissection", "lap27", 100, "s", "ldlt", 1, {}),
    "c4s": ("complex-double convection-diffusion 64^3 LU static pivoting (config 4 at single-GPU size)", "cd", 64, "z", "lu", 1, {}),
    "c4": ("complex-double convection-diffusion 128^3 LU static pivoting", "cd", 128, "z", "lu", 1, {}),
    "c5s": ("blockwise ILU(2) + 64-RHS solve, 3-D 7-point Laplacian 64^3", "lap7", 64, "d", "llt", 64,
            {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 2}),
    "c5m": ("blockwise ILU(2) + 64-RHS solve, 3-D 7-point Laplacian 100^3 (the largest size whose ILU(2) host analysis fits a GPU lease)", "lap7", 100, "d", "llt", 64,
            {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 2}),
    "c5": ("blockwise ILU(2) + 64-RHS solve, 3-D 7-point Laplacian 200^3", "lap7", 200, "d", "llt", 64,
           {"IPARM_INCOMPLETE": 1, "IPARM_LEVEL_OF_FILL": 2}),
}
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
SYM = {"llt": "yes", "ldlt": "yes", "lu": "no", "ldlh": "her"}
ESIZE = {"s": 4, "d": 8, "c": 8, "z": 16}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def case_matrix(kind, N, dt):
    from pastix_b200 import generators as G
    if kind == "lap1d":
        return G.laplacian_1d(N, dt), G.nested_dissection_perm_1d(N)
    if kind == "lap7":
        return G.laplacian_3d(N, 7, dt), G.nested_dissection_perm(N)
    if kind == "lap27":
        return G.laplacian_3d(N, 27, dt), G.nested_dissection_perm(N)
    if kind == "cd":
        return G.convection_diffusion_3d(N, dt), G.nested_dissection_perm(N)
    raise ValueError(kind)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w": float(np.median(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- FP64 roof
def measure_fp64_peak(device: int) -> dict:
    """MEASURED_PEAKS.json carries HBM GB/s and bf16 TF/s only; the FP64 denominators are measured here,
    live: our own DMMA (mma.sync.m16n8k8.f64) issue-rate probe and a cuBLAS DGEMM 8192^3 burst."""
    import torch
    from pastix_b200 import _lib
    L = _lib.lib()
    out = {"dmma_probe_tflops": L.pb200_probe_fp64_gflops(device, 2) / 1e3,
           "dfma_probe_tflops": L.pb200_probe_fp64_gflops(device, 0) / 1e3}
    n = 8192
    a = torch.randn(n, n, device=f"cuda:{device}", dtype=torch.float64)
    b = torch.randn(n, n, device=f"cuda:{device}", dtype=torch.float64)
    for _ in range(2):
        c = a @ b
    torch.cuda.synchronize(device)
    best = 1e30
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(device)
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    out["cublas_dgemm_tflops"] = 2.0 * n ** 3 / best / 1e9
    out["peak_tflops"] = max(out["dmma_probe_tflops"], out["cublas_dgemm_tflops"])
    out["source"] = "measured live in bench.py (max of own DMMA probe and cuBLAS DGEMM 8192^3 burst); MEASURED_PEAKS.json has no FP64 figure"
    return out


def measure_fp32_peak(device: int) -> dict:
    """Single-precision roofs, measured live: cuBLAS SGEMM 8192^3 in true FP32 (the SGEMM-class roof the s / c path is
    held against) and with TF32 tensor cores allowed (our kernels issue THREE tf32 MMAs per FP32 product — 3xTF32 —
    so a third of that figure is their tensor-pipe ceiling)."""
    import torch
    n = 8192
    a = torch.randn(n, n, device=f"cuda:{device}", dtype=torch.float32)
    b = torch.randn(n, n, device=f"cuda:{device}", dtype=torch.float32)
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    for name, tf in (("cublas_sgemm_fp32_tflops", False), ("cublas_sgemm_tf32_tflops", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf
        for _ in range(2):
            c = a @ b
        torch.cuda.synchronize(device)
        best = 1e30
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize(device)
            best = min(best, e0.elapsed_time(e1))
        out[name] = 2.0 * n ** 3 / best / 1e9
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b, c
    torch.cuda.empty_cache()
    out["tf32_over_3_tflops"] = out["cublas_sgemm_tf32_tflops"] / 3.0
    out["peak_tflops"] = out["cublas_sgemm_fp32_tflops"]
    out["source"] = ("measured live in bench.py: cuBLAS SGEMM 8192^3 burst in true FP32 (allow_tf32 = False); the TF32 figure / 3 is "
                     "the ceiling of the 3xTF32 tensor path")
    return out


def ncu_traffic(workload: str = "c2"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel (k_gemm_scatter: the largest
    captured launch) and of the two up_down sweeps, from the committed `ncu --set full` captures
    (profiles/r02/r02_full_gemm_scatter_*_summary.json and r02c_full_updown_c2_summary.json — the k_dag3 sweeps —, written by
    tools/gpu_profile.sh + tools/ncu_summary.py); None when absent."""
    wl = workload if workload in ("c2", "c3") else "c2"
    out = None
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02", f"r02_full_gemm_scatter_{wl}_summary.json")))
        k = max(d, key=lambda x: x.get("grid", 0))
        out = {"dram_bytes": k["dram_read_bytes"] + k["dram_write_bytes"], "launch_tiles": int(k["grid"]),
               "launch_ms": k["duration_us"] * 1e-3, "dmma_pipe_active_pct": k.get("dmma_pipe_pct"),
               "l2_hit_pct": k.get("l2_hit_pct"), "capture": f"{wl}, largest captured launch (ncu --set full, cold cache)"}
    except Exception:
        pass
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02", "r02c_full_updown_c2_summary.json")))
        if out is not None and wl == "c2":
            out["updown_sweeps"] = [{"kernel": k["kernel"].replace("void ", "").split("<")[0], "dram_bytes": k["dram_read_bytes"] + k["dram_write_bytes"],
                                     "launch_ms": k["duration_us"] * 1e-3} for k in d]
    except Exception:
        pass
    return out


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p)); d["_which"] = "measured"
            return d
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "_which": "fallback"}


def bench_config(args) -> dict:
    """`config` of the JSON line: identical in both arms (ours / --impl reference) for the same command line."""
    desc = WORKLOADS[args.workload][0]
    over = WORKLOADS[args.workload][6]
    return {"workload": f"{args.workload}: {desc}",
            "ordering": "geometric nested dissection passed as API_ORDER_PERSONAL (Scotch is not in the image)",
            "iparm": {k: int(v) for k, v in sorted(over.items())},
            "l2": "inputs larger than L2: the factor slab is rewritten by the assembly at the start of every step",
            "gpus": int(args.gpus),
            "multi_gpu": ("single GPU" if args.gpus == 1 else
                          "ONE factorization spread over the GPUs (strong scaling): proportional subtree mapping + fan-in over NVLink peer memory")}


def sample_cblks(cblknbr: int, k: int = 64):
    """Evenly spaced column blocks plus the last ones (the top of the elimination tree, where every update has landed)."""
    idx = set(np.linspace(0, cblknbr - 1, num=min(k, cblknbr)).astype(int).tolist())
    idx.update(range(max(0, cblknbr - 8), cblknbr))
    return sorted(idx)


def _sol_view(sol: dict) -> dict:
    """The two flat SolverMatrix dicts (Pastix.solver() / RefPastix.solver()) under one naming."""
    g = lambda a, b: sol[a] if a in sol else sol[b]
    return dict(cblknbr=int(sol["cblknbr"]), fcol=g("fcolnum", "fcol"), lcol=g("lcolnum", "lcol"), bloknum=sol["bloknum"],
                stride=sol["stride"], frow=g("frownum", "frow"), lrow=g("lrownum", "lrow"))


def factor_columns(sol: dict, get_panel, cols) -> dict:
    """Columns of the factor as {global column: (global rows, values)} read through `get_panel(c)` (stride x width).
    Independent of how blend split the column blocks: IPARM_THREAD_NBR changes the split, not L."""
    v = _sol_view(sol)
    out = {}
    fcol = np.asarray(v["fcol"][:v["cblknbr"]])
    for j in cols:
        c = int(np.searchsorted(fcol, j, side="right") - 1)
        P = get_panel(c)
        rows = np.concatenate([np.arange(v["frow"][b], v["lrow"][b] + 1) for b in range(int(v["bloknum"][c]), int(v["bloknum"][c + 1]))])
        col = P[:, j - int(fcol[c])]
        keep = rows >= j                                   # lower part: the strict upper triangle of a diagonal blok is undefined
        out[j] = (rows[keep], col[keep])
    return out


def columns_relerr(a: dict, b: dict) -> float:
    """max |a - b| / max |b| over the sampled columns (rows matched by global index; a row absent on one side counts
    as a structural zero)."""
    num, den = 0.0, 1e-300
    for j in a:
        ra, va = a[j]; rb, vb = b[j]
        da = dict(zip(ra.tolist(), va.tolist())); db = dict(zip(rb.tolist(), vb.tolist()))
        for r in set(da) | set(db):
            num = max(num, abs(da.get(r, 0.0) - db.get(r, 0.0))); den = max(den, abs(db.get(r, 0.0)))
    return float(num / den)


# ----------------------------------------------------------------------------- reference (CPU) arm
def run_reference_fact(wl: str, threads: int, steps: int, warmup: int, budget_s: float, N_override: int | None = None):
    """The UNMODIFIED reference (oracle/_ref) through its own pastix(): analysis once, then NUMFACT + SOLVE
    per step, timed by the reference's own DPARM_FACT_TIME / DPARM_SOLV_TIME."""
    from oracle.refpastix import RefPastix, available
    from pastix_b200 import generators as G
    desc, kind, N, prec, facto, nrhs, over = WORKLOADS[wl]
    if N_override:
        N = N_override
    if not available(prec):
        return None
    A, perm0 = case_matrix(kind, N, DT[prec])
    over = dict(over)
    r = RefPastix(prec, threads=threads).setup(A, perm0, facto, sym=SYM[facto], iparm_over=over).analyze()
    flops = r.out()["fact_flops"]
    b = G.rhs_vector(A.shape[0], nrhs, DT[prec])
    ft, st, wall = [], [], []
    t_begin = time.time()
    done = 0
    for it in range(warmup + steps):
        t0 = time.time()
        r.numfact()
        x = r.solve(b)
        o = r.out()
        if it >= warmup:
            ft.append(o["fact_time"]); st.append(o["solv_time"]); wall.append(time.time() - t0); done += 1
        if time.time() - t_begin > budget_s and done >= 1:
            break
    Af = A if SYM[facto] == "no" else None
    return {"flops": flops, "fact_s": float(np.mean(ft)), "solve_s": float(np.mean(st)), "wall_s": float(np.mean(wall)),
            "steps_run": done, "N": N, "n": A.shape[0], "nrhs": nrhs, "x": x, "A": A, "b": b, "threads": threads,
            "ref": r, "nbpivot": r.out()["static_pivoting"]}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    desc, kind, N, prec, facto, nrhs, over = WORKLOADS[args.workload]
    # a step = the workload's numeric factorization when that runs in seconds on the host cores (c1, c2, c4s, c5s), else the
    # same problem on a smaller grid (a bounded sample: GFLOP/s is the metric) — the arm must end within minutes
    Ns, what = bounded_sample(args.workload)
    r = run_reference_fact(args.workload, cores, args.steps, min(args.warmup, 1), float(os.environ.get("PB200_REF_BUDGET_S", "150")),
                           N_override=Ns)
    if r is None:
        return {"impl": "reference", "unavailable": "oracle/_ref (the reference compiled from /root/reference) is not present"}
    gf = r["flops"] / r["fact_s"] / 1e9
    unit = "GFLOP/s"
    sample = f"{what}, {r['steps_run']} timed step(s) (DPARM_FACT_TIME)"
    return {
        "impl": "reference", "metric": "numeric factorization throughput (PaStiX flop count)", "value": gf, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "steps_run": r["steps_run"], "warmup": args.warmup,
        "ms_per_step": r["wall_s"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"d": "f64", "z": "c128", "s": "f32", "c": "c64"}[prec], "data": "synthetic",
        "config": bench_config(args), "threads": cores,
        "fact_ms": r["fact_s"] * 1e3, "solve_ms_per_rhs": r["solve_s"] * 1e3 / r["nrhs"],
        "cpu_baseline": {"value": gf, "unit": unit, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": gf, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ----------------------------------------------------------------------------- our arm
def our_arm(args, with_cpu_baseline: bool = False):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — pastix_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(local)

    def maxr(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import generators as G
    from pastix_b200.csc import permute_rhs, unpermute_solution

    desc, kind, N, prec, facto, nrhs, over = WORKLOADS[args.workload]
    dt = DT[prec]
    A, perm0 = case_matrix(kind, N, dt)
    n = A.shape[0]
    b = G.rhs_vector(n, nrhs, dt)
    t0 = time.time()
    gpu = Pastix(prec, threads=1).setup(A, perm0, facto, sym=SYM[facto], iparm_over=dict(over)).analyze()
    t_analysis = time.time() - t0
    flops = gpu.out()["fact_flops"]
    nnzL = gpu.out()["nnzeros"]
    log(f"[rank {rank}] {args.workload}: n={n} nnz(A)={A.nnz} nnzL={nnzL} flops={flops:.4g} analysis {t_analysis:.1f}s")

    # ---- e2e: the reference-facing call with host buffers
    e2e_fact, e2e_solve = [], []
    x = None
    for it in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        gpu.numfact()                     # pastix(API_TASK_NUMFACT): host CSC -> HBM, assembly, factorization
        t1 = time.perf_counter()
        x = gpu.solve(b)                  # pastix(API_TASK_SOLVE): host b -> HBM, up_down, x back
        t2 = time.perf_counter()
        if it >= args.warmup:
            e2e_fact.append(t1 - t0); e2e_solve.append(t2 - t1)
    e2e_fact_s = maxr(float(np.mean(e2e_fact))); e2e_solve_s = maxr(float(np.mean(e2e_solve)))
    # backward error of the solution the user got (north_star: <= 1e-12 in double, direct factorizations)
    import scipy.sparse as sp
    Af = A if SYM[facto] == "no" else (A + (sp.tril(A, -1).conj().T if SYM[facto] == "her" else sp.tril(A, -1).T)).tocsc()
    berr = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))
    nnzA_int = Af.nnz
    h2d = (n + 1) * 8 + nnzA_int * 4 + nnzA_int * ESIZE[prec] * (2 if facto == "lu" else 1) + n * nrhs * ESIZE[prec]
    d2h = n * nrhs * ESIZE[prec] + 8

    # ---- device-resident steps (inputs already in HBM)
    s = gpu.sopalin()
    crit = gpu.critere()
    permtab, _ = gpu.order()
    xp = permute_rhs(b, permtab)
    x_src = torch.from_numpy(np.ascontiguousarray(xp.T)).to(f"cuda:{local}")      # (nrhs, n) row-major == n x nrhs column-major
    x_dev = torch.empty_like(x_src)
    fact_s, solve_s, asm_s, launches = [], [], [], 0
    sampler = ClockSampler(local)
    for it in range(args.warmup):
        s.reassemble(); s.factorize(crit); x_dev.copy_(x_src); torch.cuda.synchronize(local)
        s.solve_device(x_dev.data_ptr(), n, nrhs)
    barrier()
    if rank == 0:
        sampler.start()
    t_begin = time.perf_counter()
    for it in range(args.steps):
        ta = time.perf_counter()
        s.reassemble()
        asm_s.append(time.perf_counter() - ta)
        s.factorize(crit)
        fact_s.append(s.fact_time)
        launches += s.last_launches() + 1       # factorization (incl. the triangle inversions) + assembly
        x_dev.copy_(x_src); torch.cuda.synchronize(local)
        s.solve_device(x_dev.data_ptr(), n, nrhs)
        solve_s.append(s.solv_time)
        launches += s.last_launches()
    barrier()
    t_region = maxr(time.perf_counter() - t_begin)
    clocks = sampler.stop() if rank == 0 else None
    fact_mean = maxr(float(np.mean(fact_s)))
    solve_mean = maxr(float(np.mean(solve_s)))
    # solution check of the device-resident path
    xh = x_dev.cpu().numpy()
    xs = unpermute_solution(xh.T if xh.ndim == 2 else xh.reshape(n, 1), permtab)
    berr_dev = float(np.linalg.norm(Af @ xs.reshape(n, -1) - b) / np.linalg.norm(b))

    # ---- roofline of the dominant kernel: serialised launches, CUDA events per kernel kind
    roof = None
    prof = None
    if rank == 0:
        s.set_profile(True)
        for _ in range(2):
            s.reassemble(); s.factorize(crit)
        prof = s.get_profile()
        s.set_profile(False)
        s.reassemble(); s.factorize(crit)
        peak = measure_fp64_peak(local) if prec in ("d", "z") else measure_fp32_peak(local)
        gms = prof["ms"]["gemm_scatter"]
        tot = sum(prof["ms"].values())
        ach = prof["gemm_flops"] / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
        traf = ncu_traffic(args.workload)
        roof = {"kernel": "k_gemm_scatter (fused %s GEMM + scatter-add into facing cblks)" % ("DMMA" if prec in ("d", "z") else "3xTF32 MMA"), "bound": "tensor",
                "achieved": ach, "peak": peak["peak_tflops"], "unit": "TFLOP/s", "frac": ach / peak["peak_tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (the largest one captured by ncu --set full;
                # which launch, its duration and pipe utilisation are in traffic_capture)
                "traffic": traf["dram_bytes"] if traf else None, "traffic_capture": traf, "peak_source": peak["source"], "measured_peaks": {k: v for k, v in peak.items() if k.endswith("tflops")},
                "kernel_share_of_step": gms / tot if tot > 0 else None,
                "kind_ms_serialised": prof["ms"], "kind_launches": prof["launches"],
                "algorithmic_flops_per_factorization": prof["gemm_flops"]}
        hp = measured_peaks()
        bytes_solve = 2.0 * nnzL * ESIZE[prec] * (1 if facto != "lu" else 1) + 6.0 * n * nrhs * ESIZE[prec]
        roof["solve"] = {"bound": "hbm", "achieved": bytes_solve / solve_mean / 1e9, "peak": hp.get("hbm_gbs"),
                         "unit": "GB/s", "frac": bytes_solve / solve_mean / 1e9 / hp.get("hbm_gbs", 6650.0),
                         "peak_source": f"MEASURED_PEAKS.json ({hp['_which']})", "algorithmic_bytes": bytes_solve}
    gf_total = world * flops / fact_mean / 1e9
    line = None
    if rank == 0:
        line = {
            "metric": "numeric factorization throughput (PaStiX flop count)", "value": gf_total, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_region / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"d": "f64", "z": "c128", "s": "f32", "c": "c64"}[prec], "data": "synthetic",
            "config": bench_config(args),
            "problem": {"n": n, "nnzL": nnzL, "fact_flops": flops, "nrhs": nrhs, "cblknbr": int(s.solver.cblknbr),
                        "factor_slab_GB": s.coefnbr * ESIZE[prec] * (2 if facto == "lu" else 1) / 1e9},
            "fact_ms": fact_mean * 1e3, "assemble_ms": float(np.mean(asm_s)) * 1e3,
            "solve_ms_per_rhs": solve_mean * 1e3 / nrhs,
            ("pct_fp64_peak" if prec in ("d", "z") else "pct_fp32_sgemm_peak"): 100.0 * (flops / fact_mean / 1e12) / roof["peak"],
            "backward_error": berr_dev,
            "e2e": {"value": world * flops / e2e_fact_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "numfact_call_ms": e2e_fact_s * 1e3, "solve_call_ms": e2e_solve_s * 1e3,
                    "backward_error": berr, "host_memory": "pageable (the reference's own CSC/RHS buffers)",
                    "call": "pastix(API_TASK_NUMFACT) + pastix(API_TASK_SOLVE) on libpastix_dropin"},
            "gpu_launches": int(launches),
            "roofline": roof, "clocks": clocks, "analysis_s": t_analysis,
        }
    if line is not None and with_cpu_baseline:
        line["cpu_baseline"], line["parity"] = cpu_baseline_for(args, gpu=gpu, x_gpu=x)
    gpu.clean()                                            # API_TASK_CLEAN: releases the HBM through the intercepted solverExit
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return line


def our_arm_dist(args):
    """N > 1: ONE factorization spread over the N GPUs of the box (strong scaling): proportional subtree
    mapping of the column blocks, fan-in contributions pulled by the owner over NVLink peer memory
    (DESIGN.md §6).  Launched as one process per GPU (torch.distributed.run): every rank runs the same deterministic
    host analysis and drives its own GPU through the C ABI (pb200_create_dist + CUDA IPC); `value` = the whole job's
    DPARM_FACT_FLOPS / max over ranks of the device-timed factorization.  `e2e` is the call a PaStiX user makes:
    pastix(API_TASK_NUMFACT) with iparm[IPARM_CUDA_NBR] = N on the drop-in, ONE process (rank 0) driving the N devices
    with host buffers, while the other ranks wait on a CPU-side barrier."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — pastix_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cpu_group = dist.new_group(backend="gloo")       # waits that must not occupy a GPU
    ctx = dict(torch=torch, dist=dist, rank=rank, world=world, local=local, cpu_group=cpu_group)
    line = dist_case(args, args.workload, args.steps if args.workload != "c3" else min(args.steps, 3), args.warmup, ctx)
    if args.default_workload and not args.iparm and os.environ.get("PB200_ALSO", "1") != "0":
        # the configuration the N = 1 line is quoted on, at the same N
        try:
            l2 = dist_case(args, "c2", args.steps, args.warmup, ctx)
            if line is not None and l2 is not None:
                line["also"] = {"c2": {k: l2[k] for k in ("value", "unit", "fact_ms", "solve_ms_per_rhs", "backward_error", "ms_per_step",
                                                         "gpu_launches", "e2e", "factor_relerr_vs_n1", "n1_same_run", "load_share", "device_bytes_per_gpu")}}
                line["also"]["c2"]["workload"] = l2["config"]["workload"]
        except Exception as e:
            if line is not None:
                line["also"] = {"c2": {"error": repr(e)}}
    dist.barrier(group=cpu_group)
    dist.destroy_process_group()
    return line


def dist_case(args, workload, steps, warmup, ctx):
    torch, dist = ctx["torch"], ctx["dist"]
    rank, world, local, cpu_group = ctx["rank"], ctx["world"], ctx["local"], ctx["cpu_group"]

    def barrier():
        dist.barrier(); torch.cuda.synchronize(local)

    def maxr(v: float) -> float:
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from pastix_b200.pastix_api import Pastix
    from pastix_b200 import Sopalin, critere_from_norm, generators as G
    from pastix_b200.csc import internal_csc, permute_rhs, unpermute_solution
    import scipy.sparse as sp
    desc, kind, N, prec, facto, nrhs, over = WORKLOADS[workload]
    dt = DT[prec]
    A, perm0 = case_matrix(kind, N, dt)
    n = A.shape[0]
    b = G.rhs_vector(n, nrhs, dt)
    t0 = time.time()
    an = Pastix(prec, threads=1).setup(A, perm0, facto, sym=SYM[facto], iparm_over=dict(over)).analyze()
    t_analysis = time.time() - t0
    flops = an.out()["fact_flops"]; nnzL = an.out()["nnzeros"]
    solver = an.solver(); permtab, _ = an.order()
    csc = internal_csc(A, permtab, SYM[facto], dt)
    s = Sopalin(solver, prec, facto, device=local, rank=rank, nranks=world).attach()
    owner, contrib, load = Sopalin.dist_plan(solver, facto, world)
    crit = critere_from_norm(s.norm1(csc["colptr"], csc["values"]))
    xp = permute_rhs(b, permtab)
    Af = A if SYM[facto] == "no" else (A + (sp.tril(A, -1).conj().T if SYM[facto] == "her" else sp.tril(A, -1).T)).tocsc()
    log(f"[rank {rank}] {workload}: n={n} nnzL={nnzL} flops={flops:.4g} analysis {t_analysis:.1f}s "
        f"load share {load[rank] / load.sum():.3f} device_bytes {s.device_bytes / 1e9:.2f} GB")
    s.assemble(csc["colptr"], csc["rows"], csc["values"], csc["tvalues"])
    # ---- device-resident steps: inputs already in HBM
    x_src = torch.from_numpy(np.ascontiguousarray(xp.T)).to(f"cuda:{local}"); x_dev = torch.empty_like(x_src)
    fact_s, solve_s, launches = [], [], 0
    sampler = ClockSampler(local)
    for it in range(warmup):
        s.reassemble(); s.factorize(crit)
    barrier()
    if rank == 0:
        sampler.start()
    t_begin = time.perf_counter()
    for it in range(steps):
        s.reassemble()
        s.factorize(crit)
        fact_s.append(s.fact_time); launches += s.last_launches() + 1
        x_dev.copy_(x_src); torch.cuda.synchronize(local)
        s.solve_device(x_dev.data_ptr(), n, nrhs)      # first solve after a factorization pulls the peers' panels
        solve_s.append(s.solv_time); launches += s.last_launches() + 3
        barrier()                                      # nobody re-assembles while a peer still pulls panels
    t_region = maxr(time.perf_counter() - t_begin)
    clocks = sampler.stop() if rank == 0 else None
    fact_mean = maxr(float(np.mean(fact_s))); solve_mean = maxr(float(np.mean(solve_s)))
    xh = x_dev.cpu().numpy()
    xs = unpermute_solution(xh.T if xh.ndim == 2 else xh.reshape(n, 1), permtab)
    berr_dev = float(np.linalg.norm(Af @ xs.reshape(n, -1) - b) / np.linalg.norm(b))
    # ---- the same factors as one GPU computes: sampled cblks of this run against a single-GPU factorization
    rel_n1 = None; n1_same_run = None
    if rank == 0 and 2.2 * s.device_bytes > 170e9:
        rel_n1 = "skipped: a second, single-GPU copy of the factors does not fit beside this rank's slab"
    elif rank == 0:
        try:
            s1 = Sopalin(solver, prec, facto, device=local)
            s1.assemble(csc["colptr"], csc["rows"], csc["values"], csc["tvalues"]); s1.factorize(crit)
            t1 = []
            for _ in range(2):
                s1.reassemble(); s1.factorize(crit); t1.append(s1.fact_time)
            n1_same_run = {"fact_ms": min(t1) * 1e3, "value": flops / min(t1) / 1e9, "unit": "GFLOP/s",
                           "what": "the same factorization on ONE of these GPUs (rank 0), same process, same run"}
            num, den = 0.0, 1e-300
            for c in sample_cblks(solver["cblknbr"]):
                w = int(solver["lcolnum"][c] - solver["fcolnum"][c] + 1)
                Pn = s.get_cblk(c); P1 = s1.get_cblk(c)
                if facto != "lu":                      # the strict upper triangle of the diagonal blok is undefined
                    iu = np.triu_indices(w, 1); Pn[:w][iu] = 0; P1[:w][iu] = 0
                num = max(num, float(np.max(np.abs(Pn - P1)))); den = max(den, float(np.max(np.abs(P1))))
            rel_n1 = num / den
            s1.close()
        except Exception as e:
            rel_n1 = repr(e)
    dev_bytes = maxr(float(s.device_bytes))
    barrier()
    s.close()                                          # collective
    an.clean()
    # ---- e2e: pastix() with iparm[IPARM_CUDA_NBR] = N, one process, host buffers
    dist.barrier(group=cpu_group)
    e2e = None
    if rank == 0:
        try:
            g = Pastix(prec, threads=1).setup(A, perm0, facto, sym=SYM[facto],
                                              iparm_over=dict(over, IPARM_CUDA_NBR=world)).analyze()
            tf, ts = [], []
            for it in range(warmup + steps):
                t0 = time.perf_counter(); g.numfact(); t1 = time.perf_counter(); x = g.solve(b); t2 = time.perf_counter()
                if it >= warmup:
                    tf.append(t1 - t0); ts.append(t2 - t1)
            berr = float(np.linalg.norm(Af @ x - b) / np.linalg.norm(b))
            nnzA = Af.nnz
            e2e = {"value": flops / float(np.mean(tf)) / 1e9, "unit": "GFLOP/s",
                   "h2d_bytes_per_step": int((n + 1) * 8 + nnzA * 8 + nnzA * ESIZE[prec] + n * nrhs * ESIZE[prec]),
                   "d2h_bytes_per_step": int(n * nrhs * ESIZE[prec] + 8),
                   "numfact_call_ms": float(np.mean(tf)) * 1e3, "solve_call_ms": float(np.mean(ts)) * 1e3, "backward_error": berr,
                   "fact_ms_dparm": g.out()["fact_time"] * 1e3,
                   "host_memory": "pageable (the reference's own CSC/RHS buffers)",
                   "call": "pastix(API_TASK_NUMFACT) + pastix(API_TASK_SOLVE) on libpastix_dropin with iparm[IPARM_CUDA_NBR] = %d "
                           "(one process driving the %d GPUs)" % (world, world)}
            g.clean()
        except Exception as e:
            e2e = {"value": None, "error": repr(e)}
    dist.barrier(group=cpu_group)
    line = None
    if rank == 0:
        a2 = argparse.Namespace(**vars(args)); a2.workload = workload
        line = {
            "metric": "numeric factorization throughput (PaStiX flop count)", "value": flops / fact_mean / 1e9, "unit": "GFLOP/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t_region / steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"d": "f64", "z": "c128", "s": "f32", "c": "c64"}[prec], "data": "synthetic",
            "config": bench_config(a2),
            "problem": {"n": n, "nnzL": nnzL, "fact_flops": flops, "nrhs": nrhs, "cblknbr": int(solver["cblknbr"])},
            "load_share": [float(v) for v in load / load.sum()], "device_bytes_per_gpu": dev_bytes,
            "fact_ms": fact_mean * 1e3, "solve_ms_per_rhs": solve_mean * 1e3 / nrhs, "backward_error": berr_dev,
            "fact_ms_covers": "numeric factorization + inversion of the diagonal triangles of the owned cblks (what N = 1 times)",
            "factor_relerr_vs_n1": rel_n1, "n1_same_run": n1_same_run,
            "e2e": e2e, "gpu_launches": int(launches), "roofline": None, "cpu_baseline": None, "clocks": clocks, "analysis_s": t_analysis,
        }
    return line


def bounded_sample(workload: str):
    """(grid size override or None, description) of what the CPU arm factorizes for this workload."""
    if workload == "c3" and os.environ.get("PB200_REF_FULL"):
        return None, "the full workload's numeric factorization (PB200_REF_FULL=1)"
    if workload == "c3":
        return 64, "same 27-point LDLt problem on a 64^3 grid (bounded sample of the 100^3 workload)"
    if workload in ("c4", "c4s"):
        return 32, "same complex LU problem on a 32^3 grid (bounded sample)"
    if workload in ("c5", "c5m"):
        return 64, "same ILU(2) problem on a 64^3 grid (bounded sample)"
    return None, "the full workload's numeric factorization"


def cpu_baseline_for(args, gpu=None, x_gpu=None):
    """Reference CPU sopalin on the host cores, bounded sample (rank 0, N=1 only).  When the sample IS the workload
    (c2, c4s, c5s) the reference's results are also compared with the drop-in's: solution, pivot count and sampled
    columns of the factor (structure-independent: the two analyses run blend with different IPARM_THREAD_NBR)."""
    cores = os.cpu_count() or 1
    desc, kind, N, prec, facto, nrhs, over = WORKLOADS[args.workload]
    # bounded sample: the full workload when it is <= ~1e12 flop, else the same stencil on a smaller grid
    Ns, sample = bounded_sample(args.workload)
    sample += ", one run"
    try:
        r = run_reference_fact(args.workload, cores, 1, 0, 120.0, N_override=Ns)
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": "GFLOP/s", "cores": cores, "kind": "reference", "sample": f"failed: {e}"}, None
    if r is None:
        return {"value": None, "unit": "GFLOP/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref not present"}, None
    cb = {"value": r["flops"] / r["fact_s"] / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "reference",
          "sample": sample, "fact_s": r["fact_s"], "solve_ms_per_rhs": r["solve_s"] * 1e3 / r["nrhs"]}
    parity = None
    if Ns is None and gpu is not None and not over.get("IPARM_INCOMPLETE"):
        try:
            ref = r["ref"]
            xr = np.asarray(r["x"]).reshape(r["n"], -1); xg = np.asarray(x_gpu).reshape(r["n"], -1)
            parity = {"against": "the unmodified reference (oracle/_ref) on the same pastix() calls, %d threads" % cores,
                      "x_relerr": float(np.max(np.abs(xg - xr)) / np.max(np.abs(xr))),
                      "nbpivot_equal": bool(gpu.out()["static_pivoting"] == r["nbpivot"])}
            s = gpu.sopalin()
            sol_g, sol_r = gpu.solver(), ref.solver()
            cbl = sample_cblks(sol_g["cblknbr"])
            cols = sorted({int(sol_g["fcolnum"][c]) for c in cbl} | {int(sol_g["lcolnum"][c]) for c in cbl})
            Lr, Ur = ref.coef()
            vr = _sol_view(sol_r)
            wr = np.asarray(vr["lcol"][:vr["cblknbr"]]) - np.asarray(vr["fcol"][:vr["cblknbr"]]) + 1
            poff = np.concatenate([[0], np.cumsum(np.asarray(vr["stride"][:vr["cblknbr"]]) * wr)]).astype(np.int64)
            ref_panel = lambda M: (lambda c: M[poff[c]:poff[c + 1]].reshape(int(vr["stride"][c]), int(wr[c]), order="F"))
            parity["L_relerr_sampled_cblks"] = columns_relerr(factor_columns(sol_g, lambda c: s.get_cblk(c), cols),
                                                              factor_columns(sol_r, ref_panel(Lr), cols))
            if facto == "lu":
                parity["U_relerr_sampled_cblks"] = columns_relerr(
                    factor_columns(sol_g, lambda c: s.get_cblk(c, with_u=True)[1], cols), factor_columns(sol_r, ref_panel(Ur), cols))
            parity["sampled_columns"] = len(cols)
        except Exception as e:  # the headline line must survive
            parity = {"error": repr(e)}
    return cb, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("PB200_WORKLOAD"), choices=sorted(WORKLOADS),
                    help="default: c2 (BASELINE.json configs[1], the configuration the metric is quoted on) at --gpus 1; "
                         "c3 (configs[2], 100^3 27-point LDLt 'on 1 and 8 B200') when ONE factorization is spread over "
                         "N > 1 GPUs — C2 holds 0.43 TFLOP, 23 ms on one GPU, nothing to spread; the other one rides in `also`")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--iparm", action="append", default=[], metavar="IPARM_NAME=VALUE",
                    help="extra iparm override for the analysis (e.g. IPARM_MAX_BLOCKSIZE=240); applied to both arms")
    ap.add_argument("--tuned", action="store_true",
                    help="GPU-aware block sizes for blend (IPARM_MIN/MAX_BLOCKSIZE = 120/240 instead of the reference's 60/120; "
                         "pb200_tune_iparm) — applied to both arms, like --iparm")
    args = ap.parse_args()
    if args.tuned:
        from pastix_b200.pastix_api import TUNED_IPARM
        args.iparm = list(args.iparm) + [f"{k}={v}" for k, v in TUNED_IPARM.items()]
    args.default_workload = args.workload is None
    if args.workload is None:
        args.workload = "c2" if args.gpus == 1 else "c3"
    if args.iparm:
        d, k, N, p, f, nr, over = WORKLOADS[args.workload]
        over = dict(over)
        for kv in args.iparm:
            key, val = kv.split("=")
            over[key] = int(val)
        WORKLOADS[args.workload] = (d + " [" + ",".join(args.iparm) + "]", k, N, p, f, nr, over)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per stream (multi-GPU flag waits must not alias)
    # the reference's analysis prints to stdout: keep fd 1 for the single JSON line
    real_out = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        line = reference_arm(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world > 1 and not os.environ.get("PB200_REPLICAS"):
            line = our_arm_dist(args)
        else:
            line = our_arm(args, with_cpu_baseline=(args.gpus == 1 and not args.no_cpu_baseline))
        if line is not None:
            line.setdefault("cpu_baseline", None)
    # the north_star target config beside the headline one (N=1 default run only): C3 = 100^3 27-point LDLt
    if (args.impl == "ours" and line is not None and args.gpus == 1 and args.workload == "c2" and args.default_workload and not args.iparm
            and os.environ.get("PB200_ALSO", "1") != "0"):
        try:
            a2 = argparse.Namespace(**vars(args)); a2.workload = "c3"; a2.steps = min(args.steps, 3)
            l3 = our_arm(a2)
            line["also"] = {"c3": {k: l3[k] for k in ("value", "unit", "fact_ms", "assemble_ms", "solve_ms_per_rhs", "pct_fp64_peak",
                                                     "backward_error", "ms_per_step", "gpu_launches")}}
            line["also"]["c3"]["workload"] = l3["config"]["workload"]
            line["also"]["c3"]["e2e"] = l3["e2e"]
            line["also"]["c3"]["roofline"] = {k: l3["roofline"][k] for k in ("achieved", "peak", "unit", "frac", "kernel_share_of_step")}
            line["also"]["c3"]["solve_roofline"] = l3["roofline"]["solve"]
        except Exception as e:  # the headline line must survive a failure of the extra config
            line["also"] = {"c3": {"error": repr(e)}}
    sys.stdout.flush()
    if rank == 0 and line is not None:
        os.write(real_out, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
