// engine.cu — host side of the C ABI (include/pastix_b200.h): flattens the
// SolverMatrix into device arrays, builds the elimination-tree level schedule
// and drives the CUDA kernels.  No CPU compute path exists here: every numeric
// entry point launches kernels or fails.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <type_traits>

#include "../../include/pastix_b200.h"
#include "kernels_factor.cuh"
#include "kernels_solve.cuh"
#include "kernels_solve_dag.cuh"
#include "kernels_solve_dag2.cuh"
#include "kernels_solve_dag3.cuh"
#include "kernels_mma.cuh"
#include "kernels_dist.cuh"
#include "kernels_small.cuh"
#include "kernels_raff.cuh"
#include "dist_plan.h"
#include "csc_build.h"

using namespace pb200;

#define PB200_SLAB_PAD 4096

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(PB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct pb200_handle_s {
  int flt = 0, facto = 0, device = 0;
  size_t esize = 0;
  int64_t cblknbr = 0, bloknbr = 0, n = 0, coefnbr = 0;
  int wmax = 0, smax = 0, nlevels = 0, sm_count = 0, cc_major = 0, cc_minor = 0;
  size_t device_bytes = 0;
  std::vector<int> h_fcol, h_width, h_stride, h_fblok, h_frow, h_nrow, h_fcblk, h_coefind;
  std::vector<int64_t> h_poff;
  // level schedule (host copies of the per-level extents)
  std::vector<int> lvl_ptr;       // cblks of level l: d_lvl_cblk[lvl_ptr[l] .. lvl_ptr[l+1])
  std::vector<int> trsm_ptr, trsm_tiles, upd_ptr, slv_ptr, slv_tiles;
  std::vector<long long> upd_tiles;
  DevSym S{};
  int *d_lvl_cblk = nullptr;
  RowTask *d_trsm = nullptr, *d_slv = nullptr;
  UpdTask *d_upd = nullptr;
  void *dL = nullptr, *dU = nullptr;
  void *dW = nullptr;   // LDLt / LDLh on the tensor path: L*D beside the panels (same layout), written by the TRSM, read by the updates
  // resident CSC
  int64_t nnz = 0;
  int64_t *d_colptr = nullptr;
  int *d_rows = nullptr;
  void *d_vals = nullptr, *d_tvals = nullptr;
  unsigned long long *d_cnt = nullptr;  // [0] nbpivot [1] dropped [2] inertia
  void *d_x = nullptr; size_t x_bytes = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t last_launches = 0;
  bool assembled = false, factorized = false;
  // ---- solve schedule (all precisions)
  struct GStep { int kind; int l0, l1; };   // 0: classic level (separate launches), 1: one fused small-cblk launch, 2: single-CTA chain of thin levels
  struct SlvStep { int task0, ntasks; long long ntiles, t2t0; };
  std::vector<int> slv_lvl_ptr, slv_lvl_step;   // full (all-GPU) level lists; slv_steps of level l: [slv_lvl_step[l], slv_lvl_step[l+1])
  std::vector<GStep> sgsteps;                  // up_down schedule over levels (kinds as gsteps)
  int *d_slv_cblk = nullptr, *d_slv_lvl_ptr = nullptr;
  int64_t *d_rmbase = nullptr;
  bool slv_all_small = false;
  void *d_xt = nullptr; size_t xt_bytes = 0;
  std::vector<SlvStep> slv_steps;          // ascending (level, round)
  SlvTask *d_slvtask = nullptr; int *d_slv_t2t = nullptr;
  int64_t *d_invoff = nullptr; int64_t inv_elems = 0; int nsubpanels = 0;
  void *d_inv = nullptr, *d_inv_up = nullptr;   // inverted diagonal triangles (LU: L and U^T)
  unsigned int *d_slv_cnt = nullptr;
  // persistent counter-ordered sweeps (kernels_solve_dag.cuh)
  bool dag_ok = false; int dag_tiles = 0, dag_nbs = 0;
  int dag_rows = PB200_DAG_ROWS;                                 // panel rows per T sub-tile of the general list: 64 (k_fwd_dag / k_bwd_dag), 32 (k_dag2, PB200_DAG_V2=1)
  bool dag3_ok = false; int dag3_GD = 0, dag3_GT = 0;            // third generation (one right-hand side): D and T ticket lists
  DagTick *d_dag3_ticksD = nullptr, *d_dag3_ticksT = nullptr; int *d_dag3_tgt = nullptr;
  unsigned long long *d_dag3_xpub = nullptr; unsigned dag3_epoch = 0;   // solved unknowns published with the flag in the data (kernels_solve_dag3.cuh)
  unsigned long long *d_dag_trace = nullptr;                     // PB200_DAG_TRACE=<file>: per-ticket time stamps of the last solve
  DagTick *d_dag_ticks = nullptr; int *d_dag_tgt = nullptr;
  unsigned int *d_dag_need = nullptr, *d_dag_state = nullptr;   // state: arrived[nsp] ready[nsp] done[nsp] cnt[nsp] ticket[2] err[1]
  unsigned int *h_dag_err = nullptr;                             // pinned
  void *d_y = nullptr; size_t y_bytes = 0;
  bool inv_ready = false;
  bool solve_transposed = false;           // IPARM_TRANSPOSE_SOLVE (LU only)
  bool herm = false;                       // internal CSC of type 'H': ucoeftab takes conj(transposed values) (csc_intern_solve.c:112-115)
  // round-2 experiment (PB200_GRAPH=1, not measured yet): the launch sequence of one factorization captured once per
  // pivot threshold and replayed as a CUDA graph
  cudaGraphExec_t fact_graph = nullptr; double fact_graph_crit = 0.0; int64_t fact_graph_launches = 0;
  unsigned attr_mask = 0;                  // which cudaFuncSetAttribute groups this handle has applied on ITS device (the
                                           // attributes are per device: a process-wide flag would skip the second GPU)
  bool schur = false;                      // IPARM_SCHUR: the last cblk is never factored (it ends up holding the Schur complement) and
                                           // up_down ignores it and every blok facing it (sopalin_compute.c:767-772, updo.c:425-428)
  void *d_raff_partial = nullptr, *h_raff_partial = nullptr;   // dot-product partial sums (device / pinned host)
  std::vector<int64_t> h_rmbase;           // per cblk: first entry of its off-diagonal rows in d_rowglob
  int *d_rowglob = nullptr;                // global row of every off-diagonal panel row
  std::vector<int> inv_lvl_nbmax;          // widest sub-panel of each level (sizes the shared memory of k_tri_inverse)
  std::vector<int> inv_lvl_ptr;            // sub-panels of level l: [inv_lvl_ptr[l], inv_lvl_ptr[l+1])
  std::vector<int> inv_cls_ptr[3]; int *d_inv_order = nullptr;   // sub-panels sorted by size class (k_tri_inverse launches): all / owned / not owned
  cudaStream_t stream_i = nullptr;         // low priority: triangle inversions underneath the factorization
  cudaEvent_t ev_inv = nullptr;
  // ---- FP64 tensor-core path (double / complex double, direct factorizations)
  bool use_mma = false;
  DevMap M{};
  struct Step { int kind; int task0, ntasks; long long ntiles; int nbmax; int lvl; long long t2t0 = 0; int strm = 0, wait_ev = -1, rec_ev = -1; };  // kind: 0 diag 1 trsm 2 gemm 3 transpose 4 marker 5 fan-in
  std::vector<Step> steps;
  SubTask *d_sub = nullptr;
  GemmTask *d_gemm = nullptr;
  int *d_t2t = nullptr;
  TileDesc *d_desc = nullptr;
  bool prof_on = false;                  // pb200_set_profile: serialise the launches and time each kind
  double prof_ms[4] = {0, 0, 0, 0}; int64_t prof_n[4] = {0, 0, 0, 0};
  double gemm_flops = 0;                 // algorithmic flops of the fused GEMM+scatter launches (PaStiX's GEMM term)
  cudaStream_t stream_u = nullptr;       // second stream: the bulk of the fused GEMM+scatter updates
  std::vector<cudaEvent_t> sched_ev;     // [l] panel(l) done, [nlevels + l] bulk update of level l done
  std::vector<cudaEvent_t> tl_ev;        // PB200_TIMELINE: timed copies of the same points
  std::vector<int> h_gemm_modes;
  std::vector<void *> allocs;
  // ---- small-supernode path of the generic factorization (kernels_small.cuh)
  std::vector<GStep> gsteps;
  int *d_lvl_ptr = nullptr;
  int64_t *d_sm_pbase = nullptr, *d_sm_tabL = nullptr, *d_sm_tabU = nullptr;
  // ---- multi-GPU (one process per GPU; see kernels_dist.cuh)
  int rank = 0, nranks = 1;
  DistPlan plan;
  int *d_owner = nullptr;
  Peers peers{};
  bool attached = false, gathered = true;
  bool local_group = false;               // pb200_attach_local: the peers are handles of THIS process (no IPC mappings, no collective in destroy)
  unsigned int *d_flags = nullptr, *d_dist_err = nullptr;   // flags: [0,nlevels) level ready, [nlevels] factorization done, [nlevels+1] barrier
  unsigned int epoch = 0, bar_epoch = 0;
  struct DistLevel {
    int sig = 0; unsigned int wait_mask = 0, late_mask = 0; int task0 = 0, ntasks = 0; long long ntiles = 0;
    // fan-out of the shared top separators: psig = this GPU owns a shared cblk of the level whose targets live elsewhere
    // (publish "panel factored"); fp_* = the shared cblks of the level owned elsewhere whose targets live here (pull)
    int psig = 0; unsigned int fp_mask = 0; int fp_task0 = 0, fp_ntasks = 0; long long fp_tiles = 0;
  };
  std::vector<int> all_level, all_lvl_ptr, all_lvl_cblk, own_lvl_cblk;   // every cblk's level; level lists over all cblks / the owned ones
  bool fanout = false;                     // multi-GPU tensor path: updates of shared cblks computed by the owners of their targets
  std::vector<std::vector<int>> foreign;   // per level: shared cblks owned elsewhere that update cblks owned here
  char *d_fanout = nullptr;                // per cblk: 1 = fan-out source (device copy of plan.shared)
  FanTask *d_fpull = nullptr;
  cudaStream_t stream_g = nullptr;        // early part of the fan-in gathers (contributors that finished long before the level is due)
  std::vector<cudaEvent_t> gather_ev;     // per level
  std::vector<DistLevel> dist_lvl;
  FanTask *d_fan = nullptr, *d_pull = nullptr; int npull = 0; long long pull_tiles = 0;
  std::vector<void *> ipc_opened;
};

extern "C" const char *pb200_last_error(void) { return g_err.c_str(); }
extern "C" void pb200_set_error(const char *msg) { g_err = msg ? msg : ""; }   // csc_build.cu reports through the same channel
extern "C" const char *pb200_version(void) { return "pastix_b200 0.1 (sm_100a)"; }

template <class V>
static int upload(pb200_handle_t *h, const std::vector<V> &v, V **d) {
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(V);
  CK(cudaMalloc((void **)d, bytes));
  h->allocs.push_back(*d);
  h->device_bytes += bytes;
  if (!v.empty()) CK(cudaMemcpy(*d, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
  return 0;
}

// Launch of a panel-chain kernel, optionally with programmatic stream serialization (kernels_mma.cuh, pdl_wait): the
// launch may be scheduled while its predecessor in the stream still runs.  Only kernels that execute
// griddepcontrol.wait before touching panels may take the attribute.
template <class... KArgs, class... Args>
static cudaError_t launch_chain(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static size_t elem_size(int flt) {
  switch (flt) {
    case PB200_REALSINGLE: return 4;
    case PB200_REALDOUBLE: return 8;
    case PB200_COMPLEXSINGLE: return 8;
    case PB200_COMPLEXDOUBLE: return 16;
  }
  return 0;
}

// ------------------------------------------------------------------ persistent up_down sweeps (kernels_solve_dag.cuh)
// For every tile (sub-panel x 128 panel rows, ticket order = ascending (level, round)) the distinct sub-panels
// that own its rows: the counters they bump in the down step (the reference's UPDOWN_CTRBCNT bookkeeping,
// updo.c:631-793) and the flags they wait on in the up step (flagtab, updo_sendrecv.c:496-639).
static int build_solve_dag(pb200_handle_t *h, const std::vector<SlvTask> &tasks) {
  h->dag_ok = false;
  if (getenv("PB200_SOLVE_LEVELS") != nullptr) return PB200_SUCCESS;   // A/B switch: keep the launch-per-level sweeps
  // all-small schedules (incomplete factorizations) keep their fused warp-per-cblk path — except in Schur mode, which
  // only the persistent sweeps implement (the reference ignores the Schur cblk on every path, updo.c:425-428)
  if (h->slv_all_small && !h->schur) return PB200_SUCCESS;
  const int NB = (h->flt == PB200_COMPLEXDOUBLE) ? SlvCfg<cdouble>::NB : SlvCfg<double>::NB;
  const int64_t C = h->cblknbr;
  const int nsp = (int)tasks.size();
  const int DROWS = 32;      // finest tiling any generation uses (ticket count bound below)
  // (cblk, round) -> sub-panel id, and the sub-panel width of each cblk
  std::vector<int> sp_ptr(C + 1, 0), sw(C, 1);
  for (int64_t c = 0; c < C; ++c) {
    const int w = h->h_width[c], nsub = std::max(1, (w + NB - 1) / NB);
    sw[c] = (w + nsub - 1) / nsub; sp_ptr[c + 1] = sp_ptr[c] + nsub;
  }
  if (sp_ptr[C] != nsp) return PB200_SUCCESS;
  std::vector<int> spid(nsp, -1);
  int nbs = 1;
  long long nticks = nsp;
  // Schur mode: the last cblk gets no ticket at all and every panel ends where its rows start facing it (those bloks
  // are the last ones of a panel: the Schur cblk holds the highest rows) — the reference's "ignore schur" tests in
  // the down, diagonal and up loops (updo.c:425-428, 639-646, 711-714, 951-954, 1154-1180; updo_sendrecv.c:518-523)
  const int sc = h->schur ? (int)C - 1 : -1;
  std::vector<int> mend(C);
  for (int64_t c = 0; c < C; ++c) {
    mend[c] = h->h_stride[c];
    if (sc >= 0)
      for (int b = h->h_fblok[c] + 1; b < h->h_fblok[c + 1]; ++b)
        if (h->h_fcblk[b] == sc) { mend[c] = h->h_coefind[b]; break; }
  }
  for (int i = 0; i < nsp; ++i) {
    const SlvTask &tk = tasks[i];
    spid[sp_ptr[tk.cblk] + tk.c0 / sw[tk.cblk]] = i;
    if (tk.cblk == sc) continue;
    nbs = std::max(nbs, tk.c1 - tk.c0);
    nticks += (std::max(mend[tk.cblk], tk.c1) - tk.c1 + DROWS - 1) / DROWS;
  }
  if (nticks >= (1LL << 24)) return PB200_SUCCESS;   // millions of tiny tickets: the level sweeps batch them better
  // One ticket list.  rows: panel rows per sub-tile; blocked: per (level, round) step all D tickets, then all T tickets,
  // T tickets grown to up to 8 sub-tiles where a step has thousands of them (second and third generation); split:
  // D and T tickets in two lists (third generation).  need[] = T tickets contributing to each sub-panel.
  struct DagList { std::vector<DagTick> ticks, ticksT; std::vector<int> tgt; std::vector<unsigned int> need; };
  const long long csplit_below = getenv("PB200_DAG3_CSPLIT") ? atoll(getenv("PB200_DAG3_CSPLIT")) : 600;   // steps with at most this many sub-tiles: T tickets also split by columns
  auto build_list = [&](int rows, bool blocked, bool split, long long grow_at, DagList &L) {
    L.ticks.reserve((size_t)nticks); L.tgt.reserve((size_t)nticks * 3);
    L.need.assign(nsp, 0);
    std::vector<DagTick> &tt = split ? L.ticksT : L.ticks;
    std::vector<long long> stamp(nsp, -1);
    std::vector<int> ntk(nsp, 0);                    // T tickets of each sub-panel
    long long serial = 0;
    auto emit_D = [&](int i) {
      const SlvTask &tk = tasks[i];
      const int ld = tk.ld, nb = tk.c1 - tk.c0;
      DagTick d{};
      d.src = tk.invoff; d.aux = tk.poff + (int64_t)tk.c0 * (ld + 1); d.ld = ld; d.nb = nb; d.mrows = -1; d.sp = tk.sp;
      d.xcol = tk.fcol + tk.c0; d.nsib = 0;
      L.ticks.push_back(d);
    };
    // T tickets of sub-panel i, `trows` panel rows each
    auto emit_T = [&](int i, int trows, int ccols) {
      const SlvTask &tk = tasks[i];
      const int c = tk.cblk, w = tk.w, ld = tk.ld, nb = tk.c1 - tk.c0;
      const int me = std::max(mend[c], tk.c1);   // rows [c1, me) of the panel take part
      const int nt = (me - tk.c1 + trows - 1) / trows;
      ntk[i] = nt;
      int b = h->h_fblok[c] + 1;                    // first off-diagonal blok
      const int be = h->h_fblok[c + 1];
      for (int t = 0; t < nt; ++t) {
        const long long g = serial++;
        const int m0 = tk.c1 + t * trows, m1 = std::min(me, m0 + trows);
        DagTick k{};
        k.src = tk.poff + (int64_t)tk.c0 * ld + m0; k.aux = tk.rgbase + m0; k.ld = ld; k.nb = nb; k.mrows = m1 - m0; k.sp = tk.sp;
        k.xcol = tk.fcol + tk.c0; k.grow0 = tk.fcol + m0; k.wrem = w - m0; k.tptr = (int)L.tgt.size();
        auto add = [&](int s) { if (stamp[s] != g) { stamp[s] = g; L.tgt.push_back(s); ++L.need[s]; } };
        // rows still inside the diagonal block: later sub-panels of the same cblk
        for (int m = m0; m < std::min(m1, w); m = (m / sw[c] + 1) * sw[c]) add(spid[sp_ptr[c] + m / sw[c]]);
        // off-diagonal rows: the bloks crossing [m0, m1)
        while (b < be && h->h_coefind[b] + h->h_nrow[b] <= m0) ++b;
        for (int bb = b; bb < be && h->h_coefind[bb] < m1; ++bb) {
          const int fc = h->h_fcblk[bb];
          const int lo = std::max(m0, h->h_coefind[bb]), hi = std::min(m1, h->h_coefind[bb] + h->h_nrow[bb]);   // panel rows [lo, hi)
          const int o0 = h->h_frow[bb] + (lo - h->h_coefind[bb]) - h->h_fcol[fc], o1 = h->h_frow[bb] + (hi - 1 - h->h_coefind[bb]) - h->h_fcol[fc];
          for (int r = o0 / sw[fc]; r <= o1 / sw[fc]; ++r) add(spid[sp_ptr[fc] + r]);
        }
        k.ntgt = (int)L.tgt.size() - k.tptr;
        if (ccols <= 0 || ccols >= nb) { tt.push_back(k); continue; }
        // column split (links of the dependency chain): the same rows, ccols columns of the sub-panel per ticket — each is
        // a narrower sub-panel to the kernels (x_J slice, panel columns), all of them count for the same sub-panels
        const int ncs = (nb + ccols - 1) / ccols;
        for (int cs = 0; cs < ncs; ++cs) {
          DagTick kc = k;
          const int j0 = cs * ccols, j1 = std::min(nb, j0 + ccols);
          kc.src = k.src + (int64_t)j0 * ld; kc.xcol = k.xcol + j0; kc.nb = j1 - j0;
          tt.push_back(kc);
        }
        for (int q = 0; q < k.ntgt; ++q) L.need[L.tgt[k.tptr + q]] += (unsigned)(ncs - 1);
        ntk[i] += ncs - 1;
      }
    };
    if (!blocked) {
      // first generation: D(J) followed by its own T tickets
      for (int i = 0; i < nsp; ++i) {
        if (tasks[i].cblk == sc) continue;
        emit_D(i); emit_T(i, rows, 0);
      }
    } else {
      // per (level, round) step all the D tickets, then all the T tickets — a ticket and the ones it waits for are a whole
      // block of independent tickets apart wherever a level is wide.  T tickets grow to up to 8 sub-tiles in steps with
      // thousands of sub-tiles (the per-ticket cost — atomics, polls, fence — is paid once per 256 rows) and stay one
      // sub-tile where a step is a link of the dependency chain.
      for (const auto &st : h->slv_steps) {
        long long subt = 0;
        for (int i = st.task0; i < st.task0 + st.ntasks; ++i) {
          if (tasks[i].cblk == sc) continue;
          const int c = tasks[i].cblk;
          subt += (std::max(mend[c], tasks[i].c1) - tasks[i].c1 + rows - 1) / rows;
        }
        int nsub = 1;
        while (nsub < 8 && subt / (2 * nsub) >= grow_at) nsub *= 2;
        for (int i = st.task0; i < st.task0 + st.ntasks; ++i) if (tasks[i].cblk != sc) emit_D(i);
        const int ccols = (split && subt <= csplit_below) ? 32 : 0;
        for (int i = st.task0; i < st.task0 + st.ntasks; ++i) if (tasks[i].cblk != sc) emit_T(i, rows * nsub, ccols);
      }
    }
    for (auto &k : L.ticks) if (k.mrows < 0) { k.nsib = ntk[k.sp]; k.pad0 = (int)L.need[k.sp]; }
  };
  // which kernels: PB200_DAG_V1 = first generation for everything, PB200_DAG_V2 = second generation for everything;
  // default: third generation (kernels_solve_dag3.cuh) for one right-hand side, first generation for several
  const bool v1only = getenv("PB200_DAG_V1") != nullptr, v2only = !v1only && getenv("PB200_DAG_V2") != nullptr;
  h->dag_rows = v2only ? PB200_DAG2_ROWS : PB200_DAG_ROWS;
  h->dag3_ok = false;
  DagList LA;
  build_list(h->dag_rows, v2only, false, 1024, LA);
  if (LA.tgt.size() >= (size_t)INT32_MAX) return PB200_SUCCESS;
  for (int s : LA.tgt) if (s < 0) return fail(PB200_ERR_STRUCT, "up_down dependency table: row without an owning sub-panel");
  { int rc = upload(h, LA.ticks, &h->d_dag_ticks); if (rc) return rc; }
  { int rc = upload(h, LA.tgt, &h->d_dag_tgt); if (rc) return rc; }
  { int rc = upload(h, LA.need, &h->d_dag_need); if (rc) return rc; }
  if (!v1only && !v2only) {
    DagList L3;
    build_list(32, true, true, getenv("PB200_DAG3_GROW") ? atoll(getenv("PB200_DAG3_GROW")) : 1024, L3);
    if (L3.tgt.size() < (size_t)INT32_MAX && !L3.ticks.empty()) {
      h->dag3_GD = (int)L3.ticks.size(); h->dag3_GT = (int)L3.ticksT.size();
      { int rc = upload(h, L3.ticks, &h->d_dag3_ticksD); if (rc) return rc; }
      { int rc = upload(h, L3.ticksT, &h->d_dag3_ticksT); if (rc) return rc; }
      { int rc = upload(h, L3.tgt, &h->d_dag3_tgt); if (rc) return rc; }
      {
        const size_t pb = (size_t)h->n * (h->esize / 4) * sizeof(unsigned long long);
        CK(cudaMalloc((void **)&h->d_dag3_xpub, pb));
        CK(cudaMemset(h->d_dag3_xpub, 0, pb));     // epoch 0 = never published
        h->allocs.push_back(h->d_dag3_xpub); h->device_bytes += pb;
      }
      h->dag3_ok = true;
    }
  }
  const size_t sb = ((size_t)4 * nsp + 8) * sizeof(unsigned int);
  CK(cudaMalloc((void **)&h->d_dag_state, sb));
  h->allocs.push_back(h->d_dag_state); h->device_bytes += sb;
  CK(cudaHostAlloc((void **)&h->h_dag_err, sizeof(unsigned int), cudaHostAllocDefault));
  *h->h_dag_err = 0;
  h->dag_tiles = (int)LA.ticks.size(); h->dag_nbs = nbs; h->dag_ok = true;
  return PB200_SUCCESS;
}

// ------------------------------------------------------------------ schedule of the up_down sweeps
static int build_solve_schedule(pb200_handle_t *h, const std::vector<int> &lvl_cblk) {
  const int NB = (h->flt == PB200_COMPLEXDOUBLE) ? SlvCfg<cdouble>::NB : SlvCfg<double>::NB;
  std::vector<SlvTask> tasks; std::vector<int> t2t; std::vector<int64_t> invoff;
  int64_t inv_elems = 0; int sp = 0;
  std::vector<int64_t> rmbase((size_t)h->cblknbr + 1, 0);
  for (int64_t c = 0; c < h->cblknbr; ++c) rmbase[c + 1] = rmbase[c] + (h->h_stride[c] - h->h_width[c]);
  h->inv_lvl_ptr.assign(h->nlevels + 1, 0);
  h->inv_lvl_nbmax.assign(h->nlevels, 1);
  h->slv_lvl_ptr = h->lvl_ptr;
  h->slv_lvl_step.assign(h->nlevels + 1, 0);
  for (int l = 0; l < h->nlevels; ++l) {
    const int q0 = h->lvl_ptr[l], q1 = h->lvl_ptr[l + 1];
    int rounds = 0;
    h->inv_lvl_ptr[l] = sp;
    h->slv_lvl_step[l] = (int)h->slv_steps.size();
    for (int q = q0; q < q1; ++q) rounds = std::max(rounds, (h->h_width[lvl_cblk[q]] + NB - 1) / NB);
    for (int r = 0; r < rounds; ++r) {
      int t0 = (int)tasks.size(); long long tiles = 0; long long tt0 = (long long)t2t.size();
      for (int q = q0; q < q1; ++q) {
        int c = lvl_cblk[q], w = h->h_width[c], ld = h->h_stride[c];
        int nsub = (w + NB - 1) / NB;
        if (nsub <= r) continue;
        int sw = (w + nsub - 1) / nsub;
        int c0 = r * sw, c1 = std::min(w, c0 + sw);
        int nt = std::max(1, (ld - c1 + PB200_SLV_ROWS - 1) / PB200_SLV_ROWS);
        tasks.push_back({c, (int)tiles, c0, c1, sp, nt, ld, h->h_fcol[c], w, 0, h->h_poff[c], rmbase[c] - w, inv_elems});
        h->inv_lvl_nbmax[l] = std::max(h->inv_lvl_nbmax[l], c1 - c0);
        for (int i = 0; i < nt; ++i) t2t.push_back((int)tasks.size() - 1 - t0);
        tiles += nt;
        invoff.push_back(inv_elems); inv_elems += (int64_t)(c1 - c0) * (c1 - c0); ++sp;
      }
      if (tiles >= (1LL << 31)) return fail(PB200_ERR_STRUCT, "too many solve tiles in one level");
      h->slv_steps.push_back({t0, (int)tasks.size() - t0, tiles, tt0});
    }
  }
  h->nsubpanels = sp; h->inv_elems = inv_elems; h->inv_lvl_ptr[h->nlevels] = sp;
  h->slv_lvl_step[h->nlevels] = (int)h->slv_steps.size();
  h->h_rmbase = rmbase;
  {
    // small-cblk levels of the up_down (kernels_small.cuh): same classification as the factorization, over all cblks
    const int nl = h->nlevels;
    const int chain_warps = (h->esize >= 16) ? SmChain<cdouble>::WARPS : SmChain<double>::WARPS;
    const bool off = getenv("PB200_NO_SMALL_PATH") != nullptr;
    std::vector<char> small(nl, 0);
    bool all = !off;
    for (int l = 0; l < nl; ++l) {
      bool ok = !off;
      for (int q = h->lvl_ptr[l]; q < h->lvl_ptr[l + 1] && ok; ++q) {
        const int c = lvl_cblk[q];
        ok = h->h_width[c] <= PB200_SM_WMAX && h->h_stride[c] - h->h_width[c] <= PB200_SM_RMAX;
      }
      small[l] = ok; all = all && ok;
    }
    h->slv_all_small = all;
    for (int l = 0; l < nl;) {
      const int nc = h->lvl_ptr[l + 1] - h->lvl_ptr[l];
      if (!small[l]) { h->sgsteps.push_back({0, l, l + 1}); ++l; continue; }
      if (nc <= 2 * chain_warps) {
        int e = l + 1;
        while (e < nl && small[e] && h->lvl_ptr[e + 1] - h->lvl_ptr[e] <= 2 * chain_warps) ++e;
        h->sgsteps.push_back({2, l, e}); l = e;
      } else { h->sgsteps.push_back({1, l, l + 1}); ++l; }
    }
    { int rc = upload(h, lvl_cblk, &h->d_slv_cblk); if (rc) return rc; }
    { int rc = upload(h, h->lvl_ptr, &h->d_slv_lvl_ptr); if (rc) return rc; }
  }
  {
    // size classes of the triangle inversions (16, 32, 48, 64, 96, 128 columns): widest first is not needed, the
    // launches are independent
    const int ncls = 6; const int bound[ncls] = {16, 32, 48, 64, 96, 128};
    std::vector<std::vector<int>> byc(ncls);
    for (int i = 0; i < (int)tasks.size(); ++i) {
      if (h->schur && tasks[i].cblk == (int)h->cblknbr - 1) continue;   // never factored: nothing to invert
      const int nb = tasks[i].c1 - tasks[i].c0;
      int k = 0; while (k + 1 < ncls && nb > bound[k]) ++k;
      byc[k].push_back(i);
    }
    // three lists per size class, one after the other: every sub-panel, the ones of the cblks this GPU owns (inverted
    // inside the timed factorization of a multi-GPU run), the others (inverted after their panels have been pulled)
    std::vector<int> order;
    for (int which = 0; which < 3; ++which) {
      h->inv_cls_ptr[which].assign(1, (int)order.size());
      for (int k = 0; k < ncls; ++k) {
        for (int i : byc[k]) {
          const bool mine = h->plan.owner[tasks[i].cblk] == h->rank;
          if (which == 0 || (which == 1) == mine) order.push_back(i);
        }
        h->inv_cls_ptr[which].push_back((int)order.size());
      }
    }
    int rc = upload(h, order, &h->d_inv_order); if (rc) return rc;
  }
  { int rc = upload(h, tasks, &h->d_slvtask); if (rc) return rc; }
  { int rc = upload(h, t2t, &h->d_slv_t2t); if (rc) return rc; }
  { int rc = build_solve_dag(h, tasks); if (rc) return rc; }
  { int rc = upload(h, invoff, &h->d_invoff); if (rc) return rc; }
  size_t ib = (size_t)std::max<int64_t>(inv_elems, 1) * h->esize;
  if (cudaMalloc(&h->d_inv, ib) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(inverse triangles) failed");
  h->allocs.push_back(h->d_inv); h->device_bytes += ib;
  if (h->facto == PB200_FACT_LU) {
    if (cudaMalloc(&h->d_inv_up, ib) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(inverse triangles) failed");
    h->allocs.push_back(h->d_inv_up); h->device_bytes += ib;
  }
  CK(cudaMalloc((void **)&h->d_slv_cnt, (size_t)std::max(sp, 1) * sizeof(unsigned int)));
  h->allocs.push_back(h->d_slv_cnt);
  CK(cudaMemset(h->d_slv_cnt, 0, (size_t)std::max(sp, 1) * sizeof(unsigned int)));
  return PB200_SUCCESS;
}

// ------------------------------------------------------------------ schedule of the tensor-core path
// Levels of the elimination tree; inside a level, cblks wider than NBMAX are walked sub-panel by
// sub-panel ("rounds"): diag -> trsm -> internal update; then one fused GEMM+scatter launch for the level.
static int build_mma_schedule(pb200_handle_t *h, const std::vector<int> &level, const std::vector<int> &lvl_cblk) {
  const bool cx = (h->flt == PB200_COMPLEXDOUBLE || h->flt == PB200_COMPLEXSINGLE);
  const bool lu = (h->facto == PB200_FACT_LU);
  const int NBMAX = h->flt == PB200_REALDOUBLE ? SubCfg<double>::NBMAX : h->flt == PB200_COMPLEXDOUBLE ? SubCfg<cdouble>::NBMAX
                  : h->flt == PB200_REALSINGLE ? SubCfg<float>::NBMAX : SubCfg<cfloat>::NBMAX;
  static_assert(UpdCfg<double>::TM == 64 && UpdCfg<cdouble>::TM == 64 && UpdCfg<float>::TM == 64 && UpdCfg<cfloat>::TM == 64 &&
                UpdCfg<double>::TN == 64 && UpdCfg<cdouble>::TN == 64 && UpdCfg<float>::TN == 64 && UpdCfg<cfloat>::TN == 64,
                "one tile shape for the four precisions (the schedule and the tile descriptors are built once)");
  const int TM = 64, TN = 64;
  const int64_t C = h->cblknbr;
  // pair tables
  std::vector<int64_t> pairbase(C + 1, 0);
  for (int64_t c = 0; c < C; ++c) {
    int64_t nb = h->h_fblok[c + 1] - h->h_fblok[c] - 1;
    pairbase[c + 1] = pairbase[c] + nb * (nb + 1) / 2;
  }
  int64_t *d_pb; int *d_po = nullptr, *d_napa = nullptr;
  { int rc = upload(h, pairbase, &d_pb); if (rc) return rc; }
  size_t pbytes = (size_t)std::max<int64_t>(pairbase[C], 1) * sizeof(int);
  CK(cudaMalloc((void **)&d_po, pbytes)); h->allocs.push_back(d_po); h->device_bytes += pbytes;
  CK(cudaMalloc((void **)&d_napa, sizeof(int))); h->allocs.push_back(d_napa);
  CK(cudaMemset(d_napa, 0, sizeof(int)));
  k_build_pairs<<<(unsigned)C, 128>>>(h->S, d_pb, d_po, d_napa);
  CK(cudaGetLastError());
  int napa = 0;
  CK(cudaMemcpy(&napa, d_napa, sizeof(int), cudaMemcpyDeviceToHost));
  if (napa) { h->use_mma = false; return PB200_SUCCESS; }   // incomplete factorization: generic path
  // per blok: target column origin inside the facing cblk
  std::vector<BlokTgt> btgt(h->bloknbr);
  for (int64_t c = 0; c < C; ++c)
    for (int bb = h->h_fblok[c]; bb < h->h_fblok[c + 1]; ++bb) {
      int fc = h->h_fcblk[bb];
      if (bb == h->h_fblok[c]) fc = (int)c;
      BlokTgt t;
      t.cj0 = h->h_frow[bb] - h->h_fcol[fc]; t.tld = h->h_stride[fc]; t.tw = h->h_width[fc]; t.fc = fc;
      t.tgt = h->h_poff[fc] + (int64_t)t.cj0 * t.tld;
      btgt[bb] = t;
    }
  BlokTgt *d_bt;
  { int rc = upload(h, btgt, &d_bt); if (rc) return rc; }
  h->M.pairbase = d_pb; h->M.pairoff = d_po; h->M.btgt = d_bt;
  {
    // static row/column scatter maps, one entry per off-diagonal panel row
    std::vector<int64_t> rmbase(C + 1, 0);
    for (int64_t c = 0; c < C; ++c) rmbase[c + 1] = rmbase[c] + (h->h_stride[c] - h->h_width[c]);
    int64_t *d_rb; RowMap *d_rm = nullptr; ColMap *d_cm = nullptr;
    { int rc = upload(h, rmbase, &d_rb); if (rc) return rc; }
    const size_t nrow = (size_t)std::max<int64_t>(rmbase[C], 1);
    if (cudaMalloc((void **)&d_rm, nrow * sizeof(RowMap)) != cudaSuccess || cudaMalloc((void **)&d_cm, nrow * sizeof(ColMap)) != cudaSuccess)
      return fail(PB200_ERR_NOMEM, "cudaMalloc(scatter maps) failed");
    h->allocs.push_back(d_rm); h->allocs.push_back(d_cm); h->device_bytes += nrow * (sizeof(RowMap) + sizeof(ColMap));
    k_build_maps<<<(unsigned)C, 128>>>(h->S, d_bt, d_rb, d_rm, d_cm, h->d_owner, h->fanout ? h->d_fanout : nullptr, h->rank);
    CK(cudaGetLastError());
    h->M.rmbase = d_rb; h->M.rm = d_rm; h->M.cm = d_cm;
  }

  std::vector<SubTask> sub;
  std::vector<GemmTask> gemm;
  // complex LLt: the reference factors the 64-column blocks of PASTIX_potrf_block with a SYMMETRIC unblocked kernel
  // (csqrt + geru, compute_diag.c:140) but updates the trailing block with zherk (sopalin_compute.h:178-179), so its
  // result depends on where the block boundaries are: keep them at multiples of MAXSIZEOFBLOCKS = 64 (compute_diag.c:46)
  const bool ref_blocking = cx && h->facto == PB200_FACT_LLT;
  auto subpanel = [&](int w, int r, int &c0, int &c1) {
    int nsub = (w + NBMAX - 1) / NBMAX;
    int sw = ref_blocking ? NBMAX : ((((w + nsub - 1) / nsub) + 7) & ~7);
    c0 = r * sw; c1 = std::min(w, c0 + sw);
  };
  for (int l = 0; l < h->nlevels; ++l) {
    const int q0 = h->lvl_ptr[l], q1 = h->lvl_ptr[l + 1];
    const size_t level_first_step = h->steps.size();
    int rounds = 0;
    for (int q = q0; q < q1; ++q) rounds = std::max(rounds, (h->h_width[lvl_cblk[q]] + NBMAX - 1) / NBMAX);
    if (h->nranks > 1 && (h->dist_lvl[l].sig || h->dist_lvl[l].ntasks))
      h->steps.push_back({5, 0, 1, 0, 0, l});   // fan-in: publish our contributions, pull the peers' (kernels_dist.cuh)
    if (lu) h->steps.push_back({3, q0, q1 - q0, 0, 0, l});
    if (rounds == 0) h->steps.push_back({4, 0, 0, 0, 0, l});   // multi-GPU: no cblk of this level lives here; keeps the level's events
    for (int r = 0; r < rounds; ++r) {
      // diag and trsm, each as (at most) two launches: sub-panels of at most 64 columns and wider ones — the narrow
      // kernels hold a quarter of the registers / shared memory and keep several CTAs per SM on the fat bottom levels
      int t0 = 0; long long tiles = 0;
      for (int cls = 0; cls < 2; ++cls) {
        int nbmax = 0; t0 = (int)sub.size();
        for (int q = q0; q < q1; ++q) {
          int c = lvl_cblk[q], w = h->h_width[c];
          if ((w + NBMAX - 1) / NBMAX <= r) continue;
          int c0, c1; subpanel(w, r, c0, c1);
          if ((c1 - c0 > 64) != (cls == 1)) continue;
          sub.push_back({c, 0, c0, c1}); nbmax = std::max(nbmax, c1 - c0);
        }
        if ((int)sub.size() > t0)
          h->steps.push_back({0, t0, (int)sub.size() - t0, (long long)sub.size() - t0, nbmax, l});
      }
      for (int cls = 0; cls < 2; ++cls) {
        int nbmax = 0; t0 = (int)sub.size(); tiles = 0;
        for (int q = q0; q < q1; ++q) {
          int c = lvl_cblk[q], w = h->h_width[c], ld = h->h_stride[c];
          if ((w + NBMAX - 1) / NBMAX <= r) continue;
          int c0, c1; subpanel(w, r, c0, c1);
          if (ld - c1 <= 0 || (c1 - c0 > 64) != (cls == 1)) continue;
          sub.push_back({c, (int)tiles, c0, c1}); tiles += (ld - c1 + PB200_TRSM_TM - 1) / PB200_TRSM_TM;
          nbmax = std::max(nbmax, c1 - c0);
        }
        if (tiles > 0) h->steps.push_back({1, t0, (int)sub.size() - t0, tiles, nbmax, l});
      }
      // internal update
      t0 = (int)gemm.size(); tiles = 0;
      for (int q = q0; q < q1; ++q) {
        int c = lvl_cblk[q], w = h->h_width[c], ld = h->h_stride[c];
        if ((w + NBMAX - 1) / NBMAX <= r) continue;
        int c0, c1; subpanel(w, r, c0, c1);
        if (c1 >= w) continue;
        for (int a0 = c1; a0 < ld; a0 += TM) {
          int ncols = (lu ? w : std::min(w, a0 + TM)) - c1;
          int ntn = (ncols + TN - 1) / TN;
          gemm.push_back({c, (int)tiles, ntn, a0, ld, c1, c1 + ncols, c0, c1, 1, 0, 0});
          tiles += ntn;
        }
      }
      if (tiles > 0) h->steps.push_back({2, t0, (int)gemm.size() - t0, tiles, 0, l});
    }
    if (lu && rounds > 1) h->steps.push_back({6, q0, q1 - q0, 0, 0, l});   // complete the diagonal bloks of the multi-round cblks
    // fan-out (multi-GPU): the shared panels factored here are published; the ones factored elsewhere whose targets
    // live here are pulled into the local slab (same offsets) and take part in this GPU's update launches below
    std::vector<int> srcs(lvl_cblk.begin() + q0, lvl_cblk.begin() + q1);
    if (h->nranks > 1 && h->fanout) {
      if (h->dist_lvl[l].psig) h->steps.push_back({7, 0, 1, 0, 0, l});
      if (h->dist_lvl[l].fp_ntasks) {
        h->steps.push_back({8, 0, 1, 0, 0, l});
        srcs.insert(srcs.end(), h->foreign[l].begin(), h->foreign[l].end());
      }
    }
    // external update: fused GEMM + scatter, in two launches.  U1 = the column tiles that hit cblks of
    // the NEXT level (they gate that level's panel work) stays on the panel stream; U2 = everything else
    // runs on the second stream underneath the next level's diag/trsm chain.
    size_t first_p = level_first_step;
    for (int pass = 0; pass < 2; ++pass) {
      int t0 = (int)gemm.size(); long long tiles = 0;
      for (int c : srcs) {
        const int w = h->h_width[c], ld = h->h_stride[c];
        const bool filt = h->nranks > 1 && h->fanout && h->plan.shared[c];   // only the targets this GPU owns
        if (ld <= w) continue;
        int b = h->h_fblok[c] + 1;
        for (int a0 = w; a0 < ld; a0 += TM) {
          int imax = std::min(ld, a0 + TM) - 1;
          while (h->h_coefind[b] + h->h_nrow[b] <= imax) ++b;   // blok holding the last row of this tile
          int ncols = h->h_coefind[b] + h->h_nrow[b] - w;
          int ntn = (ncols + TN - 1) / TN;
          if (filt) {
            // fan-out source: the column tiles are cut along the ownership of the facing cblks (maximal runs of
            // consecutive bloks facing cblks owned here), so that no tile is computed on two GPUs
            const int bend = h->h_fblok[c + 1], cend = w + ncols;
            for (int bb = h->h_fblok[c] + 1; bb < bend && h->h_coefind[bb] < cend;) {
              if (h->plan.owner[h->h_fcblk[bb]] != h->rank) { ++bb; continue; }
              int be = bb;
              while (be + 1 < bend && h->h_coefind[be + 1] < cend && h->plan.owner[h->h_fcblk[be + 1]] == h->rank) ++be;
              const int s0 = h->h_coefind[bb], s1 = std::min(cend, h->h_coefind[be] + h->h_nrow[be]);
              const int nts = (s1 - s0 + TN - 1) / TN;
              int run0 = -1, cbk = bb;
              for (int tn = 0; tn <= nts; ++tn) {
                bool want = false;
                if (tn < nts) {
                  const int n_lo = s0 + tn * TN, n_hi = std::min(s1, n_lo + TN);
                  while (h->h_coefind[cbk] + h->h_nrow[cbk] <= n_lo) ++cbk;
                  bool next = false;
                  for (int q = cbk; q <= be && h->h_coefind[q] < n_hi; ++q)
                    if (level[h->h_fcblk[q]] == l + 1) next = true;
                  want = (pass == 0) ? next : !next;
                }
                if (want && run0 < 0) run0 = tn;
                if (!want && run0 >= 0) {
                  const int rn = tn - run0;
                  const int br0 = s0 + run0 * TN, br1 = std::min(s1, br0 + rn * TN);
                  gemm.push_back({c, (int)tiles, rn, a0, ld, br0, br1, 0, w, 0, b, 0});
                  tiles += rn; run0 = -1;
                }
              }
              bb = be + 1;
            }
            continue;
          }
          // classify the column tiles, emit maximal runs of the wanted class
          int run0 = -1, cbk = h->h_fblok[c] + 1;
          for (int tn = 0; tn <= ntn; ++tn) {
            bool want = false;
            if (tn < ntn) {
              int n_lo = w + tn * TN, n_hi = std::min(w + ncols, n_lo + TN);   // panel rows [n_lo, n_hi)
              while (h->h_coefind[cbk] + h->h_nrow[cbk] <= n_lo) ++cbk;
              bool next = false, mine = !filt;
              for (int bb = cbk; bb < h->h_fblok[c + 1] && h->h_coefind[bb] < n_hi; ++bb) {
                if (level[h->h_fcblk[bb]] == l + 1) next = true;
                if (h->plan.owner[h->h_fcblk[bb]] == h->rank) mine = true;
              }
              want = mine && ((pass == 0) ? next : !next);
            }
            if (want && run0 < 0) run0 = tn;
            if (!want && run0 >= 0) {
              int rn = tn - run0;
              int br0 = w + run0 * TN, br1 = std::min(w + ncols, br0 + rn * TN);
              gemm.push_back({c, (int)tiles, rn, a0, ld, br0, br1, 0, w, 0, b, 0});
              tiles += rn; run0 = -1;
            }
          }
        }
      }
      if (tiles * 2 >= (1LL << 31)) return fail(PB200_ERR_STRUCT, "too many update tiles in one level");
      pb200_handle_t::Step st{tiles > 0 ? 2 : 4, t0, (int)gemm.size() - t0, tiles, 0, l};
      if (pass == 0) {
        // the panel steps of this level precede: first one waits for the bulk updates of level l-2
        if (l >= 2) h->steps[first_p].wait_ev = h->nlevels + (l - 2);
        h->steps.back().rec_ev = l;          // panel(l) done
        st.strm = 0;
      } else {
        st.strm = 1; st.wait_ev = l; st.rec_ev = h->nlevels + l;
      }
      h->steps.push_back(st);
    }
  }
  { int rc = upload(h, sub, &h->d_sub); if (rc) return rc; }
  { int rc = upload(h, gemm, &h->d_gemm); if (rc) return rc; }
  {
    // tile -> task map of every GEMM launch (task index relative to the launch's first task)
    std::vector<int> t2t;
    for (auto &st : h->steps) {
      if (st.kind != 2) continue;
      st.t2t0 = (long long)t2t.size();
      for (int t = 0; t < st.ntasks; ++t)
        for (int q = 0; q < gemm[st.task0 + t].ntn; ++q) t2t.push_back(t);
    }
    int rc = upload(h, t2t, &h->d_t2t); if (rc) return rc;
    // one TileDesc per tile of every launch, filled on the device
    const size_t nt = std::max<size_t>(t2t.size(), 1);
    if (cudaMalloc((void **)&h->d_desc, nt * sizeof(TileDesc)) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(tile descriptors) failed");
    h->allocs.push_back(h->d_desc); h->device_bytes += nt * sizeof(TileDesc);
    for (const auto &st : h->steps) {
      if (st.kind != 2 || st.ntiles == 0) continue;
      const int n = (int)st.ntiles;
      k_build_tiledesc<64, 64><<<(n + 127) / 128, 128>>>(h->S, h->M, h->d_gemm + st.task0, h->d_t2t + st.t2t0, n, h->d_desc + st.t2t0);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  CK(cudaStreamCreateWithFlags(&h->stream_u, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->stream_i, cudaStreamNonBlocking, lo));
    CK(cudaEventCreateWithFlags(&h->ev_inv, cudaEventDisableTiming));
  }
  h->sched_ev.resize(2 * (size_t)h->nlevels);
  for (auto &e : h->sched_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  h->h_gemm_modes.resize(gemm.size());
  for (size_t i = 0; i < gemm.size(); ++i) h->h_gemm_modes[i] = gemm[i].mode;
  h->use_mma = true;
  return PB200_SUCCESS;
}

// ------------------------------------------------------------------ schedule of the generic factorization
// Levels whose (local) cblks are all small run the fused warp-per-cblk kernel; consecutive thin ones are chained
// inside one CTA (kernels_small.cuh).  Everything else keeps the three launches per level.
static int build_small_schedule(pb200_handle_t *h, const std::vector<int> &lvl_cblk) {
  const int nl = h->nlevels;
  const bool cx16 = (h->esize >= 16);
  const int chain_warps = cx16 ? SmChain<cdouble>::WARPS : SmChain<double>::WARPS;
  const bool off = getenv("PB200_NO_SMALL_PATH") != nullptr;
  std::vector<char> small(nl, 0);
  std::vector<int64_t> pbase(lvl_cblk.size() + 1, 0);
  for (int l = 0; l < nl; ++l) {
    const int q0 = h->lvl_ptr[l], q1 = h->lvl_ptr[l + 1];
    bool ok = !off && q1 > q0;
    for (int q = q0; q < q1 && ok; ++q) {
      const int c = lvl_cblk[q];
      ok = h->h_width[c] <= PB200_SM_WMAX && h->h_stride[c] - h->h_width[c] <= PB200_SM_RMAX;
    }
    small[l] = ok;
    for (int q = q0; q < q1; ++q) {
      const int c = lvl_cblk[q];
      const int64_t mr = h->h_stride[c] - h->h_width[c];
      pbase[q + 1] = pbase[q] + (ok ? mr * (mr + 1) / 2 : 0);
    }
  }
  for (int l = 0; l < nl;) {
    const int nc = h->lvl_ptr[l + 1] - h->lvl_ptr[l];
    if (!small[l]) { h->gsteps.push_back({0, l, l + 1}); ++l; continue; }
    if (h->nranks == 1 && nc <= 2 * chain_warps) {
      int e = l + 1;
      while (e < nl && small[e] && h->lvl_ptr[e + 1] - h->lvl_ptr[e] <= 2 * chain_warps) ++e;
      h->gsteps.push_back({2, l, e}); l = e;
    } else { h->gsteps.push_back({1, l, l + 1}); ++l; }
  }
  const int64_t npairs = pbase.back();
  if (npairs == 0) return PB200_SUCCESS;
  { int rc = upload(h, pbase, &h->d_sm_pbase); if (rc) return rc; }
  { int rc = upload(h, h->lvl_ptr, &h->d_lvl_ptr); if (rc) return rc; }
  const size_t tb = (size_t)npairs * sizeof(int64_t);
  if (cudaMalloc((void **)&h->d_sm_tabL, tb) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(small-cblk contribution table) failed");
  h->allocs.push_back(h->d_sm_tabL); h->device_bytes += tb;
  if (h->facto == PB200_FACT_LU) {
    if (cudaMalloc((void **)&h->d_sm_tabU, tb) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(small-cblk contribution table) failed");
    h->allocs.push_back(h->d_sm_tabU); h->device_bytes += tb;
  }
  for (int l = 0; l < nl; ++l) {
    if (!small[l]) continue;
    const int q0 = h->lvl_ptr[l], nc = h->lvl_ptr[l + 1] - q0;
    k_build_small_pairs<<<(nc * 32 + 255) / 256, 256>>>(h->S, h->d_lvl_cblk + q0, nc, h->d_sm_pbase + q0, h->d_sm_tabL, h->d_sm_tabU);
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return PB200_SUCCESS;
}

// ------------------------------------------------------------------ multi-GPU: per-level exchange lists
// Owned cblks of every level, fan-in tasks (owned cblks with remote contributors), what this GPU publishes, and — in
// fan-out mode — the shared cblks owned elsewhere whose factored panels are pulled to compute the updates of the
// cblks owned here.  Rebuilt without fan-out when the tensor path turns out not to apply (incomplete factorization).
static int build_dist_levels(pb200_handle_t *h, bool fanout) {
  const int nl = h->nlevels, rank = h->rank;
  const int64_t C = h->cblknbr;
  const std::vector<int> &level = h->all_level, &lptr = h->all_lvl_ptr, &lcblk = h->all_lvl_cblk;
  const std::vector<uint32_t> &contrib = fanout ? h->plan.contrib_priv : h->plan.contrib;
  std::vector<int> optr(nl + 1, 0), ocblk;
  std::vector<FanTask> fan, pull, fpull;
  h->dist_lvl.assign(nl, pb200_handle_t::DistLevel());
  h->foreign.assign(nl, std::vector<int>());
  h->pull_tiles = 0;
  for (int l = 0; l < nl; ++l) {
    auto &D = h->dist_lvl[l];
    D.task0 = (int)fan.size(); D.fp_task0 = (int)fpull.size();
    for (int q = lptr[l]; q < lptr[l + 1]; ++q) {
      const int c = lcblk[q];
      const int64_t len = h->h_poff[c + 1] - h->h_poff[c];
      const int nt = (int)((len + PB200_FAN_ELEMS - 1) / PB200_FAN_ELEMS);
      bool tgt_here = false, tgt_else = false;   // (shared cblks) targets owned by this GPU / by others
      if (fanout && h->plan.shared[c])
        for (int b = h->h_fblok[c] + 1; b < h->h_fblok[c + 1]; ++b)
          (h->plan.owner[h->h_fcblk[b]] == rank ? tgt_here : tgt_else) = true;
      if (h->plan.owner[c] == rank) {
        ocblk.push_back(c);
        if (contrib[c]) {
          fan.push_back({c, (int)D.ntiles, contrib[c], 0});
          D.ntiles += nt; D.wait_mask |= contrib[c];
        }
        if (tgt_else) D.psig = 1;
      } else {
        if ((contrib[c] >> rank) & 1u) D.sig = 1;
        pull.push_back({c, (int)h->pull_tiles, 0u, 0});
        h->pull_tiles += nt;
        if (tgt_here) {
          h->foreign[l].push_back(c);
          fpull.push_back({c, (int)D.fp_tiles, 0u, 0});
          D.fp_tiles += nt; D.fp_mask |= 1u << h->plan.owner[c];
        }
      }
    }
    D.ntasks = (int)fan.size() - D.task0; D.fp_ntasks = (int)fpull.size() - D.fp_task0;
    optr[l + 1] = (int)ocblk.size();
  }
  // contributors whose last contribution to a level comes from the level just before it (the hand-off on the
  // critical path); everybody else can be pulled ahead of need
  for (int64_t k = 0; k < C; ++k) {
    if (h->plan.owner[k] == rank || (fanout && h->plan.shared[k])) continue;
    for (int b = h->h_fblok[k] + 1; b < h->h_fblok[k + 1]; ++b) {
      const int fc = h->h_fcblk[b];
      if (h->plan.owner[fc] == rank && level[fc] == level[k] + 1) h->dist_lvl[level[fc]].late_mask |= 1u << h->plan.owner[k];
    }
  }
  h->lvl_ptr = optr; h->own_lvl_cblk = ocblk;
  h->npull = (int)pull.size();
  { int rc = upload(h, fan, &h->d_fan); if (rc) return rc; }
  { int rc = upload(h, pull, &h->d_pull); if (rc) return rc; }
  { int rc = upload(h, fpull, &h->d_fpull); if (rc) return rc; }
  return PB200_SUCCESS;
}

static int dist_barrier(pb200_handle_t *h);
extern "C" int pb200_create(pb200_handle_t **out, const pb200_solver_t *s, int flttype, int factotype, int device) {
  return pb200_create_dist(out, s, flttype, factotype, device, 0, 1);
}

extern "C" int pb200_create_dist(pb200_handle_t **out, const pb200_solver_t *s, int flttype, int factotype, int device,
                                 int rank, int nranks) {
  return pb200_create_opts(out, s, flttype, factotype, device, rank, nranks, nullptr);
}

extern "C" int pb200_create_opts(pb200_handle_t **out, const pb200_solver_t *s, int flttype, int factotype, int device,
                                 int rank, int nranks, const pb200_options_t *opts) {
  if (!out || !s) return fail(PB200_ERR_BADARG, "null argument");
  const bool schur = opts && opts->schur != 0;
  if (schur && nranks > 1) return fail(PB200_ERR_BADARG, "Schur mode is single-GPU (the reference's Schur cblk lives on one process too)");
  if (nranks < 1 || nranks > PB200_MAXRANKS || rank < 0 || rank >= nranks) return fail(PB200_ERR_BADARG, "bad rank / nranks");
  if (elem_size(flttype) == 0) return fail(PB200_ERR_BADARG, "bad flttype");
  if (factotype < 0 || factotype > 3) return fail(PB200_ERR_BADARG, "bad factotype");
  if (s->cblknbr <= 0 || s->bloknbr < s->cblknbr) return fail(PB200_ERR_BADARG, "empty SolverMatrix");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PB200_ERR_CUDA, "no CUDA device: pastix_b200 has no CPU fallback");
  if (device < 0) CK(cudaGetDevice(&device));
  if (device >= ndev) return fail(PB200_ERR_BADARG, "device ordinal out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));

  pb200_handle_t *h = new pb200_handle_t();
  h->flt = flttype; h->facto = factotype; h->device = device; h->esize = elem_size(flttype);
  h->sm_count = prop.multiProcessorCount; h->cc_major = prop.major; h->cc_minor = prop.minor;
  h->rank = rank; h->nranks = nranks; h->gathered = (nranks == 1);
  h->schur = schur;
  const int64_t C = s->cblknbr, B = s->bloknbr;
  h->cblknbr = C; h->bloknbr = B;
  h->h_fcol.resize(C); h->h_width.resize(C); h->h_stride.resize(C); h->h_fblok.resize(C + 1);
  h->h_frow.resize(B); h->h_nrow.resize(B); h->h_fcblk.resize(B); h->h_coefind.resize(B);
  h->h_poff.resize(C + 1);
  auto bad = [&](const std::string &m) { pb200_destroy(h); return fail(PB200_ERR_STRUCT, m); };   // also frees what was uploaded so far
  h->h_poff[0] = 0;
  for (int64_t c = 0; c < C; ++c) {
    int64_t w = s->lcolnum[c] - s->fcolnum[c] + 1;
    if (w <= 0 || s->stride[c] < w) return bad("cblk with non-positive width or stride < width");
    if (s->lcolnum[c] >= (int64_t)1 << 31 || s->stride[c] >= (int64_t)1 << 31) return bad("dimension exceeds 2^31");
    if (c > 0 && s->fcolnum[c] != s->lcolnum[c - 1] + 1) return bad("cblks do not tile the columns contiguously");
    if (c == 0 && s->fcolnum[0] != 0) return bad("baseval must be 0");
    h->h_fcol[c] = (int)s->fcolnum[c]; h->h_width[c] = (int)w; h->h_stride[c] = (int)s->stride[c];
    h->h_fblok[c] = (int)s->bloknum[c];
    h->h_poff[c + 1] = h->h_poff[c] + s->stride[c] * w;
    h->wmax = std::max(h->wmax, (int)w); h->smax = std::max(h->smax, (int)s->stride[c]);
  }
  h->h_fblok[C] = (int)s->bloknum[C];
  if (h->h_fblok[C] != B) return bad("bloknum sentinel != bloknbr");
  h->n = s->lcolnum[C - 1] + 1;
  h->coefnbr = h->h_poff[C];
  if (h->schur && s->bloknum[C] - s->bloknum[C - 1] != 1) return bad("Schur mode: the last cblk has off-diagonal bloks");
  for (int64_t c = 0; c < C; ++c) {
    int b0 = h->h_fblok[c], b1 = h->h_fblok[c + 1];
    if (b1 <= b0) return bad("cblk without diagonal blok");
    int64_t rows = 0;
    for (int b = b0; b < b1; ++b) {
      h->h_frow[b] = (int)s->frownum[b]; h->h_nrow[b] = (int)(s->lrownum[b] - s->frownum[b] + 1);
      h->h_fcblk[b] = (int)s->cblknum[b]; h->h_coefind[b] = (int)s->coefind[b];
      if (h->h_nrow[b] <= 0) return bad("blok with no rows");
      if (s->coefind[b] != rows) return bad("coefind is not the running row count of the panel");
      if (b > b0 && s->frownum[b] <= s->lrownum[b - 1]) return bad("bloks of a cblk not sorted / overlapping");
      if (b > b0 && (s->cblknum[b] <= c || s->cblknum[b] >= C)) return bad("off-diagonal blok must face a later cblk");
      rows += h->h_nrow[b];
    }
    if (rows != s->stride[c]) return bad("stride != sum of blok heights");
    if (s->frownum[b0] != s->fcolnum[c] || s->lrownum[b0] != s->lcolnum[c]) return bad("first blok is not the diagonal blok");
  }
  // facing containment: rows of an off-diagonal blok lie inside the facing cblk's columns
  for (int64_t c = 0; c < C; ++c)
    for (int b = h->h_fblok[c] + 1; b < h->h_fblok[c + 1]; ++b) {
      int fc = h->h_fcblk[b];
      if (h->h_frow[b] < h->h_fcol[fc] || h->h_frow[b] + h->h_nrow[b] > h->h_fcol[fc] + h->h_width[fc])
        return bad("blok rows not contained in the facing cblk's column range");
    }

  // algorithmic flops of the update GEMMs as PaStiX counts them (blend_symbol_cost.c:382-430:
  // sum over off-diagonal bloks of 2*M_k*N_k*w, M_k = rows from blok k down; complex x4, LU x2)
  for (int64_t c = 0; c < C; ++c)
    for (int b = h->h_fblok[c] + 1; b < h->h_fblok[c + 1]; ++b)
      h->gemm_flops += 2.0 * (double)(h->h_stride[c] - h->h_coefind[b]) * (double)h->h_nrow[b] * (double)h->h_width[c];
  if (flttype == PB200_COMPLEXSINGLE || flttype == PB200_COMPLEXDOUBLE) h->gemm_flops *= 4.0;
  if (factotype == PB200_FACT_LU) h->gemm_flops *= 2.0;

  // ---- elimination-tree levels: level(c) > level(k) for every k with a blok facing c
  std::vector<int> level(C, 0);
  for (int64_t c = 0; c < C; ++c)
    for (int b = h->h_fblok[c] + 1; b < h->h_fblok[c + 1]; ++b) {
      int fc = h->h_fcblk[b];
      level[fc] = std::max(level[fc], level[c] + 1);
    }
  int nl = 0;
  for (int64_t c = 0; c < C; ++c) nl = std::max(nl, level[c] + 1);
  h->nlevels = nl;
  h->lvl_ptr.assign(nl + 1, 0);
  for (int64_t c = 0; c < C; ++c) h->lvl_ptr[level[c] + 1]++;
  for (int l = 0; l < nl; ++l) h->lvl_ptr[l + 1] += h->lvl_ptr[l];
  std::vector<int> lvl_cblk(C), fill(h->lvl_ptr.begin(), h->lvl_ptr.end() - 1);
  for (int64_t c = 0; c < C; ++c) lvl_cblk[fill[level[c]]++] = (int)c;

  // ---- multi-GPU: proportional subtree mapping; the factorization schedule keeps the owned cblks only
  h->plan = dist_plan(C, h->h_fblok.data(), h->h_fcblk.data(), h->h_width.data(), h->h_stride.data(), h->h_nrow.data(),
                      h->h_coefind.data(), nranks, factotype == PB200_FACT_LU, (opts && nranks > 1) ? opts->owner : nullptr);

  // up_down runs over every cblk on every GPU (factors are gathered after a distributed factorization)
  { int rc = build_solve_schedule(h, lvl_cblk); if (rc) { pb200_destroy(h); return rc; } }
  if (nranks > 1) {
    // fan-out applies to the tensor path (double / complex double, direct factorizations); PB200_NO_FANOUT=1 keeps the
    // owner-computes-everything scheme of round 1 for A/B runs
    h->fanout = (flttype == PB200_REALDOUBLE || flttype == PB200_COMPLEXDOUBLE || getenv("PB200_NO_MMA_SINGLE") == nullptr) &&
                getenv("PB200_NO_FANOUT") == nullptr;
    h->all_level = level; h->all_lvl_ptr = h->lvl_ptr; h->all_lvl_cblk = lvl_cblk;
    int rc = build_dist_levels(h, h->fanout);
    if (rc) { pb200_destroy(h); return rc; }
    lvl_cblk = h->own_lvl_cblk;
    CK(cudaStreamCreateWithFlags(&h->stream_g, cudaStreamNonBlocking));
    h->gather_ev.resize(nl);
    for (auto &e : h->gather_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    { int rc2 = upload(h, h->plan.owner, &h->d_owner); if (rc2) { pb200_destroy(h); return rc2; } }
    { std::vector<char> sh(h->plan.shared.begin(), h->plan.shared.end());
      int rc2 = upload(h, sh, &h->d_fanout); if (rc2) { pb200_destroy(h); return rc2; } }
    // flags: [0, nl) contributions into level l complete, [nl] factorization done, [nl + 1] barrier,
    //        [nl + 2, 2 nl + 2) shared panels of level l factored (fan-out)
    CK(cudaMalloc((void **)&h->d_flags, (size_t)(2 * nl + 2) * sizeof(unsigned int)));
    CK(cudaMemset(h->d_flags, 0, (size_t)(2 * nl + 2) * sizeof(unsigned int)));
    CK(cudaMalloc((void **)&h->d_dist_err, sizeof(unsigned int)));
    CK(cudaMemset(h->d_dist_err, 0, sizeof(unsigned int)));
  }

  if (h->schur) {
    // IPARM_SCHUR: compute_1d returns at once for the cblk holding the last column (sopalin_compute.c:767-772) — it
    // keeps receiving contributions and is never factored.  The factorization schedule simply does not list it.
    std::vector<int> optr(nl + 1, 0), ocblk;
    for (int l = 0; l < nl; ++l) {
      for (int q = h->lvl_ptr[l]; q < h->lvl_ptr[l + 1]; ++q) if (lvl_cblk[q] != (int)C - 1) ocblk.push_back(lvl_cblk[q]);
      optr[l + 1] = (int)ocblk.size();
    }
    h->lvl_ptr = optr; lvl_cblk = ocblk;
  }

  // ---- per-level task lists
  std::vector<RowTask> trsm, slv;
  std::vector<UpdTask> upd;
  h->trsm_ptr.assign(nl + 1, 0); h->slv_ptr.assign(nl + 1, 0); h->upd_ptr.assign(nl + 1, 0);
  h->trsm_tiles.assign(nl, 0); h->slv_tiles.assign(nl, 0); h->upd_tiles.assign(nl, 0);
  for (int l = 0; l < nl; ++l) {
    int t_tiles = 0, s_tiles = 0; long long u_tiles = 0;
    for (int q = h->lvl_ptr[l]; q < h->lvl_ptr[l + 1]; ++q) {
      int c = lvl_cblk[q];
      int m = h->h_stride[c] - h->h_width[c];
      if (m <= 0) continue;
      trsm.push_back({c, t_tiles}); t_tiles += (m + PB200_TRSM_ROWS - 1) / PB200_TRSM_ROWS;
      slv.push_back({c, s_tiles}); s_tiles += (m + PB200_SLV_ROWS - 1) / PB200_SLV_ROWS;
      for (int b = h->h_fblok[c] + 1; b < h->h_fblok[c + 1]; ++b) {
        int mi = h->h_stride[c] - h->h_coefind[b];
        int ntm = (mi + PB200_UPD_TM - 1) / PB200_UPD_TM, ntn = (h->h_nrow[b] + PB200_UPD_TN - 1) / PB200_UPD_TN;
        upd.push_back({c, b, (int)u_tiles, ntn}); u_tiles += (long long)ntm * ntn;
      }
    }
    if (u_tiles * 2 >= (1LL << 31)) return bad("too many update tiles in one level");
    h->trsm_ptr[l + 1] = (int)trsm.size(); h->slv_ptr[l + 1] = (int)slv.size(); h->upd_ptr[l + 1] = (int)upd.size();
    h->trsm_tiles[l] = t_tiles; h->slv_tiles[l] = s_tiles; h->upd_tiles[l] = u_tiles;
  }

  // ---- upload
  std::vector<int> col2cblk(h->n);
  for (int64_t c = 0; c < C; ++c) for (int j = 0; j < h->h_width[c]; ++j) col2cblk[h->h_fcol[c] + j] = (int)c;
  int *d; int64_t *d64;
#define UP(vec, field) { int rc = upload(h, vec, &d); if (rc) { pb200_destroy(h); return rc; } h->S.field = d; }
  UP(h->h_fcol, fcol) UP(h->h_width, width) UP(h->h_stride, stride) UP(h->h_fblok, fblok)
  UP(h->h_frow, frow) UP(h->h_nrow, nrow) UP(h->h_fcblk, fcblk) UP(h->h_coefind, coefind) UP(col2cblk, col2cblk)
#undef UP
  { int rc = upload(h, h->h_poff, &d64); if (rc) { pb200_destroy(h); return rc; } h->S.poff = d64; }
  h->S.cblknbr = (int)C; h->S.bloknbr = (int)B;
  { int rc = upload(h, lvl_cblk, &h->d_lvl_cblk); if (rc) { pb200_destroy(h); return rc; } }
  {
    int64_t *d_rb = nullptr;
    { int rc = upload(h, h->h_rmbase, &d_rb); if (rc) { pb200_destroy(h); return rc; } }
    h->d_rmbase = d_rb;
    const size_t nrow = (size_t)std::max<int64_t>(h->h_rmbase[C], 1);
    if (cudaMalloc((void **)&h->d_rowglob, nrow * sizeof(int)) != cudaSuccess) { pb200_destroy(h); return fail(PB200_ERR_NOMEM, "cudaMalloc(row map) failed"); }
    h->allocs.push_back(h->d_rowglob); h->device_bytes += nrow * sizeof(int);
    k_build_rowglob<<<(unsigned)C, 128>>>(h->S, d_rb, h->d_rowglob);
    CK(cudaGetLastError());
  }
  { int rc = upload(h, trsm, &h->d_trsm); if (rc) { pb200_destroy(h); return rc; } }
  { int rc = upload(h, slv, &h->d_slv); if (rc) { pb200_destroy(h); return rc; } }
  { int rc = upload(h, upd, &h->d_upd); if (rc) { pb200_destroy(h); return rc; } }

  // tensor-core path for the four precisions (double: DMMA, single: 3xTF32); PB200_NO_MMA_SINGLE=1 keeps s / c on the
  // generic SIMT kernels of round 1 for A/B runs
  if (flttype == PB200_REALDOUBLE || flttype == PB200_COMPLEXDOUBLE || getenv("PB200_NO_MMA_SINGLE") == nullptr) {
    int rc = build_mma_schedule(h, level, lvl_cblk);
    if (rc) { pb200_destroy(h); return rc; }
  }
  if (!h->use_mma && h->nranks > 1 && h->fanout) {   // generic path: owner computes every update of its cblks
    h->fanout = false;
    int rc = build_dist_levels(h, false);
    if (rc) { pb200_destroy(h); return rc; }
  }
  if (!h->use_mma) {
    int rc = build_small_schedule(h, lvl_cblk);
    if (rc) { pb200_destroy(h); return rc; }
  }

  // the TMA-staged tiles of k_gemm_scatter copy whole 64-row column segments: the last tile of the last panel reads
  // (and discards) up to one segment past the slab
  const size_t slab = (size_t)h->coefnbr * h->esize, slab_alloc = slab + PB200_SLAB_PAD;
  if (cudaMalloc(&h->dL, slab_alloc) != cudaSuccess) { pb200_destroy(h); return fail(PB200_ERR_NOMEM, "cudaMalloc(L slab) failed"); }
  h->device_bytes += slab_alloc;
  CK(cudaMemset((char *)h->dL + slab, 0, PB200_SLAB_PAD));
  if (factotype == PB200_FACT_LU) {
    if (cudaMalloc(&h->dU, slab_alloc) != cudaSuccess) { pb200_destroy(h); return fail(PB200_ERR_NOMEM, "cudaMalloc(U slab) failed"); }
    h->device_bytes += slab_alloc;
    CK(cudaMemset((char *)h->dU + slab, 0, PB200_SLAB_PAD));
  }
  if (h->use_mma && (factotype == PB200_FACT_LDLT || factotype == PB200_FACT_LDLH)) {
    if (cudaMalloc(&h->dW, slab_alloc) != cudaSuccess) { pb200_destroy(h); return fail(PB200_ERR_NOMEM, "cudaMalloc(L*D workspace) failed"); }
    h->device_bytes += slab_alloc;
    CK(cudaMemset(h->dW, 0, slab_alloc));
  }
  CK(cudaMalloc((void **)&h->d_cnt, 4 * sizeof(unsigned long long)));
  CK(cudaMemset(h->d_cnt, 0, 4 * sizeof(unsigned long long)));
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi));   // panel chain: highest priority
  }
  CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1));
  *out = h;
  return PB200_SUCCESS;
}

extern "C" int pb200_destroy(pb200_handle_t *h) {
  if (!h) return PB200_SUCCESS;
  cudaSetDevice(h->device);
  if (h->attached && !h->local_group) dist_barrier(h);   // collective: no peer is still reading our slab
  for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  cudaFree(h->d_flags); cudaFree(h->d_dist_err);
  for (void *p : h->allocs) cudaFree(p);
  if (h->h_dag_err) cudaFreeHost(h->h_dag_err);
  if (h->h_raff_partial) cudaFreeHost(h->h_raff_partial);
  cudaFree(h->d_raff_partial);
  cudaFree(h->dL); cudaFree(h->dU); cudaFree(h->dW); cudaFree(h->d_colptr); cudaFree(h->d_rows); cudaFree(h->d_vals);
  cudaFree(h->d_tvals); cudaFree(h->d_cnt); cudaFree(h->d_x); cudaFree(h->d_y); cudaFree(h->d_xt);
  if (h->fact_graph) cudaGraphExecDestroy(h->fact_graph);
  for (auto e : h->sched_ev) cudaEventDestroy(e);
  if (h->stream_u) cudaStreamDestroy(h->stream_u);
  if (h->stream_i) cudaStreamDestroy(h->stream_i);
  if (h->stream_g) cudaStreamDestroy(h->stream_g);
  for (auto e : h->gather_ev) cudaEventDestroy(e);
  if (h->ev_inv) cudaEventDestroy(h->ev_inv);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return PB200_SUCCESS;
}

extern "C" int pb200_info(const pb200_handle_t *h, pb200_info_t *info) {
  if (!h || !info) return fail(PB200_ERR_BADARG, "null argument");
  info->n = h->n; info->coefnbr = h->coefnbr; info->nlevels = h->nlevels; info->device_bytes = (int64_t)h->device_bytes;
  info->device = h->device; info->sm_count = h->sm_count; info->cc_major = h->cc_major; info->cc_minor = h->cc_minor;
  return PB200_SUCCESS;
}

extern "C" int pb200_panel_offsets(const pb200_handle_t *h, int64_t *offsets) {
  if (!h || !offsets) return fail(PB200_ERR_BADARG, "null argument");
  memcpy(offsets, h->h_poff.data(), (size_t)(h->cblknbr + 1) * sizeof(int64_t));
  return PB200_SUCCESS;
}

extern "C" int64_t pb200_last_launches(const pb200_handle_t *h) { return h ? h->last_launches : 0; }

extern "C" int pb200_set_profile(pb200_handle_t *h, int on) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  h->prof_on = (on != 0);
  return PB200_SUCCESS;
}
extern "C" int pb200_get_profile(const pb200_handle_t *h, double *kind_ms, int64_t *kind_launches, double *gemm_flops) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  for (int q = 0; q < 4; ++q) { if (kind_ms) kind_ms[q] = h->prof_ms[q]; if (kind_launches) kind_launches[q] = h->prof_n[q]; }
  if (gemm_flops) *gemm_flops = h->gemm_flops;
  return PB200_SUCCESS;
}

template <class T>
static double norm1_t(int64_t n, const int64_t *colptr, const T *v) {
  double mx = 0;
  for (int64_t j = 0; j < n; ++j) {
    double s = 0;
    for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) s += (double)ST<T>::abs(v[p]);
    mx = std::max(mx, s);
  }
  return mx;
}
extern "C" double pb200_norm1(int flttype, int64_t n, const int64_t *colptr, const void *values) {
  switch (flttype) {
    case PB200_REALSINGLE: return norm1_t(n, colptr, (const float *)values);
    case PB200_REALDOUBLE: return norm1_t(n, colptr, (const double *)values);
    case PB200_COMPLEXSINGLE: return norm1_t(n, colptr, (const cfloat *)values);
    case PB200_COMPLEXDOUBLE: return norm1_t(n, colptr, (const cdouble *)values);
  }
  return -1.0;
}

static int invert_dispatch(pb200_handle_t *h, cudaStream_t sm, int which = 0);
static int dist_barrier(pb200_handle_t *h);
static int ensure_gathered(pb200_handle_t *h);
// ------------------------------------------------------------------ dispatch helpers
#define DISPATCH_T(h, FN, ...)                                                    \
  switch ((h)->flt) {                                                             \
    case PB200_REALSINGLE: return FN<float>(__VA_ARGS__);                         \
    case PB200_REALDOUBLE: return FN<double>(__VA_ARGS__);                        \
    case PB200_COMPLEXSINGLE: return FN<cfloat>(__VA_ARGS__);                     \
    case PB200_COMPLEXDOUBLE: return FN<cdouble>(__VA_ARGS__);                    \
  }                                                                               \
  return fail(PB200_ERR_BADARG, "bad flttype");

// ------------------------------------------------------------------ multi-GPU helpers
static const unsigned long long kDistTimeoutNs = 60ULL * 1000000000ULL;
template <class T>
static int64_t launch_fanin(pb200_handle_t *h, int l, cudaStream_t sm, bool split = true) {
  const auto &D = h->dist_lvl[l];
  int64_t n = 0;
  if (D.sig) { k_dist_signal<<<1, 32, 0, sm>>>(h->d_flags, l, h->epoch); ++n; }
  if (D.ntasks) {
    const unsigned int late = split ? (D.wait_mask & D.late_mask) : D.wait_mask;
    const unsigned int early = D.wait_mask & ~late;
    if (early) {
      // pulled on the side stream as soon as those GPUs have published the level — usually several levels ahead
      k_dist_wait<<<1, 32, 0, h->stream_g>>>(h->peers, early, l, h->epoch, kDistTimeoutNs, h->d_dist_err);
      k_fanin_gather<T><<<(unsigned)D.ntiles, 256, 0, h->stream_g>>>(h->S, h->peers, (T *)h->dL, (T *)h->dU, h->d_fan + D.task0, D.ntasks, early);
      cudaEventRecord(h->gather_ev[l], h->stream_g);
      n += 2;
    }
    if (late) {
      k_dist_wait<<<1, 32, 0, sm>>>(h->peers, late, l, h->epoch, kDistTimeoutNs, h->d_dist_err);
      k_fanin_gather<T><<<(unsigned)D.ntiles, 256, 0, sm>>>(h->S, h->peers, (T *)h->dL, (T *)h->dU, h->d_fan + D.task0, D.ntasks, late);
      n += 2;
    }
    if (early) cudaStreamWaitEvent(sm, h->gather_ev[l], 0);
  }
  return n;
}
static int dist_check(pb200_handle_t *h) {
  unsigned int e = 0;
  CK(cudaMemcpy(&e, h->d_dist_err, sizeof(e), cudaMemcpyDeviceToHost));
  if (e) return fail(PB200_ERR_STATE, "multi-GPU wait timed out on rank " + std::to_string(e - 1) + " (peer did not reach the same step)");
  return PB200_SUCCESS;
}
// device-side barrier over the peers' flag arrays: everything this process launched before is finished on
// every GPU when it returns (collective: every rank calls it the same number of times)
static int dist_barrier(pb200_handle_t *h) {
  if (h->nranks == 1) return PB200_SUCCESS;
  if (!h->attached) return fail(PB200_ERR_STATE, "pb200_ipc_attach has not been called");
  CK(cudaStreamSynchronize(h->stream));
  if (h->stream_u) CK(cudaStreamSynchronize(h->stream_u));
  if (h->stream_g) CK(cudaStreamSynchronize(h->stream_g));
  ++h->bar_epoch;
  k_dist_signal<<<1, 32, 0, h->stream>>>(h->d_flags, h->nlevels + 1, h->bar_epoch);
  k_dist_wait<<<1, 32, 0, h->stream>>>(h->peers, (1u << h->nranks) - 1u, h->nlevels + 1, h->bar_epoch, kDistTimeoutNs, h->d_dist_err);
  CK(cudaStreamSynchronize(h->stream));
  return dist_check(h);
}
// copy the other GPUs' factored panels into the local slab (once per factorization, before the first
// up_down / read-back)
template <class T>
static int gather_t(pb200_handle_t *h) {
  k_dist_wait<<<1, 32, 0, h->stream>>>(h->peers, (1u << h->nranks) - 1u, h->nlevels, h->epoch, kDistTimeoutNs, h->d_dist_err);
  if (h->npull > 0)
    k_pull_panels<T><<<(unsigned)h->pull_tiles, 256, 0, h->stream>>>(h->S, h->peers, (T *)h->dL, (T *)h->dU, (T *)nullptr, h->d_owner, h->d_pull, h->npull);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  { int rc = dist_check(h); if (rc) return rc; }
  return invert_dispatch(h, h->stream, 2);   // triangles of the panels that just arrived (ours were inverted at the end of the factorization)
}
static int gather_dispatch(pb200_handle_t *h) { DISPATCH_T(h, gather_t, h) }
static int ensure_gathered(pb200_handle_t *h) {
  if (h->nranks == 1 || h->gathered) return PB200_SUCCESS;
  if (!h->factorized) return fail(PB200_ERR_STATE, "not factorized");
  int rc = gather_dispatch(h);
  if (rc) return rc;
  h->gathered = true;
  return PB200_SUCCESS;
}
template <class T>
static int reassemble_t(pb200_handle_t *h) {
  size_t slab = (size_t)h->coefnbr * sizeof(T);
  { int rc = dist_barrier(h); if (rc) return rc; }   // peers may still be reading our panels / fan-in buffers
  CK(cudaMemsetAsync(h->dL, 0, slab, h->stream));
  if (h->dU) CK(cudaMemsetAsync(h->dU, 0, slab, h->stream));
  CK(cudaMemsetAsync(h->d_cnt + 1, 0, sizeof(unsigned long long), h->stream));
  int n = (int)h->n;
  k_assemble<T><<<(n + 255) / 256, 256, 0, h->stream>>>(h->S, n, h->d_colptr, h->d_rows, (const T *)h->d_vals,
                                                        (const T *)h->d_tvals, h->herm ? 1 : 0, (T *)h->dL, (T *)h->dU, h->d_cnt + 1,
                                                        h->d_owner, h->rank);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  h->assembled = true; h->factorized = false; h->inv_ready = false; h->gathered = (h->nranks == 1);
  return PB200_SUCCESS;
}

extern "C" int pb200_reassemble(pb200_handle_t *h) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  if (!h->d_colptr) return fail(PB200_ERR_STATE, "pb200_assemble has not been called");
  CK(cudaSetDevice(h->device));
  DISPATCH_T(h, reassemble_t, h)
}

extern "C" int pb200_assemble(pb200_handle_t *h, const int64_t *colptr, const int64_t *rows, const void *values,
                              const void *tvalues) {
  if (!h || !colptr || !rows || !values) return fail(PB200_ERR_BADARG, "null argument");
  if (h->facto == PB200_FACT_LU && !tvalues) return fail(PB200_ERR_BADARG, "LU needs the transposed values");
  CK(cudaSetDevice(h->device));
  int64_t nnz = colptr[h->n];
  if (colptr[0] != 0) return fail(PB200_ERR_BADARG, "colptr must be 0-based");
  if (nnz != h->nnz || !h->d_colptr) {
    cudaFree(h->d_colptr); cudaFree(h->d_rows); cudaFree(h->d_vals); cudaFree(h->d_tvals);
    h->d_colptr = nullptr; h->d_rows = nullptr; h->d_vals = nullptr; h->d_tvals = nullptr;
    CK(cudaMalloc((void **)&h->d_colptr, (size_t)(h->n + 1) * sizeof(int64_t)));
    CK(cudaMalloc((void **)&h->d_rows, (size_t)std::max<int64_t>(nnz, 1) * sizeof(int)));
    CK(cudaMalloc(&h->d_vals, (size_t)std::max<int64_t>(nnz, 1) * h->esize));
    if (h->facto == PB200_FACT_LU) CK(cudaMalloc(&h->d_tvals, (size_t)std::max<int64_t>(nnz, 1) * h->esize));
    h->nnz = nnz;
  }
  std::vector<int> r32((size_t)nnz);
  for (int64_t i = 0; i < nnz; ++i) r32[i] = (int)rows[i];
  CK(cudaMemcpyAsync(h->d_colptr, colptr, (size_t)(h->n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_rows, r32.data(), (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_vals, values, (size_t)nnz * h->esize, cudaMemcpyHostToDevice, h->stream));
  if (h->d_tvals) CK(cudaMemcpyAsync(h->d_tvals, tvalues, (size_t)nnz * h->esize, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return pb200_reassemble(h);
}

// Assembly from an internal CSC that pb200_csc_build left in HBM (csc_build.cu): no host copy, no second upload.
extern "C" int pb200_assemble_csc(pb200_handle_t *h, const pb200_csc_t *c) {
  if (!h || !c) return fail(PB200_ERR_BADARG, "null argument");
  if (!c->valid) return fail(PB200_ERR_STATE, "no internal CSC built");
  if (c->flt != h->flt || c->n != h->n) return fail(PB200_ERR_BADARG, "internal CSC does not match this SolverMatrix (precision / order)");
  if (h->facto == PB200_FACT_LU && !c->has_t) return fail(PB200_ERR_BADARG, "LU needs the transposed values");
  CK(cudaSetDevice(h->device));
  h->herm = (c->type == 'H');
  const int64_t nnz = c->nnz;
  if (nnz != h->nnz || !h->d_colptr) {
    cudaFree(h->d_colptr); cudaFree(h->d_rows); cudaFree(h->d_vals); cudaFree(h->d_tvals);
    h->d_colptr = nullptr; h->d_rows = nullptr; h->d_vals = nullptr; h->d_tvals = nullptr;
    CK(cudaMalloc((void **)&h->d_colptr, (size_t)(h->n + 1) * sizeof(int64_t)));
    CK(cudaMalloc((void **)&h->d_rows, (size_t)std::max<int64_t>(nnz, 1) * sizeof(int)));
    CK(cudaMalloc(&h->d_vals, (size_t)std::max<int64_t>(nnz, 1) * h->esize));
    if (h->facto == PB200_FACT_LU) CK(cudaMalloc(&h->d_tvals, (size_t)std::max<int64_t>(nnz, 1) * h->esize));
    h->nnz = nnz;
  }
  // the device CSC may live on another GPU of the box (one pb200_csc_build, several handles): peer copies
  CK(cudaMemcpyPeerAsync(h->d_colptr, h->device, c->d_colptr, c->device, (size_t)(h->n + 1) * sizeof(int64_t), h->stream));
  CK(cudaMemcpyPeerAsync(h->d_rows, h->device, c->d_rows, c->device, (size_t)nnz * sizeof(int), h->stream));
  CK(cudaMemcpyPeerAsync(h->d_vals, h->device, c->d_vals, c->device, (size_t)nnz * h->esize, h->stream));
  if (h->d_tvals) CK(cudaMemcpyPeerAsync(h->d_tvals, h->device, c->d_tvals, c->device, (size_t)nnz * h->esize, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return pb200_reassemble(h);
}

// ------------------------------------------------------------------ factorization
template <class T, int FACTO>
static int factorize_tf(pb200_handle_t *h, double crit) {
  T *L = (T *)h->dL, *U = (T *)h->dU;
  const int lu = (FACTO == F_LU) ? 2 : 1;
  // dynamic shared memory for the diagonal block: as much as fits
  int smem_max = 200 * 1024;
  if (!(h->attr_mask & 1u)) {   // once per handle
    CK(cudaFuncSetAttribute(k_diag_factor<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    h->attr_mask |= 1u;
  }
  const size_t sm_lvl_smem = (size_t)PB200_SM_WARPS * sizeof(SmallWs<T>), sm_chain_smem = (size_t)SmChain<T>::WARPS * sizeof(SmallWs<T>);
  if (!(h->attr_mask & 2u)) {
    CK(cudaFuncSetAttribute(k_small_level<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_lvl_smem));
    CK(cudaFuncSetAttribute(k_small_chain<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chain_smem));
    h->attr_mask |= 2u;
  }
  int64_t launches = 0;
  for (const auto &gs : h->gsteps) {
    if (gs.kind == 2) {   // run of thin small levels: one CTA walks them
      const int q0 = h->lvl_ptr[gs.l0];
      k_small_chain<T, FACTO><<<1, SmChain<T>::WARPS * 32, sm_chain_smem, h->stream>>>(
          h->S, L, U, h->d_lvl_cblk, h->d_lvl_ptr + gs.l0, gs.l1 - gs.l0, h->d_sm_pbase, h->d_sm_tabL, h->d_sm_tabU, crit, h->d_cnt);
      (void)q0;
      ++launches;
      continue;
    }
    const int l = gs.l0;
    int nc = h->lvl_ptr[l + 1] - h->lvl_ptr[l];
    if (h->nranks > 1) launches += launch_fanin<T>(h, l, h->stream, getenv("PB200_EARLY_GATHER") != nullptr);
    if (nc == 0) continue;
    if (gs.kind == 1) {   // every cblk of the level is small: diag + trsm + updates fused, one warp per cblk
      const int q0 = h->lvl_ptr[l];
      k_small_level<T, FACTO><<<(nc + PB200_SM_WARPS - 1) / PB200_SM_WARPS, PB200_SM_WARPS * 32, sm_lvl_smem, h->stream>>>(
          h->S, L, U, h->d_lvl_cblk + q0, nc, h->d_sm_pbase + q0, h->d_sm_tabL, h->d_sm_tabU, crit, h->d_cnt);
      ++launches;
      continue;
    }
    // smem sized for the widest cblk of this launch would need a per-level max; use global wmax bound
    int elems = std::min<long long>((long long)h->wmax * h->wmax, smem_max / (long long)sizeof(T));
    size_t smem = (size_t)elems * sizeof(T);
    k_diag_factor<T, FACTO><<<nc, 256, smem, h->stream>>>(h->S, L, U, h->d_lvl_cblk + h->lvl_ptr[l], crit, h->d_cnt, elems);
    ++launches;
    if (h->trsm_tiles[l] > 0) {
      k_panel_trsm<T, FACTO><<<h->trsm_tiles[l] * lu, PB200_TRSM_ROWS, 0, h->stream>>>(
          h->S, L, U, h->d_trsm + h->trsm_ptr[l], h->trsm_ptr[l + 1] - h->trsm_ptr[l]);
      ++launches;
    }
    if (h->upd_tiles[l] > 0) {
      k_update<T, FACTO><<<(unsigned)(h->upd_tiles[l] * lu), 256, 0, h->stream>>>(
          h->S, L, U, h->d_upd + h->upd_ptr[l], h->upd_ptr[l + 1] - h->upd_ptr[l]);
      ++launches;
    }
  }
  CK(cudaGetLastError());
  h->last_launches = launches;
  return PB200_SUCCESS;
}

static int h_gemm_mode(const pb200_handle_t *h, int task) { return h->h_gemm_modes[task]; }
template <class T> static int invert_range(pb200_handle_t *h, int sp0, int sp1, cudaStream_t sm, int nbmax);

// ------------------------------------------------------------------ factorization (tensor-core path)
template <class T, int FACTO>
static int factorize_mma(pb200_handle_t *h, double crit) {
  T *L = (T *)h->dL, *U = (T *)h->dU;
  const int lu = (FACTO == F_LU) ? 2 : 1;
  if (!(h->attr_mask & 4u)) {   // once per handle
    CK(cudaFuncSetAttribute(k_gemm_scatter<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem_bytes<T>()));
    CK(cudaFuncSetAttribute(k_trsm_mma<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)trsm_smem_bytes<T>(SubCfg<T>::NBMAX)));
    h->attr_mask |= 4u;
  }
  int64_t launches = 0;
  const bool prof = h->prof_on || getenv("PB200_PROFILE") != nullptr;
  double tkind[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; long long nk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double tlevel_max = 0; int lvl_max = -1;
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  if (prof) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); }
  const bool serial = prof || getenv("PB200_SERIAL") != nullptr;
  const bool diag_old = getenv("PB200_DIAG_OLD") != nullptr;
  const bool timeline = !serial && getenv("PB200_TIMELINE") != nullptr;   // A/B switch: the one-barrier-per-pivot kernel of round 1
  const bool overlap_inv = !serial && h->nranks == 1 && getenv("PB200_INV_OVERLAP") != nullptr;   // opt-in: measured slower (r01)
  // The schedule is a fixed sequence of launches on two streams joined by events: capture it once (per threshold value,
  // which is a by-value kernel argument) and replay it — kernel-to-kernel dependencies then resolve on the device
  // without the stream scheduler in between.  Multi-GPU schedules carry an epoch argument and stay on streams.
  const bool use_graph = !serial && !overlap_inv && h->nranks == 1 && h->nlevels > 0 && getenv("PB200_GRAPH") != nullptr;
  // programmatic dependent launch of the chain kernels: PB200_PDL=0 plain stream order, 1 every chain kernel, 2 only
  // launches of at most PB200_PDL_MAX CTAs (default: half the SMs; measured on C2 / 64^3 z-LU: 8: 21.66 / 133.5 ms, 32: 21.54 / 132.4,
  // 74: 21.37 / 133.9, 148: 21.50 / 137.9, 296: 21.82 / 142.5, every launch: 23.35 / 147.4, none: 22.14 / 133.5 — the CTAs of an early-scheduled large launch sit on SM resources the
  // OTHER stream's kernels could use — measured slower, profiles/README.md); per-launch event timing needs the plain order
  // compact shared-memory diagonal kernel for real LLt / LDLt: opt-in (PB200_DIAG_CMP=1) — measured SLOWER than the
  // register-resident k_diag_blk (C2 22.75 vs 21.43 ms, C3 271.1 vs 268.4 ms: one warp per panel, the rest at a barrier)
  const bool diag_cmp = getenv("PB200_DIAG_CMP") != nullptr && atoi(getenv("PB200_DIAG_CMP")) != 0;
  const int pdl_mode = (prof || use_graph) ? 0 : (getenv("PB200_PDL") ? atoi(getenv("PB200_PDL")) : 2);
  const long long pdl_max = pdl_mode == 1 ? (1LL << 40) : (getenv("PB200_PDL_MAX") ? atoll(getenv("PB200_PDL_MAX")) : (long long)h->sm_count / 2);
  if (use_graph && h->fact_graph && h->fact_graph_crit == crit) {
    CK(cudaGraphLaunch(h->fact_graph, h->stream));
    h->last_launches = h->fact_graph_launches;
    return PB200_SUCCESS;
  }
  if (use_graph) {
    if (h->fact_graph) { cudaGraphExecDestroy(h->fact_graph); h->fact_graph = nullptr; }
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  }
  for (const auto &st : h->steps) {
    cudaStream_t sm = (serial || st.strm == 0) ? h->stream : h->stream_u;
    if (!serial && st.wait_ev >= 0) CK(cudaStreamWaitEvent(sm, h->sched_ev[st.wait_ev], 0));
    if (st.ntasks == 0 || st.kind == 4) {
      if (!serial && st.rec_ev >= 0) CK(cudaEventRecord(h->sched_ev[st.rec_ev], sm));
      continue;
    }
    if (prof) cudaEventRecord(pe0, h->stream);
    switch (st.kind) {
      case 0: {
        if (diag_old)
          k_diag_sub<T, FACTO><<<st.ntasks, 256, 0, sm>>>(h->S, L, U, h->d_sub + st.task0, crit, h->d_cnt);
        else if (diag_cmp && st.nbmax <= 64 && !ST<T>::is_complex && (FACTO == F_LLT || FACTO == F_LDLT)) {
          if constexpr (!ST<T>::is_complex && (FACTO == F_LLT || FACTO == F_LDLT))
            CK(launch_chain(pdl_mode && st.ntasks <= pdl_max, k_diag_cmp<T, FACTO>, dim3(st.ntasks), dim3(256), 0, sm, h->S, L,
                            (const SubTask *)(h->d_sub + st.task0), crit, h->d_cnt));
        } else if (st.nbmax <= 64)
          CK(launch_chain(pdl_mode && st.ntasks <= pdl_max, k_diag_blk<T, FACTO, 4>, dim3(st.ntasks), dim3(256), 0, sm, h->S, L, U, (const SubTask *)(h->d_sub + st.task0), crit, h->d_cnt));
        else if constexpr (SubCfg<T>::NBMAX > 64)
          CK(launch_chain(pdl_mode && st.ntasks <= pdl_max, k_diag_blk<T, FACTO, SubCfg<T>::NBMAX / 16>, dim3(st.ntasks), dim3(256), 0, sm, h->S, L, U,
                          (const SubTask *)(h->d_sub + st.task0), crit, h->d_cnt));
      } break;
      case 1:
        CK(launch_chain(pdl_mode && st.ntiles * lu <= pdl_max, k_trsm_mma<T, FACTO>, dim3((unsigned)(st.ntiles * lu)), dim3(128), trsm_smem_bytes<T>(st.nbmax), sm,
                        h->S, L, U, (T *)h->dW, (const SubTask *)(h->d_sub + st.task0), (int)st.ntasks));
        break;
      case 2:
        CK(launch_chain(pdl_mode && st.ntiles * lu <= pdl_max, k_gemm_scatter<T, FACTO>, dim3((unsigned)(st.ntiles * lu)), dim3(UpdCfg<T>::NT), upd_smem_bytes<T>(), sm,
                        h->M, L, U, (const T *)h->dW, (const TileDesc *)(h->d_desc + st.t2t0)));
        break;
      case 3:
        if (FACTO == F_LU)
          k_diag_transpose<T><<<dim3(4, std::min(st.ntasks, 65535)), dim3(32, 8), 0, sm>>>(h->S, L, U, h->d_lvl_cblk + st.task0, st.ntasks);
        break;
      case 5:
        launches += launch_fanin<T>(h, st.lvl, sm, !serial && getenv("PB200_EARLY_GATHER") != nullptr) - 1;   // opt-in: measured neutral at N=4 (r01)
        break;
      case 7:   // fan-out: the shared panels of this level are factored (everything launched before on this stream)
        k_dist_signal<<<1, 32, 0, sm>>>(h->d_flags, h->nlevels + 2 + st.lvl, h->epoch);
        break;
      case 8: { // fan-out: pull the shared panels of this level factored elsewhere (and their L*D copies)
        const auto &D = h->dist_lvl[st.lvl];
        k_dist_wait<<<1, 32, 0, sm>>>(h->peers, D.fp_mask, h->nlevels + 2 + st.lvl, h->epoch, kDistTimeoutNs, h->d_dist_err);
        k_pull_panels<T><<<(unsigned)D.fp_tiles, 256, 0, sm>>>(h->S, h->peers, L, U, (T *)h->dW, h->d_owner, h->d_fpull + D.fp_task0, D.fp_ntasks);
        ++launches;
      } break;
      case 6:
        if (FACTO == F_LU)
          k_diag_complete_lu<T><<<dim3(8, std::min(st.ntasks, 65535)), dim3(32, 8), 0, sm>>>(h->S, L, U, h->d_lvl_cblk + st.task0, st.ntasks,
                                                                                         SubCfg<T>::NBMAX);
        break;
    }
    ++launches;
    if (!serial && st.rec_ev >= 0) CK(cudaEventRecord(h->sched_ev[st.rec_ev], sm));
    if (timeline && st.rec_ev >= 0) {   // dev aid: when the panel work (rec_ev < nlevels) / the bulk update of a level finished
      if (h->tl_ev.size() != 2 * (size_t)h->nlevels) { h->tl_ev.resize(2 * (size_t)h->nlevels, nullptr); for (auto &e : h->tl_ev) cudaEventCreate(&e); }
      cudaEventRecord(h->tl_ev[st.rec_ev], sm);
    }
    if (overlap_inv && st.rec_ev >= 0 && st.rec_ev < h->nlevels) {
      // panel(l) is final: invert its diagonal triangles (up_down preparation) underneath the rest
      const int l = st.rec_ev;
      if (h->inv_lvl_ptr[l + 1] > h->inv_lvl_ptr[l]) {
        CK(cudaStreamWaitEvent(h->stream_i, h->sched_ev[l], 0));
        int rc = invert_range<T>(h, h->inv_lvl_ptr[l], h->inv_lvl_ptr[l + 1], h->stream_i, h->inv_lvl_nbmax[l]);
        if (rc) return rc;
        launches += lu;
      }
    }
    if (prof) {
      cudaEventRecord(pe1, h->stream); cudaEventSynchronize(pe1);
      float ms = 0; cudaEventElapsedTime(&ms, pe0, pe1);
      int kd = st.kind;
      if (kd == 2 && h_gemm_mode(h, st.task0) == 1) kd = 3;   // internal update counted apart
      tkind[kd] += ms; nk[kd]++;
      if (getenv("PB200_PROFILE_VERBOSE"))
        fprintf(stderr, "  lvl %3d kind %d tasks %6d tiles %8lld nbmax %3d : %8.3f ms\n", st.lvl, kd, st.ntasks, st.ntiles, st.nbmax, ms);
    }
  }
  if (prof) {
    for (int q = 0; q < 4; ++q) { h->prof_ms[q] = tkind[q]; h->prof_n[q] = nk[q]; }
    h->prof_ms[3] += tkind[5] + tkind[6] + tkind[7] + tkind[8]; h->prof_n[3] += nk[5] + nk[6] + nk[7] + nk[8];
    if (getenv("PB200_PROFILE") != nullptr)
    fprintf(stderr, "[pb200 profile] diag %.3f ms (%lld)  trsm %.3f ms (%lld)  ext-update %.3f ms (%lld)  int-update/transpose %.3f ms (%lld)\n",
            tkind[0], nk[0], tkind[1], nk[1], tkind[2], nk[2], tkind[3], nk[3]);
    cudaEventDestroy(pe0); cudaEventDestroy(pe1);
  }
  (void)tlevel_max; (void)lvl_max;
  if (!serial && h->nlevels > 0) {
    // join: the panel stream (which carries the timing events) waits for the last bulk update
    CK(cudaStreamWaitEvent(h->stream, h->sched_ev[2 * (size_t)h->nlevels - 1], 0));
    if (overlap_inv) {
      CK(cudaEventRecord(h->ev_inv, h->stream_i));
      CK(cudaStreamWaitEvent(h->stream, h->ev_inv, 0));
      h->inv_ready = true;
    }
  }
  CK(cudaGetLastError());
  h->last_launches = launches;
  if (use_graph) {
    cudaGraph_t g = nullptr;
    CK(cudaStreamEndCapture(h->stream, &g));
    cudaError_t e = cudaGraphInstantiate(&h->fact_graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { h->fact_graph = nullptr; return fail(PB200_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    h->fact_graph_crit = crit; h->fact_graph_launches = launches;
    CK(cudaGraphLaunch(h->fact_graph, h->stream));
  }
  return PB200_SUCCESS;
}

template <class T>
static int factorize_t(pb200_handle_t *h, double crit) {
  {
    if (h->use_mma) {
      switch (h->facto) {
        case PB200_FACT_LLT: return factorize_mma<T, F_LLT>(h, crit);
        case PB200_FACT_LDLT: return factorize_mma<T, F_LDLT>(h, crit);
        case PB200_FACT_LU: return factorize_mma<T, F_LU>(h, crit);
        case PB200_FACT_LDLH:
          return ST<T>::is_complex ? factorize_mma<T, F_LDLH>(h, crit) : factorize_mma<T, F_LDLT>(h, crit);
      }
    }
  }
  switch (h->facto) {
    case PB200_FACT_LLT: return factorize_tf<T, F_LLT>(h, crit);
    case PB200_FACT_LDLT: return factorize_tf<T, F_LDLT>(h, crit);
    case PB200_FACT_LU: return factorize_tf<T, F_LU>(h, crit);
    case PB200_FACT_LDLH:
      return ST<T>::is_complex ? factorize_tf<T, F_LDLH>(h, crit) : factorize_tf<T, F_LDLT>(h, crit);
  }
  return fail(PB200_ERR_BADARG, "bad factotype");
}

static int factorize_dispatch(pb200_handle_t *h, double crit) { DISPATCH_T(h, factorize_t, h, crit) }

extern "C" int pb200_factorize(pb200_handle_t *h, double critere, int64_t *nbpivot, double *seconds) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  if (!h->assembled) return fail(PB200_ERR_STATE, "panels not assembled (pb200_assemble / pb200_set_coeftab)");
  if (h->factorized) return fail(PB200_ERR_STATE, "panels already factorized; reassemble first");
  CK(cudaSetDevice(h->device));
  if (h->nranks > 1) {
    if (!h->attached) return fail(PB200_ERR_STATE, "pb200_ipc_attach has not been called");
    ++h->epoch;
  }
  CK(cudaMemsetAsync(h->d_cnt, 0, sizeof(unsigned long long), h->stream));
  CK(cudaEventRecord(h->ev0, h->stream));
  int rc = factorize_dispatch(h, critere);
  if (rc) return rc;
  if (h->nranks == 1) {
    if (!h->inv_ready) rc = invert_dispatch(h, h->stream);   // diagonal triangles inverted once, for the up_down sweeps
    if (rc) return rc;
  } else {
    // the same preparation of the up_down, for the cblks this GPU owns (N = 1 has all of it inside DPARM_FACT_TIME)
    rc = invert_dispatch(h, h->stream, 1);
    if (rc) return rc;
    k_dist_signal<<<1, 32, 0, h->stream>>>(h->d_flags, h->nlevels, h->epoch);   // our share of the factorization is done
    h->gathered = false;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->nranks > 1) { rc = dist_check(h); if (rc) return rc; }
  float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (!h->tl_ev.empty() && getenv("PB200_TIMELINE") != nullptr) {
    if (h->stream_u) cudaStreamSynchronize(h->stream_u);
    for (int l = 0; l < h->nlevels; ++l) {
      float tp = -1.f, tu = -1.f;
      if (cudaEventQuery(h->tl_ev[l]) == cudaSuccess) cudaEventElapsedTime(&tp, h->ev0, h->tl_ev[l]);
      if (cudaEventQuery(h->tl_ev[h->nlevels + l]) == cudaSuccess) cudaEventElapsedTime(&tu, h->ev0, h->tl_ev[h->nlevels + l]);
      fprintf(stderr, "[pb200 timeline] rank %d lvl %3d panel+U1 done %9.3f ms   bulk update done %9.3f ms\n", h->rank, l, tp, tu);
    }
    (void)cudaGetLastError();
  }
  unsigned long long nb = 0;
  CK(cudaMemcpy(&nb, h->d_cnt, sizeof(nb), cudaMemcpyDeviceToHost));
  if (nbpivot) *nbpivot = (int64_t)nb;
  if (seconds) *seconds = ms * 1e-3;
  h->factorized = true;
  return PB200_SUCCESS;
}

template <class T>
static int inertia_t(pb200_handle_t *h, int64_t *out) {
  CK(cudaMemsetAsync(h->d_cnt + 2, 0, sizeof(unsigned long long), h->stream));
  k_inertia<T><<<(int)((h->cblknbr + 127) / 128), 128, 0, h->stream>>>(h->S, (const T *)h->dL, h->d_cnt + 2);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  unsigned long long v = 0;
  CK(cudaMemcpy(&v, h->d_cnt + 2, sizeof(v), cudaMemcpyDeviceToHost));
  *out = (int64_t)v;
  return PB200_SUCCESS;
}
extern "C" int pb200_inertia(pb200_handle_t *h, int64_t *inertia) {
  if (!h || !inertia) return fail(PB200_ERR_BADARG, "null argument");
  if (!h->factorized) return fail(PB200_ERR_STATE, "not factorized");
  CK(cudaSetDevice(h->device));
  { int rc = ensure_gathered(h); if (rc) return rc; }
  DISPATCH_T(h, inertia_t, h, inertia)
}

// ------------------------------------------------------------------ solve
// invert the diagonal triangles of the freshly factored panels (one CTA per sub-panel)
static const int kInvClasses[] = {16, 32, 48, 64, 96, 128};
template <class T>
static size_t inv_smem(int nbmax) { return ((size_t)nbmax * (nbmax + 1) / 2 + (size_t)tri_xelems(nbmax)) * sizeof(T); }   // packed W + per-warp X rectangles
template <class T>
static int inv_attr(pb200_handle_t *h) {
  if (!(h->attr_mask & 8u)) {   // once per handle
    CK(cudaFuncSetAttribute(k_tri_inverse<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)inv_smem<T>(SlvCfg<T>::NB)));
    h->attr_mask |= 8u;
  }
  return PB200_SUCCESS;
}
// sub-panels [sp0, sp1) (tasks are ordered by level, so a level is one contiguous range); nbmax = widest of them
template <class T>
static int invert_range(pb200_handle_t *h, int sp0, int sp1, cudaStream_t sm, int nbmax = SlvCfg<T>::NB) {
  { int rc = inv_attr<T>(h); if (rc) return rc; }
  if (sp1 <= sp0) return PB200_SUCCESS;
  const int unit_down = (h->facto != PB200_FACT_LLT);   // LDLt / LDLh / LU-L: unit lower triangle
  const int ntri = nbmax * (nbmax + 1) / 2;
  k_tri_inverse<T><<<sp1 - sp0, 128, inv_smem<T>(nbmax), sm>>>((const T *)h->dL, h->d_slvtask + sp0, nullptr, (T *)h->d_inv, unit_down, ntri);
  if (h->facto == PB200_FACT_LU)   // up sweep: lower triangle of ucoeftab's diagonal blok = U^T, non-unit
    k_tri_inverse<T><<<sp1 - sp0, 128, inv_smem<T>(nbmax), sm>>>((const T *)h->dU, h->d_slvtask + sp0, nullptr, (T *)h->d_inv_up, 0, ntri);
  return PB200_SUCCESS;
}
// sub-panels (which = 0: all, 1: of the cblks this GPU owns, 2: the others), one launch per size class (the shared memory
// of a launch fits its widest triangle)
template <class T>
static int invert_t(pb200_handle_t *h, cudaStream_t sm, int which) {
  { int rc = inv_attr<T>(h); if (rc) return rc; }
  const int unit_down = (h->facto != PB200_FACT_LLT);
  const std::vector<int> &ptr = h->inv_cls_ptr[which];
  const bool pdl_inv = getenv("PB200_PDL") == nullptr || atoi(getenv("PB200_PDL")) != 0;
  bool first = true;
  for (size_t k = 0; k + 1 < ptr.size(); ++k) {
    const int n = ptr[k + 1] - ptr[k];
    if (n == 0) continue;
    const int nbmax = std::min(kInvClasses[k], (int)SlvCfg<T>::NB), ntri = nbmax * (nbmax + 1) / 2;
    const int *ord = h->d_inv_order + ptr[k];
    // the first launch of the pass waits for everything before it in the stream (the factorization); the others only for
    // the previous class to have STARTED (programmatic launch without a griddepcontrol.wait: the classes are independent),
    // so the six launches run as one — PB200_PDL=0 keeps them in plain stream order
    CK(launch_chain(pdl_inv && !first, k_tri_inverse<T>, dim3(n), dim3(128), inv_smem<T>(nbmax), sm, (const T *)h->dL,
                    (const SlvTask *)h->d_slvtask, (const int *)ord, (T *)h->d_inv, unit_down, ntri));
    first = false;
    if (h->facto == PB200_FACT_LU)
      CK(launch_chain(pdl_inv, k_tri_inverse<T>, dim3(n), dim3(128), inv_smem<T>(nbmax), sm, (const T *)h->dU,
                      (const SlvTask *)h->d_slvtask, (const int *)ord, (T *)h->d_inv_up, 0, ntri));
    h->last_launches += (h->facto == PB200_FACT_LU) ? 2 : 1;
  }
  CK(cudaGetLastError());
  if (which != 1) h->inv_ready = true;
  return PB200_SUCCESS;
}
static int invert_dispatch(pb200_handle_t *h, cudaStream_t sm, int which) { DISPATCH_T(h, invert_t, h, sm, which) }

template <class T, int FACTO>
static int solve_tf(pb200_handle_t *h, T *x, int64_t ldx, int nrhs) {
  // LU with IPARM_TRANSPOSE_SOLVE: A^T = U^T L^T, i.e. the down step runs on the U^T panels (ucoeftab, non-unit
  // triangle) and the up step on L (unit) — what the reference obtains by rescaling and swapping coeftab/ucoeftab
  // around its ordinary sweeps (updo.c:165-260, 1553-1600)
  const bool tsolve = (FACTO == F_LU) && h->solve_transposed;
  if (h->schur && !h->dag_ok)
    return fail(PB200_ERR_STATE, "Schur mode: up_down needs the persistent sweeps (not an all-small / PB200_SOLVE_LEVELS schedule)");
  const T *L = (const T *)(tsolve ? h->dU : h->dL);
  const T *Mup = (FACTO == F_LU) ? (const T *)(tsolve ? h->dL : h->dU) : L;
  const T *inv = (const T *)(tsolve ? h->d_inv_up : h->d_inv);
  const T *inv_up = (FACTO == F_LU) ? (const T *)(tsolve ? h->d_inv : h->d_inv_up) : inv;
  T *y = (T *)h->d_y;
  int64_t launches = 0;
  const size_t sm_lvl_smem = (size_t)PB200_SM_WARPS * sizeof(SmallSolveWs<T>), sm_chain_smem = (size_t)SmChain<T>::WARPS * sizeof(SmallSolveWs<T>);
  if (!(h->attr_mask & 16u)) {
    CK(cudaFuncSetAttribute(k_small_solve_level<T, FACTO, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_lvl_smem));
    CK(cudaFuncSetAttribute(k_small_solve_level<T, FACTO, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_lvl_smem));
    CK(cudaFuncSetAttribute(k_small_solve_chain<T, FACTO, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chain_smem));
    CK(cudaFuncSetAttribute(k_small_solve_chain<T, FACTO, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_chain_smem));
    h->attr_mask |= 16u;
  }
  // several right-hand sides through an all-small schedule (ILU): transposed work copies, lanes over right-hand sides
  const bool tr = h->slv_all_small && nrhs >= 4 && !tsolve && !h->dag_ok;
  T *xs = x, *ys = y;
  int64_t rs = 1, cs = ldx;
  const int n = (int)h->n;
  if (tr) {
    const size_t need = (size_t)n * nrhs * sizeof(T);
    if (need > h->xt_bytes) {
      cudaFree(h->d_xt); h->d_xt = nullptr; h->xt_bytes = 0;
      if (cudaMalloc(&h->d_xt, need) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(transposed right-hand sides) failed");
      h->xt_bytes = need;
    }
    xs = (T *)h->d_xt; rs = nrhs; cs = 1;
    k_rhs_transpose<T><<<dim3((n + 31) / 32, (nrhs + 31) / 32), dim3(32, 8), 0, h->stream>>>(x, xs, n, nrhs, ldx, 1);
    ++launches;
  }
  if (h->dag_ok) {
    // one persistent launch per sweep, ordered by device-side contribution counters
    const size_t nsp = (size_t)h->nsubpanels;
    DagArgs A;
    A.ticks = h->d_dag_ticks; A.tgt = h->d_dag_tgt; A.need = h->d_dag_need;
    A.arrived = h->d_dag_state; A.ready = A.arrived + nsp; A.done = A.ready + nsp; A.cnt = A.done + nsp;
    A.ticket = A.cnt + nsp; A.err = A.ticket + 4; A.rowglob = h->d_rowglob;
    A.G = h->dag_tiles; A.nbs = h->dag_nbs; A.trace = nullptr;
    CK(cudaMemsetAsync(h->d_dag_state, 0, (4 * nsp + 8) * sizeof(unsigned int), h->stream));
    // one right-hand side: the independent-worker kernels (measured against the first generation: 64^3 LLt 2.51 -> 2.18 ms,
    // 100^3 LDLt 12.2 -> 7.8 ms, 64^3 complex LU 4.03 -> 3.78 ms); PB200_DAG3=0 keeps the first generation for A/B
    bool use3 = h->dag3_ok && nrhs == 1;
    if (const char *f3 = getenv("PB200_DAG3")) use3 = use3 && atoi(f3) != 0;
    if (use3) {
      // third generation (kernels_solve_dag3.cuh): one right-hand side, independent tile workers + a diagonal team per SM
      Dag3Args B;
      B.ticksD = h->d_dag3_ticksD; B.ticksT = h->d_dag3_ticksT; B.GD = h->dag3_GD; B.GT = h->dag3_GT; B.tgt = h->d_dag3_tgt;
      B.arrived = A.arrived; B.ready = A.ready; B.done = A.done; B.cnt = A.cnt; B.ticket = A.ticket; B.err = A.ticket + 4;
      B.rowglob = h->d_rowglob; B.trace = nullptr;
      B.xpub = h->d_dag3_xpub;
      h->dag3_epoch = (h->dag3_epoch + 2u) & 0x7ffffffeu;
      if (h->dag3_epoch == 0) { CK(cudaMemsetAsync(h->d_dag3_xpub, 0, (size_t)h->n * (h->esize / 4) * sizeof(unsigned long long), h->stream)); h->dag3_epoch = 2u; }
      B.epoch_down = h->dag3_epoch - 1u; B.epoch_up = h->dag3_epoch;
      const char *trf = getenv("PB200_DAG_TRACE");
      const size_t tb = (size_t)2 * ((size_t)B.GD + B.GT) * 8 * sizeof(unsigned long long);
      if (trf && !h->d_dag_trace) {
        CK(cudaMalloc((void **)&h->d_dag_trace, tb));
        h->allocs.push_back(h->d_dag_trace); h->device_bytes += tb;
      }
      if (trf) { B.trace = h->d_dag_trace; CK(cudaMemsetAsync(h->d_dag_trace, 0, tb, h->stream)); }
      const size_t smem = Dag3Cfg<T>::bytes;
      if (!(h->attr_mask & 128u)) {
        CK(cudaFuncSetAttribute(k_dag3<T, FACTO, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_dag3<T, FACTO, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(k_dag3<T, FACTO, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(k_dag3<T, FACTO, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        h->attr_mask |= 128u;
      }
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dag3<T, FACTO, 0>, Dag3Cfg<T>::NT, smem));
      if (occ < 1) return fail(PB200_ERR_CUDA, "persistent up_down kernels do not fit on this device");
      const unsigned grid = (unsigned)(h->sm_count * occ);
      k_dag3<T, FACTO, 0><<<grid, Dag3Cfg<T>::NT, smem, h->stream>>>(L, inv, x, y, B);
      k_dag3<T, FACTO, 1><<<grid, Dag3Cfg<T>::NT, smem, h->stream>>>(Mup, inv_up, x, y, B);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(h->h_dag_err, B.err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
      if (getenv("PB200_DAG_VERBOSE")) fprintf(stderr, "[pb200 dag3] D tickets %d, T tickets %d, smem %zu B, CTAs/SM %d\n", B.GD, B.GT, smem, occ);
      if (trf) {
        // debugging aid: [sweep][D tickets, then T tickets] = {taken, dependencies met, done, (sm << 32) | sub-tiles << 16 | is-diagonal,
        // taken, first data in shared memory, partial sums written, before the fence} (ns, %globaltimer)
        CK(cudaStreamSynchronize(h->stream));
        std::vector<unsigned long long> tr(tb / sizeof(unsigned long long));
        CK(cudaMemcpy(tr.data(), h->d_dag_trace, tb, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(trf, "wb")) { fwrite(tr.data(), sizeof(unsigned long long), tr.size(), f); fclose(f); }
      }
      h->last_launches = 2;
      return PB200_SUCCESS;
    }
    if (h->dag_rows == PB200_DAG2_ROWS) {
      // second generation (kernels_solve_dag2.cuh): three tickets in flight per CTA, one or NRMAX right-hand sides per pass
      const char *trf = getenv("PB200_DAG_TRACE");
      if (trf && !h->d_dag_trace) {
        const size_t tb = (size_t)2 * A.G * 8 * sizeof(unsigned long long);
        CK(cudaMalloc((void **)&h->d_dag_trace, tb));
        h->allocs.push_back(h->d_dag_trace); h->device_bytes += tb;
      }
      if (trf) { A.trace = h->d_dag_trace; CK(cudaMemsetAsync(h->d_dag_trace, 0, (size_t)2 * A.G * 8 * sizeof(unsigned long long), h->stream)); }
      constexpr int NRM = Dag2Cfg<T>::NRMAX;
      const bool one = nrhs == 1;
      const size_t smem = one ? Dag2Cfg<T>::bytes(1) : Dag2Cfg<T>::bytes(NRM);
      if (!(h->attr_mask & 64u)) {
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Dag2Cfg<T>::bytes(1)));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Dag2Cfg<T>::bytes(1)));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 0, NRM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Dag2Cfg<T>::bytes(NRM)));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 1, NRM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Dag2Cfg<T>::bytes(NRM)));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 0, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 0, NRM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaFuncSetAttribute(k_dag2<T, FACTO, 1, NRM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        h->attr_mask |= 64u;
      }
      int occ = 0;
      if (one) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dag2<T, FACTO, 0, 1>, PB200_DAG2_NT, smem));
      else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dag2<T, FACTO, 0, NRM>, PB200_DAG2_NT, smem));
      if (occ < 1) return fail(PB200_ERR_CUDA, "persistent up_down kernels do not fit on this device");
      const unsigned grid = (unsigned)std::min<long long>(A.G, (long long)h->sm_count * occ);
      if (one) {
        k_dag2<T, FACTO, 0, 1><<<grid, PB200_DAG2_NT, smem, h->stream>>>(L, inv, x, y, ldx, nrhs, A);
        k_dag2<T, FACTO, 1, 1><<<grid, PB200_DAG2_NT, smem, h->stream>>>(Mup, inv_up, x, y, ldx, nrhs, A);
      } else {
        k_dag2<T, FACTO, 0, NRM><<<grid, PB200_DAG2_NT, smem, h->stream>>>(L, inv, x, y, ldx, nrhs, A);
        k_dag2<T, FACTO, 1, NRM><<<grid, PB200_DAG2_NT, smem, h->stream>>>(Mup, inv_up, x, y, ldx, nrhs, A);
      }
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(h->h_dag_err, A.err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
      if (getenv("PB200_DAG_VERBOSE")) fprintf(stderr, "[pb200 dag2] tickets %d, widest sub-panel %d, smem %zu B, CTAs/SM %d\n", A.G, A.nbs, smem, occ);
      if (trf) {
        // debugging aid: [sweep][ticket] = {taken, dependencies met, done, (sm << 32) | sub-tiles << 16 | stages in flight << 8 | is-diagonal,
        // copies landed, input vector in shared memory, partial sums written, before the fence} (ns, %globaltimer)
        CK(cudaStreamSynchronize(h->stream));
        std::vector<unsigned long long> tr((size_t)2 * A.G * 8);
        CK(cudaMemcpy(tr.data(), h->d_dag_trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(trf, "wb")) { fwrite(tr.data(), sizeof(unsigned long long), tr.size(), f); fclose(f); }
      }
      h->last_launches = 2;
      return PB200_SUCCESS;
    }
    const size_t smem = DagSmem<T>::bytes(h->dag_nbs), belems = DagSmem<T>::buf_elems(h->dag_nbs);
    if (!(h->attr_mask & 32u)) {
      CK(cudaFuncSetAttribute(k_fwd_dag<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DagSmem<T>::bytes(SlvCfg<T>::NB)));
      CK(cudaFuncSetAttribute(k_bwd_dag<T, FACTO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DagSmem<T>::bytes(SlvCfg<T>::NB)));
      CK(cudaFuncSetAttribute(k_fwd_dag<T, FACTO>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      CK(cudaFuncSetAttribute(k_bwd_dag<T, FACTO>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      h->attr_mask |= 32u;
    }
    int occ_f = 0, occ_b = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_fwd_dag<T, FACTO>, PB200_DAG_NT, smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_bwd_dag<T, FACTO>, PB200_DAG_NT, smem));
    if (occ_f < 1 || occ_b < 1) return fail(PB200_ERR_CUDA, "persistent up_down kernels do not fit on this device");
    const unsigned gf = (unsigned)std::min<long long>(A.G, (long long)h->sm_count * occ_f);
    const unsigned gb = (unsigned)std::min<long long>(A.G, (long long)h->sm_count * occ_b);
    k_fwd_dag<T, FACTO><<<gf, PB200_DAG_NT, smem, h->stream>>>(L, inv, x, y, ldx, nrhs, A, belems);
    k_bwd_dag<T, FACTO><<<gb, PB200_DAG_NT, smem, h->stream>>>(Mup, inv_up, x, y, ldx, nrhs, A, belems);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_dag_err, A.err, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    if (getenv("PB200_DAG_VERBOSE")) fprintf(stderr, "[pb200 dag] tickets %d, widest sub-panel %d, smem %zu B, CTAs/SM %d/%d\n", A.G, A.nbs, smem, occ_f, occ_b);
    h->last_launches = 2;
    return PB200_SUCCESS;
  }
  for (int dir = 0; dir < 2; ++dir) {
    const T *M = dir == 0 ? L : Mup;
    for (size_t gi = 0; gi < h->sgsteps.size(); ++gi) {
      const auto &gs = h->sgsteps[dir == 0 ? gi : h->sgsteps.size() - 1 - gi];
      // the fused small-cblk kernels read the diagonal block of the panel they sweep (unit L / non-unit U^T fixed by
      // the direction); a transposed LU solve goes through the general kernels, which take the inverted triangles
      if (gs.kind == 0 || tsolve) {
        const int s0 = h->slv_lvl_step[gs.l0], s1 = h->slv_lvl_step[gs.l1];
        for (int k = 0; k < s1 - s0; ++k) {
          const auto &st = h->slv_steps[dir == 0 ? s0 + k : s1 - 1 - k];
          if (dir == 0)
            k_fwd<T, FACTO><<<(unsigned)st.ntiles, PB200_SLV_NT, 0, h->stream>>>(L, inv, xs, ys, ldx, nrhs, h->d_slvtask + st.task0,
                                                                               h->d_slv_t2t + st.t2t0, h->d_rowglob);
          else
            k_bwd<T, FACTO><<<(unsigned)st.ntiles, PB200_BWD_NT, 0, h->stream>>>(Mup, inv_up, xs, ys, ldx, nrhs, h->d_slvtask + st.task0,
                                                                               h->d_slv_t2t + st.t2t0, h->d_rowglob, h->d_slv_cnt);
          ++launches;
        }
      } else if (gs.kind == 1) {
        const int q0 = h->slv_lvl_ptr[gs.l0], nc = h->slv_lvl_ptr[gs.l0 + 1] - q0;
        const unsigned grid = (unsigned)((nc + PB200_SM_WARPS - 1) / PB200_SM_WARPS);
        if (dir == 0)
          k_small_solve_level<T, FACTO, 0><<<grid, PB200_SM_WARPS * 32, sm_lvl_smem, h->stream>>>(h->S, M, xs, ys, rs, cs, nrhs, h->d_slv_cblk + q0, nc,
                                                                                             h->d_rowglob, h->d_rmbase);
        else
          k_small_solve_level<T, FACTO, 1><<<grid, PB200_SM_WARPS * 32, sm_lvl_smem, h->stream>>>(h->S, M, xs, ys, rs, cs, nrhs, h->d_slv_cblk + q0, nc,
                                                                                             h->d_rowglob, h->d_rmbase);
        ++launches;
      } else {
        if (dir == 0)
          k_small_solve_chain<T, FACTO, 0><<<1, SmChain<T>::WARPS * 32, sm_chain_smem, h->stream>>>(h->S, M, xs, ys, rs, cs, nrhs, h->d_slv_cblk,
                                                                                               h->d_slv_lvl_ptr + gs.l0, gs.l1 - gs.l0, h->d_rowglob, h->d_rmbase);
        else
          k_small_solve_chain<T, FACTO, 1><<<1, SmChain<T>::WARPS * 32, sm_chain_smem, h->stream>>>(h->S, M, xs, ys, rs, cs, nrhs, h->d_slv_cblk,
                                                                                               h->d_slv_lvl_ptr + gs.l0, gs.l1 - gs.l0, h->d_rowglob, h->d_rmbase);
        ++launches;
      }
    }
  }
  if (tr) {
    k_rhs_transpose<T><<<dim3((n + 31) / 32, (nrhs + 31) / 32), dim3(32, 8), 0, h->stream>>>(xs, x, n, nrhs, ldx, 0);
    ++launches;
  }
  CK(cudaGetLastError());
  h->last_launches = launches;
  return PB200_SUCCESS;
}

template <class T>
static int solve_t(pb200_handle_t *h, void *x, int64_t ldx, int nrhs) {
  switch (h->facto) {
    case PB200_FACT_LLT: return solve_tf<T, F_LLT>(h, (T *)x, ldx, nrhs);
    case PB200_FACT_LDLT: return solve_tf<T, F_LDLT>(h, (T *)x, ldx, nrhs);
    case PB200_FACT_LU: return solve_tf<T, F_LU>(h, (T *)x, ldx, nrhs);
    case PB200_FACT_LDLH:
      return ST<T>::is_complex ? solve_tf<T, F_LDLH>(h, (T *)x, ldx, nrhs) : solve_tf<T, F_LDLT>(h, (T *)x, ldx, nrhs);
  }
  return fail(PB200_ERR_BADARG, "bad factotype");
}
static int solve_dispatch(pb200_handle_t *h, void *x, int64_t ldx, int nrhs) { DISPATCH_T(h, solve_t, h, x, ldx, nrhs) }

// Hermitian-typed internal CSC ('H'): the U^T panels of an LU factorization are filled with the CONJUGATE of the
// transposed values (csc_intern_solve.c:110-116).  pb200_assemble_csc sets it from the device CSC's type.
extern "C" int pb200_set_hermitian(pb200_handle_t *h, int hermitian) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  h->herm = hermitian != 0;
  return PB200_SUCCESS;
}

extern "C" int pb200_set_transpose_solve(pb200_handle_t *h, int transposed) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  h->solve_transposed = transposed != 0;
  return PB200_SUCCESS;
}

extern "C" int pb200_solve_device(pb200_handle_t *h, void *x_dev, int64_t ldx, int64_t nrhs, double *seconds) {
  if (!h || !x_dev) return fail(PB200_ERR_BADARG, "null argument");
  if (!h->factorized) return fail(PB200_ERR_STATE, "not factorized");
  if (ldx < h->n || nrhs <= 0) return fail(PB200_ERR_BADARG, "bad ldx / nrhs");
  CK(cudaSetDevice(h->device));
  { int rc = ensure_gathered(h); if (rc) return rc; }
  {
    size_t yb = (size_t)ldx * nrhs * h->esize;
    if (yb > h->y_bytes) {
      cudaFree(h->d_y); h->d_y = nullptr; h->y_bytes = 0;
      if (cudaMalloc(&h->d_y, yb) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMalloc(work vector) failed");
      h->y_bytes = yb;
    }
    if (!h->inv_ready) { int rc0 = invert_dispatch(h, h->stream); if (rc0) return rc0; }
  }
  CK(cudaEventRecord(h->ev0, h->stream));
  int rc = solve_dispatch(h, x_dev, ldx, (int)nrhs);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  if (seconds) *seconds = ms * 1e-3;
  if (h->dag_ok && h->h_dag_err && *h->h_dag_err) {
    *h->h_dag_err = 0;
    return fail(PB200_ERR_CUDA, "up_down: a contribution counter never reached its target (dependency table / device fault)");
  }
  return PB200_SUCCESS;
}

extern "C" int pb200_solve(pb200_handle_t *h, void *x, int64_t ldx, int64_t nrhs, double *seconds) {
  if (!h || !x) return fail(PB200_ERR_BADARG, "null argument");
  if (!h->factorized) return fail(PB200_ERR_STATE, "not factorized");
  if (ldx < h->n || nrhs <= 0) return fail(PB200_ERR_BADARG, "bad ldx / nrhs");
  CK(cudaSetDevice(h->device));
  size_t bytes = (size_t)ldx * nrhs * h->esize;
  if (bytes > h->x_bytes) {
    cudaFree(h->d_x); h->d_x = nullptr; h->x_bytes = 0;
    CK(cudaMalloc(&h->d_x, bytes));
    h->x_bytes = bytes;
  }
  CK(cudaMemcpyAsync(h->d_x, x, bytes, cudaMemcpyHostToDevice, h->stream));
  int rc = pb200_solve_device(h, h->d_x, ldx, nrhs, seconds);
  if (rc) return rc;
  CK(cudaMemcpyAsync(x, h->d_x, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}

// ------------------------------------------------------------------ refinement back end (kernels_raff.cuh)
// Vectors are managed allocations: kernels work on them in HBM, and the handful of scalars / pointer tables the
// reference's drivers allocate through the same call stay host-dereferenceable.
extern "C" int pb200_vec_alloc(pb200_handle_t *h, void **p, int64_t bytes) {
  if (!h || !p || bytes < 0) return fail(PB200_ERR_BADARG, "bad argument");
  CK(cudaSetDevice(h->device));
  *p = nullptr;
  const size_t nb = (size_t)std::max<int64_t>(bytes, 8);
  if (cudaMallocManaged(p, nb) != cudaSuccess) return fail(PB200_ERR_NOMEM, "cudaMallocManaged(refinement vector) failed");
  if (nb >= (size_t)(1 << 16)) {   // a vector: make HBM its home before the first kernel touches it
    CK(cudaMemsetAsync(*p, 0, nb, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  } else {
    memset(*p, 0, nb);
  }
  return PB200_SUCCESS;
}
extern "C" int pb200_vec_free(pb200_handle_t *h, void *p) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  CK(cudaSetDevice(h->device));
  if (p) CK(cudaFree(p));
  return PB200_SUCCESS;
}
extern "C" int pb200_vec_set(pb200_handle_t *h, void *dst, const void *src_host, int64_t nelem) {
  if (!h || !dst || (!src_host && nelem > 0)) return fail(PB200_ERR_BADARG, "null argument");
  CK(cudaSetDevice(h->device));
  if (src_host) CK(cudaMemcpyAsync(dst, src_host, (size_t)nelem * h->esize, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
extern "C" int pb200_vec_zero(pb200_handle_t *h, void *dst, int64_t nelem) {
  if (!h || !dst) return fail(PB200_ERR_BADARG, "null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(dst, 0, (size_t)nelem * h->esize, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
extern "C" int pb200_vec_get(pb200_handle_t *h, void *dst_host, const void *src, int64_t nelem) {
  if (!h || !dst_host || !src) return fail(PB200_ERR_BADARG, "null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst_host, src, (size_t)nelem * h->esize, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
extern "C" int pb200_vec_copy(pb200_handle_t *h, void *dst, const void *src, int64_t nelem) {
  if (!h || !dst || !src) return fail(PB200_ERR_BADARG, "null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(dst, src, (size_t)nelem * h->esize, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
template <class T>
static int vec_axpy_t(pb200_handle_t *h, const void *alpha, const void *x, void *y, int64_t n) {
  T a; memcpy(&a, alpha, sizeof(T));   // the caller's scalar is a C99 complex: 8-byte aligned, not alignof(T)
  k_raff_axpy<T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, a, (const T *)x, (T *)y);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
template <class T>
static int vec_scal_t(pb200_handle_t *h, const void *alpha, void *x, int64_t n) {
  T a; memcpy(&a, alpha, sizeof(T));
  k_raff_scal<T><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(n, a, (T *)x);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
template <class T>
static int vec_dot_t(pb200_handle_t *h, int conjy, const void *x, const void *y, int64_t n, void *result) {
  if (!h->d_raff_partial) {
    CK(cudaMalloc(&h->d_raff_partial, PB200_RAFF_BLOCKS * 16));
    CK(cudaHostAlloc(&h->h_raff_partial, PB200_RAFF_BLOCKS * 16, cudaHostAllocDefault));
  }
  k_raff_dot<T><<<PB200_RAFF_BLOCKS, 256, 0, h->stream>>>(n, (const T *)x, (const T *)y, conjy, (T *)h->d_raff_partial);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(h->h_raff_partial, h->d_raff_partial, PB200_RAFF_BLOCKS * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  T s = ST<T>::zero();
  const T *p = (const T *)h->h_raff_partial;
  for (int i = 0; i < PB200_RAFF_BLOCKS; ++i) s = s + p[i];
  memcpy(result, &s, sizeof(T));
  return PB200_SUCCESS;
}
template <class T>
static int csc_ax_t(pb200_handle_t *h, int type, int trans, const void *b, const void *x, void *r) {
  // row c of A read off column c of the symmetric-pattern CSC (see kernels_raff.cuh); A^T x is the column itself
  const T *vals = (const T *)h->d_vals; int conjv = 0;
  if (!trans) {
    if (type == 'H') conjv = 1;
    else if (type == 'U') {
      if (!h->d_tvals) return fail(PB200_ERR_STATE, "unsymmetric matrix without the transposed values in HBM");
      vals = (const T *)h->d_tvals;
    }
  }
  k_raff_spmv<T><<<(unsigned)((h->n + 127) / 128), 128, 0, h->stream>>>(h->n, h->d_colptr, h->d_rows, vals, conjv, (const T *)x, (const T *)b, (T *)r);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return PB200_SUCCESS;
}
static int vec_check(pb200_handle_t *h, const void *a, const void *b) {
  if (!h || !a || !b) return fail(PB200_ERR_BADARG, "null argument");
  CK(cudaSetDevice(h->device));
  return PB200_SUCCESS;
}
extern "C" int pb200_vec_axpy(pb200_handle_t *h, const void *alpha, const void *x, void *y, int64_t nelem) {
  { int rc = vec_check(h, x, y); if (rc) return rc; }
  if (!alpha) return fail(PB200_ERR_BADARG, "null argument");
  DISPATCH_T(h, vec_axpy_t, h, alpha, x, y, nelem)
}
extern "C" int pb200_vec_scal(pb200_handle_t *h, const void *alpha, void *x, int64_t nelem) {
  { int rc = vec_check(h, alpha, x); if (rc) return rc; }
  DISPATCH_T(h, vec_scal_t, h, alpha, x, nelem)
}
extern "C" int pb200_vec_dot(pb200_handle_t *h, int conj_y, const void *x, const void *y, int64_t nelem, void *result) {
  { int rc = vec_check(h, x, y); if (rc) return rc; }
  if (!result) return fail(PB200_ERR_BADARG, "null argument");
  DISPATCH_T(h, vec_dot_t, h, conj_y, x, y, nelem, result)
}
extern "C" int pb200_csc_ax(pb200_handle_t *h, char type, int trans, const void *b, const void *x, void *r) {
  { int rc = vec_check(h, x, r); if (rc) return rc; }
  if (!h->d_colptr || !h->assembled) return fail(PB200_ERR_STATE, "no internal CSC resident in HBM");
  if (type != 'S' && type != 'H' && type != 'U') return fail(PB200_ERR_BADARG, "matrix type must be S, H or U");
  DISPATCH_T(h, csc_ax_t, h, (int)type, trans, b, x, r)
}
extern "C" int pb200_precond(pb200_handle_t *h, const void *s, void *d) {
  { int rc = vec_check(h, s, d); if (rc) return rc; }
  if (d != s) CK(cudaMemcpyAsync(d, s, (size_t)h->n * h->esize, cudaMemcpyDeviceToDevice, h->stream));
  return pb200_solve_device(h, d, h->n, 1, nullptr);
}

// ------------------------------------------------------------------ factor slabs <-> host
extern "C" int pb200_get_coeftab(pb200_handle_t *h, void *L, void *U) {
  if (!h || !L) return fail(PB200_ERR_BADARG, "null argument");
  CK(cudaSetDevice(h->device));
  if (h->factorized) { int rc = ensure_gathered(h); if (rc) return rc; }
  size_t slab = (size_t)h->coefnbr * h->esize;
  CK(cudaMemcpy(L, h->dL, slab, cudaMemcpyDeviceToHost));
  if (U) {
    if (!h->dU) return fail(PB200_ERR_STATE, "no U slab (not an LU factorization)");
    CK(cudaMemcpy(U, h->dU, slab, cudaMemcpyDeviceToHost));
  }
  return PB200_SUCCESS;
}

// one panel: coeftab[c] / ucoeftab[c] as the reference lays them out (stride x width, column-major)
extern "C" int pb200_get_cblk(pb200_handle_t *h, int64_t c, void *L, void *U) {
  if (!h || !L) return fail(PB200_ERR_BADARG, "null argument");
  if (c < 0 || c >= h->cblknbr) return fail(PB200_ERR_BADARG, "cblk index out of range");
  CK(cudaSetDevice(h->device));
  if (h->factorized) { int rc = ensure_gathered(h); if (rc) return rc; }
  const size_t off = (size_t)h->h_poff[c] * h->esize, bytes = (size_t)(h->h_poff[c + 1] - h->h_poff[c]) * h->esize;
  CK(cudaMemcpy(L, (const char *)h->dL + off, bytes, cudaMemcpyDeviceToHost));
  if (U) {
    if (!h->dU) return fail(PB200_ERR_STATE, "no U slab (not an LU factorization)");
    CK(cudaMemcpy(U, (const char *)h->dU + off, bytes, cudaMemcpyDeviceToHost));
  }
  return PB200_SUCCESS;
}

extern "C" int pb200_set_coeftab(pb200_handle_t *h, const void *L, const void *U) {
  if (!h || !L) return fail(PB200_ERR_BADARG, "null argument");
  if (h->dU && !U) return fail(PB200_ERR_BADARG, "LU needs both slabs");
  CK(cudaSetDevice(h->device));
  size_t slab = (size_t)h->coefnbr * h->esize;
  CK(cudaMemcpy(h->dL, L, slab, cudaMemcpyHostToDevice));
  if (h->dU) CK(cudaMemcpy(h->dU, U, slab, cudaMemcpyHostToDevice));
  h->assembled = true; h->factorized = false; h->inv_ready = false; h->gathered = true;
  return PB200_SUCCESS;
}

// ------------------------------------------------------------------ multi-GPU plumbing (C ABI)
extern "C" int pb200_dist_plan(const pb200_solver_t *s, int factotype, int nranks, int32_t *owner, uint32_t *contrib, double *load) {
  if (!s || !owner || nranks < 1 || nranks > PB200_MAXRANKS) return fail(PB200_ERR_BADARG, "bad argument");
  const int64_t C = s->cblknbr, B = s->bloknbr;
  std::vector<int> fblok(C + 1), fcblk(B), width(C), stride(C), nrow(B), coefind(B);
  for (int64_t c = 0; c < C; ++c) { fblok[c] = (int)s->bloknum[c]; width[c] = (int)(s->lcolnum[c] - s->fcolnum[c] + 1); stride[c] = (int)s->stride[c]; }
  fblok[C] = (int)s->bloknum[C];
  for (int64_t b = 0; b < B; ++b) { fcblk[b] = (int)s->cblknum[b]; nrow[b] = (int)(s->lrownum[b] - s->frownum[b] + 1); coefind[b] = (int)s->coefind[b]; }
  DistPlan P = dist_plan(C, fblok.data(), fcblk.data(), width.data(), stride.data(), nrow.data(), coefind.data(), nranks,
                         factotype == PB200_FACT_LU);
  for (int64_t c = 0; c < C; ++c) { owner[c] = P.owner[c]; if (contrib) contrib[c] = P.contrib[c]; }
  if (load) for (int p = 0; p < nranks; ++p) load[p] = P.load[p];
  return PB200_SUCCESS;
}

extern "C" int pb200_ipc_size(void) { return (int)(4 * sizeof(cudaIpcMemHandle_t)); }

extern "C" int pb200_ipc_export(pb200_handle_t *h, void *buf) {
  if (!h || !buf) return fail(PB200_ERR_BADARG, "null argument");
  if (h->nranks == 1) return fail(PB200_ERR_STATE, "not a multi-GPU handle");
  CK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t *o = (cudaIpcMemHandle_t *)buf;
  memset(o, 0, 4 * sizeof(cudaIpcMemHandle_t));
  CK(cudaIpcGetMemHandle(&o[0], h->dL));
  if (h->dU) CK(cudaIpcGetMemHandle(&o[1], h->dU));
  CK(cudaIpcGetMemHandle(&o[2], h->d_flags));
  if (h->dW) CK(cudaIpcGetMemHandle(&o[3], h->dW));
  return PB200_SUCCESS;
}

extern "C" int pb200_ipc_attach(pb200_handle_t *h, const void *all_handles) {
  if (!h || !all_handles) return fail(PB200_ERR_BADARG, "null argument");
  if (h->nranks == 1) return fail(PB200_ERR_STATE, "not a multi-GPU handle");
  if (h->attached) return PB200_SUCCESS;
  CK(cudaSetDevice(h->device));
  const cudaIpcMemHandle_t *in = (const cudaIpcMemHandle_t *)all_handles;
  h->peers.rank = h->rank; h->peers.nranks = h->nranks;
  for (int p = 0; p < h->nranks; ++p) {
    if (p == h->rank) { h->peers.L[p] = h->dL; h->peers.U[p] = h->dU; h->peers.W[p] = h->dW; h->peers.flags[p] = h->d_flags; continue; }
    void *q = nullptr;
    CK(cudaIpcOpenMemHandle(&q, in[4 * p + 0], cudaIpcMemLazyEnablePeerAccess)); h->ipc_opened.push_back(q); h->peers.L[p] = q;
    h->peers.U[p] = nullptr; h->peers.W[p] = nullptr;
    if (h->dU) { CK(cudaIpcOpenMemHandle(&q, in[4 * p + 1], cudaIpcMemLazyEnablePeerAccess)); h->ipc_opened.push_back(q); h->peers.U[p] = q; }
    CK(cudaIpcOpenMemHandle(&q, in[4 * p + 2], cudaIpcMemLazyEnablePeerAccess)); h->ipc_opened.push_back(q);
    h->peers.flags[p] = (unsigned int *)q;
    if (h->dW) { CK(cudaIpcOpenMemHandle(&q, in[4 * p + 3], cudaIpcMemLazyEnablePeerAccess)); h->ipc_opened.push_back(q); h->peers.W[p] = q; }
  }
  h->attached = true;
  return PB200_SUCCESS;
}

// All `n` handles of one distributed factorization live in THIS process (one host process driving the GPUs of the
// box, e.g. pastix() with iparm[IPARM_CUDA_NBR] = n): peer access is enabled pairwise and the peers' slabs / flags
// are addressed directly — no IPC blobs.  hs[r] must have been created with rank r of n.  The blocking calls
// (pb200_reassemble, pb200_factorize, ...) are collective: the caller drives every handle from its own host thread.
extern "C" int pb200_attach_local(pb200_handle_t **hs, int n) {
  if (!hs || n < 2 || n > PB200_MAXRANKS) return fail(PB200_ERR_BADARG, "bad handle group");
  for (int r = 0; r < n; ++r) {
    if (!hs[r] || hs[r]->nranks != n || hs[r]->rank != r) return fail(PB200_ERR_BADARG, "handle r must be rank r of n");
    if (hs[r]->attached) return fail(PB200_ERR_STATE, "handle already attached");
    for (int q = 0; q < r; ++q)
      if (hs[q]->device == hs[r]->device) return fail(PB200_ERR_BADARG, "two ranks on one device");
  }
  for (int r = 0; r < n; ++r) {
    CK(cudaSetDevice(hs[r]->device));
    for (int q = 0; q < n; ++q) {
      if (q == r) continue;
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, hs[r]->device, hs[q]->device));
      if (!can) return fail(PB200_ERR_CUDA, "no peer access between the devices of the group");
      cudaError_t e = cudaDeviceEnablePeerAccess(hs[q]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
      else if (e != cudaSuccess) return fail(PB200_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
    }
  }
  for (int r = 0; r < n; ++r) {
    pb200_handle_t *h = hs[r];
    h->peers.rank = r; h->peers.nranks = n;
    for (int q = 0; q < n; ++q) { h->peers.L[q] = hs[q]->dL; h->peers.U[q] = hs[q]->dU; h->peers.W[q] = hs[q]->dW; h->peers.flags[q] = hs[q]->d_flags; }
    h->attached = true; h->local_group = true;
  }
  return PB200_SUCCESS;
}

// destroy the handles of a pb200_attach_local group (all of them, in one call: no peer may outlive the others)
extern "C" int pb200_destroy_group(pb200_handle_t **hs, int n) {
  if (!hs) return PB200_SUCCESS;
  for (int r = 0; r < n; ++r)
    if (hs[r]) { cudaSetDevice(hs[r]->device); cudaDeviceSynchronize(); }
  for (int r = 0; r < n; ++r)
    if (hs[r]) { hs[r]->local_group = true; pb200_destroy(hs[r]); hs[r] = nullptr; }
  return PB200_SUCCESS;
}

extern "C" int pb200_dist_barrier(pb200_handle_t *h) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  CK(cudaSetDevice(h->device));
  return dist_barrier(h);
}

// set by tests that upload already-factored panels (solve-only parity)
extern "C" int pb200_mark_factorized(pb200_handle_t *h) {
  if (!h) return fail(PB200_ERR_BADARG, "null handle");
  h->factorized = true;
  return PB200_SUCCESS;
}
