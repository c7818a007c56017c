// kernels_solve_dag3.cuh — persistent up_down sweeps for ONE right-hand side, third generation: independent workers.
//
// Reference (src/sopalin/src): up_down_smp updo.c:114-1664; ticket / counter protocol of kernels_solve_dag.cuh
// (UPDOWN_CTRBCNT updo.c:631-793, flagtab updo_sendrecv.c:496-639).
//
// What the time stamps of the second generation showed (tools/dag_trace.py, profiles/README.md round 2): a ticket is a
// CHAIN OF L2 ROUND TRIPS — ticket counter, 64-byte record, dependency flag, the 128 values of x_J, reductions, fence,
// signal: 0.4-0.6 us each under load — wrapped around 0.1-0.4 us of arithmetic.  Eight warps of a CTA working on one
// tile wait at barriers for the one warp that walks the chain; prefetching tiles (cp.async rings) fills shared
// memory but not the HBM pipe, because tickets are retired at 1 per ~4 us and CTA.  The sweeps are bound by the number
// of independent chains in flight, not by bytes in flight per chain.
//
// So here nothing is CTA-wide.  One CTA per SM, no __syncthreads after the prologue:
//   * warps 4.. are eight (four for complex double) independent TILE WORKERS.  A worker takes T tickets from its own counter (the next ticket
//     number and record are always on their way while the current ticket is worked on), streams the ticket's
//     32-row sub-tiles through its private double buffer in chunks of 32 rows x 32 columns (cp.async, the next chunk
//     in flight while the current one is multiplied), lane = panel row in the down step (one RED per row and
//     sub-tile), lane = columns {lane, lane+32, ..} in the up step (sums kept in registers for the whole ticket),
//     then fences and signals — warp-level synchronisation only.  1184 chains on the device instead of 296 or 444.
//   * warps 0..3 are the DIAGONAL TEAM: D tickets from a second counter, the packed inverted triangle of a sub-panel
//     (up to 66 KB) in a slot of their own, the product as ten 32 x 32 block tasks over four warps with independent
//     accumulators, a 128-thread named barrier.
// Both lists are in ascending (level, round) order and every worker processes what it took in order: the lowest
// unfinished ticket of either list is always at the head of a running worker, so the sweeps cannot deadlock and need
// no residency assumption between CTAs.
// Several right-hand sides keep the first-generation kernels (a panel tile is reused across right-hand sides there).
#pragma once
#include "kernels_solve_dag2.cuh"

namespace pb200 {


struct Dag3Args {
  const DagTick *ticksD, *ticksT;   // forward order
  int GD, GT;
  const int *tgt;
  unsigned *arrived, *ready, *done, *cnt;   // [nsp]
  unsigned *ticket;                 // [4]: down D, down T, up D, up T
  unsigned *err;
  const int *rowglob;
  unsigned long long *trace;        // optional [2][GD + GT][8]
  unsigned long long *xpub;         // [n][sizeof(T) / 4] {32-bit word of x, 32-bit epoch}: x published with the flag in the data
  unsigned epoch_down, epoch_up;    // tags of this solve's two sweeps (never 0)
};

template <class T> struct Dag3Cfg {
  static constexpr int NB = SlvCfg<T>::NB;
  static constexpr int CC = 32;                                       // columns per chunk
  static constexpr int WORKERS = sizeof(T) >= 16 ? 4 : 8;             // tile-worker warps per CTA
  static constexpr int NT = 128 + 32 * WORKERS;
  static constexpr int LDT = 33;
  static constexpr int HALF = CC * LDT;                               // elements of one chunk buffer
  static constexpr int DSLOT = (NB * (NB + 1)) / 2 + NB;              // packed triangle + LDLt diagonal
  // per worker: two chunk buffers, x_J [NB], x[rows] [32], record
  static constexpr size_t worker_bytes = ((((size_t)2 * HALF + NB + 32) * sizeof(T) + sizeof(DagTick)) + 63) / 64 * 64;
  static constexpr size_t team_bytes = ((((size_t)DSLOT + NB + 10 * 32) * sizeof(T) + sizeof(DagTick) + 16) + 63) / 64 * 64;
  static constexpr size_t bytes = team_bytes + WORKERS * worker_bytes;
};

__device__ __forceinline__ void dag3_team_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---- flag-in-data publication of solved unknowns.  A D ticket used to store x_J, fence, and bump a flag; the T tickets
// waiting for it polled the flag (acquire) and only then loaded x_J: three L2 round trips on the D -> T hop of the
// dependency chain.  Here every 32-bit word of a published value travels in ONE 8-byte store together with the
// sweep's epoch ({word, epoch}: 8-byte accesses are single-copy atomic), so the consumer's load of the value IS its
// poll — no fence, no flag, one round trip — and a value is valid exactly when all its words carry the epoch
// (the LL protocol of NCCL, applied to a solution vector).
template <class T> struct PubCfg { static constexpr int W = (int)sizeof(T) / 4; };
template <class T>
__device__ __forceinline__ void pub_store(unsigned long long *xpub, size_t idx, T v, unsigned epoch) {
  constexpr int W = PubCfg<T>::W;
  unsigned w[W];
  memcpy(w, &v, sizeof(T));
  unsigned long long *p = xpub + idx * W;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    const unsigned long long u = ((unsigned long long)epoch << 32) | w[k];
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p + k), "l"(u) : "memory");
  }
}
// one look: true (and v filled) when every word of element idx carries `epoch`
template <class T>
__device__ __forceinline__ bool pub_try_load(const unsigned long long *xpub, size_t idx, unsigned epoch, T &v) {
  constexpr int W = PubCfg<T>::W;
  unsigned w[W];
  bool ok = true;
  const unsigned long long *p = xpub + idx * W;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    unsigned long long u;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(u) : "l"(p + k) : "memory");
    ok = ok && (unsigned)(u >> 32) == epoch;
    w[k] = (unsigned)u;
  }
  memcpy(&v, w, sizeof(T));
  return ok;
}
// all lanes of a warp spin until each has its element (lanes with !want pass); raises *err after the watchdog time
template <class T>
__device__ __forceinline__ T pub_wait_load(const unsigned long long *xpub, size_t idx, bool want, unsigned epoch, unsigned *err) {
  T v = ST<T>::zero();
  bool have = !want;
  if (!have) have = pub_try_load<T>(xpub, idx, epoch, v);
  if (__all_sync(0xffffffffu, have)) return v;
  const long long t0 = clock64();
  for (unsigned it = 1;; ++it) {
    if (!have) have = pub_try_load<T>(xpub, idx, epoch, v);
    if (__all_sync(0xffffffffu, have)) return v;
    if ((it & 63u) == 0) {
      bool stop = *reinterpret_cast<volatile unsigned *>(err) != 0;
      if (!stop && clock64() - t0 > PB200_DAG_TIMEOUT) { atomicExch(err, 1u); stop = true; }
      if (__any_sync(0xffffffffu, stop)) return v;
    }
  }
}

// 32 x 32 block task of a triangle product with four independent accumulators (see dag2_tri_task for the layouts).
// Branch-free: every lane walks all 32 summation indices; where the term does not exist (above the diagonal of a
// diagonal block, past the sub-panel's width) the load goes to a valid element anyway and a select drops it.
template <class T, int DIR, bool CONJ>
__device__ __forceinline__ T dag3_tri_task(const T *buf, const T *xs, int nb, int rb, int cb, int lane) {
  const T zero = ST<T>::zero();
  T a0 = zero, a1 = zero, a2 = zero, a3 = zero;
  if (32 * rb >= nb) return zero;
  const int j0 = 32 * (DIR == 0 ? cb : rb), j1 = min(j0 + 32, nb);
  const int p = min(32 * (DIR == 0 ? rb : cb) + lane, nb - 1);      // lanes past the width redo the last output (never stored)
  auto term = [&](int j, T &acc) {
    const int jj = min(j, j1 - 1);
    T a = DIR == 0 ? buf[jj * nb - ((jj * (jj + 1)) >> 1) + p] : buf[((jj * (jj + 1)) >> 1) + p];
    if (CONJ) a = ST<T>::conj(a);
    const bool ok = j < j1 && (DIR == 0 ? j <= p : j >= p);
    fma_acc(acc, ok ? a : zero, xs[jj]);
  };
#pragma unroll
  for (int j = 0; j < 32; j += 4) { term(j0 + j, a0); term(j0 + j + 1, a1); term(j0 + j + 2, a2); term(j0 + j + 3, a3); }
  return (a0 + a1) + (a2 + a3);
}

template <class T, int FACTO, int DIR>
__global__ void __launch_bounds__(Dag3Cfg<T>::NT, 1)
k_dag3(const T *__restrict__ M, const T *__restrict__ inv, T *x, T *y, Dag3Args A) {
  using C = Dag3Cfg<T>;
  constexpr int NB = C::NB, CC = C::CC, LDT = C::LDT, HALF = C::HALF, ROWS = 32;
  constexpr int CPL = CC / 32, NBL = NB / 32;     // columns per lane and chunk, column sums per lane (up step)
  constexpr bool LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  constexpr bool CONJ = (DIR == 1 && FACTO == F_LDLH);
  extern __shared__ __align__(16) unsigned char dag3_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T zero = ST<T>::zero();

  if (warp < 4) {
    // ================================================================== diagonal team
    T *buf = reinterpret_cast<T *>(dag3_smem);
    T *dxs = buf + C::DSLOT;
    T *dparts = dxs + NB;
    DagTick *drec = reinterpret_cast<DagTick *>(dparts + 10 * 32);
    int *dg = reinterpret_cast<int *>(drec + 1);
    unsigned nx_g = (unsigned)A.GD; int nx_w = 0;
    auto pretake = [&]() {
      unsigned g = 0;
      if (lane == 0) g = atomicAdd(A.ticket + 2 * DIR, 1u);
      g = __shfl_sync(0xffffffffu, g, 0);
      nx_g = min(g, (unsigned)A.GD);
      if (g < (unsigned)A.GD && lane < 16) nx_w = __ldg(reinterpret_cast<const int *>(A.ticksD + (DIR ? A.GD - 1 - (int)g : (int)g)) + lane);
    };
    if (warp == 0) pretake();
    for (;;) {
      unsigned long long t_take = 0, t_dep = 0, t_b1 = 0, t_b2 = 0, t_fin = 0;
      if (warp == 0) {
        if (nx_g < (unsigned)A.GD && lane < 16) reinterpret_cast<int *>(drec)[lane] = nx_w;
        if (lane == 0) { *dg = (int)nx_g; if (A.trace) t_take = dag_gtime(); }
      }
      dag3_team_bar();
      const int g = *dg;
      if (g >= A.GD) return;
      const DagTick tk = *drec;
      if (warp == 0) pretake();
      const int nb = tk.nb, tri = (nb * (nb + 1)) >> 1;
      {
        const T *Inv = inv + tk.src;
        for (int col = warp; col < nb; col += 4) {
          const T *src = Inv + (size_t)col * nb;
          if (DIR == 0) {
            T *dst = buf + (col * nb - ((col * (col - 1)) >> 1)) - col;      // packed by columns
            for (int r = col + lane; r < nb; r += 32) dag_cp_async<sizeof(T)>(dst + r, src + r);
          } else {
            for (int r = col + lane; r < nb; r += 32) dag_cp_async<sizeof(T)>(buf + ((r * (r + 1)) >> 1) + col, src + r);   // packed by rows
          }
        }
        if (LDL && DIR == 0 && tid < nb) dag_cp_async<sizeof(T)>(buf + tri + tid, M + tk.aux + (size_t)tid * (tk.ld + 1));
        dag_cp_commit();
      }
      if (warp == 0) {
        if (lane == 0) {
          if (DIR == 0) dag_wait_ge(A.arrived + tk.sp, (unsigned)tk.pad0, A.err);
          else dag_wait_ge(A.cnt + tk.sp, (unsigned)tk.nsib, A.err);
          if (A.trace) t_dep = dag_gtime();
        }
        __syncwarp();
        const T *vin = DIR == 0 ? x : y;
        for (int j = lane; j < NB; j += 32) dxs[j] = j < nb ? ld_cg(&vin[tk.xcol + j]) : zero;
      }
      dag_cp_wait<0>();
      dag3_team_bar();
      if (A.trace && tid == 0) t_b1 = dag_gtime();
      // ten block tasks (every one costs the same: branch-free), 3 + 3 + 2 + 2 over the four warps:
      //   w0: (3,0) (2,0) (0,0)   w1: (3,1) (2,1) (1,1)   w2: (3,2) (1,0)   w3: (3,3) (2,2)
      // part ids: (3,c) -> c, (2,c) -> 4 + c, (1,0) -> 7, (0,0) -> 8, (1,1) -> 9
      T dreg[4];
#pragma unroll
      for (int ob = 0; ob < 4; ++ob) dreg[ob] = (LDL && DIR == 0 && warp == 0 && 32 * ob + lane < nb) ? buf[tri + 32 * ob + lane] : ST<T>::from_real(1.0);
      if (warp == 0) {
        dparts[0 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 3, 0, lane); dparts[4 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 2, 0, lane);
        dparts[8 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 0, 0, lane);
      }
      if (warp == 1) {
        dparts[1 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 3, 1, lane); dparts[5 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 2, 1, lane);
        dparts[9 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 1, 1, lane);
      }
      if (warp == 2) { dparts[2 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 3, 2, lane); dparts[7 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 1, 0, lane); }
      if (warp == 3) { dparts[3 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 3, 3, lane); dparts[6 * 32 + lane] = dag3_tri_task<T, DIR, CONJ>(buf, dxs, nb, 2, 2, lane); }
      dag3_team_bar();
      if (warp == 0) {
        if (A.trace && lane == 0) t_b2 = dag_gtime();
#pragma unroll
        for (int ob = 0; ob < 4; ++ob) {
          const int p = 32 * ob + lane;
          if (p >= nb) continue;
          // down (row blocks): 3 <- 0,1,2,3;  2 <- 4,5,6;  1 <- 7,9;  0 <- 8     up (column blocks): 0 <- 0,4,7,8;  1 <- 1,5,9;  2 <- 2,6;  3 <- 3
          T v;
          if (DIR == 0) {
            if (ob == 3) v = (dparts[lane] + dparts[32 + lane]) + (dparts[64 + lane] + dparts[96 + lane]);
            else if (ob == 2) v = (dparts[4 * 32 + lane] + dparts[5 * 32 + lane]) + dparts[6 * 32 + lane];
            else if (ob == 1) v = dparts[7 * 32 + lane] + dparts[9 * 32 + lane];
            else v = dparts[8 * 32 + lane];
          } else {
            if (ob == 0) v = (dparts[lane] + dparts[4 * 32 + lane]) + (dparts[7 * 32 + lane] + dparts[8 * 32 + lane]);
            else if (ob == 1) v = (dparts[32 + lane] + dparts[5 * 32 + lane]) + dparts[9 * 32 + lane];
            else if (ob == 2) v = dparts[64 + lane] + dparts[6 * 32 + lane];
            else v = dparts[96 + lane];
          }
          x[tk.xcol + p] = v;
          // what the T tickets of the sweep wait for, value and flag in one store each
          pub_store<T>(A.xpub, (size_t)(tk.xcol + p), v, DIR == 0 ? A.epoch_down : A.epoch_up);
          // LDLt / LDLh: the diagonal step x_k /= D_kk folded into the write-back (updo.c:948-984)
          if (DIR == 0) y[tk.xcol + p] = LDL ? v / dreg[ob] : v;
        }
        if (A.trace && lane == 0) t_fin = dag_gtime();
        if (lane == 0) {
          if (A.trace) {
            unsigned long long *tr = A.trace + ((size_t)DIR * (A.GD + A.GT) + (size_t)g) * 8;
            unsigned sm;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            tr[0] = t_take; tr[1] = t_dep; tr[2] = dag_gtime(); tr[3] = ((unsigned long long)sm << 32) | 1u | (1u << 16);
            tr[4] = t_take; tr[5] = t_b1; tr[6] = t_b2; tr[7] = t_fin;
          }
        }
      }
    }
  }

  // ==================================================================== tile worker (one warp)
  const int wk = warp - 4;
  unsigned char *wbase = dag3_smem + C::team_bytes + (size_t)wk * C::worker_bytes;
  T *hs = reinterpret_cast<T *>(wbase);            // [2][HALF]
  T *xs = hs + 2 * HALF;                           // down: x_J [NB]
  T *xr = xs + NB;                                 // up: x[rows of the sub-tile] [32]
  int *rec = reinterpret_cast<int *>(xr + 32);     // the ticket record
  unsigned nx_g = (unsigned)A.GT; int nx_w = 0;
  auto pretake = [&]() {
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(A.ticket + 2 * DIR + 1, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    nx_g = min(g, (unsigned)A.GT);
    if (g < (unsigned)A.GT && lane < 16) nx_w = __ldg(reinterpret_cast<const int *>(A.ticksT + (DIR ? A.GT - 1 - (int)g : (int)g)) + lane);
  };
  pretake();
  for (;;) {
    if (nx_g >= (unsigned)A.GT) return;
    const int g = (int)nx_g;
    __syncwarp();
    if (lane < 16) rec[lane] = nx_w;
    __syncwarp();
    const DagTick tk = *reinterpret_cast<const DagTick *>(rec);
    unsigned long long t_take = 0, t_dep = 0, t_b1 = 0, t_fin = 0;
    if (A.trace && lane == 0) t_take = dag_gtime();
    pretake();
    const int nb = tk.nb, nsub = (tk.mrows + ROWS - 1) / ROWS, ncc = (nb + CC - 1) / CC, total = nsub * ncc;
    const T *P0 = M + tk.src;
    // sub-panels owning rows of the ticket: counters to bump (down) / flags to wait for (up)
    int my_tgt = lane < tk.ntgt ? __ldg(A.tgt + tk.tptr + lane) : 0;
    auto issue = [&](int q) {
      const int k = q / ncc, c = q - k * ncc;
      const int rk = k * ROWS, mr = min(ROWS, tk.mrows - rk);
      const int j0 = c * CC, j1 = min(nb, j0 + CC);
      T *dst = hs + (q & 1) * HALF + lane;
      const T *src = P0 + (size_t)j0 * tk.ld + rk + lane;
      if (lane < mr)
        for (int j = j0; j < j1; ++j, dst += LDT, src += tk.ld) dag_cp_async<sizeof(T)>(dst, src);
      dag_cp_commit();
    };
    auto grow_of = [&](int k) -> int {
      const int r = k * ROWS + lane;
      if (r >= tk.mrows) return 0;
      return r < tk.wrem ? tk.grow0 + r : __ldg(A.rowglob + tk.aux + r);
    };
    issue(0);
    int grow = grow_of(0);
    // ---- dependency = the data itself: x_J of the sub-panel (down), x[rows] per sub-tile (up, below)
    if (DIR == 0) {
      for (int j = lane; j < NB; j += 32) {          // NB / 32 uniform rounds
        const T v = pub_wait_load<T>(A.xpub, (size_t)(tk.xcol + j), j < nb, A.epoch_down, A.err);
        xs[j] = j < nb ? v : zero;
      }
      if (A.trace && lane == 0) t_dep = dag_gtime();
    }
    T acc0 = zero, acc1 = zero, acc2 = zero, acc3 = zero;     // down: row sums of the current sub-tile (4 chains)
    T bacc[NBL];                                              // up: column sums of the ticket
#pragma unroll
    for (int t = 0; t < NBL; ++t) bacc[t] = zero;
    int grow_next = 0;
    for (int q = 0; q < total; ++q) {
      const int k = q / ncc, c = q - k * ncc;
      const int mr = min(ROWS, tk.mrows - k * ROWS);
      const int j0 = c * CC, ncol = min(nb, j0 + CC) - j0;
      if (q + 1 < total) { issue(q + 1); dag_cp_wait<1>(); } else dag_cp_wait<0>();
      if (c == 0) {
        if (k > 0) grow = grow_next;
        if (k + 1 < nsub) grow_next = grow_of(k + 1);
        if (DIR == 1) {
          const T v = pub_wait_load<T>(A.xpub, (size_t)grow, lane < mr, A.epoch_up, A.err);
          xr[lane] = lane < mr ? v : zero;
          if (A.trace && lane == 0 && q == 0) t_dep = dag_gtime();
        }
      }
      __syncwarp();
      if (A.trace && lane == 0 && q == 0) t_b1 = dag_gtime();
      const T *h = hs + (q & 1) * HALF;
      if (DIR == 0) {
        // lane = panel row: sum over the columns of the chunk
        const T *a = h + lane;
        const T *xv = xs + j0;
        int j = 0;
        for (; j + 4 <= ncol; j += 4) {
          fma_acc(acc0, a[(j + 0) * LDT], xv[j + 0]);
          fma_acc(acc1, a[(j + 1) * LDT], xv[j + 1]);
          fma_acc(acc2, a[(j + 2) * LDT], xv[j + 2]);
          fma_acc(acc3, a[(j + 3) * LDT], xv[j + 3]);
        }
        for (; j < ncol; ++j) fma_acc(acc0, a[j * LDT], xv[j]);
        if (c == ncc - 1) {
          if (lane < mr) atomic_sub(&x[grow], (acc0 + acc1) + (acc2 + acc3));
          acc0 = acc1 = acc2 = acc3 = zero;
        }
      } else {
        // lane = columns lane, lane + 32, .. of the chunk: sum over the rows of the sub-tile
#pragma unroll
        for (int t = 0; t < NBL; ++t) {
          if (t / CPL != c) continue;
          const int pl = (t % CPL) * 32 + lane;
          if (pl >= ncol) continue;
          const T *a = h + pl * LDT;
          T s0 = zero, s1 = zero;
          int i = 0;
          for (; i + 2 <= mr; i += 2) {
            T v0 = a[i], v1 = a[i + 1];
            if (CONJ) { v0 = ST<T>::conj(v0); v1 = ST<T>::conj(v1); }
            fma_acc(s0, v0, xr[i]); fma_acc(s1, v1, xr[i + 1]);
          }
          if (i < mr) { T v0 = a[i]; if (CONJ) v0 = ST<T>::conj(v0); fma_acc(s0, v0, xr[i]); }
          bacc[t] += s0 + s1;
        }
      }
      __syncwarp();            // every lane is done with this chunk buffer (and x[rows]) before they are refilled
    }
    if (DIR == 1) {
#pragma unroll
      for (int t = 0; t < NBL; ++t) if (t * 32 + lane < nb) atomic_sub(&y[tk.xcol + t * 32 + lane], bacc[t]);
    }
    if (A.trace && lane == 0) t_fin = dag_gtime();
    __threadfence();
    if (DIR == 0) {
      if (lane < min(tk.ntgt, 32)) atomicAdd(A.arrived + my_tgt, 1u);
      for (int q = 32 + lane; q < tk.ntgt; q += 32) atomicAdd(A.arrived + __ldg(A.tgt + tk.tptr + q), 1u);
    } else {
      if (lane == 0) atomicAdd(A.cnt + tk.sp, 1u);
    }
    if (A.trace && lane == 0) {
      unsigned long long *tr = A.trace + ((size_t)DIR * (A.GD + A.GT) + (size_t)A.GD + (size_t)g) * 8;
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      tr[0] = t_take; tr[1] = t_dep; tr[2] = dag_gtime(); tr[3] = ((unsigned long long)sm << 32) | ((unsigned)nsub << 16);
      tr[4] = t_take; tr[5] = t_b1; tr[6] = t_b1; tr[7] = t_fin;
    }
  }
}

}  // namespace pb200
