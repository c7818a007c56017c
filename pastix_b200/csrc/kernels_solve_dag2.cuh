// kernels_solve_dag2.cuh — the persistent up_down sweeps, second generation: every CTA keeps up to three tickets
// in flight.
//
// Reference (src/sopalin/src): up_down_smp updo.c:114-1664 — same ticket / counter protocol as
// kernels_solve_dag.cuh (UPDOWN_CTRBCNT updo.c:631-793, flagtab updo_sendrecv.c:496-639); what changes is how a
// CTA spends its time.  ncu on the first generation (profiles/r02/r02_full_updown_c2_summary.json): DRAM 13 %,
// issue slots 15-21 %, 64 % of the warp samples parked at the barrier behind the polling thread.  A CTA fetched a
// tile, waited for it, multiplied, reduced, fenced, signalled and only then asked for the next ticket: the HBM
// pipe was busy for ~2 of the ~9 us a 64 KB tile took, and every hop of the dependency chain paid a 64-term loop
// with four right-hand-side accumulators whether or not four right-hand sides existed.
//
// Here
//   * a CTA owns a ring of three shared-memory slots (32 panel rows x NB columns each).  It takes tickets ahead
//     of the one it is working on and issues their copies (cp.async: the tile or the packed inverted triangle,
//     the LDLt diagonal, the global row numbers and the list of dependent sub-panels) before it waits on
//     anything: two tiles per CTA, four to six per SM, are always on their way from HBM, and by the time a
//     ticket becomes the head of the queue nothing it needs is outside shared memory except the values other
//     CTAs produce.  Tickets are processed in the order they were taken, so a ticket still only waits on tickets
//     held by running CTAs in front of it: no residency assumption, no deadlock (kernels_solve_dag.cuh).
//   * the dependent path is cut down: one warp polls, loads the nb values of x_J / y_J from L2 and later does the
//     reductions, the fence and the signal itself (no CTA barrier between the last reduction and the signal);
//     the products are templated on the number of right-hand sides carried per pass (1 or NRMAX); the triangle
//     product is dealt to the warps as 32 x 32 blocks (ten block tasks, 32 terms per thread instead of 64);
//     index arithmetic is a running pointer.
//   * an inverted triangle wider than ~90 columns does not fit one slot: it takes slots 0-1 once the queue has
//     drained (the only bubble in the pipeline, and only in front of a diagonal ticket).
#pragma once
#include "kernels_solve_dag.cuh"

namespace pb200 {

#define PB200_DAG2_NT 256
#define PB200_DAG2_ROWS 32
#define PB200_DAG2_LDT 33
#define PB200_DAG2_DEPTH 3
#define PB200_DAG2_PARTS 320     // partial sums per right-hand side: 10 block tasks x 32 (triangle), 8 warps x 32 / H x NB (tiles)

template <class T> struct Dag2Cfg {
  static constexpr int NB = SlvCfg<T>::NB;
  static constexpr int SLOT = NB * PB200_DAG2_LDT;           // elements per slot
  static constexpr int NRMAX = sizeof(T) >= 16 ? 2 : 4;      // right-hand sides per pass when there are several
  // work region: input vector + partial sums; with several right-hand sides the partial sums overlay the vector
  __host__ __device__ static constexpr size_t work_elems(int NR) { return NR == 1 ? (size_t)NB + PB200_DAG2_PARTS : (size_t)PB200_DAG2_PARTS * NR; }
  __host__ __device__ static constexpr size_t bytes(int NR) {
    return ((size_t)PB200_DAG2_DEPTH * SLOT + work_elems(NR)) * sizeof(T) + PB200_DAG2_DEPTH * (sizeof(DagTick) + 2 * 32 * sizeof(int) + sizeof(int) + sizeof(unsigned long long)) + 32;
  }
};

__device__ __forceinline__ unsigned long long dag_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <int N> __device__ __forceinline__ void dag_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void dag_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }


// one 32 x 32 block task of a triangle product.  down (DIR 0): out[p] += sum_{j <= p} A(p, j) v[j], A packed by
// columns (column j = rows j..nb-1);  up (DIR 1): out[p] += sum_{j >= p} A(j, p) v[j], A packed by rows (row j =
// columns 0..j).  Lane = output index inside the block; every lane walks the same j, so the reads of A are
// consecutive words and v[j] is a broadcast.
template <class T, int NR, int DIR, bool CONJ>
__device__ __forceinline__ void dag2_tri_task(T (&acc)[NR], const T *buf, const T *xs, int nb, int rb, int cb, int lane) {
  constexpr int NB = Dag2Cfg<T>::NB;
  if (32 * rb >= nb) return;
  const int j0 = 32 * (DIR == 0 ? cb : rb), j1 = min(j0 + 32, nb);     // summation index: column (down) / row (up)
  const int p = 32 * (DIR == 0 ? rb : cb) + lane;                      // output index: row (down) / column (up)
  if (p >= nb) return;
  if (DIR == 0) {
    int idx = j0 * nb - ((j0 * (j0 - 1)) >> 1) - j0 + p;
    for (int j = j0; j < j1; ++j) {
      if (j <= p) {
        const T a = buf[idx];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a, xs[rr * NB + j]);
      }
      idx += nb - j - 1;
    }
  } else {
    int idx = ((j0 * (j0 + 1)) >> 1) + p;
    for (int j = j0; j < j1; ++j) {
      if (j >= p) {
        T a = buf[idx];
        if (CONJ) a = ST<T>::conj(a);
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a, xs[rr * NB + j]);
      }
      idx += j + 1;
    }
  }
}

// DIR 0: down + diagonal step over coeftab;  DIR 1: up step over coeftab (ucoeftab for LU), tickets in reverse.
// A T ticket covers up to 8 sub-tiles of 32 panel rows (the host picks the size per level: large where a level has
// thousands of tiles, one sub-tile where the level is a link of the dependency chain); the ring holds sub-tiles.
//
// Three warps carry the latency chains, so that they overlap instead of adding up (measured on the first cut of this
// kernel, tools/dag_trace.py: 1.2 us to take a ticket, 0.5-1 us from "dependency met" to "vector in shared memory",
// 0.6-0.7 us for fence + signal — 4.4 us per ticket and CTA when one warp did them in turn):
//   warp 2 (take)  has the NEXT ticket number and its 64-byte record on their way while the current ones are worked on;
//   warp 1 (head)  polls the dependency of the next stage and loads its input vector from L2 — with one right-hand side
//                  it does so while
//   warp 0 (tail)  is still reducing, fencing and signalling the current one.
template <class T, int FACTO, int DIR, int NR>
__global__ void __launch_bounds__(PB200_DAG2_NT, 2)
k_dag2(const T *__restrict__ M, const T *__restrict__ inv, T *x, T *y, int64_t ldx, int nrhs, DagArgs A) {
  using C = Dag2Cfg<T>;
  constexpr int NB = C::NB, SLOT = C::SLOT, ROWS = PB200_DAG2_ROWS, LDT = PB200_DAG2_LDT, NT = PB200_DAG2_NT, DEPTH = PB200_DAG2_DEPTH;
  constexpr bool LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  constexpr bool CONJ = (DIR == 1 && FACTO == F_LDLH);
  constexpr bool OVERLAY = (NR > 1);
  constexpr bool PIPE = (NR == 1);                            // head of the next stage overlaps the tail of this one
  constexpr int WT = 0, WH = 1, WK = 2;
  extern __shared__ __align__(16) unsigned char dag2_smem[];
  T *slots = reinterpret_cast<T *>(dag2_smem);
  T *xs = slots + (size_t)DEPTH * SLOT;                       // input vector [NR][NB] (tiles of the up step: [NR][32])
  T *parts = OVERLAY ? xs : xs + NB;                          // [task][NR][32] or [H][NR][NB]
  DagTick *ent = reinterpret_cast<DagTick *>(xs + C::work_elems(NR));
  int *s_grow = reinterpret_cast<int *>(ent + DEPTH);         // [slot][32] global row of each row of the sub-tile in the slot
  int *e_tgt = s_grow + DEPTH * 32;                           // [entry][32] first 32 sub-panels owning rows of the ticket
  int *e_g = e_tgt + DEPTH * 32;                              // [entry] ticket number (>= G: none left)
  unsigned long long *e_dep = reinterpret_cast<unsigned long long *>(e_g + DEPTH + 1);   // [entry] trace: dependencies met
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T zero = ST<T>::zero();

  // tickets taken: entries (ehd + k) % DEPTH, k < enq; the newest one is being issued sub-tile by sub-tile (iss_*);
  // stages = sub-tiles / triangles in flight, oldest first, in slots [off_k, off_k + use_k); pk = sub-tile of the head
  // ticket the oldest stage holds
  int ehd = 0, enq = 0, iss_e = 0, iss_k = 0, iss_n = 0, pk = 0;
  int snq = 0, used = 0, wr = 0;
  int off0 = 0, off1 = 0, off2 = 0, use0 = 0, use1 = 0, use2 = 0;
  bool exhausted = false, head_done = false;
  T bacc[NR];                                                 // up step: column sums carried across the sub-tiles of a ticket
#pragma unroll
  for (int rr = 0; rr < NR; ++rr) bacc[rr] = zero;
  unsigned long long t_dep = 0, t_cpw = 0, t_b1 = 0, t_b2 = 0, t_fin = 0;

  // warp WK: ticket taken ahead (number + one word of its record per lane)
  unsigned nx_g = (unsigned)A.G; int nx_w = 0;
  auto pretake = [&]() {
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(A.ticket + DIR, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    nx_g = min(g, (unsigned)A.G);
    if (g < (unsigned)A.G && lane < 16) {
      const int gi = DIR ? A.G - 1 - (int)g : (int)g;
      nx_w = __ldg(reinterpret_cast<const int *>(A.ticks + gi) + lane);
    }
  };
  if (warp == WK) pretake();

  // head of a stage (warp WH): dependency of the ticket (first sub-tile), input vector of pass r0 into shared memory
  auto head = [&](const DagTick &tk, int eh, int sh, int kh, int r0) {
    const int nb = tk.nb, nr = min(NR, nrhs - r0);
    if (tk.mrows < 0) {
      if (r0 == 0) {
        if (lane == 0) {
          if (DIR == 0) dag_wait_ge(A.arrived + tk.sp, (unsigned)tk.pad0, A.err);
          else dag_wait_ge(A.cnt + tk.sp, (unsigned)tk.nsib, A.err);
          if (A.trace) e_dep[eh] = dag_gtime();
        }
        __syncwarp();
      }
      const T *vin = DIR == 0 ? x : y;
      for (int k = lane; k < NR * NB; k += 32) {
        const int rr = k / NB, j = k % NB;
        xs[k] = (rr < nr && j < nb) ? ld_cg(&vin[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : zero;
      }
    } else if (DIR == 0) {
      if (kh == 0 && r0 == 0) {
        if (lane == 0) {
          dag_wait_ge(A.ready + tk.sp, 1u, A.err);
          if (A.trace) e_dep[eh] = dag_gtime();
        }
        __syncwarp();
      }
      if (kh == 0 || NR > 1)          // one right-hand side: x_J stays in shared memory for all the sub-tiles of the ticket
        for (int k = lane; k < NR * NB; k += 32) {
          const int rr = k / NB, j = k % NB;
          xs[k] = (rr < nr && j < nb) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : zero;
        }
    } else {
      const int mr = min(ROWS, tk.mrows - kh * ROWS);
      if (kh == 0 && r0 == 0) {
        for (int q = lane; q < tk.ntgt; q += 32) dag_wait_ge(A.done + (q < 32 ? e_tgt[eh * 32 + q] : A.tgt[tk.tptr + q]), 1u, A.err);
        __syncwarp();
        if (A.trace && lane == 0) e_dep[eh] = dag_gtime();
      }
      const int grow = lane < mr ? s_grow[sh * 32 + lane] : 0;
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) xs[rr * 32 + lane] = (lane < mr && rr < nr) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + grow]) : zero;
    }
  };

  for (;;) {
    // ------------------------------------------------------------ fill: take tickets, issue copies
    for (;;) {
      if (iss_k == iss_n) {
        if (exhausted || enq == DEPTH) break;
        const int e = (ehd + enq) % DEPTH;
        if (warp == WK) {
          const bool valid = nx_g < (unsigned)A.G;
          if (valid && lane < 16) reinterpret_cast<int *>(ent + e)[lane] = nx_w;
          if (lane == 0) {
            e_g[e] = (int)nx_g;
            if (A.trace && valid) A.trace[(size_t)(DIR * (size_t)A.G + nx_g) * 8 + 0] = dag_gtime();
          }
          if (valid) pretake();
        }
        __syncthreads();
        if (e_g[e] >= A.G) { exhausted = true; break; }
        ++enq; iss_e = e; iss_k = 0;
        const int mr = ent[e].mrows;
        iss_n = mr < 0 ? 1 : (mr + ROWS - 1) / ROWS;
      }
      const DagTick tk = ent[iss_e];
      const int nb = tk.nb;
      const bool isD = tk.mrows < 0;
      const int tri = (nb * (nb + 1)) >> 1;
      const bool wide = isD && (tri + (LDL && DIR == 0 ? nb : 0)) > SLOT;
      if (wide ? snq > 0 : used + 1 > DEPTH) break;             // no room yet
      int off, use;
      if (wide) { off = 0; use = 2; wr = 2; }
      else { off = wr; use = 1; wr = (wr + 1) % DEPTH; }
      T *buf = slots + (size_t)off * SLOT;
      if (isD) {
        const T *Inv = inv + tk.src;
        // column `col` of the nb x nb inverse (column-major), rows col..nb-1
        for (int col = warp; col < nb; col += NT / 32) {
          const T *src = Inv + (size_t)col * nb;
          if (DIR == 0) {
            T *dst = buf + (col * nb - ((col * (col - 1)) >> 1)) - col;      // packed by columns
            for (int r = col + lane; r < nb; r += 32) dag_cp_async<sizeof(T)>(dst + r, src + r);
          } else {
            for (int r = col + lane; r < nb; r += 32) dag_cp_async<sizeof(T)>(buf + ((r * (r + 1)) >> 1) + col, src + r);   // packed by rows
          }
        }
        if (LDL && DIR == 0 && tid < nb) dag_cp_async<sizeof(T)>(buf + tri + tid, M + tk.aux + (size_t)tid * (tk.ld + 1));
      } else {
        const int rk = iss_k * ROWS, mr = min(ROWS, tk.mrows - rk);
        const T *P0 = M + tk.src + rk;
        if (lane < mr)
          for (int j = warp; j < nb; j += NT / 32) dag_cp_async<sizeof(T)>(buf + j * LDT + lane, P0 + (size_t)j * tk.ld + lane);
        if (warp == WH) {                        // read back by the head warp before any barrier: its own copies
          if (lane < mr) {
            if (rk + lane < tk.wrem) s_grow[off * 32 + lane] = tk.grow0 + rk + lane;
            else dag_cp_async<4>(s_grow + off * 32 + lane, A.rowglob + tk.aux + rk + lane);
          }
          if (iss_k == 0 && lane < tk.ntgt) dag_cp_async<4>(e_tgt + iss_e * 32 + lane, A.tgt + tk.tptr + lane);
        }
      }
      dag_cp_commit();
      if (snq == 0) { off0 = off; use0 = use; } else if (snq == 1) { off1 = off; use1 = use; } else { off2 = off; use2 = use; }
      ++snq; used += use; ++iss_k;
    }
    if (snq == 0) {
      if (exhausted && enq == 0) return;
      continue;
    }
    // ------------------------------------------------------------ process the oldest stage
    const int e = ehd;
    const DagTick tk = ent[e];
    const int my_g = e_g[e];
    const int nb = tk.nb;
    const int sl = off0;
    T *buf = slots + (size_t)sl * SLOT;
    if (snq == 1) dag_cp_wait<0>(); else if (snq == 2) dag_cp_wait<1>(); else dag_cp_wait<2>();
    const bool isD = tk.mrows < 0;
    const int nsub = isD ? 1 : (tk.mrows + ROWS - 1) / ROWS;
    const bool first = pk == 0, last = pk == nsub - 1;
    const int mr = isD ? 0 : min(ROWS, tk.mrows - pk * ROWS);
    if (A.trace && tid == 0 && first) t_cpw = dag_gtime();
    // what the tail needs from tables the next stages may overwrite while it runs: read after the first barrier
    int my_grow = 0, my_tgt = 0;
    T dreg[4];
#pragma unroll
    for (int ob = 0; ob < 4; ++ob) dreg[ob] = ST<T>::from_real(1.0);

    for (int r0 = 0; r0 < nrhs; r0 += NR) {
      const int nr = min(NR, nrhs - r0);
      if (warp == WH && !(r0 == 0 && head_done)) head(tk, e, sl, pk, r0);
      __syncthreads();
      if (A.trace && tid == 0 && r0 == 0 && last) { t_b1 = dag_gtime(); t_dep = e_dep[e]; }
      if (warp == WT && r0 == 0) {
        if (!isD && lane < mr) my_grow = s_grow[sl * 32 + lane];
        if (!isD && lane < min(tk.ntgt, 32)) my_tgt = e_tgt[e * 32 + lane];
      }
      if (isD) {
        // ---------------- D(J).  down: x_J <- inv(L_JJ) x_J, y_J <- x_J (/ D_JJ);  up: x_J <- inv(W_JJ)^T y_J
        // block tasks (row block, column block) of the lower triangle, 32 x 32 each; warp w: first task -> part w,
        // second task (the two warps that hold two triangles) -> parts 8, 9
        //   w: 0 (3,0)  1 (3,1)  2 (3,2)  3 (3,3)+(0,0)  4 (2,0)  5 (2,1)  6 (2,2)+(1,1)  7 (1,0)
        const int tri = (nb * (nb + 1)) >> 1;
        T acc0[NR], acc1[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc0[rr] = acc1[rr] = zero;
        dag2_tri_task<T, NR, DIR, CONJ>(acc0, buf, xs, nb, warp < 4 ? 3 : (warp < 7 ? 2 : 1), warp < 4 ? warp : (warp < 7 ? warp - 4 : 0), lane);
        if (warp == 3) dag2_tri_task<T, NR, DIR, CONJ>(acc1, buf, xs, nb, 0, 0, lane);
        if (warp == 6) dag2_tri_task<T, NR, DIR, CONJ>(acc1, buf, xs, nb, 1, 1, lane);
        // the diagonal of LDLt sits behind the triangle in the slot: read it before the slot can be handed on
        if (LDL && DIR == 0 && warp == WT && r0 == 0) {
#pragma unroll
          for (int ob = 0; ob < 4; ++ob) if (32 * ob + lane < nb) dreg[ob] = buf[tri + 32 * ob + lane];
        }
        if (OVERLAY) __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          parts[(warp * NR + rr) * 32 + lane] = acc0[rr];
          if (warp == 3) parts[(8 * NR + rr) * 32 + lane] = acc1[rr];
          if (warp == 6) parts[(9 * NR + rr) * 32 + lane] = acc1[rr];
        }
      } else if (DIR == 0) {
        // ---------------- T(J,t), down: x[rows] -= P[rows, J] x_J
        const int cw = (nb + NT / 32 - 1) / (NT / 32);
        const int j0 = warp * cw, j1 = min(nb, j0 + cw);
        T acc[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc[rr] = zero;
        {
          const T *a = buf + j0 * LDT + lane;
#pragma unroll 4
          for (int j = j0; j < j1; ++j, a += LDT) {
            const T av = *a;
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], av, xs[rr * NB + j]);
          }
        }
        if (OVERLAY) __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) parts[(warp * NR + rr) * 32 + lane] = acc[rr];
      } else {
        // ---------------- T(J,t), up: y_J -= P[rows, J]^T x[rows]
        constexpr int H = NT / NB, RG = ROWS / H;
        const int p = tid % NB, hh = tid / NB;
        const int i0 = hh * RG, i1 = min(mr, i0 + RG);
        const bool single = nrhs <= NR;        // one pass: the column sums stay in registers until the last sub-tile
        if (!single || first) {
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) bacc[rr] = zero;
        }
        if (p < nb) {
          const T *a = buf + p * LDT;
#pragma unroll 4
          for (int i = i0; i < i1; ++i) {
            T av = a[i];
            if (CONJ) av = ST<T>::conj(av);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(bacc[rr], av, xs[rr * 32 + i]);
          }
        }
        if (!(single && !last)) {
          if (OVERLAY) __syncthreads();
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) parts[(hh * NR + rr) * NB + p] = bacc[rr];
        }
      }
      __syncthreads();       // products done: the slot and the vector may be reused, the partial sums are visible
      if (A.trace && tid == 0 && r0 == 0 && last) t_b2 = dag_gtime();
      // ---- reductions and stores of this pass (tail warp)
      if (warp == WT) {
        if (isD) {
#pragma unroll
          for (int ob = 0; ob < 4; ++ob) {
            const int p = 32 * ob + lane;
            if (p >= nb) continue;
            // parts feeding output block ob.  down (row blocks): 3 <- 0,1,2,3;  2 <- 4,5,6;  1 <- 7,9;  0 <- 8
            //                                 up (column blocks): 0 <- 0,4,7,8;  1 <- 1,5,9;  2 <- 2,6;  3 <- 3
            int l0, l1, l2, l3, n;
            if (DIR == 0) {
              if (ob == 3) { l0 = 0; l1 = 1; l2 = 2; l3 = 3; n = 4; } else if (ob == 2) { l0 = 4; l1 = 5; l2 = 6; l3 = 0; n = 3; }
              else if (ob == 1) { l0 = 7; l1 = 9; l2 = 0; l3 = 0; n = 2; } else { l0 = 8; l1 = 0; l2 = 0; l3 = 0; n = 1; }
            } else {
              if (ob == 0) { l0 = 0; l1 = 4; l2 = 7; l3 = 8; n = 4; } else if (ob == 1) { l0 = 1; l1 = 5; l2 = 9; l3 = 0; n = 3; }
              else if (ob == 2) { l0 = 2; l1 = 6; l2 = 0; l3 = 0; n = 2; } else { l0 = 3; l1 = 0; l2 = 0; l3 = 0; n = 1; }
            }
            const T d = dreg[ob];
            for (int rr = 0; rr < nr; ++rr) {
              T v = parts[(l0 * NR + rr) * 32 + lane];
              if (n > 1) v += parts[(l1 * NR + rr) * 32 + lane];
              if (n > 2) v += parts[(l2 * NR + rr) * 32 + lane];
              if (n > 3) v += parts[(l3 * NR + rr) * 32 + lane];
              x[(size_t)(r0 + rr) * ldx + tk.xcol + p] = v;
              // LDLt / LDLh: the diagonal step x_k /= D_kk folded into the write-back (updo.c:948-984)
              if (DIR == 0) y[(size_t)(r0 + rr) * ldx + tk.xcol + p] = LDL ? v / d : v;
            }
          }
        } else if (DIR == 0) {
          if (lane < mr)
            for (int rr = 0; rr < nr; ++rr) {
              T v = parts[rr * 32 + lane];
#pragma unroll
              for (int q = 1; q < NT / 32; ++q) v += parts[(q * NR + rr) * 32 + lane];
              atomic_sub(&x[(size_t)(r0 + rr) * ldx + my_grow], v);
            }
        } else if (!(nrhs <= NR && !last)) {
          constexpr int H = NT / NB;
          for (int c = lane; c < nb; c += 32)
            for (int rr = 0; rr < nr; ++rr) {
              T v = parts[rr * NB + c];
#pragma unroll
              for (int q = 1; q < H; ++q) v += parts[(q * NR + rr) * NB + c];
              atomic_sub(&y[(size_t)(r0 + rr) * ldx + tk.xcol + c], v);
            }
        }
      }
      if (OVERLAY || r0 + NR < nrhs) __syncthreads();
    }
    // ---- the ticket is complete with its last sub-tile: fence + signal (tail warp) ...
    if (last && warp == WT) {
      if (A.trace && lane == 0) t_fin = dag_gtime();
      __threadfence();
      if (isD) {
        if (lane == 0) atomicAdd((DIR == 0 ? A.ready : A.done) + tk.sp, 1u);
      } else if (DIR == 0) {
        if (lane < min(tk.ntgt, 32)) atomicAdd(A.arrived + my_tgt, 1u);
        for (int q = 32 + lane; q < tk.ntgt; q += 32) atomicAdd(A.arrived + A.tgt[tk.tptr + q], 1u);
      } else {
        if (lane == 0) atomicAdd(A.cnt + tk.sp, 1u);
      }
      if (A.trace && lane == 0) {
        unsigned long long *tr = A.trace + (size_t)(DIR * (size_t)A.G + (size_t)my_g) * 8;
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        tr[1] = t_dep; tr[2] = dag_gtime(); tr[3] = ((unsigned long long)sm << 32) | (unsigned)(isD ? 1 : 0) | ((unsigned)snq << 8) | ((unsigned)nsub << 16);
        tr[4] = t_cpw; tr[5] = t_b1; tr[6] = t_b2; tr[7] = t_fin;
      }
    }
    // ---- ... while the head warp already polls and loads for the next stage in the queue (one right-hand side)
    head_done = false;
    if (PIPE && snq >= 2) {
      const int e2 = last ? (e + 1) % DEPTH : e, k2 = last ? 0 : pk + 1;
      if (warp == WH) {
        if (snq == 2) dag_cp_wait<0>(); else dag_cp_wait<1>();      // its own copies of that stage (row numbers, dependents)
        const DagTick tk2 = ent[e2];
        head(tk2, e2, off1, k2, 0);
      }
      head_done = true;
    }
    // pop the stage; the ticket with its last sub-tile
    used -= use0; off0 = off1; use0 = use1; off1 = off2; use1 = use2; --snq;
    if (last) { pk = 0; ehd = (ehd + 1) % DEPTH; --enq; } else ++pk;
  }
}

}  // namespace pb200
