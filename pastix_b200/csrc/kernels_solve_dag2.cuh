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
    return ((size_t)PB200_DAG2_DEPTH * SLOT + work_elems(NR)) * sizeof(T) + PB200_DAG2_DEPTH * (sizeof(DagTick) + 2 * 32 * sizeof(int) + sizeof(int)) + 16;
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

// DIR 0: down + diagonal step over coeftab;  DIR 1: up step over coeftab (ucoeftab for LU), tickets in reverse
template <class T, int FACTO, int DIR, int NR>
__global__ void __launch_bounds__(PB200_DAG2_NT, 2)
k_dag2(const T *__restrict__ M, const T *__restrict__ inv, T *x, T *y, int64_t ldx, int nrhs, DagArgs A) {
  using C = Dag2Cfg<T>;
  constexpr int NB = C::NB, SLOT = C::SLOT, ROWS = PB200_DAG2_ROWS, LDT = PB200_DAG2_LDT, NT = PB200_DAG2_NT, DEPTH = PB200_DAG2_DEPTH;
  constexpr int PARTS = PB200_DAG2_PARTS;
  constexpr bool LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  constexpr bool CONJ = (DIR == 1 && FACTO == F_LDLH);
  constexpr bool OVERLAY = (NR > 1);
  extern __shared__ __align__(16) unsigned char dag2_smem[];
  T *slots = reinterpret_cast<T *>(dag2_smem);
  T *xs = slots + (size_t)DEPTH * SLOT;                       // input vector [NR][NB] (tiles of the up step: [NR][32])
  T *parts = OVERLAY ? xs : xs + NB;                          // [task][NR][32] or [H][NR][NB]
  DagTick *ent = reinterpret_cast<DagTick *>(xs + C::work_elems(NR));
  int *e_grow = reinterpret_cast<int *>(ent + DEPTH);         // [DEPTH][32] global row of each tile row
  int *e_tgt = e_grow + DEPTH * 32;                           // [DEPTH][32] sub-panels owning the rows of the tile
  int *e_g = e_tgt + DEPTH * 32;                              // [DEPTH] ticket number (>= G: none left)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const T zero = ST<T>::zero();

  // queue of taken tickets: entry (hd + k) % DEPTH, k < nq, occupies slots [off_k, off_k + use_k)
  int hd = 0, nq = 0, used = 0, wr = 0;
  int off0 = 0, off1 = 0, off2 = 0, use0 = 0, use1 = 0, use2 = 0;
  bool pending = false, exhausted = false;

  for (;;) {
    // ------------------------------------------------------------ top up: take tickets, issue their copies
    for (;;) {
      const int e = (hd + nq) % DEPTH;
      if (!pending) {
        if (exhausted || nq == DEPTH) break;
        if (warp == 0) {
          unsigned g = 0;
          if (lane == 0) g = atomicAdd(A.ticket + DIR, 1u);
          g = __shfl_sync(0xffffffffu, g, 0);
          if (g < (unsigned)A.G && lane < 16) {
            const int gi = DIR ? A.G - 1 - (int)g : (int)g;
            reinterpret_cast<int *>(ent + e)[lane] = __ldg(reinterpret_cast<const int *>(A.ticks + gi) + lane);
          }
          if (lane == 0) {
            e_g[e] = (int)min(g, (unsigned)A.G);
            if (A.trace && g < (unsigned)A.G) A.trace[(size_t)(DIR * (size_t)A.G + g) * 4 + 0] = dag_gtime();
          }
        }
        __syncthreads();
        if (e_g[e] >= A.G) { exhausted = true; break; }
        pending = true;
      }
      const DagTick tk = ent[e];
      const int nb = tk.nb;
      const bool isD = tk.mrows < 0;
      const int tri = (nb * (nb + 1)) >> 1;
      const bool wide = isD && (tri + (LDL && DIR == 0 ? nb : 0)) > SLOT;
      if (wide ? nq > 0 : used + 1 > DEPTH) break;              // no room yet: keep it pending
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
        const T *P0 = M + tk.src;
        const int mrows = tk.mrows;
        if (lane < mrows)
          for (int j = warp; j < nb; j += NT / 32) dag_cp_async<sizeof(T)>(buf + j * LDT + lane, P0 + (size_t)j * tk.ld + lane);
        if (warp == 0) {
          if (lane < mrows) {
            if (lane < tk.wrem) e_grow[e * 32 + lane] = tk.grow0 + lane;
            else dag_cp_async<4>(e_grow + e * 32 + lane, A.rowglob + tk.aux + lane);
          }
          if (lane < tk.ntgt) dag_cp_async<4>(e_tgt + e * 32 + lane, A.tgt + tk.tptr + lane);
        }
      }
      dag_cp_commit();
      if (nq == 0) { off0 = off; use0 = use; } else if (nq == 1) { off1 = off; use1 = use; } else { off2 = off; use2 = use; }
      ++nq; used += use; pending = false;
    }
    if (nq == 0) {
      if (exhausted) return;
      continue;
    }
    // ------------------------------------------------------------ process the head of the queue
    const int e = hd;
    const DagTick tk = ent[e];
    const int nb = tk.nb;
    T *buf = slots + (size_t)off0 * SLOT;
    if (nq == 1) dag_cp_wait<0>(); else if (nq == 2) dag_cp_wait<1>(); else dag_cp_wait<2>();
    const bool isD = tk.mrows < 0;
    unsigned long long t_dep = 0;

    if (isD) {
      // ---------------- D(J).  down: x_J <- inv(L_JJ) x_J, y_J <- x_J (/ D_JJ);  up: x_J <- inv(W_JJ)^T y_J
      const T *vin = DIR == 0 ? x : y;
      const int tri = (nb * (nb + 1)) >> 1;
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        if (warp == 0) {
          if (r0 == 0) {
            if (lane == 0) {
              if (DIR == 0) dag_wait_ge(A.arrived + tk.sp, (unsigned)tk.pad0, A.err);
              else dag_wait_ge(A.cnt + tk.sp, (unsigned)tk.nsib, A.err);
              if (A.trace) t_dep = dag_gtime();
            }
            __syncwarp();
          }
          for (int k = lane; k < NR * NB; k += 32) {
            const int rr = k / NB, j = k % NB;
            xs[k] = (rr < nr && j < nb) ? ld_cg(&vin[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : zero;
          }
        }
        __syncthreads();
        // block tasks (row block, column block) of the lower triangle, 32 x 32 each; warp w: first task -> part w,
        // second task (the two warps that hold two triangles) -> parts 8, 9
        //   w: 0 (3,0)  1 (3,1)  2 (3,2)  3 (3,3)+(0,0)  4 (2,0)  5 (2,1)  6 (2,2)+(1,1)  7 (1,0)
        T acc0[NR], acc1[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc0[rr] = acc1[rr] = zero;
        dag2_tri_task<T, NR, DIR, CONJ>(acc0, buf, xs, nb, warp < 4 ? 3 : (warp < 7 ? 2 : 1), warp < 4 ? warp : (warp < 7 ? warp - 4 : 0), lane);
        if (warp == 3) dag2_tri_task<T, NR, DIR, CONJ>(acc1, buf, xs, nb, 0, 0, lane);
        if (warp == 6) dag2_tri_task<T, NR, DIR, CONJ>(acc1, buf, xs, nb, 1, 1, lane);
        // the diagonal of LDLt sits behind the triangle in the slot: read it before the slot can be handed on
        T dreg[4];
#pragma unroll
        for (int ob = 0; ob < 4; ++ob) dreg[ob] = (LDL && DIR == 0 && warp == 0 && 32 * ob + lane < nb) ? buf[tri + 32 * ob + lane] : ST<T>::from_real(1.0);
        if (OVERLAY) __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          parts[(warp * NR + rr) * 32 + lane] = acc0[rr];
          if (warp == 3) parts[(8 * NR + rr) * 32 + lane] = acc1[rr];
          if (warp == 6) parts[(9 * NR + rr) * 32 + lane] = acc1[rr];
        }
        __syncthreads();
        if (warp == 0) {
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
        }
        if (OVERLAY || r0 + NR < nrhs) __syncthreads();
      }
      if (warp == 0) {
        __threadfence();
        if (lane == 0) atomicAdd((DIR == 0 ? A.ready : A.done) + tk.sp, 1u);
      }
    } else if (DIR == 0) {
      // ---------------- T(J,t), down: x[rows] -= P[rows, J] x_J
      const int mrows = tk.mrows;
      const int cw = (nb + NT / 32 - 1) / (NT / 32);
      const int j0 = warp * cw, j1 = min(nb, j0 + cw);
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        if (warp == 0) {
          if (r0 == 0) {
            if (lane == 0) {
              dag_wait_ge(A.ready + tk.sp, 1u, A.err);
              if (A.trace) t_dep = dag_gtime();
            }
            __syncwarp();
          }
          for (int k = lane; k < NR * NB; k += 32) {
            const int rr = k / NB, j = k % NB;
            xs[k] = (rr < nr && j < nb) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : zero;
          }
        }
        __syncthreads();
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
        __syncthreads();
        if (warp == 0 && lane < mrows) {
          const int grow = e_grow[e * 32 + lane];
          for (int rr = 0; rr < nr; ++rr) {
            T v = parts[rr * 32 + lane];
#pragma unroll
            for (int q = 1; q < NT / 32; ++q) v += parts[(q * NR + rr) * 32 + lane];
            atomic_sub(&x[(size_t)(r0 + rr) * ldx + grow], v);
          }
        }
        if (OVERLAY || r0 + NR < nrhs) __syncthreads();
      }
      if (warp == 0) {
        __threadfence();
        if (lane < tk.ntgt) atomicAdd(A.arrived + e_tgt[e * 32 + lane], 1u);
      }
    } else {
      // ---------------- T(J,t), up: y_J -= P[rows, J]^T x[rows]
      constexpr int H = NT / NB, RG = ROWS / H;
      const int mrows = tk.mrows;
      const int p = tid % NB, hh = tid / NB;
      const int i0 = hh * RG, i1 = min(mrows, i0 + RG);
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        if (warp == 0) {
          if (r0 == 0) {
            if (lane < tk.ntgt) dag_wait_ge(A.done + e_tgt[e * 32 + lane], 1u, A.err);
            __syncwarp();
            if (A.trace && lane == 0) t_dep = dag_gtime();
          }
          const int grow = lane < mrows ? e_grow[e * 32 + lane] : 0;
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) xs[rr * 32 + lane] = (lane < mrows && rr < nr) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + grow]) : zero;
        }
        __syncthreads();
        T acc[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc[rr] = zero;
        if (p < nb) {
          const T *a = buf + p * LDT;
#pragma unroll 4
          for (int i = i0; i < i1; ++i) {
            T av = a[i];
            if (CONJ) av = ST<T>::conj(av);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], av, xs[rr * 32 + i]);
          }
        }
        if (OVERLAY) __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) parts[(hh * NR + rr) * NB + p] = acc[rr];
        __syncthreads();
        if (warp == 0)
          for (int c = lane; c < nb; c += 32)
            for (int rr = 0; rr < nr; ++rr) {
              T v = parts[rr * NB + c];
#pragma unroll
              for (int q = 1; q < H; ++q) v += parts[(q * NR + rr) * NB + c];
              atomic_sub(&y[(size_t)(r0 + rr) * ldx + tk.xcol + c], v);
            }
        if (OVERLAY || r0 + NR < nrhs) __syncthreads();
      }
      if (warp == 0) {
        __threadfence();
        if (lane == 0) atomicAdd(A.cnt + tk.sp, 1u);
      }
    }
    if (A.trace && tid == 0) {
      unsigned long long *tr = A.trace + (size_t)(DIR * (size_t)A.G + (size_t)e_g[e]) * 4;
      unsigned sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      tr[1] = t_dep; tr[2] = dag_gtime(); tr[3] = ((unsigned long long)sm << 32) | (unsigned)(isD ? 1 : 0) | ((unsigned)nq << 8);
    }
    // pop
    used -= use0; off0 = off1; use0 = use1; off1 = off2; use1 = use2;
    hd = (hd + 1) % DEPTH; --nq;
  }
}

}  // namespace pb200
