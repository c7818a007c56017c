// kernels_solve_dag.cuh — up_down as ONE persistent launch per sweep, ordered by contribution counters.
//
// Reference (src/sopalin/src): up_down_smp updo.c:114-1664.  The reference orders the down step with a
// per-cblk contribution counter (UPDOWN_CTRBCNT, decremented under mutex_task, updo.c:631-793) and the up
// step with a per-cblk "solved" flag that consumers wait on (flagtab, updo_sendrecv.c:496-639).  The
// level-scheduled kernels of kernels_solve.cuh replace both by kernel boundaries, which costs one launch +
// one drained GPU per (level, round): 250 launches of a few microseconds for C2, ten times the HBM time.
//
// Here the same two mechanisms come back as device-side counters.  The work is cut into tickets, in
// ascending (level, round) order; per sub-panel J of a cblk (at most SlvCfg::NB columns):
//   D(J)    the diagonal product   down: x_J <- inv(L_JJ) x_J, y_J <- x_J (/ D_JJ)   up: x_J <- inv(W_JJ)^T y_J
//   T(J,t)  64 panel rows below J  down: x[rows] -= P[rows,J] x_J                    up: y_J -= P[rows,J]^T x[rows]
// CTAs take tickets with one atomicAdd (the up sweep walks them in reverse), so a ticket only ever waits
// on tickets that running CTAs already hold: no residency assumption, no deadlock.
//   down: D(J) waits arrived[J] == need[J] (every tile owning rows in J has subtracted its product = CTRBCNT),
//         then releases ready[J]; T(J,t) waits ready[J] and afterwards bumps arrived[] of the sub-panels
//         its rows live in.
//   up:   T(J,t) waits done[] of those sub-panels (= flagtab) and bumps cnt[J]; D(J) waits cnt[J] == tiles(J)
//         and releases done[J].
// Everything that does not depend on the right-hand side — the panel tile or the packed inverted
// triangle — is requested with cp.async BEFORE the wait, three CTAs per SM keep ~190 KB of panel data in
// flight per SM, and the dependent path of a level is "flag -> 128 values out of L2 -> one shared-memory
// product -> L2 reductions -> flag".  With several right-hand sides the tile stays in shared memory for
// all of them (a panel is read once per sweep whatever nrhs is).  x and y are only touched with L2 (.cg)
// accesses and reductions.
#pragma once
#include "kernels_solve.cuh"

namespace pb200 {

#define PB200_DAG_NT 256
#define PB200_DAG_ROWS 64                 // panel rows per T ticket
#define PB200_DAG_LDT (PB200_DAG_ROWS + 1) // shared-memory leading dimension of a tile (both products conflict-free)
#define PB200_DAG_TIMEOUT 20000000000LL   // cycles (~10 s): a dependency that never arrives is an error, not a hang

struct DagTick {
  int64_t src;     // T: slab offset of the tile's first element (first row, column c0);  D: offset of the inverted triangle
  int64_t aux;     // T: rowglob index of the tile's first row;                            D: slab offset of the first diagonal entry of J
  int ld, nb;      // panel leading dimension, columns of the sub-panel
  int mrows;       // T: rows of the tile (1..64); D: -1
  int sp;          // sub-panel
  int xcol;        // global index of the sub-panel's first column
  int grow0;       // T: global row of the tile's first row while it is still inside the diagonal block
  int wrem;        // T: leading tile rows that are inside the diagonal block (<= 0: none)
  int nsib;        // D: T tickets of the sub-panel
  int tptr, ntgt;  // T: sub-panels owning the rows of the tile: tgt[tptr .. tptr+ntgt)
  int pad0, pad1;  // pad0: D: contributions x_J waits for in the down step (= need[sp], carried in the record for k_dag2)
};
static_assert(sizeof(DagTick) == 64, "DagTick is one 64-byte record");

struct DagArgs {
  const DagTick *ticks;     // [G] forward ticket order
  const int *tgt;
  const unsigned *need;     // [nsp] T tickets contributing to x_J of each sub-panel
  unsigned *arrived;        // [nsp] down: contributions received
  unsigned *ready;          // [nsp] down: x_J substituted
  unsigned *done;           // [nsp] up: x_J final
  unsigned *cnt;            // [nsp] up: T tickets finished
  unsigned *ticket;         // [2] down / up
  unsigned *err;
  const int *rowglob;
  int G, nbs;               // tickets; widest sub-panel
  unsigned long long *trace; // optional [2][G][4] time stamps per ticket (PB200_DAG_TRACE, k_dag2 only), else null
};

template <int BYTES>
__device__ __forceinline__ void dag_cp_async(void *smem_dst, const void *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(gmem_src), "n"(BYTES));
}
__device__ __forceinline__ void dag_cp_commit_wait() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ unsigned dag_ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// spin until *p >= want; gives up (and raises *err) after PB200_DAG_TIMEOUT cycles or when another CTA did
__device__ __forceinline__ void dag_wait_ge(const unsigned *p, unsigned want, unsigned *err) {
  if (dag_ld_acquire(p) >= want) return;
  const long long t0 = clock64();
  for (unsigned it = 1;; ++it) {
    if (dag_ld_acquire(p) >= want) return;
    if ((it & 63u) == 0) {
      if (*reinterpret_cast<volatile unsigned *>(err)) return;
      if (clock64() - t0 > PB200_DAG_TIMEOUT) { atomicExch(err, 1u); return; }
    }
  }
}
// release: everything this CTA did before the preceding barrier is visible before the counter moves
__device__ __forceinline__ void dag_signal_add(unsigned *p) { __threadfence(); atomicAdd(p, 1u); }

template <class T>
struct DagSmem {
  static constexpr int NR = PB200_SLV_NR;
  static size_t buf_elems(int nbs) { return std::max((size_t)nbs * PB200_DAG_LDT, (size_t)nbs * (nbs + 1) / 2); }
  static size_t bytes(int nbs) { return (buf_elems(nbs) + (size_t)PB200_DAG_NT * NR) * sizeof(T); }
};

__device__ __forceinline__ int dag_pow2_ge(int v) { return v <= 1 ? 1 : 1 << (32 - __clz(v - 1)); }

// shared-memory product skeleton: thread (p, q) = (tid % P, tid / P), P a power of two >= the number of outputs,
// sums the terms k = q, q+Q, ... < K of output p; the Q partial sums are combined through `pv`.
// On return (after the barrier) pv[rr * P + p] ... is NOT what holds the result: the caller reads
// dag_combined() for the threads with q == 0.
template <class T, int NR>
__device__ __forceinline__ T dag_combined(const T *pv, int P, int Q, int p, int rr) {
  T v = pv[rr * P + p];
  for (int q = 1; q < Q; ++q) v += pv[(q * NR + rr) * P + p];
  return v;
}

// ---- forward (down step + diagonal step)
template <class T, int FACTO>
__global__ void __launch_bounds__(PB200_DAG_NT, 3)
k_fwd_dag(const T *__restrict__ L, const T *__restrict__ inv, T *x, T *y, int64_t ldx, int nrhs, DagArgs A, size_t buf_elems) {
  constexpr int NB = SlvCfg<T>::NB, NR = PB200_SLV_NR, ROWS = PB200_DAG_ROWS, LDT = PB200_DAG_LDT, NT = PB200_DAG_NT;
  constexpr bool LDL = (FACTO == F_LDLT || FACTO == F_LDLH);
  extern __shared__ __align__(16) unsigned char dag_smem[];
  T *buf = reinterpret_cast<T *>(dag_smem);   // T: tile [nb][LDT]; D: packed lower triangle, column j = rows j..nb-1
  T *pv = buf + buf_elems;                    // [NR][NB] input vector, then [Q][NR][P] partial sums
  __shared__ int s_g;
  const int tid = threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_g = (int)atomicAdd(A.ticket, 1u);
    __syncthreads();
    const int g = s_g;
    if (g >= A.G) return;
    const DagTick tk = A.ticks[g];
    const int nb = tk.nb, ld = tk.ld;
    if (tk.mrows < 0) {
      // ---------------- D(J): x_J <- inv(L_JJ) x_J ; y_J <- x_J (/ D_JJ)
      const T *Inv = inv + tk.src;
      for (int e = tid; e < nb * nb; e += NT) {
        const int j = e / nb, i = e - j * nb;          // Inv(i, j), i >= j, column j contiguous
        if (i >= j) dag_cp_async<sizeof(T)>(buf + (j * nb - ((j * (j - 1)) >> 1)) + (i - j), Inv + e);
      }
      const int P = dag_pow2_ge(nb), Q = NT / P, p = tid & (P - 1), q = tid / P;
      T d = ST<T>::from_real(1.0);
      if (LDL && q == 0 && p < nb) d = L[tk.aux + (size_t)p * (ld + 1)];
      if (tid == 0) dag_wait_ge(A.arrived + tk.sp, A.need[tk.sp], A.err);
      dag_cp_commit_wait();
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        __syncthreads();
        for (int e = tid; e < NR * NB; e += NT) {
          const int rr = e / NB, j = e % NB;
          pv[e] = (rr < nr && j < nb) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : ST<T>::zero();
        }
        __syncthreads();
        T acc[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc[rr] = ST<T>::zero();
        if (p < nb)
          for (int j = q; j <= p; j += Q) {
            const T a = buf[(j * nb - ((j * (j - 1)) >> 1)) + (p - j)];
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a, pv[rr * NB + j]);
          }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) pv[(q * NR + rr) * P + p] = acc[rr];
        __syncthreads();
        if (q == 0 && p < nb)
          for (int rr = 0; rr < nr; ++rr) {
            const T v = dag_combined<T, NR>(pv, P, Q, p, rr);
            x[(size_t)(r0 + rr) * ldx + tk.xcol + p] = v;
            // LDLt / LDLh: the diagonal step x_k /= D_kk folded into the write-back (updo.c:948-984)
            y[(size_t)(r0 + rr) * ldx + tk.xcol + p] = LDL ? v / d : v;
          }
      }
      __syncthreads();
      if (tid == 0) dag_signal_add(A.ready + tk.sp);
    } else {
      // ---------------- T(J,t): x[rows] -= P[rows, J] x_J
      const T *P0 = L + tk.src;
      const int mrows = tk.mrows;
      for (int e = tid; e < nb * ROWS; e += NT) {
        const int j = e / ROWS, r = e % ROWS;
        if (r < mrows) dag_cp_async<sizeof(T)>(buf + j * LDT + r, P0 + (size_t)j * ld + r);
      }
      const int P = dag_pow2_ge(mrows), Q = NT / P, p = tid & (P - 1), q = tid / P;
      int grow = 0;
      if (q == 0 && p < mrows) grow = (p < tk.wrem) ? tk.grow0 + p : A.rowglob[tk.aux + p];
      if (tid == 0) dag_wait_ge(A.ready + tk.sp, 1u, A.err);
      dag_cp_commit_wait();
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        __syncthreads();
        for (int e = tid; e < NR * NB; e += NT) {
          const int rr = e / NB, j = e % NB;
          pv[e] = (rr < nr && j < nb) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : ST<T>::zero();
        }
        __syncthreads();
        T acc[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc[rr] = ST<T>::zero();
        if (p < mrows)
          for (int j = q; j < nb; j += Q) {
            const T a = buf[j * LDT + p];
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a, pv[rr * NB + j]);
          }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) pv[(q * NR + rr) * P + p] = acc[rr];
        __syncthreads();
        if (q == 0 && p < mrows)
          for (int rr = 0; rr < nr; ++rr)
            atomic_sub(&x[(size_t)(r0 + rr) * ldx + grow], dag_combined<T, NR>(pv, P, Q, p, rr));
      }
      __syncthreads();
      for (int k = tid; k < tk.ntgt; k += NT) dag_signal_add(A.arrived + A.tgt[tk.tptr + k]);
    }
  }
}

// ---- backward (up step).  M is coeftab (ucoeftab for LU); tickets run in reverse.
template <class T, int FACTO>
__global__ void __launch_bounds__(PB200_DAG_NT, 3)
k_bwd_dag(const T *__restrict__ M, const T *__restrict__ inv, T *x, T *y, int64_t ldx, int nrhs, DagArgs A, size_t buf_elems) {
  constexpr int NB = SlvCfg<T>::NB, NR = PB200_SLV_NR, ROWS = PB200_DAG_ROWS, LDT = PB200_DAG_LDT, NT = PB200_DAG_NT;
  constexpr bool CONJ = (FACTO == F_LDLH);
  extern __shared__ __align__(16) unsigned char dag_smem[];
  T *buf = reinterpret_cast<T *>(dag_smem);   // T: tile [nb][LDT]; D: packed lower triangle by rows, row j = columns 0..j
  T *pv = buf + buf_elems;
  __shared__ int s_g;
  const int tid = threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_g = (int)atomicAdd(A.ticket + 1, 1u);
    __syncthreads();
    if (s_g >= A.G) return;
    const DagTick tk = A.ticks[A.G - 1 - s_g];
    const int nb = tk.nb, ld = tk.ld;
    if (tk.mrows >= 0) {
      // ---------------- T(J,t): y_J -= P[rows, J]^T x[rows]
      const T *P0 = M + tk.src;
      const int mrows = tk.mrows;
      for (int e = tid; e < nb * ROWS; e += NT) {
        const int j = e / ROWS, r = e % ROWS;
        if (r < mrows) dag_cp_async<sizeof(T)>(buf + j * LDT + r, P0 + (size_t)j * ld + r);
      }
      int grow = 0;
      if (tid < mrows) grow = (tid < tk.wrem) ? tk.grow0 + tid : A.rowglob[tk.aux + tid];
      for (int k = tid; k < tk.ntgt; k += NT) dag_wait_ge(A.done + A.tgt[tk.tptr + k], 1u, A.err);
      dag_cp_commit_wait();
      const int P = dag_pow2_ge(nb), Q = NT / P, p = tid & (P - 1), q = tid / P;   // thread (column p, row group q)
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        __syncthreads();
        if (tid < ROWS)
#pragma unroll
          for (int rr = 0; rr < NR; ++rr)
            pv[rr * ROWS + tid] = (tid < mrows && rr < nr) ? ld_cg(&x[(size_t)(r0 + rr) * ldx + grow]) : ST<T>::zero();
        __syncthreads();
        T acc[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc[rr] = ST<T>::zero();
        if (p < nb)
          for (int i = q; i < mrows; i += Q) {
            T a = buf[p * LDT + i];
            if (CONJ) a = ST<T>::conj(a);
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a, pv[rr * ROWS + i]);
          }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) pv[(q * NR + rr) * P + p] = acc[rr];
        __syncthreads();
        if (q == 0 && p < nb)
          for (int rr = 0; rr < nr; ++rr)
            atomic_sub(&y[(size_t)(r0 + rr) * ldx + tk.xcol + p], dag_combined<T, NR>(pv, P, Q, p, rr));
      }
      __syncthreads();
      if (tid == 0) dag_signal_add(A.cnt + tk.sp);
    } else {
      // ---------------- D(J): x_J <- inv(W_JJ)^T y_J once every tile of J is in
      const T *Inv = inv + tk.src;
      for (int e = tid; e < nb * nb; e += NT) {
        const int i = e / nb, j = e - i * nb;          // Inv(j, i), j >= i, stored by rows: row j = columns 0..j
        if (j >= i) dag_cp_async<sizeof(T)>(buf + ((j * (j + 1)) >> 1) + i, Inv + e);
      }
      const int P = dag_pow2_ge(nb), Q = NT / P, p = tid & (P - 1), q = tid / P;
      if (tid == 0) dag_wait_ge(A.cnt + tk.sp, (unsigned)tk.nsib, A.err);
      dag_cp_commit_wait();
      for (int r0 = 0; r0 < nrhs; r0 += NR) {
        const int nr = min(NR, nrhs - r0);
        __syncthreads();
        for (int e = tid; e < NR * NB; e += NT) {
          const int rr = e / NB, j = e % NB;
          pv[e] = (rr < nr && j < nb) ? ld_cg(&y[(size_t)(r0 + rr) * ldx + tk.xcol + j]) : ST<T>::zero();
        }
        __syncthreads();
        T acc[NR];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) acc[rr] = ST<T>::zero();
        if (p < nb)
          for (int j = q; j < nb; j += Q)
            if (j >= p) {
              T a = buf[((j * (j + 1)) >> 1) + p];
              if (CONJ) a = ST<T>::conj(a);
#pragma unroll
              for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a, pv[rr * NB + j]);
            }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) pv[(q * NR + rr) * P + p] = acc[rr];
        __syncthreads();
        if (q == 0 && p < nb)
          for (int rr = 0; rr < nr; ++rr)
            x[(size_t)(r0 + rr) * ldx + tk.xcol + p] = dag_combined<T, NR>(pv, P, Q, p, rr);
      }
      __syncthreads();
      if (tid == 0) dag_signal_add(A.done + tk.sp);
    }
  }
}

}  // namespace pb200
