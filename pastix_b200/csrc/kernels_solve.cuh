// kernels_solve.cuh — level-scheduled up_down, bandwidth-oriented.
//
// Reference (src/sopalin/src): up_down_smp updo.c:114-1664
//   down : x_c <- L_cc^{-1} x_c (TRSV, unit diagonal except LLt)          updo.c:574-596
//          x[rows(b)] -= L_b x_c for every off-diagonal blok               updo.c:631-793
//   diag : x_c[k] /= D_kk (LDLt / LDLh)                                    updo.c:948-984
//   up   : x_c -= L_b^T x[rows(b)] (LU: U^T panel, LDLh: conjugate)        updo_sendrecv.c:496-639
//          x_c <- L_cc^{-T} x_c                                            updo.c:1309-1342
//
// Here every cblk is walked by sub-panels J = [c0,c1) of at most SLV_NB columns.  The sequential
// TRSV on the diagonal block is replaced by a product with the explicitly inverted nb x nb triangle
// (k_tri_inverse, run once after the factorization), so one launch per (level, round) does
//   forward : y_J = inv(L_JJ) x_J  (recomputed by every CTA of the sub-panel, 128 KB out of L2),
//             x[rows below] -= P[rows, J] y_J   (one thread per panel row, coalesced column walk)
//   backward: y_J -= P[rows, J]^T x[rows]       (warp per column, lanes over rows, L2 reductions)
//             last CTA to finish: x_J = inv(W_JJ)^T y_J
// with the LDLt diagonal scaling folded into the forward write-back.  Each panel is read exactly
// once per sweep.  x is the permuted right-hand side, column-major n x nrhs, leading dimension ldx
// (UpDownVector.sm2xtab, blend/src/updown.h:69-72); y is a work vector of the same shape.
#pragma once
#include "scalar.cuh"
#include "symbol.cuh"
#include "kernels_factor.cuh"

namespace pb200 {

#define PB200_SLV_ROWS 128
#define PB200_SLV_NR 4     // right-hand sides carried per pass over a panel
template <class T> struct SlvCfg { static constexpr int NB = 128; };
template <> struct SlvCfg<cdouble> { static constexpr int NB = 64; };

struct SlvTask {
  int cblk, tile0, c0, c1;
  int sp;        // sub-panel id: index into invoff / counters
  int ntiles;
  // static copies of what the kernels need, so that a CTA starts after two dependent loads
  int ld, fcol, w, pad;
  int64_t poff;    // panel offset in the slab
  int64_t rgbase;  // rowglob index of panel row 0 (= rmbase[cblk] - w); rows < w are the cblk's own columns
  int64_t invoff;  // offset of the inverted triangle of this sub-panel
};

// L2 (cache-global) loads: values other CTAs produced with L2 reductions
__device__ __forceinline__ float ld_cg(const float *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ cfloat ld_cg(const cfloat *p) { float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); return cfloat(v.x, v.y); }
__device__ __forceinline__ cdouble ld_cg(const cdouble *p) { double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); return cdouble(v.x, v.y); }

// global row of panel row m of cblk c
__device__ __forceinline__ int panel_row_to_global(const DevSym &S, int c, int m) {
  const int w = S.width[c];
  if (m < w) return S.fcol[c] + m;
  const int b = upper_le(S.coefind, S.fblok[c], S.fblok[c + 1], m);
  return S.frow[b] + (m - S.coefind[b]);
}

// ---- inverse of the lower triangle of every diagonal sub-block: X = W^{-1}, W nb x nb (ld), unit or not.
// One CTA per sub-panel (`order`, when given, lists the sub-panels of one size class so that the launch can be
// sized for them).  Thread j builds column j of X by forward substitution, x_ij = -(sum_{k<i} w_ik x_kj) / w_ii.
// W sits in shared memory as a row-packed triangle (row i = columns 0..i).  X is stored per warp: the 32 columns
// [j0, j0+32) of warp j0/32 as a rectangle of rows j0..nb-1, zero above the diagonal — so the k loop of a row runs
// from the warp's first column without predicates: w_ik is a warp-wide broadcast, x_kj (own column, written by the
// same thread) falls on consecutive words, the loop unrolls into independent FMA chains, and no barrier separates
// the rows.
__device__ __host__ __forceinline__ int tri_row(int i) { return (i * (i + 1)) >> 1; }
// elements of the per-warp X rectangles of an nb-wide triangle, and the offset of warp block wb
__device__ __host__ __forceinline__ int tri_xbase(int wb, int nb) { return 32 * (wb * nb - 16 * wb * (wb - 1)); }
__device__ __host__ __forceinline__ int tri_xelems(int nb) { return tri_xbase((nb + 31) / 32, nb); }

template <class T>
__global__ void __launch_bounds__(128)
k_tri_inverse(const T *__restrict__ M, const SlvTask *__restrict__ tasks, const int *__restrict__ order, T *inv, int unit, int ntri) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the size-classed launches of one inversion pass are independent of each other: a launch with the programmatic
  // attribute (engine.cu, invert_t) may start as soon as the CTAs of the previous class have all started
  asm volatile("griddepcontrol.launch_dependents;");
  T *Ws = reinterpret_cast<T *>(smem_raw);
  T *Xs = Ws + ntri;
  const SlvTask tk = tasks[order ? order[blockIdx.x] : blockIdx.x];
  const int ld = tk.ld, nb = tk.c1 - tk.c0;
  const T *W = M + tk.poff + (size_t)tk.c0 * (ld + 1);
  const T one = ST<T>::from_real(1.0), zero = ST<T>::zero();
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int k = e / nb, i = e - k * nb;
    if (i >= k) Ws[tri_row(i) + k] = W[(size_t)k * ld + i];
  }
  for (int e = threadIdx.x; e < tri_xelems(nb); e += blockDim.x) Xs[e] = zero;
  __syncthreads();
  const int j = threadIdx.x, j0 = j & ~31, lj = j & 31;          // j0: first column of this warp
  if (j0 < nb) {
    T *xc = Xs + tri_xbase(j0 >> 5, nb) + lj;                      // X(j0 + t, j) = xc[32 t]
    if (j < nb) xc[32 * (j - j0)] = unit ? one : one / Ws[tri_row(j) + j];
    for (int i = j0 + 1; i < nb; ++i) {
      const T *wr = Ws + tri_row(i) + j0;                          // W(i, j0 + t) = wr[t]
      const int cnt = i - j0;
      T s0 = zero, s1 = zero, s2 = zero, s3 = zero;
      int t = 0;
      for (; t + 3 < cnt; t += 4) {
        fma_acc(s0, wr[t], xc[32 * t]);
        fma_acc(s1, wr[t + 1], xc[32 * (t + 1)]);
        fma_acc(s2, wr[t + 2], xc[32 * (t + 2)]);
        fma_acc(s3, wr[t + 3], xc[32 * (t + 3)]);
      }
      for (; t < cnt; ++t) fma_acc(s0, wr[t], xc[32 * t]);
      if (i > j && j < nb) {
        const T sm = zero - ((s0 + s1) + (s2 + s3));
        xc[32 * cnt] = unit ? sm : sm / wr[cnt];
      }
    }
  }
  __syncthreads();
  T *out = inv + tk.invoff;
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int jj = e / nb, i = e - jj * nb, b0 = jj & ~31;
    out[e] = (i >= jj) ? Xs[tri_xbase(b0 >> 5, nb) + 32 * (i - b0) + (jj & 31)] : zero;
  }
}

// global row of every off-diagonal panel row, built once (one CTA per cblk)
__global__ void k_build_rowglob(DevSym S, const int64_t *__restrict__ rmbase, int *rowglob) {
  const int k = blockIdx.x;
  const int w = S.width[k], ld = S.stride[k], bf = S.fblok[k] + 1, be = S.fblok[k + 1];
  const int64_t base = rmbase[k];
  for (int m = w + threadIdx.x; m < ld; m += blockDim.x) {
    const int b = upper_le(S.coefind, bf, be, m);
    rowglob[base + m - w] = S.frow[b] + (m - S.coefind[b]);
  }
}

#define PB200_SLV_CG 4                      // column groups per CTA (k_fwd): 128 rows x 4 groups = 512 threads
#define PB200_SLV_NT (PB200_SLV_ROWS * PB200_SLV_CG)
#define PB200_SLV_MLP 16                    // independent loads in flight per thread

// ---- forward: one launch per (level, round); CTA = (sub-panel, 128-row tile of rows [c1, stride)).
// Thread (r, cg) owns panel row r of the tile and every 4th... the cg-th quarter of the nb columns; the
// loads of a batch are issued together (latency of a level = a few memory round trips, not nb of them).
template <class T, int FACTO>
__global__ void __launch_bounds__(PB200_SLV_NT, 1)
k_fwd(const T *__restrict__ L, const T *__restrict__ inv, T *x, T *y, int64_t ldx, int nrhs,
      const SlvTask *__restrict__ tasks, const int *__restrict__ tile2task, const int *__restrict__ rowglob) {
  constexpr int NB = SlvCfg<T>::NB, NR = PB200_SLV_NR, CG = PB200_SLV_CG, ML = sizeof(T) >= 16 ? PB200_SLV_MLP / 2 : PB200_SLV_MLP;
  __shared__ T xs[NR][NB];
  __shared__ T ys[NR][NB];
  __shared__ T part[CG][NR][PB200_SLV_ROWS];
  const SlvTask tk = tasks[tile2task[blockIdx.x]];
  const int ld = tk.ld, fcol = tk.fcol, nb = tk.c1 - tk.c0;
  const int tile = blockIdx.x - tk.tile0;
  const int tid = threadIdx.x, r = tid & (PB200_SLV_ROWS - 1), cg = tid / PB200_SLV_ROWS;
  const T *P = L + tk.poff;
  const T *Inv = inv + tk.invoff;
  const int m = tk.c1 + tile * PB200_SLV_ROWS + r;
  const bool rowok = m < ld;
  const int cgn = (nb + CG - 1) / CG, j_lo = cg * cgn, j_hi = min(nb, j_lo + cgn);
  int grow = 0;
  if (rowok && cg == 0) grow = (m < tk.w) ? fcol + m : rowglob[tk.rgbase + m];
  for (int r0 = 0; r0 < nrhs; r0 += NR) {
    const int nr = min(NR, nrhs - r0);
    __syncthreads();
    for (int e = tid; e < NR * NB; e += PB200_SLV_NT) {
      const int rr = e / NB, j = e % NB;
      xs[rr][j] = (rr < nr && j < nb) ? x[(size_t)(r0 + rr) * ldx + fcol + tk.c0 + j] : ST<T>::zero();
    }
    __syncthreads();
    // y_J = Inv * x_J (the strict upper triangle of Inv is stored as zeros: uniform trip count)
    {
      T acc[NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) acc[rr] = ST<T>::zero();
      if (r < nb)
        for (int j0 = j_lo; j0 < j_hi; j0 += ML) {
          T a[ML];
#pragma unroll
          for (int q = 0; q < ML; ++q) a[q] = (j0 + q < j_hi) ? Inv[(size_t)(j0 + q) * nb + r] : ST<T>::zero();
#pragma unroll
          for (int q = 0; q < ML; ++q)
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a[q], xs[rr][min(j0 + q, NB - 1)]);
        }
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) part[cg][rr][r] = acc[rr];
    }
    __syncthreads();
    if (cg == 0 && r < nb) {
      T d = ST<T>::from_real(1.0);
      if ((FACTO == F_LDLT || FACTO == F_LDLH) && tile == 0) d = P[(size_t)(tk.c0 + r) * (ld + 1)];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        T v = part[0][rr][r];
#pragma unroll
        for (int g = 1; g < CG; ++g) v += part[g][rr][r];
        ys[rr][r] = v;
        // write-back; LDLt/LDLh: the diagonal step x_k /= D_kk folded in (updo.c:948-984)
        if (tile == 0 && rr < nr)
          y[(size_t)(r0 + rr) * ldx + fcol + tk.c0 + r] = (FACTO == F_LDLT || FACTO == F_LDLH) ? v / d : v;
      }
    }
    __syncthreads();
    {
      T acc[NR];
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) acc[rr] = ST<T>::zero();
      if (rowok) {
        const T *col = P + (size_t)tk.c0 * ld + m;
        for (int j0 = j_lo; j0 < j_hi; j0 += ML) {
          T a[ML];
#pragma unroll
          for (int q = 0; q < ML; ++q) a[q] = (j0 + q < j_hi) ? col[(size_t)(j0 + q) * ld] : ST<T>::zero();
#pragma unroll
          for (int q = 0; q < ML; ++q)
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) fma_acc(acc[rr], a[q], ys[rr][min(j0 + q, NB - 1)]);
        }
      }
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) part[cg][rr][r] = acc[rr];
    }
    __syncthreads();
    if (cg == 0 && rowok) {
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        T v = part[0][rr][r];
#pragma unroll
        for (int g = 1; g < CG; ++g) v += part[g][rr][r];
        if (rr < nr) atomic_sub(&x[(size_t)(r0 + rr) * ldx + grow], v);
      }
    }
  }
}

// ---- backward: one launch per (level, round), descending.  M is coeftab (ucoeftab for LU).
#define PB200_BWD_NT 512
template <class T, int FACTO>
__global__ void __launch_bounds__(PB200_BWD_NT, 1)
k_bwd(const T *__restrict__ M, const T *__restrict__ inv, T *x, T *y, int64_t ldx, int nrhs,
      const SlvTask *__restrict__ tasks, const int *__restrict__ tile2task, const int *__restrict__ rowglob,
      unsigned int *counters) {
  constexpr int NB = SlvCfg<T>::NB, NR = PB200_SLV_NR;
  constexpr bool CONJ = (FACTO == F_LDLH);
  constexpr int RI = PB200_SLV_ROWS / 32;   // row chunks per lane
  __shared__ T xr[NR][PB200_SLV_ROWS];
  __shared__ T yj[NR][NB];
  __shared__ int s_last;
  const SlvTask tk = tasks[tile2task[blockIdx.x]];
  const int ld = tk.ld, fcol = tk.fcol, nb = tk.c1 - tk.c0;
  const int tile = blockIdx.x - tk.tile0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = PB200_BWD_NT / 32;
  const T *P = M + tk.poff;
  const int mbase = tk.c1 + tile * PB200_SLV_ROWS;
  const int mrows = max(0, min(PB200_SLV_ROWS, ld - mbase));
  int grow = 0;
  if (tid < mrows) { const int m = mbase + tid; grow = (m < tk.w) ? fcol + m : rowglob[tk.rgbase + m]; }
  if (mrows > 0) {
    for (int r0 = 0; r0 < nrhs; r0 += NR) {
      const int nr = min(NR, nrhs - r0);
      __syncthreads();
      if (tid < PB200_SLV_ROWS)
#pragma unroll
        for (int rr = 0; rr < NR; ++rr)
          xr[rr][tid] = (tid < mrows && rr < nr) ? x[(size_t)(r0 + rr) * ldx + grow] : ST<T>::zero();
      __syncthreads();
      constexpr int JC = 4;   // columns per warp pass: JC * RI loads in flight per lane
      for (int j0 = warp * JC; j0 < nb; j0 += nwarp * JC) {
        T a[JC][RI];
#pragma unroll
        for (int q = 0; q < JC; ++q)
#pragma unroll
          for (int ii = 0; ii < RI; ++ii) {
            const int i = lane + 32 * ii;
            a[q][ii] = (j0 + q < nb && i < mrows) ? P[(size_t)(tk.c0 + j0 + q) * ld + mbase + i] : ST<T>::zero();
            if (CONJ) a[q][ii] = ST<T>::conj(a[q][ii]);
          }
        T acc[JC][NR];
#pragma unroll
        for (int q = 0; q < JC; ++q)
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            acc[q][rr] = ST<T>::zero();
#pragma unroll
            for (int ii = 0; ii < RI; ++ii) fma_acc(acc[q][rr], a[q][ii], xr[rr][lane + 32 * ii]);
          }
#pragma unroll
        for (int q = 0; q < JC; ++q)
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            typedef typename ST<T>::real R;
            R *p = reinterpret_cast<R *>(&acc[q][rr]);
            for (int o = 16; o > 0; o >>= 1) {
              p[0] += __shfl_down_sync(0xffffffffu, p[0], o);
              if (ST<T>::is_complex) p[1] += __shfl_down_sync(0xffffffffu, p[1], o);
            }
            if (lane == 0 && rr < nr && j0 + q < nb) atomic_sub(&y[(size_t)(r0 + rr) * ldx + fcol + tk.c0 + j0 + q], acc[q][rr]);
          }
      }
    }
  }
  // last CTA of the sub-panel applies the inverted triangle: x_J = Inv^T y_J
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&counters[tk.sp], 1u);
    s_last = (prev == (unsigned int)tk.ntiles - 1);
    if (s_last) counters[tk.sp] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const T *Inv = inv + tk.invoff;
  constexpr int NJ = NB / 32;
  for (int r0 = 0; r0 < nrhs; r0 += NR) {
    const int nr = min(NR, nrhs - r0);
    __syncthreads();
    for (int e = tid; e < NR * NB; e += PB200_BWD_NT) {
      const int rr = e / NB, j = e % NB;
      yj[rr][j] = (rr < nr && j < nb) ? ld_cg(&y[(size_t)(r0 + rr) * ldx + fcol + tk.c0 + j]) : ST<T>::zero();
    }
    __syncthreads();
    // x_i = sum_{j >= i} op(Inv[j][i]) y_j : warp per output, lanes along the contiguous j
    constexpr int IC = 4;
    for (int i0 = warp * IC; i0 < nb; i0 += nwarp * IC) {
      T a[IC][NJ];
#pragma unroll
      for (int q = 0; q < IC; ++q)
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const int j = lane + 32 * jj;   // entries with j < i are stored zeros
          a[q][jj] = (i0 + q < nb && j < nb) ? Inv[(size_t)(i0 + q) * nb + j] : ST<T>::zero();
          if (CONJ) a[q][jj] = ST<T>::conj(a[q][jj]);
        }
#pragma unroll
      for (int q = 0; q < IC; ++q)
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          T acc = ST<T>::zero();
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) fma_acc(acc, a[q][jj], yj[rr][lane + 32 * jj]);
          typedef typename ST<T>::real R;
          R *p = reinterpret_cast<R *>(&acc);
          for (int o = 16; o > 0; o >>= 1) {
            p[0] += __shfl_down_sync(0xffffffffu, p[0], o);
            if (ST<T>::is_complex) p[1] += __shfl_down_sync(0xffffffffu, p[1], o);
          }
          if (lane == 0 && rr < nr && i0 + q < nb) x[(size_t)(r0 + rr) * ldx + fcol + tk.c0 + i0 + q] = acc;
        }
    }
  }
}

}  // namespace pb200
