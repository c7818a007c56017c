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
// One CTA per sub-panel, thread j builds column j by forward substitution into shared memory.
template <class T>
__global__ void __launch_bounds__(128)
k_tri_inverse(DevSym S, const T *__restrict__ M, const SlvTask *__restrict__ tasks, const int64_t *__restrict__ invoff,
              T *inv, int unit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *Xs = reinterpret_cast<T *>(smem_raw);
  const SlvTask tk = tasks[blockIdx.x];
  const int c = tk.cblk, ld = S.stride[c], nb = tk.c1 - tk.c0;
  const T *W = M + S.poff[c] + (size_t)tk.c0 * (ld + 1);
  const int ldx = nb | 1;
  const int j = threadIdx.x;
  const T one = ST<T>::from_real(1.0), zero = ST<T>::zero();
  if (j < nb) {
    T *x = Xs + (size_t)j * ldx;
    x[j] = unit ? one : one / W[(size_t)j * (ld + 1)];
    for (int i = j + 1; i < nb; ++i) {
      T s = zero;
      for (int k = j; k < i; ++k) fma_acc(s, W[(size_t)k * ld + i], x[k]);
      s = zero - s;
      x[i] = unit ? s : s / W[(size_t)i * (ld + 1)];
    }
  }
  __syncthreads();
  T *out = inv + invoff[tk.sp];
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int jj = e / nb, i = e % nb;
    out[e] = (i >= jj) ? Xs[(size_t)jj * ldx + i] : zero;
  }
}

// ---- forward: one launch per (level, round); CTA = (sub-panel, 128-row tile of rows [c1, stride))
template <class T, int FACTO>
__global__ void __launch_bounds__(PB200_SLV_ROWS)
k_fwd(DevSym S, const T *__restrict__ L, const T *__restrict__ inv, const int64_t *__restrict__ invoff,
      T *x, T *y, int64_t ldx, int nrhs, const SlvTask *__restrict__ tasks, const int *__restrict__ tile2task) {
  constexpr int NB = SlvCfg<T>::NB, NR = PB200_SLV_NR;
  __shared__ T xs[NR][NB];
  __shared__ T ys[NR][NB];
  const SlvTask tk = tasks[tile2task[blockIdx.x]];
  const int c = tk.cblk, ld = S.stride[c], fcol = S.fcol[c], nb = tk.c1 - tk.c0;
  const int tile = blockIdx.x - tk.tile0;
  const int tid = threadIdx.x;
  const T *P = L + S.poff[c];
  const T *Inv = inv + invoff[tk.sp];
  const int m = tk.c1 + tile * PB200_SLV_ROWS + tid;
  const bool rowok = m < ld;
  const int grow = rowok ? panel_row_to_global(S, c, m) : 0;
  for (int r0 = 0; r0 < nrhs; r0 += NR) {
    const int nr = min(NR, nrhs - r0);
    __syncthreads();
    for (int e = tid; e < NR * NB; e += PB200_SLV_ROWS) {
      const int r = e / NB, j = e % NB;
      xs[r][j] = (r < nr && j < nb) ? x[(size_t)(r0 + r) * ldx + fcol + tk.c0 + j] : ST<T>::zero();
    }
    __syncthreads();
    // y_J = Inv * x_J (lower triangular): thread i builds row i, coalesced column walk
    if (tid < nb) {
      T acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = ST<T>::zero();
      // (the strict upper triangle of Inv is stored as zeros: uniform trip count, deep unrolling)
#pragma unroll 16
      for (int j = 0; j < nb; ++j) {
        const T a = Inv[(size_t)j * nb + tid];
#pragma unroll
        for (int r = 0; r < NR; ++r) fma_acc(acc[r], a, xs[r][j]);
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) ys[r][tid] = acc[r];
      if (tile == 0) {
        // write-back; LDLt/LDLh: the diagonal step x_k /= D_kk folded in (updo.c:948-984)
        T d = ST<T>::from_real(1.0);
        if (FACTO == F_LDLT || FACTO == F_LDLH) d = P[(size_t)(tk.c0 + tid) * (ld + 1)];
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (r < nr) y[(size_t)(r0 + r) * ldx + fcol + tk.c0 + tid] =
              (FACTO == F_LDLT || FACTO == F_LDLH) ? acc[r] / d : acc[r];
      }
    }
    __syncthreads();
    if (rowok) {
      T acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = ST<T>::zero();
      const T *col = P + (size_t)tk.c0 * ld + m;
#pragma unroll 16
      for (int j = 0; j < nb; ++j) {
        const T a = col[(size_t)j * ld];
#pragma unroll
        for (int r = 0; r < NR; ++r) fma_acc(acc[r], a, ys[r][j]);
      }
#pragma unroll
      for (int r = 0; r < NR; ++r)
        if (r < nr) atomic_sub(&x[(size_t)(r0 + r) * ldx + grow], acc[r]);
    }
  }
}

// ---- backward: one launch per (level, round), descending.  M is coeftab (ucoeftab for LU).
template <class T, int FACTO>
__global__ void __launch_bounds__(256)
k_bwd(DevSym S, const T *__restrict__ M, const T *__restrict__ inv, const int64_t *__restrict__ invoff,
      T *x, T *y, int64_t ldx, int nrhs, const SlvTask *__restrict__ tasks, const int *__restrict__ tile2task,
      unsigned int *counters) {
  constexpr int NB = SlvCfg<T>::NB, NR = PB200_SLV_NR;
  constexpr bool CONJ = (FACTO == F_LDLH);
  __shared__ T xr[NR][PB200_SLV_ROWS];
  __shared__ T yj[NR][NB];
  __shared__ int s_last;
  const SlvTask tk = tasks[tile2task[blockIdx.x]];
  const int c = tk.cblk, ld = S.stride[c], fcol = S.fcol[c], nb = tk.c1 - tk.c0;
  const int tile = blockIdx.x - tk.tile0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const T *P = M + S.poff[c];
  const int mbase = tk.c1 + tile * PB200_SLV_ROWS;
  const int mrows = max(0, min(PB200_SLV_ROWS, ld - mbase));
  int grow = 0;
  if (tid < mrows) grow = panel_row_to_global(S, c, mbase + tid);
  if (mrows > 0) {
    for (int r0 = 0; r0 < nrhs; r0 += NR) {
      const int nr = min(NR, nrhs - r0);
      __syncthreads();
      if (tid < PB200_SLV_ROWS)
#pragma unroll
        for (int r = 0; r < NR; ++r)
          xr[r][tid] = (tid < mrows && r < nr) ? x[(size_t)(r0 + r) * ldx + grow] : ST<T>::zero();
      __syncthreads();
      constexpr int JC = 4;   // columns per warp pass: JC * (rows/32) loads in flight
      for (int j0 = warp * JC; j0 < nb; j0 += nwarp * JC) {
        T acc[JC][NR];
#pragma unroll
        for (int q = 0; q < JC; ++q)
#pragma unroll
          for (int r = 0; r < NR; ++r) acc[q][r] = ST<T>::zero();
        for (int i = lane; i < mrows; i += 32) {
          T a[JC];
#pragma unroll
          for (int q = 0; q < JC; ++q) {
            a[q] = (j0 + q < nb) ? P[(size_t)(tk.c0 + j0 + q) * ld + mbase + i] : ST<T>::zero();
            if (CONJ) a[q] = ST<T>::conj(a[q]);
          }
#pragma unroll
          for (int q = 0; q < JC; ++q)
#pragma unroll
            for (int r = 0; r < NR; ++r) fma_acc(acc[q][r], a[q], xr[r][i]);
        }
#pragma unroll
        for (int q = 0; q < JC; ++q)
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            typedef typename ST<T>::real R;
            R *p = reinterpret_cast<R *>(&acc[q][r]);
            for (int o = 16; o > 0; o >>= 1) {
              p[0] += __shfl_down_sync(0xffffffffu, p[0], o);
              if (ST<T>::is_complex) p[1] += __shfl_down_sync(0xffffffffu, p[1], o);
            }
            if (lane == 0 && r < nr && j0 + q < nb) atomic_sub(&y[(size_t)(r0 + r) * ldx + fcol + tk.c0 + j0 + q], acc[q][r]);
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
  const T *Inv = inv + invoff[tk.sp];
  for (int r0 = 0; r0 < nrhs; r0 += NR) {
    const int nr = min(NR, nrhs - r0);
    __syncthreads();
    for (int e = tid; e < NR * NB; e += blockDim.x) {
      const int r = e / NB, j = e % NB;
      yj[r][j] = (r < nr && j < nb) ? ld_cg(&y[(size_t)(r0 + r) * ldx + fcol + tk.c0 + j]) : ST<T>::zero();
    }
    __syncthreads();
    // x_i = sum_{j >= i} op(Inv[j][i]) y_j : warp per output, lanes along the contiguous j
    constexpr int IC = 4;
    for (int i0 = warp * IC; i0 < nb; i0 += nwarp * IC) {
      T acc[IC][NR];
#pragma unroll
      for (int q = 0; q < IC; ++q)
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[q][r] = ST<T>::zero();
      for (int j = lane; j < nb; j += 32) {   // entries with j < i are stored zeros
        T a[IC];
#pragma unroll
        for (int q = 0; q < IC; ++q) {
          a[q] = (i0 + q < nb) ? Inv[(size_t)(i0 + q) * nb + j] : ST<T>::zero();
          if (CONJ) a[q] = ST<T>::conj(a[q]);
        }
#pragma unroll
        for (int q = 0; q < IC; ++q)
#pragma unroll
          for (int r = 0; r < NR; ++r) fma_acc(acc[q][r], a[q], yj[r][j]);
      }
#pragma unroll
      for (int q = 0; q < IC; ++q)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          typedef typename ST<T>::real R;
          R *p = reinterpret_cast<R *>(&acc[q][r]);
          for (int o = 16; o > 0; o >>= 1) {
            p[0] += __shfl_down_sync(0xffffffffu, p[0], o);
            if (ST<T>::is_complex) p[1] += __shfl_down_sync(0xffffffffu, p[1], o);
          }
          if (lane == 0 && r < nr && i0 + q < nb) x[(size_t)(r0 + r) * ldx + fcol + tk.c0 + i0 + q] = acc[q][r];
        }
    }
  }
}

}  // namespace pb200
