// kernels_solve.cuh — level-scheduled up_down (generic scalar path).
//
// Reference (src/sopalin/src): up_down_smp updo.c:114-1664
//   down : x_c <- L_cc^{-1} x_c (TRSV/TRSM, unit diagonal except LLt)  updo.c:574-596
//          x[rows(b)] -= L_b x_c for every off-diagonal blok            updo.c:631-793
//   diag : x_c[k] /= D_kk (LDLt / LDLh)                                 updo.c:948-984
//   up   : x_c -= L_b^T x[rows(b)] (LU: U^T panel, LDLh: conjugate)     updo_sendrecv.c:496-639
//          x_c <- L_cc^{-T} x_c                                         updo.c:1309-1342
// x is the permuted right-hand side, column-major n x nrhs with leading dimension ldx
// (UpDownVector.sm2xtab, blend/src/updown.h:69-72).  Rows of a blok are global
// indices, so x[frow(b) + i] is addressed directly.
#pragma once
#include "scalar.cuh"
#include "symbol.cuh"
#include "kernels_factor.cuh"

namespace pb200 {

#define PB200_SLV_ROWS 128
#define PB200_SLV_NR 4   // right-hand sides handled per pass over the panel

// ---- forward: diagonal solve, one CTA per cblk
template <class T, int FACTO>
__global__ void k_fwd_diag(DevSym S, const T *__restrict__ L, T *x, int64_t ldx, int nrhs,
                           const int *__restrict__ cblks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *xs = reinterpret_cast<T *>(smem_raw);  // w * nrhs
  const int c = cblks[blockIdx.x];
  const int w = S.width[c], ld = S.stride[c], fcol = S.fcol[c];
  const T *A = L + S.poff[c];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < w * nrhs; e += nt) xs[e] = x[(size_t)(e / w) * ldx + fcol + e % w];
  __syncthreads();
  for (int j = 0; j < w; ++j) {
    if (FACTO == F_LLT) {
      for (int r = tid; r < nrhs; r += nt) xs[r * w + j] = xs[r * w + j] / A[(size_t)j * (ld + 1)];
      __syncthreads();
    }
    const int nn = w - j - 1;
    for (int e = tid; e < nn * nrhs; e += nt) {
      const int r = e / nn, i = j + 1 + e % nn;
      xs[r * w + i] -= A[(size_t)j * ld + i] * xs[r * w + j];
    }
    __syncthreads();
  }
  for (int e = tid; e < w * nrhs; e += nt) x[(size_t)(e / w) * ldx + fcol + e % w] = xs[e];
}

// ---- forward: off-diagonal update, one thread per panel row
template <class T>
__global__ void k_fwd_update(DevSym S, const T *__restrict__ L, T *x, int64_t ldx, int nrhs,
                             const RowTask *__restrict__ tasks, int ntasks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *xs = reinterpret_cast<T *>(smem_raw);  // w * PB200_SLV_NR
  const int t = find_task(tasks, ntasks, blockIdx.x);
  const int c = tasks[t].cblk;
  const int w = S.width[c], ld = S.stride[c], fcol = S.fcol[c];
  const int m = w + (blockIdx.x - tasks[t].tile0) * PB200_SLV_ROWS + threadIdx.x;
  const T *A = L + S.poff[c];
  int grow = -1;
  if (m < ld) {
    const int sb = upper_le(S.coefind, S.fblok[c], S.fblok[c + 1], m);
    grow = S.frow[sb] + (m - S.coefind[sb]);
  }
  for (int r0 = 0; r0 < nrhs; r0 += PB200_SLV_NR) {
    const int nr = min(PB200_SLV_NR, nrhs - r0);
    __syncthreads();
    for (int e = threadIdx.x; e < w * nr; e += blockDim.x) xs[e] = x[(size_t)(r0 + e / w) * ldx + fcol + e % w];
    __syncthreads();
    if (m < ld) {
      T acc[PB200_SLV_NR];
#pragma unroll
      for (int r = 0; r < PB200_SLV_NR; ++r) acc[r] = ST<T>::zero();
      for (int l = 0; l < w; ++l) {
        const T a = A[(size_t)l * ld + m];
#pragma unroll
        for (int r = 0; r < PB200_SLV_NR; ++r) if (r < nr) fma_acc(acc[r], a, xs[r * w + l]);
      }
#pragma unroll
      for (int r = 0; r < PB200_SLV_NR; ++r) if (r < nr) atomic_sub(&x[(size_t)(r0 + r) * ldx + grow], acc[r]);
    }
  }
}

// ---- diagonal scaling (LDLt / LDLh), one thread per unknown
template <class T>
__global__ void k_diag_scale(DevSym S, const T *__restrict__ L, T *x, int64_t ldx, int nrhs, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int c = S.col2cblk[j];
  const T d = L[S.poff[c] + (size_t)(j - S.fcol[c]) * (S.stride[c] + 1)];
  for (int r = 0; r < nrhs; ++r) x[(size_t)r * ldx + j] = x[(size_t)r * ldx + j] / d;
}

// ---- backward: x_c -= B^T x[rows], one CTA per 128-row chunk, a warp per column subset
template <class T, int FACTO>
__global__ void k_bwd_update(DevSym S, const T *__restrict__ M, T *x, int64_t ldx, int nrhs,
                             const RowTask *__restrict__ tasks, int ntasks) {
  __shared__ T xr[PB200_SLV_ROWS];
  const int t = find_task(tasks, ntasks, blockIdx.x);
  const int c = tasks[t].cblk;
  const int w = S.width[c], ld = S.stride[c], fcol = S.fcol[c];
  const int mbase = w + (blockIdx.x - tasks[t].tile0) * PB200_SLV_ROWS;
  const int mrows = min(PB200_SLV_ROWS, ld - mbase);
  const T *A = M + S.poff[c];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  int grow = -1;
  if (tid < mrows) {
    const int m = mbase + tid;
    const int sb = upper_le(S.coefind, S.fblok[c], S.fblok[c + 1], m);
    grow = S.frow[sb] + (m - S.coefind[sb]);
  }
  for (int r = 0; r < nrhs; ++r) {
    __syncthreads();
    if (tid < PB200_SLV_ROWS) xr[tid] = (tid < mrows) ? x[(size_t)r * ldx + grow] : ST<T>::zero();
    __syncthreads();
    for (int l = warp; l < w; l += nwarp) {
      T acc = ST<T>::zero();
      for (int i = lane; i < mrows; i += 32) {
        T a = A[(size_t)l * ld + mbase + i];
        if (FACTO == F_LDLH) a = ST<T>::conj(a);
        fma_acc(acc, a, xr[i]);
      }
      // warp reduction
      if (ST<T>::is_complex) {
        typedef typename ST<T>::real R;
        R *p = reinterpret_cast<R *>(&acc);
        for (int o = 16; o > 0; o >>= 1) { p[0] += __shfl_down_sync(0xffffffffu, p[0], o); p[1] += __shfl_down_sync(0xffffffffu, p[1], o); }
      } else {
        typedef typename ST<T>::real R;
        R *p = reinterpret_cast<R *>(&acc);
        for (int o = 16; o > 0; o >>= 1) p[0] += __shfl_down_sync(0xffffffffu, p[0], o);
      }
      if (lane == 0) atomic_sub(&x[(size_t)r * ldx + fcol + l], acc);
    }
  }
}

// ---- backward: diagonal solve x_c <- A_cc^{-T} x_c, one CTA per cblk
template <class T, int FACTO>
__global__ void k_bwd_diag(DevSym S, const T *__restrict__ M, T *x, int64_t ldx, int nrhs,
                           const int *__restrict__ cblks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *xs = reinterpret_cast<T *>(smem_raw);  // w * nrhs
  const int c = cblks[blockIdx.x];
  const int w = S.width[c], ld = S.stride[c], fcol = S.fcol[c];
  const T *A = M + S.poff[c];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < w * nrhs; e += nt) xs[e] = x[(size_t)(e / w) * ldx + fcol + e % w];
  __syncthreads();
  for (int j = w - 1; j >= 0; --j) {
    if (FACTO == F_LLT || FACTO == F_LU) {
      for (int r = tid; r < nrhs; r += nt) xs[r * w + j] = xs[r * w + j] / A[(size_t)j * (ld + 1)];
      __syncthreads();
    }
    // x_i -= conj?(A[j,i]) * x_j for i < j  (row j of the lower triangle)
    for (int e = tid; e < j * nrhs; e += nt) {
      const int r = e / j, i = e % j;
      T a = A[(size_t)i * ld + j];
      if (FACTO == F_LDLH) a = ST<T>::conj(a);
      xs[r * w + i] -= a * xs[r * w + j];
    }
    __syncthreads();
  }
  for (int e = tid; e < w * nrhs; e += nt) x[(size_t)(e / w) * ldx + fcol + e % w] = xs[e];
}

}  // namespace pb200
