// kernels_factor.cuh — numeric factorization kernels (generic scalar path).
//
// What each kernel replaces in the reference (paths under src/sopalin/src):
//   k_assemble     Csc2solv_cblk                     csc_intern_solve.c:65-125
//   k_diag_factor  factor_diag -> PASTIX_{potrf,sytrf,hetrf,getrf}_block + DimTrans
//                                                    compute_diag.c:124-605
//   k_panel_trsm   factor_trsm1d / kernel_trsm       compute_trsm.c:40-171
//   k_update       compute_1dgemm = compute_contrib_compact + add_contrib_local
//                                                    sopalin_compute.c:270-598, 865-1032
// Static pivoting is the reference's rule: |pivot| < critere => pivot := critere,
// nbpivot++ (compute_diag.c:133-137, 232-236, 444-448).
#pragma once
#include "scalar.cuh"
#include "symbol.cuh"

namespace pb200 {

enum { F_LLT = 0, F_LDLT = 1, F_LU = 2, F_LDLH = 3 };

// ---------------------------------------------------------------- assemble
// one thread per column of the permuted CSC
template <class T>
__global__ void k_assemble(DevSym S, int n, const int64_t *__restrict__ colptr, const int *__restrict__ rows,
                           const T *__restrict__ vals, const T *__restrict__ tvals, int herm, T *L, T *U,
                           unsigned long long *dropped, const int *__restrict__ owner, int rank) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int c = S.col2cblk[j];
  if (owner != nullptr && owner[c] != rank) return;   // multi-GPU: that panel is a (zero) fan-in buffer here
  int fcol = S.fcol[c], ld = S.stride[c], b0 = S.fblok[c], b1 = S.fblok[c + 1];
  int64_t base = S.poff[c] + (int64_t)ld * (j - fcol);
  for (int64_t p = colptr[j]; p < colptr[j + 1]; ++p) {
    int r = rows[p];
    if (r < fcol) continue;
    int b = upper_le(S.frow, b0, b1, r);
    if (b >= b0 && r < S.frow[b] + S.nrow[b]) {
      int64_t idx = base + S.coefind[b] + (r - S.frow[b]);
      L[idx] = vals[p];
      if (U != nullptr && b != b0) U[idx] = herm ? ST<T>::conj(tvals[p]) : tvals[p];
    } else {
      atomicAdd(dropped, 1ULL);  // ILU: entry outside the incomplete pattern
    }
  }
}

// ---------------------------------------------------------------- diagonal block
// Right-looking unblocked factorization of a w x w block (leading dimension ld)
// by one CTA.  Two barriers per pivot.
template <class T, int FACTO>
__device__ void factor_block(T *A, int w, int ld, double crit, unsigned long long *nbpivot, T *s_piv) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k < w; ++k) {
    if (tid == 0) {
      T d = A[(size_t)k * (ld + 1)];
      if ((double)ST<T>::abs(d) < crit) { d = ST<T>::from_real(crit); atomicAdd(nbpivot, 1ULL); }
      if (FACTO == F_LLT) d = ST<T>::sqrt(d);
      A[(size_t)k * (ld + 1)] = d;
      *s_piv = d;
    }
    __syncthreads();
    const T d = *s_piv;
    const T inv = ST<T>::from_real(1.0) / d;
    T *colk = A + (size_t)k * ld;
    for (int i = k + 1 + tid; i < w; i += nt) colk[i] = colk[i] * inv;
    __syncthreads();
    const int nn = w - k - 1;
    for (int idx = tid; idx < nn * nn; idx += nt) {
      int j = k + 1 + idx / nn, i = k + 1 + idx % nn;
      if (FACTO == F_LU) {
        A[(size_t)j * ld + i] -= colk[i] * A[(size_t)j * ld + k];
      } else if (i >= j) {
        if (FACTO == F_LLT) A[(size_t)j * ld + i] -= colk[i] * colk[j];  // SYR 'L' (complex: geru, symmetric)
        else if (FACTO == F_LDLT) A[(size_t)j * ld + i] -= d * colk[i] * colk[j];
        else A[(size_t)j * ld + i] -= d * colk[i] * ST<T>::conj(colk[j]);  // HER
      }
    }
    __syncthreads();
  }
}

template <class T, int FACTO>
__global__ void k_diag_factor(DevSym S, T *L, T *U, const int *__restrict__ cblks, double crit,
                              unsigned long long *nbpivot, int smem_elems) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *>(smem_raw);
  __shared__ T s_piv;
  const int c = cblks[blockIdx.x];
  const int w = S.width[c], ld = S.stride[c];
  T *A = L + S.poff[c];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (w * w <= smem_elems) {
    for (int idx = tid; idx < w * w; idx += nt) sm[idx] = A[(size_t)(idx / w) * ld + idx % w];
    __syncthreads();
    factor_block<T, FACTO>(sm, w, w, crit, nbpivot, &s_piv);
    for (int idx = tid; idx < w * w; idx += nt) A[(size_t)(idx / w) * ld + idx % w] = sm[idx];
    if (FACTO == F_LU) {  // DimTrans: ucoeftab diag blok = (LU)^T
      T *UA = U + S.poff[c];
      for (int idx = tid; idx < w * w; idx += nt) UA[(size_t)(idx / w) * ld + idx % w] = sm[(idx % w) * w + idx / w];
    }
  } else {
    factor_block<T, FACTO>(A, w, ld, crit, nbpivot, &s_piv);
    if (FACTO == F_LU) {
      T *UA = U + S.poff[c];
      for (int idx = tid; idx < w * w; idx += nt) UA[(size_t)(idx / w) * ld + idx % w] = A[(size_t)(idx % w) * ld + idx / w];
    }
  }
}

// ---------------------------------------------------------------- panel TRSM
// One thread per off-diagonal row; right-looking column sweep.
// part 0: L panel; part 1 (LU only): U^T panel against the unit upper part of ucoeftab's diag blok.
#define PB200_TRSM_ROWS 128
template <class T, int FACTO>
__global__ void k_panel_trsm(DevSym S, T *L, T *U, const RowTask *__restrict__ tasks, int ntasks) {
  int tile = blockIdx.x, part = 0;
  if (FACTO == F_LU) { part = tile & 1; tile >>= 1; }
  const int t = find_task(tasks, ntasks, tile);
  const int c = tasks[t].cblk;
  const int w = S.width[c], ld = S.stride[c];
  const int m0 = w + (tile - tasks[t].tile0) * PB200_TRSM_ROWS + threadIdx.x;
  if (m0 >= ld) return;
  const T *A = L + S.poff[c];                      // diag blok of coeftab (all variants read it)
  T *P = (part == 0 ? L : U) + S.poff[c] + m0;     // this thread's row
  for (int j = 0; j < w; ++j) {
    T x = P[(size_t)j * ld];
    if (FACTO == F_LLT || (FACTO == F_LU && part == 0)) {
      x = x / A[(size_t)j * (ld + 1)];
      P[(size_t)j * ld] = x;
    }
    for (int l = j + 1; l < w; ++l) {
      T coef;
      if (FACTO == F_LU && part == 0) coef = A[(size_t)l * ld + j];        // U[j,l]
      else if (FACTO == F_LDLH) coef = ST<T>::conj(A[(size_t)j * ld + l]); // conj(L[l,j])
      else coef = A[(size_t)j * ld + l];                                   // L[l,j]
      P[(size_t)l * ld] -= x * coef;
    }
  }
  if (FACTO == F_LDLT || FACTO == F_LDLH) {
    for (int j = 0; j < w; ++j) {
      T inv = ST<T>::from_real(1.0) / A[(size_t)j * (ld + 1)];
      P[(size_t)j * ld] = P[(size_t)j * ld] * inv;
    }
  }
}

// ---------------------------------------------------------------- update (GEMM + scatter)
#define PB200_UPD_TM 64
#define PB200_UPD_TN 64
#define PB200_UPD_TK 16
template <class T, int FACTO>
__global__ void __launch_bounds__(256)
k_update(DevSym S, T *L, T *U, const UpdTask *__restrict__ tasks, int ntasks) {
  constexpr int TM = PB200_UPD_TM, TN = PB200_UPD_TN, TK = PB200_UPD_TK;
  __shared__ T As[TK][TM + 1];
  __shared__ T Bs[TK][TN + 1];
  __shared__ int s_ro[TM];   // row offset inside the facing panel, -1 = no facing blok (ILU)
  __shared__ int s_sb[TM];   // source blok of the row

  int tile = blockIdx.x, part = 0;
  if (FACTO == F_LU) { part = tile & 1; tile >>= 1; }
  const int t = find_task(tasks, ntasks, tile);
  const UpdTask tk = tasks[t];
  const int k = tk.cblk, b1 = tk.blok;
  const int local = tile - tk.tile0;
  const int tm = local / tk.ntn, tn = local % tk.ntn;
  const int w = S.width[k], ld = S.stride[k];
  const int r0 = S.coefind[b1], nb1 = S.nrow[b1];
  const int m0 = r0 + tm * TM, mrows = min(TM, ld - m0);
  const int n0 = r0 + tn * TN, ncols = min(TN, nb1 - tn * TN);
  const int fc = S.fcblk[b1];
  const int tid = threadIdx.x;

  const T *Ap = ((FACTO == F_LU && part == 1) ? U : L) + S.poff[k];
  const T *Bp = ((FACTO == F_LU && part == 0) ? U : L) + S.poff[k];
  const T *Dp = L + S.poff[k];

  if (tid < TM) {
    int ro = -1, sb = -1;
    if (tid < mrows) {
      const int m = m0 + tid;
      sb = upper_le(S.coefind, S.fblok[k], S.fblok[k + 1], m);
      const int r = S.frow[sb] + (m - S.coefind[sb]);
      const int tb = upper_le(S.frow, S.fblok[fc], S.fblok[fc + 1], r);
      if (tb >= S.fblok[fc] && r < S.frow[tb] + S.nrow[tb]) ro = S.coefind[tb] + (r - S.frow[tb]);
    }
    s_ro[tid] = ro; s_sb[tid] = sb;
  }

  T acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = ST<T>::zero();
  const int ti = (tid & 15) * 4, tj = (tid >> 4) * 4;

  for (int k0 = 0; k0 < w; k0 += TK) {
    for (int e = tid; e < TK * TM; e += 256) {
      const int kk = e / TM, i = e % TM;
      T v = ST<T>::zero();
      if (i < mrows && k0 + kk < w) v = Ap[(size_t)(k0 + kk) * ld + m0 + i];
      As[kk][i] = v;
    }
    for (int e = tid; e < TK * TN; e += 256) {
      const int kk = e / TN, j = e % TN;
      T v = ST<T>::zero();
      if (j < ncols && k0 + kk < w) {
        v = Bp[(size_t)(k0 + kk) * ld + n0 + j];
        if (FACTO == F_LDLT) v = v * Dp[(size_t)(k0 + kk) * (ld + 1)];
        else if (FACTO == F_LDLH) v = ST<T>::conj(v * Dp[(size_t)(k0 + kk) * (ld + 1)]);
        else if (FACTO == F_LLT) v = ST<T>::conj(v);  // GEMM "N","C" (sopalin_compute.c:327-332)
      }
      Bs[kk][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { a[x] = As[kk][ti + x]; b[x] = Bs[kk][tj + x]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) fma_acc(acc[x][y], a[x], b[y]);
    }
    __syncthreads();
  }

  // scatter-subtract into the facing cblk (add_contrib_local)
  const int tld = S.stride[fc], tw = S.width[fc];
  const int cj0 = S.frow[b1] + (n0 - r0) - S.fcol[fc];
  T *TL = L + S.poff[fc];
  T *TU = (FACTO == F_LU) ? U + S.poff[fc] : nullptr;
#pragma unroll
  for (int y = 0; y < 4; ++y) {
    const int j = tj + y;
    if (j >= ncols) continue;
    const int cj = cj0 + j;
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int i = ti + x;
      if (i >= mrows) continue;
      const int ro = s_ro[i];
      if (ro < 0) continue;
      if (FACTO != F_LU || part == 0) {
        atomic_sub(&TL[(size_t)cj * tld + ro], acc[x][y]);
      } else if (ro >= tw) {
        atomic_sub(&TU[(size_t)cj * tld + ro], acc[x][y]);
      } else if (s_sb[i] != b1) {
        // facing blok is the diagonal one: U part stored transposed into coeftab
        // (sopalin_compute.c:431-435, 572-575); b1 == b2 is skipped
        atomic_sub(&TL[(size_t)ro * tld + cj], acc[x][y]);
      }
    }
  }
}

// ---------------------------------------------------------------- inertia
template <class T>
__global__ void k_inertia(DevSym S, const T *L, unsigned long long *count) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S.cblknbr) return;
  const T *A = L + S.poff[c];
  int w = S.width[c], ld = S.stride[c], cnt = 0;
  for (int k = 0; k < w; ++k) if (ST<T>::re(A[(size_t)k * (ld + 1)]) > 0) cnt++;
  atomicAdd(count, (unsigned long long)cnt);
}

}  // namespace pb200
