// kernels_small.cuh — small supernodes: one WARP per column block, the whole compute_1d task
// (factor_diag + factor_trsm1d + every compute_1dgemm / add_contrib_local of the cblk,
// sopalin_compute.c:747-1032) fused in one pass, batched by elimination level.
//
// This is the regime of incomplete factorizations (ILU(k): tens of thousands of cblks of width 1-5 with a
// dozen one-row bloks each, sopalin_compute.c:468-597 handles their partially facing bloks) and of the
// leaves of direct factorizations in the precisions that do not use the tensor-core path.  The dependency
// graph is deep and thin (32^3 ILU(2): 1284 levels, median 2 cblks per level), so what matters is the latency
// of one level, not throughput:
//   * everything structural is precomputed: the target address of every (row m, row n) contribution of a
//     cblk is a table entry (k_build_small_pairs) — no searches on the critical path, and contributions
//     without a facing row (dropped by the incomplete pattern) are -1;
//   * levels with many cblks are one launch (k_small_level, 8 warps per CTA);
//   * runs of consecutive thin levels are ONE single-CTA launch (k_small_chain): 16 warps take the cblks
//     of a level, a CTA barrier + fence separates levels — no kernel boundary, no grid synchronisation.
#pragma once
#include "scalar.cuh"
#include "symbol.cuh"
#include "kernels_factor.cuh"
#include "kernels_solve.cuh"

namespace pb200 {

#define PB200_SM_WMAX 8     // widest cblk handled here
#define PB200_SM_RMAX 64    // most off-diagonal rows handled here
#define PB200_SM_LDW (PB200_SM_WMAX + 1)

// shared memory of one warp
template <class T>
struct SmallWs {
  T W[PB200_SM_WMAX * PB200_SM_LDW];       // factored diagonal block, column-major, ld = WMAX+1
  T Xa[PB200_SM_RMAX * PB200_SM_WMAX];     // first operand rows  (L)
  T Xb[PB200_SM_RMAX * PB200_SM_WMAX];     // second operand rows (LLt: L, LDLt: L*D, LU: U^T)
};

// pair p = tri(mi, ni), mi >= ni  ->  (mi, ni)
__device__ __forceinline__ void tri_decode(int p, int &mi, int &ni) {
  mi = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
  while (mi * (mi + 1) / 2 > p) --mi;
  while ((mi + 1) * (mi + 2) / 2 <= p) ++mi;
  ni = p - mi * (mi + 1) / 2;
}

// target of every contribution of the small cblks (one warp per cblk).  tabL[p]: slab offset in coeftab of
// element (row m, column n of the facing cblk) or -1.  tabU (LU only): >= 0 offset in ucoeftab; <= -2: the
// contribution lands TRANSPOSED in coeftab at offset -(v+2) (diagonal target, sopalin_compute.c:431-435);
// -1: skipped (the diagonal element itself, or no facing row).  For a pair inside ONE blok the mirrored write is the
// upper half of the square the reference's L-part GEMM adds in full (the U part is skipped there, :572-575).
__global__ void k_build_small_pairs(DevSym S, const int *__restrict__ cblks, int ncblk, const int64_t *__restrict__ pbase,
                                    int64_t *tabL, int64_t *tabU) {
  const int wi = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (wi >= ncblk) return;
  const int c = cblks[wi];
  const int w = S.width[c], ld = S.stride[c], bf = S.fblok[c] + 1, be = S.fblok[c + 1];
  const int mr = ld - w;
  const int64_t base = pbase[wi];
  for (int p = lane; p < mr * (mr + 1) / 2; p += 32) {
    int mi, ni;
    tri_decode(p, mi, ni);
    const int m = w + mi, n = w + ni;
    const int b1 = upper_le(S.coefind, bf, be, n), b2 = upper_le(S.coefind, bf, be, m);
    const int fc = S.fcblk[b1];
    const int cj = S.frow[b1] + (n - S.coefind[b1]) - S.fcol[fc];
    const int r = S.frow[b2] + (m - S.coefind[b2]);
    const int tld = S.stride[fc], tw = S.width[fc];
    const int tb = upper_le(S.frow, S.fblok[fc], S.fblok[fc + 1], r);
    int64_t oL = -1, oU = -1;
    if (tb >= S.fblok[fc] && r < S.frow[tb] + S.nrow[tb]) {
      const int ro = S.coefind[tb] + (r - S.frow[tb]);
      oL = S.poff[fc] + (int64_t)cj * tld + ro;
      if (ro >= tw) oU = oL;
      else if (m != n) oU = -(S.poff[fc] + (int64_t)ro * tld + cj) - 2;   // mirror image in coeftab's diagonal blok
    }
    tabL[base + p] = oL;
    if (tabU != nullptr) tabU[base + p] = oU;
  }
}

// the whole task of cblk c by one warp
template <class T, int FACTO>
__device__ void small_cblk(const DevSym &S, T *L, T *U, int c, double crit, unsigned long long *nbpivot,
                           const int64_t *__restrict__ tabL, const int64_t *__restrict__ tabU, int64_t pbase,
                           SmallWs<T> &ws, int lane) {
  const int w = S.width[c], ld = S.stride[c];
  const int mr = ld - w;
  T *P = L + S.poff[c];
  T *Q = (FACTO == F_LU) ? U + S.poff[c] : nullptr;
  const T one = ST<T>::from_real(1.0);
  // ---- diagonal block -> shared memory
  for (int e = lane; e < w * w; e += 32) {
    const int j = e / w, i = e % w;
    ws.W[j * PB200_SM_LDW + i] = ld_cg(P + (size_t)j * ld + i);
  }
  // off-diagonal rows -> shared memory (row r at X[r*WMAX + k])
  for (int e = lane; e < mr * w; e += 32) {
    const int k = e / mr, r = e % mr;
    ws.Xa[r * PB200_SM_WMAX + k] = ld_cg(P + (size_t)k * ld + w + r);
    if (FACTO == F_LU) ws.Xb[r * PB200_SM_WMAX + k] = ld_cg(Q + (size_t)k * ld + w + r);
  }
  __syncwarp();
  // ---- factor_diag: right-looking, static pivoting (compute_diag.c:124-160, 223-250, 432-467)
  for (int k = 0; k < w; ++k) {
    if (lane == 0) {
      T d = ws.W[k * PB200_SM_LDW + k];
      if ((double)ST<T>::abs(d) < crit) { d = ST<T>::from_real(crit); atomicAdd(nbpivot, 1ULL); }
      if (FACTO == F_LLT) d = ST<T>::sqrt(d);
      ws.W[k * PB200_SM_LDW + k] = d;
    }
    __syncwarp();
    const T d = ws.W[k * PB200_SM_LDW + k];
    const T inv = one / d;
    if (lane > k && lane < w) ws.W[k * PB200_SM_LDW + lane] = ws.W[k * PB200_SM_LDW + lane] * inv;
    __syncwarp();
    const int nn = w - k - 1;
    for (int e = lane; e < nn * nn; e += 32) {
      const int j = k + 1 + e / nn, i = k + 1 + e % nn;
      T *a = &ws.W[j * PB200_SM_LDW + i];
      const T li = ws.W[k * PB200_SM_LDW + i];
      if (FACTO == F_LU) *a = *a - li * ws.W[j * PB200_SM_LDW + k];
      else if (i >= j) {
        const T lj = ws.W[k * PB200_SM_LDW + j];
        if (FACTO == F_LLT) *a = *a - li * lj;
        else if (FACTO == F_LDLT) *a = *a - d * li * lj;
        else *a = *a - d * li * ST<T>::conj(lj);
      }
    }
    __syncwarp();
  }
  // write the factored block back (LU: and its transpose into ucoeftab, DimTrans compute_diag.c:521-536)
  for (int e = lane; e < w * w; e += 32) {
    const int j = e / w, i = e % w;
    if (FACTO == F_LU || i >= j) P[(size_t)j * ld + i] = ws.W[j * PB200_SM_LDW + i];
    if (FACTO == F_LU) Q[(size_t)i * ld + j] = ws.W[j * PB200_SM_LDW + i];
  }
  // ---- factor_trsm1d: lane per off-diagonal row (compute_trsm.c:58-171)
  for (int r = lane; r < mr; r += 32) {
    T *x = &ws.Xa[r * PB200_SM_WMAX];
    for (int j = 0; j < w; ++j) {
      T xj = x[j];
      if (FACTO == F_LLT || FACTO == F_LU) { xj = xj / ws.W[j * PB200_SM_LDW + j]; x[j] = xj; }
      for (int l = j + 1; l < w; ++l) {
        T coef;
        if (FACTO == F_LU) coef = ws.W[l * PB200_SM_LDW + j];                        // U[j][l]
        else if (FACTO == F_LDLH) coef = ST<T>::conj(ws.W[j * PB200_SM_LDW + l]);   // conj(L[l][j])
        else coef = ws.W[j * PB200_SM_LDW + l];                                      // L[l][j]
        x[l] = x[l] - xj * coef;
      }
    }
    if (FACTO == F_LU) {          // U^T rows against the unit lower triangle L_kk
      T *y = &ws.Xb[r * PB200_SM_WMAX];
      for (int j = 0; j < w; ++j) {
        const T yj = y[j];
        for (int l = j + 1; l < w; ++l) y[l] = y[l] - yj * ws.W[j * PB200_SM_LDW + l];
      }
      for (int j = 0; j < w; ++j) Q[(size_t)j * ld + w + r] = y[j];
    }
    if (FACTO == F_LDLT || FACTO == F_LDLH) {
      // Xb keeps L*D (the reference's maxbloktab1 copy, compute_trsm.c:86-114), Xa the final L = (L*D) D^-1
      T *y = &ws.Xb[r * PB200_SM_WMAX];
      for (int j = 0; j < w; ++j) { y[j] = x[j]; x[j] = x[j] / ws.W[j * PB200_SM_LDW + j]; }
    }
    for (int j = 0; j < w; ++j) P[(size_t)j * ld + w + r] = x[j];
  }
  __syncwarp();
  // ---- every contribution of the cblk: C[m][n] = sum_k A[m][k] op(B[n][k]), m >= n over the off-diagonal rows
  const int npairs = mr * (mr + 1) / 2;
  for (int p = lane; p < npairs; p += 32) {
    const int64_t oL = tabL[pbase + p];
    int64_t oU = -1;
    if (FACTO == F_LU) oU = tabU[pbase + p];
    if (oL < 0 && oU == -1) continue;
    int mi, ni;
    tri_decode(p, mi, ni);
    const T *am = &ws.Xa[mi * PB200_SM_WMAX], *an = &ws.Xa[ni * PB200_SM_WMAX];
    const T *bm = &ws.Xb[mi * PB200_SM_WMAX], *bn = &ws.Xb[ni * PB200_SM_WMAX];
    T v = ST<T>::zero(), vu = ST<T>::zero();
    for (int k = 0; k < w; ++k) {
      if (FACTO == F_LLT) fma_acc(v, am[k], ST<T>::conj(an[k]));
      else if (FACTO == F_LDLT) fma_acc(v, am[k], bn[k]);
      else if (FACTO == F_LDLH) fma_acc(v, am[k], ST<T>::conj(bn[k]));
      else { fma_acc(v, am[k], bn[k]); fma_acc(vu, bm[k], an[k]); }
    }
    if (oL >= 0) atomic_sub(L + oL, v);
    if (FACTO == F_LU) {
      if (oU >= 0) atomic_sub(U + oU, vu);
      else if (oU <= -2) atomic_sub(L + (-(oU + 2)), vu);
    }
  }
}

#define PB200_SM_WARPS 8
// one launch per level with many small cblks
template <class T, int FACTO>
__global__ void __launch_bounds__(PB200_SM_WARPS * 32)
k_small_level(DevSym S, T *L, T *U, const int *__restrict__ cblks, int ncblk, const int64_t *__restrict__ pbase,
              const int64_t *__restrict__ tabL, const int64_t *__restrict__ tabU, double crit, unsigned long long *nbpivot) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallWs<T> *ws = reinterpret_cast<SmallWs<T> *>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = blockIdx.x * PB200_SM_WARPS + warp;
  if (wi >= ncblk) return;
  small_cblk<T, FACTO>(S, L, U, cblks[wi], crit, nbpivot, tabL, tabU, pbase[wi], ws[warp], lane);
}

// a run of consecutive thin levels in ONE CTA: lvl_ptr[q] .. lvl_ptr[q+1] index cblks / pbase
template <class T> struct SmChain { static constexpr int WARPS = sizeof(T) >= 16 ? 8 : 16; };
template <class T, int FACTO>
__global__ void __launch_bounds__(SmChain<T>::WARPS * 32)
k_small_chain(DevSym S, T *L, T *U, const int *__restrict__ cblks, const int *__restrict__ lvl_ptr, int nlev,
              const int64_t *__restrict__ pbase, const int64_t *__restrict__ tabL, const int64_t *__restrict__ tabU,
              double crit, unsigned long long *nbpivot) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallWs<T> *ws = reinterpret_cast<SmallWs<T> *>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = 0; q < nlev; ++q) {
    const int i0 = lvl_ptr[q], i1 = lvl_ptr[q + 1];
    for (int wi = i0 + warp; wi < i1; wi += SmChain<T>::WARPS)
      small_cblk<T, FACTO>(S, L, U, cblks[wi], crit, nbpivot, tabL, tabU, pbase[wi], ws[warp], lane);
    __threadfence();      // the level's reductions are ordered before the next level's (L1-bypassing) loads
    __syncthreads();
  }
}

}  // namespace pb200
