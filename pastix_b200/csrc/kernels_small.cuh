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

// ================================================================ up_down for small cblks
// Same organisation for the triangular solves: one warp per cblk, levels with many cblks in one launch,
// runs of thin levels in one CTA.  The right-hand sides are addressed as v[row * rs + rhs * cs]: column-major
// (rs = 1, cs = ldx, the layout of sm2xtab) or, when every level is small and several right-hand sides are
// solved together, transposed work copies (rs = nrhs, cs = 1) so that the lanes of a warp — which then run
// over right-hand sides — touch contiguous memory.
template <class T>
struct SmallSolveWs {
  T W[PB200_SM_WMAX * PB200_SM_LDW];
  T Xa[PB200_SM_RMAX * PB200_SM_WMAX];
  T Y[32 * PB200_SM_WMAX];      // per right-hand side of the current chunk: solved block / accumulators
  int grow[PB200_SM_RMAX];
};

__device__ __forceinline__ void smem_atomic_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void smem_atomic_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void smem_atomic_add(cdouble *p, cdouble v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }
__device__ __forceinline__ void smem_atomic_add(cfloat *p, cfloat v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }

template <class T>
__device__ __forceinline__ void small_stage(const DevSym &S, const T *M, int c, const int *__restrict__ rowglob,
                                            const int64_t *__restrict__ rmbase, SmallSolveWs<T> &ws, int lane, int &w, int &mr, int &fcol) {
  w = S.width[c]; fcol = S.fcol[c];
  const int ld = S.stride[c];
  mr = ld - w;
  const T *P = M + S.poff[c];
  for (int e = lane; e < w * w; e += 32) { const int j = e / w, i = e % w; ws.W[j * PB200_SM_LDW + i] = P[(size_t)j * ld + i]; }
  for (int e = lane; e < mr * w; e += 32) { const int k = e / mr, r = e % mr; ws.Xa[r * PB200_SM_WMAX + k] = P[(size_t)k * ld + w + r]; }
  const int64_t rb = rmbase[c];
  for (int r = lane; r < mr; r += 32) ws.grow[r] = rowglob[rb + r];
}

// down: x_c <- L_cc^-1 x_c (unit unless LLt), y_c = (D^-1) x_c, x[rows] -= L[rows, c] x_c     (updo.c:574-793, 948-984)
template <class T, int FACTO>
__device__ void small_fwd(const DevSym &S, const T *L, T *x, T *y, int64_t rs, int64_t cs, int nrhs, int c,
                          const int *__restrict__ rowglob, const int64_t *__restrict__ rmbase, SmallSolveWs<T> &ws, int lane) {
  int w, mr, fcol;
  small_stage<T>(S, L, c, rowglob, rmbase, ws, lane, w, mr, fcol);
  __syncwarp();
  for (int r0 = 0; r0 < nrhs; r0 += 32) {
    const int nrc = min(32, nrhs - r0);
    if (lane < nrc) {
      T xv[PB200_SM_WMAX];
#pragma unroll
      for (int k = 0; k < PB200_SM_WMAX; ++k) xv[k] = (k < w) ? ld_cg(&x[(int64_t)(fcol + k) * rs + (r0 + lane) * cs]) : ST<T>::zero();
#pragma unroll
      for (int j = 0; j < PB200_SM_WMAX; ++j) {
        if (j < w) {
          if (FACTO == F_LLT) xv[j] = xv[j] / ws.W[j * PB200_SM_LDW + j];
#pragma unroll
          for (int l = j + 1; l < PB200_SM_WMAX; ++l)
            if (l < w) xv[l] = xv[l] - ws.W[j * PB200_SM_LDW + l] * xv[j];
        }
      }
#pragma unroll
      for (int k = 0; k < PB200_SM_WMAX; ++k)
        if (k < w) {
          ws.Y[lane * PB200_SM_WMAX + k] = xv[k];
          y[(int64_t)(fcol + k) * rs + (r0 + lane) * cs] =
              (FACTO == F_LDLT || FACTO == F_LDLH) ? xv[k] / ws.W[k * PB200_SM_LDW + k] : xv[k];
        }
    }
    __syncwarp();
    for (int it = lane; it < mr * nrc; it += 32) {
      const int m = it / nrc, r = it % nrc;
      T s = ST<T>::zero();
      for (int k = 0; k < w; ++k) fma_acc(s, ws.Xa[m * PB200_SM_WMAX + k], ws.Y[r * PB200_SM_WMAX + k]);
      atomic_sub(&x[(int64_t)ws.grow[m] * rs + (r0 + r) * cs], s);
    }
    __syncwarp();
  }
}

// up: y_c -= op(M[rows, c])^T x[rows], x_c <- op(M_cc)^-T y_c     (updo_sendrecv.c:496-639, updo.c:1309-1342)
template <class T, int FACTO>
__device__ void small_bwd(const DevSym &S, const T *M, T *x, T *y, int64_t rs, int64_t cs, int nrhs, int c,
                          const int *__restrict__ rowglob, const int64_t *__restrict__ rmbase, SmallSolveWs<T> &ws, int lane) {
  constexpr bool CONJ = (FACTO == F_LDLH);
  constexpr bool UNIT = (FACTO == F_LDLT || FACTO == F_LDLH);
  int w, mr, fcol;
  small_stage<T>(S, M, c, rowglob, rmbase, ws, lane, w, mr, fcol);
  __syncwarp();
  for (int r0 = 0; r0 < nrhs; r0 += 32) {
    const int nrc = min(32, nrhs - r0);
    for (int e = lane; e < 32 * PB200_SM_WMAX; e += 32) ws.Y[e] = ST<T>::zero();
    __syncwarp();
    for (int it = lane; it < mr * nrc; it += 32) {
      const int m = it / nrc, r = it % nrc;
      const T xv = ld_cg(&x[(int64_t)ws.grow[m] * rs + (r0 + r) * cs]);
      for (int k = 0; k < w; ++k) {
        T a = ws.Xa[m * PB200_SM_WMAX + k];
        if (CONJ) a = ST<T>::conj(a);
        smem_atomic_add(&ws.Y[r * PB200_SM_WMAX + k], ST<T>::zero() - a * xv);
      }
    }
    __syncwarp();
    if (lane < nrc) {
      T yv[PB200_SM_WMAX];
#pragma unroll
      for (int k = 0; k < PB200_SM_WMAX; ++k)
        yv[k] = (k < w) ? ld_cg(&y[(int64_t)(fcol + k) * rs + (r0 + lane) * cs]) + ws.Y[lane * PB200_SM_WMAX + k] : ST<T>::zero();
#pragma unroll
      for (int j = PB200_SM_WMAX - 1; j >= 0; --j) {
        if (j < w) {
          T xj = yv[j];
#pragma unroll
          for (int l = j + 1; l < PB200_SM_WMAX; ++l)
            if (l < w) { T a = ws.W[j * PB200_SM_LDW + l]; if (CONJ) a = ST<T>::conj(a); xj = xj - a * yv[l]; }
          if (!UNIT) xj = xj / ws.W[j * PB200_SM_LDW + j];
          yv[j] = xj;
          x[(int64_t)(fcol + j) * rs + (r0 + lane) * cs] = xj;
        }
      }
    }
    __syncwarp();
  }
}

template <class T, int FACTO, int DIR>
__global__ void __launch_bounds__(PB200_SM_WARPS * 32)
k_small_solve_level(DevSym S, const T *M, T *x, T *y, int64_t rs, int64_t cs, int nrhs, const int *__restrict__ cblks, int ncblk,
                    const int *__restrict__ rowglob, const int64_t *__restrict__ rmbase) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallSolveWs<T> *ws = reinterpret_cast<SmallSolveWs<T> *>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = blockIdx.x * PB200_SM_WARPS + warp;
  if (wi >= ncblk) return;
  if (DIR == 0) small_fwd<T, FACTO>(S, M, x, y, rs, cs, nrhs, cblks[wi], rowglob, rmbase, ws[warp], lane);
  else small_bwd<T, FACTO>(S, M, x, y, rs, cs, nrhs, cblks[wi], rowglob, rmbase, ws[warp], lane);
}

// levels lvl_ptr[0..nlev] walked upwards (DIR 0) or downwards (DIR 1) by one CTA
template <class T, int FACTO, int DIR>
__global__ void __launch_bounds__(SmChain<T>::WARPS * 32)
k_small_solve_chain(DevSym S, const T *M, T *x, T *y, int64_t rs, int64_t cs, int nrhs, const int *__restrict__ cblks,
                    const int *__restrict__ lvl_ptr, int nlev, const int *__restrict__ rowglob, const int64_t *__restrict__ rmbase) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmallSolveWs<T> *ws = reinterpret_cast<SmallSolveWs<T> *>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qq = 0; qq < nlev; ++qq) {
    const int q = (DIR == 0) ? qq : nlev - 1 - qq;
    const int i0 = lvl_ptr[q], i1 = lvl_ptr[q + 1];
    for (int wi = i0 + warp; wi < i1; wi += SmChain<T>::WARPS) {
      if (DIR == 0) small_fwd<T, FACTO>(S, M, x, y, rs, cs, nrhs, cblks[wi], rowglob, rmbase, ws[warp], lane);
      else small_bwd<T, FACTO>(S, M, x, y, rs, cs, nrhs, cblks[wi], rowglob, rmbase, ws[warp], lane);
    }
    __threadfence();
    __syncthreads();
  }
}

// column-major (ld) <-> row-major work copy of the right-hand sides
template <class T>
__global__ void k_rhs_transpose(const T *__restrict__ src, T *__restrict__ dst, int n, int nrhs, int64_t ld, int to_rowmajor) {
  __shared__ T tile[32][33];
  const int i0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (to_rowmajor) {
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
      const int i = i0 + threadIdx.x, r = r0 + rr;
      if (i < n && r < nrhs) tile[rr][threadIdx.x] = src[(size_t)r * ld + i];
    }
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
      const int i = i0 + ii, r = r0 + threadIdx.x;
      if (i < n && r < nrhs) dst[(size_t)i * nrhs + r] = tile[threadIdx.x][ii];
    }
  } else {
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
      const int i = i0 + ii, r = r0 + threadIdx.x;
      if (i < n && r < nrhs) tile[ii][threadIdx.x] = src[(size_t)i * nrhs + r];
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
      const int i = i0 + threadIdx.x, r = r0 + rr;
      if (i < n && r < nrhs) dst[(size_t)r * ld + i] = tile[threadIdx.x][rr];
    }
  }
}

}  // namespace pb200
