/*
 * TEST INFRASTRUCTURE — the CPU oracle.  Not part of the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference leg may
 * load this.  Parity is PINNED: tests/test_oracle.py checks every
 * function below against the unmodified reference built in oracle/_ref and
 * against the committed golden dumps under tests/golden/.
 *
 * A plain-C, sequential restatement of the reference's sopalin numeric phase
 * on flat arrays.  Compiled four times (-DPREC_S/D/C/Z) into
 * oracle/libsopalin_oracle.so with prefixes s_/d_/c_/z_.
 *
 * Reference algorithm followed (all under /root/reference/src/sopalin/src):
 *   assemble      Csc2solv_cblk            csc_intern_solve.c:65-125
 *   norm1         CscNorm1                 csc_intern_compute.c:120-176
 *   potrf/sytrf/hetrf/getrf (+64-blocked)  compute_diag.c:124-518
 *   factor_diag / DimTrans                 compute_diag.c:521-605
 *   kernel_trsm / factor_trsm1d            compute_trsm.c:40-171
 *   compute_contrib_compact                sopalin_compute.c:270-374
 *   add_contrib_local (incl. ILU clipping) sopalin_compute.c:391-598
 *   compute_1d loop                        sopalin_compute.c:747-863
 *   up_down down/diag/up                   updo.c:574-793, 948-984, 1309-1342;
 *                                          updo_sendrecv.c:496-639
 * Panel layout (blend/src/solver.h:94-117): cblk c is a column-major
 * stride(c) x width(c) array; blok b starts at row offset coefind(b).
 * Here all panels live in one slab; panel c starts at poff[c] =
 * sum_{k<c} stride(k)*width(k).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>

#if defined(PREC_S)
typedef float T; typedef float R;
#define PFX(x) s_##x
#define ABS(x) fabsf(x)
#define CONJ(x) (x)
#define SQRT(x) sqrtf(x)
#elif defined(PREC_D)
typedef double T; typedef double R;
#define PFX(x) d_##x
#define ABS(x) fabs(x)
#define CONJ(x) (x)
#define SQRT(x) sqrt(x)
#elif defined(PREC_C)
typedef float complex T; typedef float R;
#define PFX(x) c_##x
#define ABS(x) cabsf(x)
#define CONJ(x) conjf(x)
#define SQRT(x) csqrtf(x)
#define CPLX 1
#elif defined(PREC_Z)
typedef double complex T; typedef double R;
#define PFX(x) z_##x
#define ABS(x) cabs(x)
#define CONJ(x) conj(x)
#define SQRT(x) csqrt(x)
#define CPLX 1
#else
#error "define PREC_S/D/C/Z"
#endif

enum { FACT_LLT = 0, FACT_LDLT = 1, FACT_LU = 2, FACT_LDLH = 3 }; /* api.h:381-384 */
#define MAXSIZEOFBLOCKS 64 /* compute_diag.c:46 */

typedef struct {
  int64_t cblknbr, bloknbr;
  const int64_t *fcol, *lcol, *bloknum, *stride; /* cblknbr+1 (bloknum) */
  const int64_t *frow, *lrow, *fcblk, *coefind;  /* bloknbr */
  const int64_t *poff;                           /* cblknbr+1 panel offsets */
} osolver;

void PFX(oracle_panel_offsets)(int64_t cblknbr, const int64_t *fcol, const int64_t *lcol,
                               const int64_t *stride, int64_t *poff)
{
  int64_t c; poff[0] = 0;
  for (c = 0; c < cblknbr; c++) poff[c + 1] = poff[c] + stride[c] * (lcol[c] - fcol[c] + 1);
}

double PFX(oracle_norm1)(int64_t n, const int64_t *colptr, const T *vals)
{
  double themax = 0; int64_t j, p;
  for (j = 0; j < n; j++) {
    double s = 0;
    for (p = colptr[j]; p < colptr[j + 1]; p++) s += ABS(vals[p]);
    if (s > themax) themax = s;
  }
  return themax;
}

/* Csc2solv_cblk: scatter the permuted CSC into the (zeroed) panels */
int64_t PFX(oracle_assemble)(const osolver *s, const int64_t *colptr, const int64_t *rows,
                             const T *vals, const T *tvals, int herm, T *L, T *U)
{
  int64_t c, j, p, b, dropped = 0;
  int64_t coefnbr = s->poff[s->cblknbr];
  memset(L, 0, (size_t)coefnbr * sizeof(T));
  if (U) memset(U, 0, (size_t)coefnbr * sizeof(T));
  for (c = 0; c < s->cblknbr; c++) {
    for (j = s->fcol[c]; j <= s->lcol[c]; j++) {
      for (p = colptr[j]; p < colptr[j + 1]; p++) {
        int64_t r = rows[p];
        if (r < s->fcol[c]) continue;
        b = s->bloknum[c];
        while (b < s->bloknum[c + 1] && (s->lrow[b] < r || s->frow[b] > r)) b++;
        if (b < s->bloknum[c + 1]) {
          int64_t idx = s->poff[c] + s->coefind[b] + (r - s->frow[b]) + s->stride[c] * (j - s->fcol[c]);
          L[idx] = vals[p];
          if (U && tvals && b != s->bloknum[c]) U[idx] = herm ? CONJ(tvals[p]) : tvals[p];
        } else dropped++; /* ILU: entry outside the incomplete pattern */
      }
    }
  }
  return dropped;
}

/* ---- dense kernels with static pivoting (compute_diag.c) ---- */
static void pivot_fix(T *d, double crit, int64_t *nbpivot)
{
  if (ABS(*d) < crit) { *d = (T)crit; (*nbpivot)++; }
}

static int potrf_unb(T *A, int64_t n, int64_t ld, int64_t *nbpivot, double crit)
{
  int64_t k, i, j;
  for (k = 0; k < n; k++) {
    T *d = A + k * (ld + 1);
    pivot_fix(d, crit, nbpivot);
    *d = SQRT(*d);
#ifndef CPLX
    if (*d < 0) return 1; /* "Negative diagonal term" (compute_diag.c:143-147); NB sqrt(<0)=NaN never trips it */
#endif
    { T inv = (T)1 / *d; for (i = k + 1; i < n; i++) A[k * ld + i] *= inv; }
    for (j = k + 1; j < n; j++)           /* SYR 'L' (complex: symmetric, geru) */
      for (i = j; i < n; i++) A[j * ld + i] -= A[k * ld + i] * A[k * ld + j];
  }
  return 0;
}

static void sytrf_unb(T *A, int64_t n, int64_t ld, int64_t *nbpivot, double crit, int herm)
{
  int64_t k, i, j;
  for (k = 0; k < n; k++) {
    T *d = A + k * (ld + 1);
    pivot_fix(d, crit, nbpivot);
    { T inv = (T)1 / *d; for (i = k + 1; i < n; i++) A[k * ld + i] *= inv; }
    for (j = k + 1; j < n; j++)           /* SYR/HER 'L' with alpha = -d */
      for (i = j; i < n; i++)
        A[j * ld + i] -= (*d) * A[k * ld + i] * (herm ? CONJ(A[k * ld + j]) : A[k * ld + j]);
  }
}

static void getrf_unb(T *A, int64_t m, int64_t n, int64_t ld, int64_t *nbpivot, double crit)
{
  int64_t k, i, j, mn = m < n ? m : n;
  for (k = 0; k < mn; k++) {
    T *d = A + k * (ld + 1);
    pivot_fix(d, crit, nbpivot);
    { T inv = (T)1 / *d; for (i = k + 1; i < m; i++) A[k * ld + i] *= inv; }
    if (k + 1 < mn)
      for (j = k + 1; j < n; j++)
        for (i = k + 1; i < m; i++) A[j * ld + i] -= A[k * ld + i] * A[j * ld + k];
  }
  pivot_fix(A + (n - 1) * (ld + 1), crit, nbpivot); /* compute_diag.c:461-467 */
}

/* 64-blocked right-looking drivers (PASTIX_*_block) */
static int potrf_block(T *A, int64_t n, int64_t ld, int64_t *nbpivot, double crit)
{
  int64_t k0, i, j, l;
  for (k0 = 0; k0 < n; k0 += MAXSIZEOFBLOCKS) {
    int64_t bs = (n - k0 < MAXSIZEOFBLOCKS) ? n - k0 : MAXSIZEOFBLOCKS, ms = n - k0 - bs;
    T *D = A + k0 * (ld + 1), *P = D + bs, *S = P + ld * bs;
    if (potrf_unb(D, bs, ld, nbpivot, crit)) return 1;
    if (ms <= 0) continue;
    /* TRSM R,L,T,N : P <- P * D^{-T} */
    for (j = 0; j < bs; j++) {
      for (l = 0; l < j; l++) for (i = 0; i < ms; i++) P[j * ld + i] -= P[l * ld + i] * D[l * ld + j];
      for (i = 0; i < ms; i++) P[j * ld + i] /= D[j * ld + j];
    }
    /* SYRK 'L' (complex: zherk, sopalin_compute.h:178-179): S -= P P^H, lower only */
    for (j = 0; j < ms; j++) for (l = 0; l < bs; l++) for (i = j; i < ms; i++)
      S[j * ld + i] -= P[l * ld + i] * CONJ(P[l * ld + j]);
#ifdef CPLX
    /* zherk treats S as Hermitian: with beta = 1 it starts from DBLE(S(j,j)) and adds real products only, so the
     * imaginary part of every trailing diagonal entry is dropped (reference BLAS zherk.f, "C(J,J) = DBLE(C(J,J))") */
    for (j = 0; j < ms; j++) S[j * ld + j] = (T)(__real__ S[j * ld + j]);
#endif
  }
  return 0;
}

static void sytrf_block(T *A, int64_t n, int64_t ld, int64_t *nbpivot, double crit, int herm)
{
  int64_t k0, i, j, l;
  T *W = (T *)malloc(sizeof(T) * (size_t)(n > 0 ? n : 1) * MAXSIZEOFBLOCKS);
  for (k0 = 0; k0 < n; k0 += MAXSIZEOFBLOCKS) {
    int64_t bs = (n - k0 < MAXSIZEOFBLOCKS) ? n - k0 : MAXSIZEOFBLOCKS, ms = n - k0 - bs;
    T *D = A + k0 * (ld + 1), *P = D + bs, *S = P + ld * bs;
    sytrf_unb(D, bs, ld, nbpivot, crit, herm);
    if (ms <= 0) continue;
    /* TRSM R,L,T|C,U then copy (=L*D) and scale by 1/d */
    for (j = 0; j < bs; j++) {
      for (l = 0; l < j; l++) {
        T f = herm ? CONJ(D[l * ld + j]) : D[l * ld + j];
        for (i = 0; i < ms; i++) P[j * ld + i] -= P[l * ld + i] * f;
      }
    }
    for (j = 0; j < bs; j++) {
      T inv = (T)1 / D[j * (ld + 1)];
      for (i = 0; i < ms; i++) { W[j * ms + i] = P[j * ld + i]; P[j * ld + i] *= inv; }
    }
    /* GEMM N,T|C : S -= W * P^T (full square, as the reference) */
    for (j = 0; j < ms; j++) for (l = 0; l < bs; l++) {
      T f = herm ? CONJ(P[l * ld + j]) : P[l * ld + j];
      for (i = 0; i < ms; i++) S[j * ld + i] -= W[l * ms + i] * f;
    }
  }
  free(W);
}

static void getrf_block(T *A, int64_t n, int64_t ld, int64_t *nbpivot, double crit)
{
  int64_t k0, i, j, l;
  for (k0 = 0; k0 < n; k0 += MAXSIZEOFBLOCKS) {
    int64_t bs = (n - k0 < MAXSIZEOFBLOCKS) ? n - k0 : MAXSIZEOFBLOCKS, ms = n - k0 - bs;
    T *D = A + k0 * (ld + 1), *P = D + bs, *Q = D + ld * bs, *S = D + (ld + 1) * bs;
    getrf_unb(D, n - k0, bs, ld, nbpivot, crit);
    if (ms <= 0) continue;
    /* TRSM L,L,N,U : Q <- D_L^{-1} Q */
    for (j = 0; j < ms; j++) for (l = 0; l < bs; l++) for (i = l + 1; i < bs; i++)
      Q[j * ld + i] -= D[l * ld + i] * Q[j * ld + l];
    /* GEMM N,N : S -= P * Q */
    for (j = 0; j < ms; j++) for (l = 0; l < bs; l++) for (i = 0; i < ms; i++)
      S[j * ld + i] -= P[l * ld + i] * Q[j * ld + l];
  }
}

/* find the blok of cblk c that contains global row r (or -1) */
static int64_t find_blok(const osolver *s, int64_t c, int64_t r)
{
  int64_t b;
  for (b = s->bloknum[c]; b < s->bloknum[c + 1]; b++)
    if (s->frow[b] <= r && r <= s->lrow[b]) return b;
  return -1;
}

/*
 * Factorization: cblks in index order (a valid sequential schedule: every
 * contributor of c has a smaller index), each as compute_1d does.
 * Returns 0, or 1 on a negative pivot in real LLt.
 */
static int factorize_impl(const osolver *s, int facto, T *L, T *U, double crit, int64_t *nbpivot, int schur)
{
  int64_t c, b1, b2, i, j, l;
  int herm = (facto == FACT_LDLH);
  int64_t wmax = 0, smax = 0;
  T *W1, *W2, *LD;
  *nbpivot = 0;
  for (c = 0; c < s->cblknbr; c++) {
    int64_t w = s->lcol[c] - s->fcol[c] + 1;
    if (w > wmax) wmax = w;
    if (s->stride[c] > smax) smax = s->stride[c];
  }
  W1 = (T *)malloc(sizeof(T) * (size_t)(smax * wmax + 1));
  W2 = (T *)malloc(sizeof(T) * (size_t)(smax * wmax + 1));
  LD = (T *)malloc(sizeof(T) * (size_t)(smax * wmax + 1));
  for (c = 0; c < s->cblknbr; c++) {
    int64_t w = s->lcol[c] - s->fcol[c] + 1, ld = s->stride[c], m = ld - w;
    int64_t fb = s->bloknum[c], lb = s->bloknum[c + 1];
    T *A = L + s->poff[c];            /* diag blok has coefind 0 */
    /* IPARM_SCHUR: the cblk holding the last column is left as assembled + updated = the Schur complement
       (compute_1d returns at once, sopalin_compute.c:767-772) */
    if (schur && c == s->cblknbr - 1) continue;
    T *P = A + w;                     /* off-diagonal panel */
    T *UA = U ? U + s->poff[c] : NULL, *UP = UA ? UA + w : NULL;
    /* factor_diag */
    if (facto == FACT_LLT) { if (potrf_block(A, w, ld, nbpivot, crit)) { free(W1); free(W2); free(LD); return 1; } }
    else if (facto == FACT_LU) {
      getrf_block(A, w, ld, nbpivot, crit);
      for (i = 0; i < w; i++) for (j = 0; j < w; j++) UA[i * ld + j] = A[j * ld + i]; /* DimTrans */
    } else sytrf_block(A, w, ld, nbpivot, crit, herm);
    if (m <= 0) continue;
    /* factor_trsm1d / kernel_trsm */
    if (facto == FACT_LLT) {          /* R,L,T,N */
      for (j = 0; j < w; j++) {
        for (l = 0; l < j; l++) for (i = 0; i < m; i++) P[j * ld + i] -= P[l * ld + i] * A[l * ld + j];
        for (i = 0; i < m; i++) P[j * ld + i] /= A[j * ld + j];
      }
    } else if (facto == FACT_LU) {
      /* L <- L * U_kk^{-1}  (R,U,N,N on coeftab diag) */
      for (j = 0; j < w; j++) {
        for (l = 0; l < j; l++) for (i = 0; i < m; i++) P[j * ld + i] -= P[l * ld + i] * A[j * ld + l];
        for (i = 0; i < m; i++) P[j * ld + i] /= A[j * ld + j];
      }
      /* U^T <- U^T * (L_kk^T)^{-1}  (R,U,N,U on ucoeftab diag = (LU)^T) */
      for (j = 0; j < w; j++)
        for (l = 0; l < j; l++) for (i = 0; i < m; i++) UP[j * ld + i] -= UP[l * ld + i] * UA[j * ld + l];
    } else {                          /* R,L,T|C,U then LD copy and scale */
      for (j = 0; j < w; j++)
        for (l = 0; l < j; l++) {
          T f = herm ? CONJ(A[l * ld + j]) : A[l * ld + j];
          for (i = 0; i < m; i++) P[j * ld + i] -= P[l * ld + i] * f;
        }
      for (j = 0; j < w; j++) {
        T inv = (T)1 / A[j * (ld + 1)];
        for (i = 0; i < m; i++) { LD[j * m + i] = P[j * ld + i]; P[j * ld + i] *= inv; }
      }
    }
    /* compute_1dgemm for every off-diagonal blok b1 */
    for (b1 = fb + 1; b1 < lb; b1++) {
      int64_t dimj = s->lrow[b1] - s->frow[b1] + 1, dimi = ld - s->coefind[b1];
      int64_t r0 = s->coefind[b1];    /* panel row offset of b1 */
      int64_t fc = s->fcblk[b1];      /* facing cblk */
      const T *Ai = A + r0;           /* rows b1..end, ld */
      /* compute_contrib_compact: W2 = A_i * B^T ; LU: W1 = U_i * L_b1^T */
      for (j = 0; j < dimj; j++) for (i = 0; i < dimi; i++) { W2[j * dimi + i] = 0; W1[j * dimi + i] = 0; }
      for (l = 0; l < w; l++) for (j = 0; j < dimj; j++) {
        T bj;
        if (facto == FACT_LLT) bj = CONJ(Ai[l * ld + j]);                       /* GEMM N,C */
        else if (facto == FACT_LU) bj = UA[r0 + l * ld + j];                    /* GEMM N,T with U */
        else bj = herm ? CONJ(LD[l * m + (r0 - w) + j]) : LD[l * m + (r0 - w) + j]; /* L*D workspace */
        for (i = 0; i < dimi; i++) W2[j * dimi + i] += Ai[l * ld + i] * bj;
        if (facto == FACT_LU) {
          T lj = Ai[l * ld + j];
          for (i = 0; i < dimi; i++) W1[j * dimi + i] += UA[r0 + l * ld + i] * lj;
        }
      }
      if (fc < 0) continue;
      /* add_contrib_local, row by row (covers the ILU partial-overlap loop) */
      {
        T *TL = L + s->poff[fc], *TU = U ? U + s->poff[fc] : NULL;
        int64_t tld = s->stride[fc], tfcol = s->fcol[fc], dblok = s->bloknum[fc];
        int64_t step = 0;
        for (b2 = b1; b2 < lb; b2++) {
          int64_t nr = s->lrow[b2] - s->frow[b2] + 1;
          for (i = 0; i < nr; i++) {
            int64_t r = s->frow[b2] + i, b3 = find_blok(s, fc, r), ro;
            if (b3 < 0) continue;     /* ILU: no facing blok, contribution dropped */
            ro = s->coefind[b3] + (r - s->frow[b3]);
            for (j = 0; j < dimj; j++) {
              int64_t cj = s->frow[b1] + j - tfcol;
              TL[ro + cj * tld] -= W2[j * dimi + step + i];
              if (facto == FACT_LU) {
                if (b3 != dblok) TU[ro + cj * tld] -= W1[j * dimi + step + i];
                else if (b1 != b2) TL[cj + (r - tfcol) * tld] -= W1[j * dimi + step + i];
              }
            }
          }
          step += nr;
        }
      }
    }
  }
  free(W1); free(W2); free(LD);
  return 0;
}

int PFX(oracle_factorize)(const osolver *s, int facto, T *L, T *U, double crit, int64_t *nbpivot)
{
  return factorize_impl(s, facto, L, U, crit, nbpivot, 0);
}
int PFX(oracle_factorize_schur)(const osolver *s, int facto, T *L, T *U, double crit, int64_t *nbpivot)
{
  return factorize_impl(s, facto, L, U, crit, nbpivot, 1);
}

/* number of positive diagonal terms of D (sopalin3d.c:1145-1161) */
int64_t PFX(oracle_inertia)(const osolver *s, const T *L)
{
  int64_t c, k, cnt = 0;
  for (c = 0; c < s->cblknbr; c++) {
    int64_t w = s->lcol[c] - s->fcol[c] + 1, ld = s->stride[c];
    for (k = 0; k < w; k++) {
#ifdef CPLX
      if (creal(L[s->poff[c] + k * (ld + 1)]) > 0) cnt++;
#else
      if (L[s->poff[c] + k * (ld + 1)] > 0) cnt++;
#endif
    }
  }
  return cnt;
}

/* up_down on x (n x nrhs, column-major, leading dimension ldx, permuted order) */
static void solve_impl(const osolver *s, int facto, const T *L, const T *U, T *x, int64_t ldx, int64_t nrhs, int schur)
{
  /* IPARM_SCHUR: the last cblk and every blok facing it are ignored by all three steps (updo.c:425-428, 639-646,
     1154-1180; updo_sendrecv.c:518-523): the interior system is solved, the Schur unknowns keep their right-hand side */
  const int64_t ncb = schur ? s->cblknbr - 1 : s->cblknbr, skip = schur ? s->cblknbr - 1 : -1;
  int64_t c, b, i, j, k;
  int herm = (facto == FACT_LDLH);
  int unit = (facto != FACT_LLT); /* updo.c:582-594: non-unit only for LLt */
  for (k = 0; k < nrhs; k++) {
    T *xk = x + k * ldx;
    /* DOWN */
    for (c = 0; c < ncb; c++) {
      int64_t w = s->lcol[c] - s->fcol[c] + 1, ld = s->stride[c];
      const T *A = L + s->poff[c];
      T *xc = xk + s->fcol[c];
      for (j = 0; j < w; j++) {
        if (!unit) xc[j] /= A[j * (ld + 1)];
        for (i = j + 1; i < w; i++) xc[i] -= A[j * ld + i] * xc[j];
      }
      for (b = s->bloknum[c] + 1; b < s->bloknum[c + 1]; b++) {
        int64_t nr = s->lrow[b] - s->frow[b] + 1;
        const T *B = A + s->coefind[b];
        T *xt = xk + s->frow[b];
        if (s->fcblk[b] == skip) continue;
        for (j = 0; j < w; j++) for (i = 0; i < nr; i++) xt[i] -= B[j * ld + i] * xc[j];
      }
    }
    /* DIAG (LDLt / LDLh) */
    if (facto == FACT_LDLT || facto == FACT_LDLH)
      for (c = 0; c < ncb; c++) {
        int64_t w = s->lcol[c] - s->fcol[c] + 1, ld = s->stride[c];
        for (j = 0; j < w; j++) xk[s->fcol[c] + j] /= L[s->poff[c] + j * (ld + 1)];
      }
    /* UP */
    for (c = ncb - 1; c >= 0; c--) {
      int64_t w = s->lcol[c] - s->fcol[c] + 1, ld = s->stride[c];
      const T *A = ((facto == FACT_LU) ? U : L) + s->poff[c];
      T *xc = xk + s->fcol[c];
      for (b = s->bloknum[c + 1] - 1; b > s->bloknum[c]; b--) {
        int64_t nr = s->lrow[b] - s->frow[b] + 1;
        const T *B = A + s->coefind[b];
        const T *xt = xk + s->frow[b];
        if (s->fcblk[b] == skip) continue;
        for (j = 0; j < w; j++) {
          T acc = 0;
          for (i = 0; i < nr; i++) acc += (herm ? CONJ(B[j * ld + i]) : B[j * ld + i]) * xt[i];
          xc[j] -= acc;
        }
      }
      /* TRSV L,T|C, N (LLt, LU on ucoeftab) or U (LDLt/LDLh) */
      for (j = w - 1; j >= 0; j--) {
        T acc = xc[j];
        for (i = j + 1; i < w; i++) acc -= (herm ? CONJ(A[j * ld + i]) : A[j * ld + i]) * xc[i];
        if (facto == FACT_LLT || facto == FACT_LU) acc /= A[j * (ld + 1)];
        xc[j] = acc;
      }
    }
  }
}

void PFX(oracle_solve)(const osolver *s, int facto, const T *L, const T *U, T *x, int64_t ldx, int64_t nrhs)
{
  solve_impl(s, facto, L, U, x, ldx, nrhs, 0);
}
void PFX(oracle_solve_schur)(const osolver *s, int facto, const T *L, const T *U, T *x, int64_t ldx, int64_t nrhs)
{
  solve_impl(s, facto, L, U, x, ldx, nrhs, 1);
}
