/*
 * sopalin_b200_shim.c — the drop-in: PaStiX's numeric-phase entry points on the B200.
 *
 * Compiled against the UNMODIFIED reference headers (read where they lie under
 * $PASTIX_REFERENCE/src, nothing is copied) and linked in place of the reference's
 * sopalin3d.o, four times like it (-DCHOL_SOPALIN / -DSOPALIN_LU / none / -DHERMITIAN give the
 * po_/ge_/sy_/he_ variants, src/CMakeLists.txt:40-62; precision prefix S_/D_/C_/Z_ from
 * common/src/redefine_functions.h:84-101).  Everything else of libpastix — pastix(),
 * pastix_fortran(), ordering, fax/kass, blend, the internal CSC, refinement — stays the
 * reference's own code, so iparm/dparm and the API_TASK_* semantics are untouched.
 *
 * Entry points replaced (reference: src/sopalin/src/sopalin3d.c):
 *   API_CALL(sopalin_thread)        :1388   numeric factorization        (API_TASK_NUMFACT)
 *   API_CALL(sopalin_updo_thread)   :1467   factorization + up_down      (NUMFACT..SOLVE)
 *   API_CALL(updo_thread)           updo.c:67  up_down                   (API_TASK_SOLVE)
 *   API_CALL(up_down_smp)           updo.c:114 up_down inside the reference's thread pool — the
 *                                   preconditioner call of the refinement drivers
 *                                   (raff_functions.c:408-437)
 *   API_CALL(sopalin_updo_{gmres,grad,pivot,bicgstab}_thread) :1549-1900 = ours + the reference's
 *                                   host refinement loop (raff_*.c, included below as sopalin3d.c does)
 * What it does: flattens the SolverMatrix (blend/src/solver.h:94-168) and the internal CSC
 * (blend/src/csc.h) into the plain arrays of include/pastix_b200.h, computes the static-pivot
 * threshold exactly as init_struct_sopalin (sopalin3d.c:586-606), runs the CUDA layer, and writes
 * back sopar->diagchange, DPARM_FACT_TIME, DPARM_SOLV_TIME, IPARM_INERTIA and the solution in
 * updovct.sm2xtab.  Factors stay resident in HBM, keyed by SolverMatrix*; set PB200_HOST_COEFTAB=1
 * to also mirror them into cblktab[].coeftab/ucoeftab (malloc'ed like CoefMatrix_Allocate,
 * coefinit.c:104, so CoefMatrix_Free keeps working) for dump consumers.  With IPARM_SCHUR the last
 * cblk's panel (= the Schur complement) is always copied back, for pastix_getSchur.
 * No CPU fallback: a CUDA failure is a fatal error (errorPrint + EXIT, like the reference's own
 * fatal paths, common/src/errors.h:161-165).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <assert.h>
#include <pthread.h>
#include <math.h>
#include <stdint.h>

#ifdef FORCE_NOMPI
#include "nompi.h"
#else
#include <mpi.h>
#endif
#include <signal.h>
#include "common_pastix.h"
#include "tools.h"
#include "trace.h"
#include "sopalin_define.h"
#include "symbol.h"
#include "ftgt.h"
#include "csc.h"
#include "updown.h"
#include "queue.h"
#include "bulles.h"
#include "solver.h"
#include "sopalin_thread.h"
#include "stack.h"
#include "sopalin3d.h"
#include "sopalin_init.h"
#include "perf.h"
#include "out.h"
#include "coefinit.h"
#include "ooc.h"
#include "order.h"
#include "debug_dump.h"
#include "sopalin_acces.h"
#include "csc_intern_compute.h"

#include "pastix_b200.h"

#if defined(TYPE_COMPLEX) && defined(PREC_DOUBLE)
#define PB200_FLT PB200_COMPLEXDOUBLE
#elif defined(TYPE_COMPLEX)
#define PB200_FLT PB200_COMPLEXSINGLE
#elif defined(PREC_DOUBLE)
#define PB200_FLT PB200_REALDOUBLE
#else
#define PB200_FLT PB200_REALSINGLE
#endif

#if defined(SOPALIN_LU)
#define PB200_FACTO PB200_FACT_LU
#elif defined(CHOL_SOPALIN)
#define PB200_FACTO PB200_FACT_LLT
#elif defined(HERMITIAN)
#define PB200_FACTO PB200_FACT_LDLH
#else
#define PB200_FACTO PB200_FACT_LDLT
#endif

/* ---- side table: SolverMatrix* -> device handle (owned by shim_hooks.c) */
#include "shim_table.h"

static void shim_fatal(const char *what)
{
  errorPrint("pastix_b200: %s: %s", what, pb200_last_error());
  EXIT(MOD_SOPALIN, INTERNAL_ERR);
}

static pb200_shim_entry_t *shim_find(const SolverMatrix *m, int create) { return pb200_shim_entry(m, create); }

/* release the HBM held for one SolverMatrix (the reference's own release points call it: shim_hooks.c) */
#define pb200_shim_release PASTIX_PREFIX_F(API_CALL(pb200_shim_release))
void pb200_shim_release(const SolverMatrix *m) { pb200_shim_entry_drop(m); }

/* SolverMatrix -> flat arrays -> device handle */
static pb200_handle_t *shim_create(SolverMatrix *datacode, int schur, int rank, int nranks)
{
  pb200_solver_t s; pb200_handle_t *h = NULL; pb200_options_t opts;
  int64_t *buf, *fcol, *lcol, *bnum, *strd, *frow, *lrow, *fcb, *cind;
  PASTIX_INT i, C = SYMB_CBLKNBR, B = SYMB_BLOKNBR;
  buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)(4 * (C + 1) + 4 * B + 8));
  if (buf == NULL) { errorPrint("pastix_b200: out of memory"); EXIT(MOD_SOPALIN, OUTOFMEMORY_ERR); }
  fcol = buf; lcol = fcol + C + 1; bnum = lcol + C + 1; strd = bnum + C + 1;
  frow = strd + C + 1; lrow = frow + B; fcb = lrow + B; cind = fcb + B;
  for (i = 0; i < C; i++) {
    fcol[i] = SYMB_FCOLNUM(i); lcol[i] = SYMB_LCOLNUM(i); bnum[i] = SYMB_BLOKNUM(i); strd[i] = SOLV_STRIDE(i);
  }
  bnum[C] = SYMB_BLOKNUM(C);
  for (i = 0; i < B; i++) {
    frow[i] = SYMB_FROWNUM(i); lrow[i] = SYMB_LROWNUM(i); fcb[i] = SYMB_CBLKNUM(i); cind[i] = SOLV_COEFIND(i);
  }
  s.cblknbr = C; s.bloknbr = B;
  s.fcolnum = fcol; s.lcolnum = lcol; s.bloknum = bnum; s.stride = strd;
  s.frownum = frow; s.lrownum = lrow; s.cblknum = fcb; s.coefind = cind;
  memset(&opts, 0, sizeof(opts));
  opts.schur = schur;
  /* PB200_DIST_MAP=blend: the GPUs take the reference's own proportional mapping — the thread blend assigned the
   * COMP_1D task of every column block to (SolverMatrix.ttsktab, blend/src/solver.h:158-159; built by the same
   * propMappTree / distribPart code that maps tasks to processes, splitpart.c:752-1012) — instead of the mapping
   * computed in the CUDA layer.  Meaningful when IPARM_THREAD_NBR was the number of GPUs at analysis time. */
  int32_t *own = NULL;
  if (nranks > 1 && getenv("PB200_DIST_MAP") != NULL && strcmp(getenv("PB200_DIST_MAP"), "blend") == 0 &&
      datacode->ttsktab != NULL && datacode->thrdnbr > 0) {
    PASTIX_INT t, k;
    own = (int32_t *)malloc(sizeof(int32_t) * (size_t)(C + 1));
    if (own == NULL) { errorPrint("pastix_b200: out of memory"); EXIT(MOD_SOPALIN, OUTOFMEMORY_ERR); }
    for (i = 0; i < C; i++) own[i] = -1;
    for (t = 0; t < datacode->thrdnbr; t++)
      for (k = 0; k < datacode->ttsknbr[t]; k++) {
        const Task *tk = &datacode->tasktab[datacode->ttsktab[t][k]];
        if ((tk->taskid == COMP_1D || tk->taskid == DIAG) && tk->cblknum >= 0 && tk->cblknum < C)
          own[tk->cblknum] = (int32_t)((t * nranks) / datacode->thrdnbr);
      }
    for (i = 0; i < C; i++) if (own[i] < 0) { free(own); own = NULL; break; }   /* not a complete 1-D task list: keep ours */
    opts.owner = own;
  }
  /* one GPU: the current device; a group: devices 0 .. nranks-1 of the box */
  if (pb200_create_opts(&h, &s, PB200_FLT, PB200_FACTO, nranks > 1 ? rank : -1, rank, nranks, &opts) != PB200_SUCCESS)
    shim_fatal("pb200_create_opts");
  if (own) free(own);
  free(buf);
  return h;
}

/* internal block-CSC (CscOrdistrib, csc_intern_build.c:352) -> flat 0-based colptr / rows (host) */
static void shim_flatten_csc(SopalinParam *sopar, int64_t **colptr_out, int64_t **rows_out)
{
  const CscMatrix *csc = sopar->cscmtx;
  PASTIX_INT i, j, ncol = 0, nnz = 0, col = 0;
  int64_t *colptr, *rows;
  for (i = 0; i < CSC_FNBR(csc); i++) { ncol += CSC_COLNBR(csc, i); nnz = CSC_COL(csc, i, CSC_COLNBR(csc, i)); }
  colptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(ncol + 1));
  rows   = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nnz + 1));
  if (colptr == NULL || rows == NULL) { errorPrint("pastix_b200: out of memory"); EXIT(MOD_SOPALIN, OUTOFMEMORY_ERR); }
  for (i = 0; i < CSC_FNBR(csc); i++)
    for (j = 0; j < CSC_COLNBR(csc, i); j++) colptr[col++] = CSC_COL(csc, i, j);
  colptr[col] = nnz;
  for (i = 0; i < nnz; i++) rows[i] = CSC_ROW(csc, i);
  *colptr_out = colptr; *rows_out = rows;
}

/* one rank's share of a numeric factorization: panels filled from the internal CSC (device copy when CscOrdistrib
 * left one in HBM, host arrays otherwise), then the factorization.  With several GPUs every rank runs this in its
 * own host thread: pb200_reassemble / pb200_factorize are collective (include/pastix_b200.h, pb200_attach_local). */
typedef struct shim_job_s {
  pb200_handle_t *h;
  pb200_csc_t    *devcsc;           /* NULL: host arrays below */
  const int64_t  *colptr, *rows;
  const void     *vals, *tvals;
  int             herm;
  double          crit;
  int64_t         nbpivot;
  double          seconds;
  int             rc;
  char            err[256];
} shim_job_t;

static void *shim_job_run(void *arg)
{
  shim_job_t *j = (shim_job_t *)arg;
  j->rc = PB200_SUCCESS; j->err[0] = 0;
  if (j->devcsc != NULL) j->rc = pb200_assemble_csc(j->h, j->devcsc);
  else {
    j->rc = pb200_set_hermitian(j->h, j->herm);
    if (j->rc == PB200_SUCCESS) j->rc = pb200_assemble(j->h, j->colptr, j->rows, j->vals, j->tvals);
  }
  if (j->rc == PB200_SUCCESS) j->rc = pb200_factorize(j->h, j->crit, &j->nbpivot, &j->seconds);
  if (j->rc != PB200_SUCCESS) { strncpy(j->err, pb200_last_error(), sizeof(j->err) - 1); j->err[sizeof(j->err) - 1] = 0; }
  return NULL;
}

/* static-pivot threshold, init_struct_sopalin (sopalin3d.c:586-606) */
static double shim_critere(SolverMatrix *datacode, SopalinParam *sopar, pb200_csc_t *devcsc)
{
  double crit = sopar->espilondiag;
  if (crit < 0.0) return -crit;
  if (sopar->usenocsc == 1) return crit;
  if (sopar->fakefact == 1)
    return (double)(UPDOWN_GNODENBR * UPDOWN_GNODENBR + UPDOWN_GNODENBR) * sqrt(sopar->espilondiag);
  if (devcsc != NULL) {                        /* same sums in the same order on the CSC already in HBM: identical for real
                                                  types; complex: |z| is the device hypot (within an ulp of cabs) */
    double nrm = 0.0;
    if (pb200_csc_norm1(devcsc, &nrm) != PB200_SUCCESS) shim_fatal("pb200_csc_norm1");
    return nrm * sqrt(sopar->espilondiag);
  }
  return CscNorm1(sopar->cscmtx, sopar->pastix_comm) * sqrt(sopar->espilondiag);
}

/* mirror the factors into cblktab[].coeftab / .ucoeftab (only when asked: PB200_HOST_COEFTAB=1) */
static void shim_mirror_coeftab(pb200_handle_t *h, SolverMatrix *datacode)
{
  PASTIX_INT c; size_t off = 0, total = 0;
  PASTIX_FLOAT *L, *U = NULL;
  for (c = 0; c < SYMB_CBLKNBR; c++) total += (size_t)SOLV_STRIDE(c) * (size_t)(SYMB_LCOLNUM(c) - SYMB_FCOLNUM(c) + 1);
  L = (PASTIX_FLOAT *)malloc(total * sizeof(PASTIX_FLOAT));
  if (PB200_FACTO == PB200_FACT_LU) U = (PASTIX_FLOAT *)malloc(total * sizeof(PASTIX_FLOAT));
  if (L == NULL || (PB200_FACTO == PB200_FACT_LU && U == NULL)) { errorPrint("pastix_b200: out of memory"); EXIT(MOD_SOPALIN, OUTOFMEMORY_ERR); }
  if (pb200_get_coeftab(h, L, U) != PB200_SUCCESS) shim_fatal("pb200_get_coeftab");
  for (c = 0; c < SYMB_CBLKNBR; c++) {
    size_t sz = (size_t)SOLV_STRIDE(c) * (size_t)(SYMB_LCOLNUM(c) - SYMB_FCOLNUM(c) + 1);
    if (SOLV_COEFTAB(c) == NULL) { MALLOC_INTERN(SOLV_COEFTAB(c), sz, PASTIX_FLOAT); }
    memcpy(SOLV_COEFTAB(c), L + off, sz * sizeof(PASTIX_FLOAT));
    if (U != NULL) {
      if (SOLV_UCOEFTAB(c) == NULL) { MALLOC_INTERN(SOLV_UCOEFTAB(c), sz, PASTIX_FLOAT); }
      memcpy(SOLV_UCOEFTAB(c), U + off, sz * sizeof(PASTIX_FLOAT));
    }
    off += sz;
  }
  free(L); if (U) free(U);
}

/* IPARM_SCHUR: the Schur complement is what the never-factored last cblk holds after the factorization.  The reference
 * leaves it in SOLV_COEFTAB(last cblk) — user memory when pastix_setSchurArray was called (pastix.c:3400-3412), else
 * allocated like every panel (coefinit.c:141-150) — where pastix_getSchur reads it (pastix.c:6434-6475). */
static void shim_fetch_schur(pb200_handle_t *h, SolverMatrix *datacode)
{
  PASTIX_INT c = SYMB_CBLKNBR - 1;
  size_t sz = (size_t)SOLV_STRIDE(c) * (size_t)(SYMB_LCOLNUM(c) - SYMB_FCOLNUM(c) + 1);
  if (SOLV_COEFTAB(c) == NULL) { MALLOC_INTERN(SOLV_COEFTAB(c), sz, PASTIX_FLOAT); }
  if (pb200_get_cblk(h, (int64_t)c, SOLV_COEFTAB(c), NULL) != PB200_SUCCESS) shim_fatal("pb200_get_cblk");
}

static void shim_numfact(SolverMatrix *datacode, SopalinParam *sopar)
{
  pb200_shim_entry_t *e = shim_find(datacode, 1);
  int64_t nbpivot = 0, *colptr = NULL, *rows = NULL; double seconds = 0.0, crit; int dev_csc = 0, r, G;
  shim_job_t jobs[8]; pthread_t thr[8];
  if (e == NULL) { errorPrint("pastix_b200: out of memory (side table)"); EXIT(MOD_SOPALIN, OUTOFMEMORY_ERR); }
  const int schur = (sopar->schur == API_YES);
  if (sopar->iparm[IPARM_DISTRIBUTION_LEVEL] != 0 || SOLV_PROCNBR > 1) {
    errorPrint("pastix_b200: 2D distribution / multi-process SolverMatrix are not handled by this shim");
    EXIT(MOD_SOPALIN, BADPARAMETER_ERR);
  }
  if (sopar->fakefact == API_YES) {             /* IPARM_FILL_MATRIX: synthetic coefficients (coefinit.c:343-437) */
    errorPrint("pastix_b200: IPARM_FILL_MATRIX (fake factorization) is not handled by this shim");
    EXIT(MOD_SOPALIN, BADPARAMETER_ERR);
  }
  /* iparm[IPARM_CUDA_NBR] (api.h:115-120, "number of cuda devices", default 0): the GPUs of the box this
   * factorization is spread over.  The Schur cblk lives on one process in the reference too: one GPU. */
  G = (int)sopar->iparm[IPARM_CUDA_NBR];
  if (G < 1 || schur) G = 1;
  if (G > 8) G = 8;
  {
  double t0 = clockGet(), t1, t2, t3;
  /* the handle must have been built for THIS structure: the key (the SolverMatrix address) survives a new
   * API_TASK_ANALYSE on the same pastix_data and can be reused by malloc after API_TASK_CLEAN */
  { uint64_t fp[2];
    pb200_shim_fingerprint(datacode, fp);
    if (e->h != NULL && (e->facto != PB200_FACTO || e->schur != schur || e->ngpu != G || e->fp[0] != fp[0] || e->fp[1] != fp[1])) {
      if (e->ngpu > 1) pb200_destroy_group(e->hs, e->ngpu); else pb200_destroy(e->h);
      e->h = NULL; memset(e->hs, 0, sizeof(e->hs)); e->ngpu = 0;
    }
    if (e->h == NULL) {
      for (r = 0; r < G; r++) e->hs[r] = shim_create(datacode, schur, r, G);
      if (G > 1 && pb200_attach_local(e->hs, G) != PB200_SUCCESS) shim_fatal("pb200_attach_local");
      e->h = e->hs[0]; e->ngpu = G;
      e->facto = PB200_FACTO; e->schur = schur; e->fp[0] = fp[0]; e->fp[1] = fp[1];
    } }
  e->factorized = 0;
  t1 = clockGet();
  if (e->csc != NULL && e->csc_fresh) {        /* CscOrdistrib of this call left the internal CSC in HBM (shim_csc.c) */
    dev_csc = 1;
    e->csc_fresh = 0;
  } else {
    pb200_shim_csc_host(datacode);
    shim_flatten_csc(sopar, &colptr, &rows);
  }
  t2 = clockGet();
  crit = shim_critere(datacode, sopar, dev_csc ? e->csc : NULL);
  t3 = clockGet();
  e->critere = crit;
  if (sopar->iparm[IPARM_VERBOSE] > API_VERBOSE_YES)
    fprintf(stdout, "Pivoting criterium (||A||*sqrt(epsilon)) = %g\n", crit);
  for (r = 0; r < G; r++) {
    memset(&jobs[r], 0, sizeof(jobs[r]));
    jobs[r].h = e->hs[r]; jobs[r].devcsc = dev_csc ? e->csc : NULL;
    jobs[r].colptr = colptr; jobs[r].rows = rows; jobs[r].vals = CSC_VALTAB(sopar->cscmtx); jobs[r].tvals = sopar->transcsc;
    jobs[r].herm = (sopar->cscmtx->type == 'H'); jobs[r].crit = crit;
  }
  for (r = 1; r < G; r++)
    if (pthread_create(&thr[r], NULL, shim_job_run, &jobs[r]) != 0) { errorPrint("pastix_b200: pthread_create failed"); EXIT(MOD_SOPALIN, INTERNAL_ERR); }
  shim_job_run(&jobs[0]);
  for (r = 1; r < G; r++) pthread_join(thr[r], NULL);
  for (r = 0; r < G; r++) {
    if (jobs[r].rc != PB200_SUCCESS) {
      errorPrint("pastix_b200: numeric factorization (GPU %d of %d): %s", r, G, jobs[r].err);
      EXIT(MOD_SOPALIN, INTERNAL_ERR);
    }
    nbpivot += jobs[r].nbpivot;                                  /* the reference's MPI_Allreduce, sopalin3d.c:1138 */
    if (jobs[r].seconds > seconds) seconds = jobs[r].seconds;
  }
  if (colptr) free(colptr);
  if (rows) free(rows);
  /* CoefMatrix_Init releases the transposed values once the panels are filled (coefinit.c:327-341) */
  if (sopar->transcsc != NULL) {
    if (PB200_FACTO == PB200_FACT_LU && (sopar->iparm[IPARM_SYM] == API_SYM_YES || sopar->iparm[IPARM_SYM] == API_SYM_HER))
      sopar->transcsc = NULL;                  /* alias of CSC_VALTAB (forcetrans) */
    else
      memFree_null(sopar->transcsc);
    e->lazy_tvals = NULL;                      /* a later on-demand copy of the internal CSC has nowhere to put them */
  }
  if (getenv("PB200_SHIM_TIMING") != NULL)
    fprintf(stderr, "[pb200 shim] create %.1f ms, CSC flatten %.1f ms, CscNorm1 %.1f ms, assembly + factorization on %d GPU(s) %.1f ms\n",
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, G, (clockGet() - t3) * 1e3);
  }
  e->factorized = 1;
  sopar->diagchange = (PASTIX_INT)nbpivot;                       /* -> IPARM_STATIC_PIVOTING (pastix.c:3853) */
  sopar->dparm[DPARM_FACT_TIME] = seconds;                       /* sopalin3d.c:1125-1132 */
#if !defined(TYPE_COMPLEX) && !defined(CHOL_SOPALIN)
  { int64_t inertia = 0;                                         /* sopalin3d.c:1145-1161 */
    if (pb200_inertia(e->h, &inertia) != PB200_SUCCESS) shim_fatal("pb200_inertia");
    sopar->iparm[IPARM_INERTIA] = (PASTIX_INT)inertia; }
#endif
  if (getenv("PB200_HOST_COEFTAB") != NULL) shim_mirror_coeftab(e->h, datacode);
  else if (schur) shim_fetch_schur(e->h, datacode);
}

static void shim_updown(SolverMatrix *datacode, SopalinParam *sopar)
{
  pb200_shim_entry_t *e = shim_find(datacode, 0);
  double seconds = 0.0;
  if (e == NULL || e->h == NULL || !e->factorized) {
    errorPrint("pastix_b200: up_down called before a numeric factorization on this SolverMatrix");
    EXIT(MOD_SOPALIN, BADPARAMETER_ERR);
  }
  /* IPARM_TRANSPOSE_SOLVE only acts on LU in the reference (updo.c:165, 1553 are under SOPALIN_LU) */
  if (pb200_set_transpose_solve(e->h, PB200_FACTO == PB200_FACT_LU && sopar->iparm[IPARM_TRANSPOSE_SOLVE] == API_YES) != PB200_SUCCESS)
    shim_fatal("pb200_set_transpose_solve");
  if (pb200_solve(e->h, UPDOWN_SM2XTAB, (int64_t)UPDOWN_SM2XSZE, (int64_t)UPDOWN_SM2XNBR, &seconds) != PB200_SUCCESS)
    shim_fatal("pb200_solve");
  sopar->dparm[DPARM_SOLV_TIME] = seconds;                       /* updo.c:1495 */
}

/* ---------------------------------------------------------------- entry points */
void API_CALL(sopalin_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_numfact(m, sopaparam);
}

void API_CALL(updo_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_updown(m, sopaparam);
}

void API_CALL(sopalin_updo_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_numfact(m, sopaparam);
  shim_updown(m, sopaparam);
}

/* up_down from inside the reference's thread pool (refinement preconditioner): thread 0 drives the GPU,
 * the callers bracket this with SYNCHRO_THREAD (raff_functions.c:423-428) */
void *API_CALL(up_down_smp)(void *arg)
{
  sopthread_data_t *argument     = (sopthread_data_t *)arg;
  Sopalin_Data_t   *sopalin_data = (Sopalin_Data_t *)(argument->data);
  if (argument->me == 0) shim_updown(sopalin_data->datacode, sopalin_data->sopar);
  return NULL;
}

/* no communication thread: one process drives the GPUs */
void *API_CALL(sopalin_updo_comm)(void *arg)
{
  (void)arg;
  return NULL;
}

/* ---------------------------------------------------------------- refinement: the reference's host loops,
 * included exactly as sopalin3d.c:409-434 does, running on top of our up_down_smp */
void *API_CALL(pivotstatique_smp)(void *arg);
void *API_CALL(gmres_smp)(void *arg);
void *API_CALL(grad_smp)(void *arg);
void *API_CALL(bicgstab_smp)(void *arg);
void  API_CALL(pivot_thread)(SolverMatrix *datacode, SopalinParam *sopaparam);
void  API_CALL(gmres_thread)(SolverMatrix *datacode, SopalinParam *sopaparam);
void  API_CALL(grad_thread)(SolverMatrix *datacode, SopalinParam *sopaparam);
void  API_CALL(bicgstab_thread)(SolverMatrix *datacode, SopalinParam *sopaparam);

/* file-scope constants the included refinement sources expect from sopalin3d.c:171-186 */
#include "sopalin_compute.h"
static PASTIX_INT   iun   = 1;
#ifdef TYPE_COMPLEX
static PASTIX_FLOAT fun   = 1.0 + 0.0 * I;
#else
static PASTIX_FLOAT fun   = 1.0;
#endif
static PASTIX_FLOAT fzero = 0.0;

#define RAFF_CLOCK_INIT {clockInit(&raff_clk);clockStart(&raff_clk);}
#define RAFF_CLOCK_STOP {clockStop(&(raff_clk));}
#define RAFF_CLOCK_GET  clockVal(&(raff_clk))
#include "raff_functions.h"
#include "raff_grad.c"
#include "raff_gmres.c"
#include "raff_pivot.c"
#include "raff_bicgstab.c"

void API_CALL(sopalin_updo_gmres_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_numfact(m, sopaparam);
  shim_updown(m, sopaparam);
  API_CALL(gmres_thread)(m, sopaparam);
}
void API_CALL(sopalin_updo_grad_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_numfact(m, sopaparam);
  shim_updown(m, sopaparam);
  API_CALL(grad_thread)(m, sopaparam);
}
void API_CALL(sopalin_updo_pivot_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_numfact(m, sopaparam);
  shim_updown(m, sopaparam);
  API_CALL(pivot_thread)(m, sopaparam);
}
void API_CALL(sopalin_updo_bicgstab_thread)(SolverMatrix *m, SopalinParam *sopaparam)
{
  shim_numfact(m, sopaparam);
  shim_updown(m, sopaparam);
  API_CALL(bicgstab_thread)(m, sopaparam);
}

/* ---- host-side readers of the internal CSC inside the reference code that stays linked.  CscbMAx / CscAxPb
 * (csc_intern_compute.c) walk the HOST CscMatrix; the static-pivot refinement (raff_pivot.c:133-139, included above)
 * and the host statistics of the other drivers (shim_raff.c) call them.  The reference objects keep their routines as
 * <variant>_CscbMAx_hostref / _CscAxPb_hostref (build_dropin.sh, objcopy); these bring the host copy up to date first
 * (shim_csc.c leaves rows / values in HBM, pb200_shim_csc_host). */
void API_CALL(CscbMAx_hostref)(Sopalin_Data_t *sopalin_data, int me, volatile PASTIX_FLOAT *r, const volatile PASTIX_FLOAT *b,
                               const CscMatrix *cscmtx, const UpDownVector *updovct, const SolverMatrix *solvmtx, MPI_Comm comm,
                               PASTIX_INT transpose);
void API_CALL(CscAxPb_hostref)(Sopalin_Data_t *sopalin_data, int me, PASTIX_FLOAT *r, const PASTIX_FLOAT *b, const CscMatrix *cscmtx,
                               const UpDownVector *updovct, const SolverMatrix *solvmtx, MPI_Comm comm, PASTIX_INT transpose);
void CscbMAx(Sopalin_Data_t *sopalin_data, int me, volatile PASTIX_FLOAT *r, const volatile PASTIX_FLOAT *b, const CscMatrix *cscmtx,
             const UpDownVector *updovct, const SolverMatrix *solvmtx, MPI_Comm comm, PASTIX_INT transpose)
{
  pb200_shim_csc_host(solvmtx);
  API_CALL(CscbMAx_hostref)(sopalin_data, me, r, b, cscmtx, updovct, solvmtx, comm, transpose);
}
void CscAxPb(Sopalin_Data_t *sopalin_data, int me, PASTIX_FLOAT *r, const PASTIX_FLOAT *b, const CscMatrix *cscmtx,
             const UpDownVector *updovct, const SolverMatrix *solvmtx, MPI_Comm comm, PASTIX_INT transpose)
{
  pb200_shim_csc_host(solvmtx);
  API_CALL(CscAxPb_hostref)(sopalin_data, me, r, b, cscmtx, updovct, solvmtx, comm, transpose);
}
