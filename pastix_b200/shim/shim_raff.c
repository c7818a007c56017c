/*
 * shim_raff.c — the vector back end of PaStiX's refinement drivers, on the B200.
 *
 * The reference's GMRES, conjugate gradient and BiCGSTAB (src/sopalin/src/raff_gmres.c, raff_grad.c,
 * raff_bicgstab.c — included UNCHANGED by sopalin_b200_shim.c, as sopalin3d.c:409-434 does) are written against a
 * table of vector operations, `struct solver` (raff_functions.h:189-224), which the reference fills with host
 * implementations (raff_functions.c:100-650: CscAx, CscbMAx, CscGradBeta, CscNormFro, ... on host vectors, the
 * preconditioner going through UPDOWN_SM2XTAB).  This file replaces that one object (compiled four times like it):
 * the same table, every entry served by the CUDA layer on vectors that live in HBM — SpMV on the internal CSC that is
 * already resident for the assembly, dot products / axpy / scal as kernels, the preconditioner = pb200_solve_device
 * in place.  One Krylov iteration no longer crosses PCIe; only the scalars (dot products, norms) come back.
 * Static-pivot refinement (raff_pivot.c) does not use the table and keeps its host vectors around the GPU up_down.
 *
 * Thread model: the drivers are SPMD over SOLV_THRDNBR threads; thread 0 drives the GPU, results are broadcast
 * through sopalin_data->common_flt / common_dbl with the reference's own barrier (SYNCHRO_THREAD).
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
#include "csc.h"
#include "updown.h"
#include "queue.h"
#include "bulles.h"
#include "ftgt.h"
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

#define up_down_smp API_CALL(up_down_smp)
void *up_down_smp(void *arg);
#define sopalin_updo_comm API_CALL(sopalin_updo_comm)
void *sopalin_updo_comm(void *arg);

#include "raff_functions.h"
#include "shim_table.h"

#if defined(SOPALIN_LU)
#define RAFF_IS_LU 1
#else
#define RAFF_IS_LU 0
#endif

/* the locals every operation needs; SYNCHRO_THREAD reads sopalin_data and datacode */
#define RAFF_CTX                                                            \
  sopthread_data_t *argument     = (sopthread_data_t *)arg;                 \
  Sopalin_Data_t   *sopalin_data = (Sopalin_Data_t *)(argument->data);      \
  SolverMatrix     *datacode     = sopalin_data->datacode;                  \
  SopalinParam     *sopar        = sopalin_data->sopar;                     \
  const PASTIX_INT  me           = argument->me;                            \
  (void)sopar; (void)datacode

static pb200_handle_t *raff_handle(const SolverMatrix *m)
{
  pb200_shim_entry_t *e = pb200_shim_entry(m, 0);
  pb200_handle_t *h = e ? e->h : NULL;
  if (h == NULL) {
    errorPrint("pastix_b200: refinement called before a numeric factorization on this SolverMatrix");
    EXIT(MOD_SOPALIN, BADPARAMETER_ERR);
  }
  return h;
}
#define RAFF_DO(call) do { if ((call) != PB200_SUCCESS) {                                  \
    errorPrint("pastix_b200: %s: %s", #call, pb200_last_error()); EXIT(MOD_SOPALIN, INTERNAL_ERR); } } while (0)

/* ---- allocation / sharing between the driver threads */
PASTIX_FLOAT *Pastix_Synchro_Vect(void *arg, void *x, int nb)
{
  RAFF_CTX;
  if (me == 0) sopalin_data->ptr_raff[nb] = x;
  SYNCHRO_THREAD;
  return (PASTIX_FLOAT *)sopalin_data->ptr_raff[nb];
}
void *Pastix_Malloc(void *arg, size_t size)
{
  RAFF_CTX; void *p = NULL;
  if (me == 0) RAFF_DO(pb200_vec_alloc(raff_handle(datacode), &p, (int64_t)size));
  return p;
}
void Pastix_Free(void *arg, void *x)
{
  RAFF_CTX;
  if (me == 0) RAFF_DO(pb200_vec_free(raff_handle(datacode), x));
}

/* ---- interface with the caller */
void Pastix_Verbose(void *arg, double t0, double t3, double tmp, PASTIX_INT nb_iter)
{
  RAFF_CTX;
  sopalin_data->count_iter = nb_iter;
  sopalin_data->stop = tmp;
  if (me == 0 && sopar->iparm[IPARM_VERBOSE] > API_VERBOSE_NOT && SOLV_PROCNUM == 0) {
    fprintf(stdout, OUT_ITERRAFF_ITER, (int)nb_iter);
    if (sopar->iparm[IPARM_ONLY_RAFF] == API_NO) fprintf(stdout, OUT_ITERRAFF_TTS, 0.0);
    fprintf(stdout, OUT_ITERRAFF_TTT, t3 - t0);
    fprintf(stdout, OUT_ITERRAFF_ERR, tmp);
  }
}
void Pastix_End(void *arg, PASTIX_FLOAT tmp, PASTIX_INT nb_iter, double t, PASTIX_FLOAT *x)
{
  RAFF_CTX;
  sopalin_data->stop = tmp;
  if (me == 0) RAFF_DO(pb200_vec_get(raff_handle(datacode), UPDOWN_SM2XTAB, x, (int64_t)UPDOWN_SM2XSZE));
  SYNCHRO_THREAD;
  sopar->rberror = tmp;
  sopar->itermax = nb_iter;
  if (sopar->iparm[IPARM_PRODUCE_STATS] == API_YES) {
    /* scaled residual of the solution now in UPDOWN_SM2XTAB: the reference's host statistics, as Pastix_End runs them */
    PASTIX_FLOAT *r, *s;
    if (me == 0) {
      MALLOC_INTERN(r, UPDOWN_SM2XSZE, PASTIX_FLOAT);
      MALLOC_INTERN(s, UPDOWN_SM2XSZE, PASTIX_FLOAT);
      sopalin_data->ptr_raff[0] = (void *)r;
      sopalin_data->ptr_raff[1] = (void *)s;
    }
    SYNCHRO_THREAD;
    r = (PASTIX_FLOAT *)sopalin_data->ptr_raff[0];
    s = (PASTIX_FLOAT *)sopalin_data->ptr_raff[1];
    if (me == 0) pb200_shim_csc_host(datacode);       /* the host statistics read the host CscMatrix */
    SYNCHRO_THREAD;
    MULTITHREAD_BEGIN;
    CscbMAx(sopalin_data, me, r, sopar->b, sopar->cscmtx, &(datacode->updovct), datacode, PASTIX_COMM,
            sopar->iparm[IPARM_TRANSPOSE_SOLVE]);
    CscAxPb(sopalin_data, me, s, sopar->b, sopar->cscmtx, &(datacode->updovct), datacode, PASTIX_COMM,
            sopar->iparm[IPARM_TRANSPOSE_SOLVE]);
    CscBerr(sopalin_data, me, r, s, UPDOWN_SM2XSZE, 1, &(sopar->dparm[DPARM_SCALED_RESIDUAL]), PASTIX_COMM);
    MULTITHREAD_END(1);
    SYNCHRO_THREAD;
    if (me == 0) { memFree_null(r); memFree_null(s); }
  }
  if (me == 0) set_dparm(sopar->dparm, DPARM_RAFF_TIME, t);
  SYNCHRO_THREAD;
}
void Pastix_X(void *arg, PASTIX_FLOAT *x)
{
  RAFF_CTX;
  if (me == 0) {
    pb200_handle_t *h = raff_handle(datacode);
    if (sopar->iparm[IPARM_ONLY_RAFF] == API_NO) {   /* the drivers start from zero (raff_functions.c:252-254) */
      memset(UPDOWN_SM2XTAB, 0, sizeof(PASTIX_FLOAT) * (size_t)(UPDOWN_SM2XSZE * UPDOWN_SM2XNBR));
      RAFF_DO(pb200_vec_zero(h, x, (int64_t)UPDOWN_SM2XSZE));
    } else
      RAFF_DO(pb200_vec_set(h, x, UPDOWN_SM2XTAB, (int64_t)UPDOWN_SM2XSZE));
  }
  SYNCHRO_THREAD;
}
PASTIX_INT Pastix_n(void *arg) { RAFF_CTX; (void)me; return UPDOWN_SM2XSZE; }
PASTIX_INT Pastix_m(void *arg) { RAFF_CTX; (void)me; return UPDOWN_SM2XNBR; }
void Pastix_B(void *arg, PASTIX_FLOAT *b)
{
  RAFF_CTX;
  if (me == 0) RAFF_DO(pb200_vec_set(raff_handle(datacode), b, sopar->b, (int64_t)UPDOWN_SM2XSZE));
  SYNCHRO_THREAD;
}
PASTIX_FLOAT Pastix_Eps(void *arg) { RAFF_CTX; (void)me; return sopar->epsilonraff; }
PASTIX_INT Pastix_Itermax(void *arg) { RAFF_CTX; (void)me; return sopar->itermax; }
PASTIX_INT Pastix_Krylov_Space(void *arg) { RAFF_CTX; (void)me; return sopar->gmresim; }
PASTIX_INT Pastix_me(void *arg) { return ((sopthread_data_t *)arg)->me; }

/* ---- scalars (host side: the drivers keep them in vectors of one element) */
void Pastix_Mult(void *arg, PASTIX_FLOAT *alpha, PASTIX_FLOAT *beta, PASTIX_FLOAT *zeta, int flag)
{
  RAFF_CTX;
  if (me == 0) zeta[0] = alpha[0] * beta[0];
  if (flag) SYNCHRO_THREAD;
}
void Pastix_Div(void *arg, PASTIX_FLOAT *alpha, PASTIX_FLOAT *beta, PASTIX_FLOAT *zeta, int flag)
{
  RAFF_CTX;
  if (me == 0) zeta[0] = alpha[0] / beta[0];
  if (flag) SYNCHRO_THREAD;
}

/* ---- vector operations on the device */
static void raff_dot(void *arg, int conj_y, PASTIX_FLOAT *x, PASTIX_FLOAT *y, PASTIX_FLOAT *r)
{
  RAFF_CTX;
  if (me == 0) {
    PASTIX_FLOAT v;
    RAFF_DO(pb200_vec_dot(raff_handle(datacode), conj_y, x, y, (int64_t)UPDOWN_SM2XSZE, &v));
    sopalin_data->common_flt[0] = v;
  }
  SYNCHRO_THREAD;
  *r = sopalin_data->common_flt[0];
  SYNCHRO_THREAD;
}
PASTIX_FLOAT Pastix_Norm2(void *arg, PASTIX_FLOAT *x)
{
  RAFF_CTX; double nrm;
  if (me == 0) {
    PASTIX_FLOAT v;
    RAFF_DO(pb200_vec_dot(raff_handle(datacode), 1, x, x, (int64_t)UPDOWN_SM2XSZE, &v));
#ifdef TYPE_COMPLEX
    sopalin_data->common_dbl[0] = sqrt((double)creal(v));
#else
    sopalin_data->common_dbl[0] = sqrt((double)v);
#endif
  }
  SYNCHRO_THREAD;
  nrm = sopalin_data->common_dbl[0];
  SYNCHRO_THREAD;
  return (PASTIX_FLOAT)nrm;
}
void Pastix_Copy(void *arg, PASTIX_FLOAT *s, PASTIX_FLOAT *d, int flag)
{
  RAFF_CTX;
  if (me == 0) RAFF_DO(pb200_vec_copy(raff_handle(datacode), d, s, (int64_t)UPDOWN_SM2XSZE));
  if (flag) SYNCHRO_THREAD;
}
/* d = M^{-1} s: the up_down of the factors in HBM, in place on d (raff_functions.c:408-437 goes through UPDOWN_SM2XTAB) */
void Pastix_Precond(void *arg, PASTIX_FLOAT *s, PASTIX_FLOAT *d, int flag)
{
  RAFF_CTX;
  SYNCHRO_THREAD;
  if (me == 0) {
    pb200_handle_t *h = raff_handle(datacode);
    if (sopar->iparm[IPARM_ONLY_RAFF] == API_NO) {
      RAFF_DO(pb200_set_transpose_solve(h, RAFF_IS_LU && sopar->iparm[IPARM_TRANSPOSE_SOLVE] == API_YES));
      RAFF_DO(pb200_precond(h, s, d));
    } else
      RAFF_DO(pb200_vec_copy(h, d, s, (int64_t)UPDOWN_SM2XSZE));
  }
  SYNCHRO_THREAD;
  (void)flag;
}
void Pastix_Scal(void *arg, PASTIX_FLOAT alpha, PASTIX_FLOAT *x, int flag)
{
  RAFF_CTX;
  if (me == 0) RAFF_DO(pb200_vec_scal(raff_handle(datacode), &alpha, x, (int64_t)UPDOWN_SM2XSZE));
  if (flag) SYNCHRO_THREAD;
}
/* sum r_i z_i, z conjugated only in the Hermitian build (CONJ_JJP, csc_intern_compute.c:95-103) */
void Pastix_Dotc(void *arg, PASTIX_FLOAT *x, PASTIX_FLOAT *y, PASTIX_FLOAT *r, int flag)
{
#ifdef HERMITIAN
  raff_dot(arg, 1, x, y, r);
#else
  raff_dot(arg, 0, x, y, r);
#endif
  (void)flag;
}
/* sum r_i conj(z_i) (CscGmresBeta, csc_intern_compute.c:1451-1560) */
void Pastix_Dotc_Gmres(void *arg, PASTIX_FLOAT *x, PASTIX_FLOAT *y, PASTIX_FLOAT *r, int flag)
{
  raff_dot(arg, 1, x, y, r);
  (void)flag;
}
void Pastix_Ax(void *arg, PASTIX_FLOAT *x, PASTIX_FLOAT *r)
{
  RAFF_CTX;
  if (me == 0)
    RAFF_DO(pb200_csc_ax(raff_handle(datacode), sopar->cscmtx->type, sopar->iparm[IPARM_TRANSPOSE_SOLVE] == API_YES, NULL, x, r));
  SYNCHRO_THREAD;
}
void Pastix_bMAx(void *arg, PASTIX_FLOAT *b, PASTIX_FLOAT *x, PASTIX_FLOAT *r)
{
  RAFF_CTX;
  if (me == 0)
    RAFF_DO(pb200_csc_ax(raff_handle(datacode), sopar->cscmtx->type, sopar->iparm[IPARM_TRANSPOSE_SOLVE] == API_YES, b, x, r));
  SYNCHRO_THREAD;
}
/* x <- beta x + y */
void Pastix_BYPX(void *arg, PASTIX_FLOAT *beta, PASTIX_FLOAT *y, PASTIX_FLOAT *x, int flag)
{
  RAFF_CTX;
  if (me == 0) {
    pb200_handle_t *h = raff_handle(datacode);
    PASTIX_FLOAT one = 1.0;
    RAFF_DO(pb200_vec_scal(h, &beta[0], x, (int64_t)UPDOWN_SM2XSZE));
    RAFF_DO(pb200_vec_axpy(h, &one, y, x, (int64_t)UPDOWN_SM2XSZE));
  }
  if (flag) SYNCHRO_THREAD;
}
/* x <- x + coeff alpha y */
void Pastix_AXPY(void *arg, double coeff, PASTIX_FLOAT *alpha, PASTIX_FLOAT *x, PASTIX_FLOAT *y, int flag)
{
  RAFF_CTX;
  if (me == 0) {
    PASTIX_FLOAT a = (PASTIX_FLOAT)alpha[0] * coeff;
    RAFF_DO(pb200_vec_axpy(raff_handle(datacode), &a, y, x, (int64_t)UPDOWN_SM2XSZE));
  }
  if (flag) SYNCHRO_THREAD;
}

void Pastix_Solveur(struct solver *solveur)
{
  solveur->Synchro = &Pastix_Synchro_Vect; solveur->Malloc = &Pastix_Malloc; solveur->Free = &Pastix_Free;
  solveur->Verbose = &Pastix_Verbose;      solveur->End = &Pastix_End;       solveur->X = &Pastix_X;
  solveur->N = &Pastix_n;                  solveur->B = &Pastix_B;           solveur->Eps = &Pastix_Eps;
  solveur->Itermax = &Pastix_Itermax;      solveur->me = &Pastix_me;         solveur->Krylov_Space = &Pastix_Krylov_Space;
  solveur->Mult = &Pastix_Mult;            solveur->Div = &Pastix_Div;       solveur->Dotc_Gmres = &Pastix_Dotc_Gmres;
  solveur->Norm = &Pastix_Norm2;           solveur->Copy = &Pastix_Copy;     solveur->Precond = &Pastix_Precond;
  solveur->Scal = &Pastix_Scal;            solveur->Dotc = &Pastix_Dotc;     solveur->Ax = &Pastix_Ax;
  solveur->AXPY = &Pastix_AXPY;            solveur->bMAx = &Pastix_bMAx;     solveur->BYPX = &Pastix_BYPX;
}

/* runs one refinement driver in the reference's thread pool (raff_functions.c:651-671) */
void raff_thread(SolverMatrix *datacode, SopalinParam *sopaparam, void *(*method)(void *))
{
  Sopalin_Data_t *sopalin_data = NULL;
  BackupSolve_t   saved;
  MALLOC_INTERN(sopalin_data, 1, Sopalin_Data_t);
  solve_backup(datacode, &saved);
  sopalin_init(sopalin_data, datacode, sopaparam, 0);
  sopalin_launch_thread(sopalin_data, SOLV_PROCNUM, SOLV_PROCNBR, datacode->btree, sopaparam->iparm[IPARM_VERBOSE],
                        SOLV_THRDNBR, method, sopalin_data,
                        sopaparam->nbthrdcomm, API_CALL(sopalin_updo_comm), sopalin_data,
                        OOC_THREAD_NBR, ooc_thread, sopalin_data);
  sopalin_clean(sopalin_data, 2);
  solve_restore(datacode, &saved);
  memFree_null(sopalin_data);
}
