/*
 * shim_hooks.c — measurement / lifetime hooks of the drop-in library, compiled once per precision.
 * They expose the device handle that sopalin_b200_shim.c keeps for the SolverMatrix inside a
 * pastix_data_t (src/sopalin/src/pastixstr.h), so that bench.py can re-run the kernels on inputs
 * already resident in HBM, and let a host application release the HBM explicitly.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include "nompi.h"
#include "common_pastix.h"
#include "tools.h"
#include "sopalin_define.h"
#include "dof.h"
#include "ftgt.h"
#include "symbol.h"
#include "csc.h"
#include "updown.h"
#include "queue.h"
#include "bulles.h"
#include "solver.h"
#include "assembly.h"
#include "param_blend.h"
#include "order.h"
#include "fax.h"
#include "kass.h"
#include "blend.h"
#include "solverRealloc.h"
#include "sopalin_thread.h"
#include "stack.h"
#include "sopalin3d.h"
#include "sopalin_init.h"
#include "sopalin_option.h"
#include "csc_intern_updown.h"
#include "csc_intern_build.h"
#include "coefinit.h"
#include "out.h"
#include "pastix.h"
#include "pastix_internal.h"
#include "pastixstr.h"
#include "shim_table.h"

/* ---- the side table: a growing array of pointers to heap entries (an entry never moves once handed out) */
static pb200_shim_entry_t **shim_tab = NULL;
static int                  shim_cap = 0;
static pthread_mutex_t      shim_mutex = PTHREAD_MUTEX_INITIALIZER;

pb200_shim_entry_t *pb200_shim_entry(const SolverMatrix *m, int create)
{
  int i, slot = -1; pb200_shim_entry_t *e = NULL;
  pthread_mutex_lock(&shim_mutex);
  for (i = 0; i < shim_cap; i++) {
    if (shim_tab[i] != NULL && shim_tab[i]->m == m) { e = shim_tab[i]; break; }
    if (shim_tab[i] == NULL && slot < 0) slot = i;
  }
  if (e == NULL && create) {
    if (slot < 0) {
      int ncap = shim_cap ? 2 * shim_cap : 16;
      pb200_shim_entry_t **nt = (pb200_shim_entry_t **)realloc(shim_tab, sizeof(*nt) * (size_t)ncap);
      if (nt != NULL) {
        for (i = shim_cap; i < ncap; i++) nt[i] = NULL;
        slot = shim_cap; shim_tab = nt; shim_cap = ncap;
      }
    }
    if (slot >= 0 && (e = (pb200_shim_entry_t *)calloc(1, sizeof(*e))) != NULL) { e->m = m; shim_tab[slot] = e; }
  }
  pthread_mutex_unlock(&shim_mutex);
  return e;
}

void pb200_shim_entry_drop(const SolverMatrix *m)
{
  int i; pb200_shim_entry_t *e = NULL;
  pthread_mutex_lock(&shim_mutex);
  for (i = 0; i < shim_cap; i++)
    if (shim_tab[i] != NULL && shim_tab[i]->m == m) { e = shim_tab[i]; shim_tab[i] = NULL; break; }
  pthread_mutex_unlock(&shim_mutex);
  if (e == NULL) return;
  if (e->ngpu > 1) pb200_destroy_group(e->hs, e->ngpu);
  else if (e->h) pb200_destroy(e->h);
  if (e->csc) pb200_csc_destroy(e->csc);
  free(e);
}

int pb200_shim_live_entries(void)
{
  int i, n = 0;
  pthread_mutex_lock(&shim_mutex);
  for (i = 0; i < shim_cap; i++) if (shim_tab[i] != NULL) n++;
  pthread_mutex_unlock(&shim_mutex);
  return n;
}

/* FNV-1a over the fields shim_create flattens (sopalin_b200_shim.c): what the device schedule, the panel layout and the
 * scatter maps are derived from */
void pb200_shim_fingerprint(const SolverMatrix *m, uint64_t fp[2])
{
  /* four independent FNV-1a chains (one per field): the serial chain cost 0.9 ms per NUMFACT on the 64^3 problem */
  uint64_t h0 = 1469598103934665603ULL, h1 = h0 ^ 0x9e3779b97f4a7c15ULL, h2 = h0 ^ 0xc2b2ae3d27d4eb4fULL, h3 = h0 ^ 0x165667b19e3779f9ULL;
  PASTIX_INT i; uint64_t coefnbr = 0;
#define FP_MIX(h, v) do { h ^= (uint64_t)(v); h *= 1099511628211ULL; } while (0)
  for (i = 0; i < m->cblknbr; i++) {
    FP_MIX(h0, m->cblktab[i].fcolnum); FP_MIX(h1, m->cblktab[i].lcolnum); FP_MIX(h2, m->cblktab[i].bloknum); FP_MIX(h3, m->cblktab[i].stride);
    coefnbr += (uint64_t)m->cblktab[i].stride * (uint64_t)(m->cblktab[i].lcolnum - m->cblktab[i].fcolnum + 1);
  }
  for (i = 0; i < m->bloknbr; i++) {
    FP_MIX(h0, m->bloktab[i].frownum); FP_MIX(h1, m->bloktab[i].lrownum); FP_MIX(h2, m->bloktab[i].cblknum); FP_MIX(h3, m->bloktab[i].coefind);
  }
#undef FP_MIX
  fp[0] = h0 ^ (h1 * 0x9e3779b97f4a7c15ULL) ^ (h2 << 21 | h2 >> 43) ^ (h3 << 42 | h3 >> 22);
  fp[1] = ((uint64_t)m->cblknbr << 40) ^ ((uint64_t)m->bloknbr << 16) ^ coefnbr;
}

/* ---- the reference's release points.  The reference objects keep their own routines under *_hostref names. */
void CoefMatrix_Free_hostref(SopalinParam *sopar, SolverMatrix *datacode, PASTIX_INT factotype);
void solverExit_hostref(SolverMatrix *solvmtx);
void Csc2updown_hostref(const CscMatrix *cscmtx, UpDownVector *updovct, const SolverMatrix *solvmtx, int mode, MPI_Comm comm);

/* CoefMatrix_Free (coefinit.c:479): the host panels go away — before a new blend on the same pastix_data
 * (pastix.c:2716) and before every re-fill of the coefficients (pastix.c:3391).  The factors on the device are void
 * from here on; the handle itself is kept (a re-factorization on the same analysis reuses schedule and slabs) and
 * is checked against the structure's fingerprint at the next numeric factorization. */
void CoefMatrix_Free(SopalinParam *sopar, SolverMatrix *datacode, PASTIX_INT factotype)
{
  pb200_shim_entry_t *e = pb200_shim_entry(datacode, 0);
  if (e != NULL) e->factorized = 0;
  CoefMatrix_Free_hostref(sopar, datacode, factotype);
}

/* solverExit (solverRealloc.c:217): the SolverMatrix is destroyed (API_TASK_CLEAN, pastix.c:4539) — all HBM held for
 * it is released here.  Temporaries of the analysis never have an entry. */
void solverExit(SolverMatrix *solvmtx)
{
  pb200_shim_entry_drop(solvmtx);
  solverExit_hostref(solvmtx);
}

static pb200_shim_entry_t *hook_find(const SolverMatrix *m) { return pb200_shim_entry(m, 0); }

/* ---- host copy of the internal CSC on demand (shim_csc.c leaves rows / values in HBM) */
static pthread_mutex_t csc_host_mutex = PTHREAD_MUTEX_INITIALIZER;
void pb200_shim_csc_host(const SolverMatrix *m)
{
  pb200_shim_entry_t *e = hook_find(m);
  int64_t *gcol;
  if (e == NULL || !e->host_stale || e->csc == NULL) return;
  pthread_mutex_lock(&csc_host_mutex);          /* the readers may be the threads of a refinement (raff_pivot.c) */
  if (e->host_stale) {
    gcol = (int64_t *)malloc(sizeof(int64_t) * (size_t)(e->lazy_ncol + 1));
    if (gcol == NULL || pb200_csc_fetch(e->csc, gcol, (int64_t *)e->lazy_rows, e->lazy_vals, e->lazy_tvals) != PB200_SUCCESS) {
      errorPrint("pastix_b200: internal CSC -> host: %s", gcol ? pb200_last_error() : "out of memory");
      EXIT(MOD_SOPALIN, INTERNAL_ERR);
    }
    free(gcol);
    e->host_stale = 0;
  }
  pthread_mutex_unlock(&csc_host_mutex);
}
/* Csc2updown (csc_intern_updown.c:339): b = A * (1 | i) read off the HOST CscMatrix when IPARM_RHS_MAKING asks for a
 * generated right-hand side (pastix.c:716) — the one reader of the CSC values inside the unchanged pastix.c */
void Csc2updown(const CscMatrix *cscmtx, UpDownVector *updovct, const SolverMatrix *solvmtx, int mode, MPI_Comm comm)
{
  pb200_shim_csc_host(solvmtx);
  Csc2updown_hostref(cscmtx, updovct, solvmtx, mode, comm);
}

/* pb200_handle_t* behind a pastix_data_t (NULL before the first API_TASK_NUMFACT) */
void *pb200_shim_get_handle(void *pastix_data)
{
  pb200_shim_entry_t *e = hook_find(&((pastix_data_t *)pastix_data)->solvmatr);
  return e ? (void *)e->h : NULL;
}

/* static-pivot threshold used by the last factorization (sopalin3d.c:586-606) */
double pb200_shim_get_critere(void *pastix_data)
{
  pb200_shim_entry_t *e = hook_find(&((pastix_data_t *)pastix_data)->solvmatr);
  return e ? e->critere : 0.0;
}

/* final permutation kept by the reference (order.h:52-57), 0-based, n entries each */
void pb200_shim_get_order(void *pastix_data, int64_t *permtab, int64_t *peritab)
{
  pastix_data_t *pd = (pastix_data_t *)pastix_data;
  PASTIX_INT i, n = pd->n2 > 0 ? pd->n2 : pd->n;
  for (i = 0; i < n; i++) { permtab[i] = pd->ordemesh.permtab[i]; peritab[i] = pd->ordemesh.peritab[i]; }
}

/* flat copy of the SolverMatrix the analysis left in this pastix_data_t (blend/src/solver.h:94-168), in the
 * layout of pb200_solver_t — what a multi-GPU host (one process per GPU, same analysis everywhere) hands to
 * pb200_create_dist.  sizes[0..1] = cblknbr, bloknbr; cblk arrays hold cblknbr+1 entries. */
void pb200_shim_solver_sizes(void *pastix_data, int64_t *sizes)
{
  SolverMatrix *m = &((pastix_data_t *)pastix_data)->solvmatr;
  sizes[0] = m->cblknbr; sizes[1] = m->bloknbr;
}
void pb200_shim_solver_get(void *pastix_data, int64_t *fcolnum, int64_t *lcolnum, int64_t *bloknum, int64_t *stride,
                           int64_t *frownum, int64_t *lrownum, int64_t *cblknum, int64_t *coefind)
{
  SolverMatrix *m = &((pastix_data_t *)pastix_data)->solvmatr;
  PASTIX_INT i;
  for (i = 0; i <= m->cblknbr; i++) {
    fcolnum[i] = m->cblktab[i].fcolnum; lcolnum[i] = m->cblktab[i].lcolnum;
    bloknum[i] = m->cblktab[i].bloknum; stride[i] = (i < m->cblknbr) ? m->cblktab[i].stride : 0;
  }
  for (i = 0; i < m->bloknbr; i++) {
    frownum[i] = m->bloktab[i].frownum; lrownum[i] = m->bloktab[i].lrownum;
    cblknum[i] = m->bloktab[i].cblknum; coefind[i] = m->bloktab[i].coefind;
  }
}

/* release the HBM held for this pastix_data_t explicitly (API_TASK_CLEAN does it too, through solverExit) */
void pb200_shim_release_data(void *pastix_data)
{
  pb200_shim_entry_drop(&((pastix_data_t *)pastix_data)->solvmatr);
}
int pb200_shim_live(void) { return pb200_shim_live_entries(); }

/* GPU-aware block sizes for blend.  The reference's defaults IPARM_MIN/MAX_BLOCKSIZE =
 * 60 / 120 (pastix.c:375-376) were chosen for CPU BLAS; on the B200 every scattered update tile pays a fixed
 * prologue / epilogue whatever the contraction length, so wider column blocks (K = 120 .. 240) raise the useful work per
 * tile — measured in profiles/r02/README.md.  Same analysis code, different parameters: call this between
 * API_TASK_INIT (which fills the defaults) and API_TASK_ANALYSE, or set the two iparm entries yourself. */
void pb200_tune_iparm(PASTIX_INT *iparm)
{
  iparm[IPARM_MIN_BLOCKSIZE] = 120;
  iparm[IPARM_MAX_BLOCKSIZE] = 240;
}

/* The factors where the reference leaves them: cblktab[].coeftab / .ucoeftab (solver.h:94-117) filled from HBM on
 * demand, allocated like CoefMatrix_Allocate does (coefinit.c:104-160) so that CoefMatrix_Free / solverExit release
 * them.  For consumers that read the panels directly (dump_all / PASTIX_DUMP_FACTO, user code walking the
 * SolverMatrix); up_down and refinement never need it.  Returns 0, or -1 when nothing is factorized. */
int pb200_shim_fetch_coeftab(void *pastix_data)
{
  SolverMatrix *m = &((pastix_data_t *)pastix_data)->solvmatr;
  pb200_shim_entry_t *e = hook_find(m);
  PASTIX_INT c;
  if (e == NULL || e->h == NULL || !e->factorized) return -1;
  for (c = 0; c < m->cblknbr; c++) {
    size_t sz = (size_t)m->cblktab[c].stride * (size_t)(m->cblktab[c].lcolnum - m->cblktab[c].fcolnum + 1);
    const int lu = (e->facto == PB200_FACT_LU);
    if (m->cblktab[c].coeftab == NULL) { MALLOC_INTERN(m->cblktab[c].coeftab, sz, PASTIX_FLOAT); }
    if (lu && m->cblktab[c].ucoeftab == NULL) { MALLOC_INTERN(m->cblktab[c].ucoeftab, sz, PASTIX_FLOAT); }
    if (pb200_get_cblk(e->h, (int64_t)c, m->cblktab[c].coeftab, lu ? m->cblktab[c].ucoeftab : NULL) != PB200_SUCCESS) {
      errorPrint("pastix_b200: pb200_get_cblk: %s", pb200_last_error());
      return -1;
    }
  }
  return 0;
}

/* the reference's own task-to-thread mapping as a cblk -> rank map for `nranks` GPUs (what PB200_DIST_MAP=blend hands to
 * pb200_create_opts, sopalin_b200_shim.c): owner[cblknbr]; returns the number of blend threads, -1 when the task list is
 * not a complete 1-D one.  Host only (tests/test_dist.py compares it with the mapping computed in the CUDA layer). */
int pb200_shim_blend_owner(void *pastix_data, int nranks, int32_t *owner)
{
  pastix_data_t *pd = (pastix_data_t *)pastix_data;
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT i, t, k, C = m->cblknbr;
  if (m->ttsktab == NULL || m->thrdnbr <= 0 || nranks <= 0) return -1;
  for (i = 0; i < C; i++) owner[i] = -1;
  for (t = 0; t < m->thrdnbr; t++)
    for (k = 0; k < m->ttsknbr[t]; k++) {
      const Task *tk = &m->tasktab[m->ttsktab[t][k]];
      if ((tk->taskid == COMP_1D || tk->taskid == DIAG) && tk->cblknum >= 0 && tk->cblknum < C)
        owner[tk->cblknum] = (int32_t)((t * nranks) / m->thrdnbr);
    }
  for (i = 0; i < C; i++) if (owner[i] < 0) return -1;
  return (int)m->thrdnbr;
}

/* internal CSC of this pastix_data_t (CscMatrix, blend/src/csc.h) flattened: sizes = {ncol, nnz, has transcsc, filled};
 * colptr 0-based with ncol+1 entries.  Used by the parity tests of the device-side CscOrdistrib (shim_csc.c). */
void pb200_shim_csc_sizes(void *pastix_data, int64_t *out)
{
  pastix_data_t *pd = (pastix_data_t *)pastix_data;
  CscMatrix *c = &pd->cscmtx;
  PASTIX_INT i, ncol = 0, nnz = 0;
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!pd->malcsc || CSC_FTAB(c) == NULL) return;
  for (i = 0; i < CSC_FNBR(c); i++) { ncol += CSC_COLNBR(c, i); nnz = CSC_COL(c, i, CSC_COLNBR(c, i)); }
  out[0] = ncol; out[1] = nnz; out[2] = (pd->sopar.transcsc != NULL); out[3] = 1;
}
int pb200_shim_csc_get(void *pastix_data, int64_t *colptr, int64_t *rows, void *vals, void *tvals)
{
  pastix_data_t *pd = (pastix_data_t *)pastix_data;
  CscMatrix *c = &pd->cscmtx;
  PASTIX_INT i, j, col = 0, nnz = 0;
  pb200_shim_csc_host(&pd->solvmatr);
  for (i = 0; i < CSC_FNBR(c); i++) {
    for (j = 0; j < CSC_COLNBR(c, i); j++) colptr[col++] = CSC_COL(c, i, j);
    nnz = CSC_COL(c, i, CSC_COLNBR(c, i));
  }
  colptr[col] = nnz;
  for (i = 0; i < nnz; i++) rows[i] = CSC_ROW(c, i);
  memcpy(vals, CSC_VALTAB(c), (size_t)nnz * sizeof(PASTIX_FLOAT));
  if (tvals && pd->sopar.transcsc) memcpy(tvals, pd->sopar.transcsc, (size_t)nnz * sizeof(PASTIX_FLOAT));
  return (int)c->type;
}

/* the internal CSC as it sits in HBM (what the device assembly read): same layout as pb200_shim_csc_get;
 * returns nnz, or -1 when CscOrdistrib did not run on the device for this pastix_data_t */
int64_t pb200_shim_csc_device_get(void *pastix_data, int64_t *colptr, int64_t *rows, void *vals, void *tvals)
{
  pastix_data_t *pd = (pastix_data_t *)pastix_data;
  pb200_shim_entry_t *e = hook_find(&pd->solvmatr);
  if (e == NULL || e->csc == NULL) return -1;
  if (pb200_csc_fetch(e->csc, colptr, rows, vals, tvals) != PB200_SUCCESS) return -1;
  return colptr[pd->n2 > 0 ? pd->n2 : pd->n];
}
