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

static pb200_shim_entry_t *hook_find(const SolverMatrix *m)
{
  int i; pb200_shim_entry_t *e = NULL;
  pthread_mutex_lock(&shim_mutex);
  for (i = 0; i < PB200_SHIM_MAX; i++)
    if (shim_table[i].m == m) { e = &shim_table[i]; break; }
  pthread_mutex_unlock(&shim_mutex);
  return e;
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

/* release the HBM held for this pastix_data_t (call before API_TASK_CLEAN) */
void pb200_shim_release_data(void *pastix_data)
{
  pb200_shim_entry_t *e = hook_find(&((pastix_data_t *)pastix_data)->solvmatr);
  if (e == NULL) return;
  if (e->h) pb200_destroy(e->h);
  if (e->csc) pb200_csc_destroy(e->csc);
  pthread_mutex_lock(&shim_mutex);
  memset(e, 0, sizeof(*e));
  pthread_mutex_unlock(&shim_mutex);
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
