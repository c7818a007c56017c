/*
 * TEST INFRASTRUCTURE — not part of the product path.
 *
 * ref_driver.c: a window into the UNMODIFIED reference library.  It is compiled
 * together with the reference's own sources (oracle/build_ref.sh) into
 * oracle/_ref/libpastix_ref_<p>.so and only *reads* the reference's internal
 * structures so that Python (ctypes) can
 *   - fetch the SolverMatrix produced by the reference's order/fax/blend
 *     analysis (blend/src/solver.h:94-168) as flat int64 arrays,
 *   - fetch the internal block-CSC built by CscOrdistrib
 *     (sopalin/src/csc_intern_build.c:352),
 *   - fetch / overwrite the factor panels coeftab/ucoeftab after the
 *     reference's own sopalin ran (sopalin/src/coefinit.c:104),
 *   - evaluate the static-pivot threshold exactly as init_struct_sopalin does
 *     (sopalin/src/sopalin3d.c:586-606).
 * The numeric work itself is done by calling the reference's public pastix()
 * entry point straight from ctypes.
 */
#include <assert.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <math.h>
#include <sys/stat.h>
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
#include "csc_intern_compute.h"

int64_t refdrv_int_size(void)   { return (int64_t)sizeof(PASTIX_INT); }
int64_t refdrv_float_size(void) { return (int64_t)sizeof(PASTIX_FLOAT); }
int64_t refdrv_iparm_size(void) { return (int64_t)IPARM_SIZE; }
int64_t refdrv_dparm_size(void) { return (int64_t)DPARM_SIZE; }

/* sizes: [cblknbr, bloknbr, nodenbr, coefnbr, coefmax, ftgtnbr, tasknbr, indnbr,
 *         thrdnbr, clustnbr, clustnum, procnbr, gcblknbr, gnodenbr, sm2xsze, sm2xnbr] */
void refdrv_solver_sizes(pastix_data_t *pd, int64_t *out)
{
  SolverMatrix *m = &pd->solvmatr;
  out[0] = m->cblknbr;  out[1] = m->bloknbr;  out[2] = m->nodenbr;  out[3] = m->coefnbr;
  out[4] = m->coefmax;  out[5] = m->ftgtnbr;  out[6] = m->tasknbr;  out[7] = m->indnbr;
  out[8] = m->thrdnbr;  out[9] = m->clustnbr; out[10] = m->clustnum; out[11] = m->procnbr;
  out[12] = m->updovct.gcblknbr; out[13] = m->updovct.gnodenbr;
  out[14] = m->updovct.sm2xsze;  out[15] = m->updovct.sm2xnbr;
}

/* cblk arrays have cblknbr+1 entries (the sentinel carries bloknum = bloknbr) */
void refdrv_solver_get(pastix_data_t *pd,
                       int64_t *fcol, int64_t *lcol, int64_t *bloknum, int64_t *stride,
                       int64_t *frow, int64_t *lrow, int64_t *fcblk, int64_t *levf, int64_t *coefind)
{
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT i;
  for (i = 0; i <= m->cblknbr; i++) {
    fcol[i] = m->cblktab[i].fcolnum; lcol[i] = m->cblktab[i].lcolnum;
    bloknum[i] = m->cblktab[i].bloknum; stride[i] = (i < m->cblknbr) ? m->cblktab[i].stride : 0;
  }
  for (i = 0; i < m->bloknbr; i++) {
    frow[i] = m->bloktab[i].frownum; lrow[i] = m->bloktab[i].lrownum;
    fcblk[i] = m->bloktab[i].cblknum; levf[i] = m->bloktab[i].levfval;
    coefind[i] = m->bloktab[i].coefind;
  }
}

/* task table: taskid, prionum, cblknum, bloknum, ftgtcnt, ctrbcnt, indnum (tasknbr each) + indtab */
void refdrv_tasks_get(pastix_data_t *pd, int64_t *taskid, int64_t *prionum, int64_t *cblknum,
                      int64_t *bloknum, int64_t *ftgtcnt, int64_t *ctrbcnt, int64_t *indnum,
                      int64_t *indtab)
{
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT i;
  for (i = 0; i < m->tasknbr; i++) {
    taskid[i] = m->tasktab[i].taskid; prionum[i] = m->tasktab[i].prionum;
    cblknum[i] = m->tasktab[i].cblknum; bloknum[i] = m->tasktab[i].bloknum;
    ftgtcnt[i] = m->tasktab[i].ftgtcnt; ctrbcnt[i] = m->tasktab[i].ctrbcnt;
    indnum[i] = m->tasktab[i].indnum;
  }
  if (indtab) for (i = 0; i < m->indnbr; i++) indtab[i] = m->indtab[i];
}

/* fan-in targets: infotab rows of 10 ints (ftgt.h:68-82) */
void refdrv_ftgt_get(pastix_data_t *pd, int64_t *info)
{
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT i, k;
  for (i = 0; i < m->ftgtnbr; i++)
    for (k = 0; k < MAXINFO; k++) info[i * MAXINFO + k] = m->ftgttab[i].infotab[k];
}
int64_t refdrv_ftgt_infosize(void) { return MAXINFO; }

/* up_down indexing: sm2xind and ctrbnbr per cblk (updown.h:54-64) */
void refdrv_updown_get(pastix_data_t *pd, int64_t *sm2xind, int64_t *ctrbnbr)
{
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT i;
  for (i = 0; i < m->cblknbr; i++) {
    sm2xind[i] = m->updovct.cblktab[i].sm2xind;
    ctrbnbr[i] = m->updovct.cblktab[i].ctrbnbr;
  }
}

/* internal CSC (blend/src/csc.h): [ncol_total, nnz, has_trans, filled] */
void refdrv_csc_sizes(pastix_data_t *pd, int64_t *out)
{
  CscMatrix *c = &pd->cscmtx;
  PASTIX_INT i, ncol = 0, nnz = 0;
  out[0] = out[1] = out[2] = out[3] = 0;
  if (!pd->malcsc || c->cscftab == NULL) return;
  for (i = 0; i < c->cscfnbr; i++) {
    ncol += c->cscftab[i].colnbr;
    nnz = c->cscftab[i].coltab[c->cscftab[i].colnbr];
  }
  out[0] = ncol; out[1] = nnz; out[2] = (pd->sopar.transcsc != NULL); out[3] = 1;
}
void refdrv_csc_get(pastix_data_t *pd, int64_t *colptr, int64_t *rows, void *vals, void *tvals)
{
  CscMatrix *c = &pd->cscmtx;
  PASTIX_INT i, j, col = 0, nnz = 0;
  for (i = 0; i < c->cscfnbr; i++) {
    for (j = 0; j < c->cscftab[i].colnbr; j++) colptr[col++] = c->cscftab[i].coltab[j];
    nnz = c->cscftab[i].coltab[c->cscftab[i].colnbr];
  }
  colptr[col] = nnz;
  for (i = 0; i < nnz; i++) rows[i] = c->rowtab[i];
  memcpy(vals, c->valtab, (size_t)nnz * sizeof(PASTIX_FLOAT));
  if (tvals && pd->sopar.transcsc) memcpy(tvals, pd->sopar.transcsc, (size_t)nnz * sizeof(PASTIX_FLOAT));
}
int refdrv_csc_type(pastix_data_t *pd) { return (int)pd->cscmtx.type; }

/* factor panels, flattened: panel of cblk c at offset sum_{k<c} stride_k * width_k */
static size_t panel_size(SolverMatrix *m, PASTIX_INT c)
{
  return (size_t)m->cblktab[c].stride * (size_t)(m->cblktab[c].lcolnum - m->cblktab[c].fcolnum + 1);
}
int refdrv_coef_get(pastix_data_t *pd, void *L, void *U)
{
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT c; size_t off = 0;
  for (c = 0; c < m->cblknbr; c++) {
    size_t sz = panel_size(m, c);
    if (m->cblktab[c].coeftab == NULL) return 1;
    memcpy((PASTIX_FLOAT *)L + off, m->cblktab[c].coeftab, sz * sizeof(PASTIX_FLOAT));
    if (U) {
      if (m->cblktab[c].ucoeftab == NULL) return 2;
      memcpy((PASTIX_FLOAT *)U + off, m->cblktab[c].ucoeftab, sz * sizeof(PASTIX_FLOAT));
    }
    off += sz;
  }
  return 0;
}
int refdrv_coef_set(pastix_data_t *pd, const void *L, const void *U)
{
  SolverMatrix *m = &pd->solvmatr;
  PASTIX_INT c; size_t off = 0;
  for (c = 0; c < m->cblknbr; c++) {
    size_t sz = panel_size(m, c);
    if (m->cblktab[c].coeftab == NULL) return 1;
    memcpy(m->cblktab[c].coeftab, (const PASTIX_FLOAT *)L + off, sz * sizeof(PASTIX_FLOAT));
    if (U) {
      if (m->cblktab[c].ucoeftab == NULL) return 2;
      memcpy(m->cblktab[c].ucoeftab, (const PASTIX_FLOAT *)U + off, sz * sizeof(PASTIX_FLOAT));
    }
    off += sz;
  }
  return 0;
}

/* ||A||_1 of the internal CSC as the reference computes it (csc_intern_compute.c:120) */
double refdrv_norm1(pastix_data_t *pd)
{
  return CscNorm1(&pd->cscmtx, pd->pastix_comm);
}

/* permutation kept by the reference (order.h:52-57): permtab/peritab, 0-based, size n */
void refdrv_order_get(pastix_data_t *pd, int64_t *permtab, int64_t *peritab)
{
  PASTIX_INT i, n = pd->n2 > 0 ? pd->n2 : pd->n;
  for (i = 0; i < n; i++) { permtab[i] = pd->ordemesh.permtab[i]; peritab[i] = pd->ordemesh.peritab[i]; }
}
