/*
 * shim_csc.c — CscOrdistrib of the drop-in library: the internal CSC built on the B200.
 *
 * Replaces (same name, same arguments, same result) CscOrdistrib, src/sopalin/src/csc_intern_build.c:352-570,
 * which pastix_fillin_csc calls at the start of every API_TASK_NUMFACT (pastix.c:3326).  The reference's own
 * object is still linked, with this one symbol renamed to CscOrdistrib_hostref (build_dropin.sh, objcopy); it
 * serves the configurations the device path does not cover (more than one dof per node) and PB200_HOST_CSC=1
 * (the parity tests compare the two).  Compiled once per precision.
 *
 * What it does: hands the user's CSC and Order.permtab to pb200_csc_build (sort-based, csrc/csc_build.cu),
 * fills the CscMatrix from the result exactly as the reference lays it out — one CSC_COLTAB per column block
 * holding global offsets into CSC_ROWTAB / CSC_VALTAB (csc_intern_build.c:100-150, 520-532), MALLOC_INTERNed so
 * that CscExit keeps working — and leaves the device copy attached to the SolverMatrix entry of the side table,
 * where the next numeric factorization picks it up (sopalin_b200_shim.c, shim_assemble) instead of uploading
 * the host arrays again.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "nompi.h"
#include "common_pastix.h"
#include "tools.h"
#include "sopalin_define.h"
#include "symbol.h"
#include "ftgt.h"
#include "csc.h"
#include "updown.h"
#include "queue.h"
#include "bulles.h"
#include "solver.h"
#include "order.h"
#include "csc_intern_build.h"
#include "shim_table.h"

#if defined(TYPE_COMPLEX) && defined(PREC_DOUBLE)
#define PB200_FLT PB200_COMPLEXDOUBLE
#elif defined(TYPE_COMPLEX)
#define PB200_FLT PB200_COMPLEXSINGLE
#elif defined(PREC_DOUBLE)
#define PB200_FLT PB200_REALDOUBLE
#else
#define PB200_FLT PB200_REALSINGLE
#endif

void CscOrdistrib_hostref(CscMatrix *thecsc, char *Type, PASTIX_FLOAT **transcsc, const Order *ord,
                          PASTIX_INT Nrow, PASTIX_INT Ncol, PASTIX_INT Nnzero, PASTIX_INT *colptr, PASTIX_INT *rowind,
                          PASTIX_FLOAT *val, PASTIX_INT forcetrans, const SolverMatrix *solvmtx, PASTIX_INT procnum, PASTIX_INT dof);

static pb200_shim_entry_t *csc_entry(const SolverMatrix *m) { return pb200_shim_entry(m, 1); }

void CscOrdistrib(CscMatrix *thecsc, char *Type, PASTIX_FLOAT **transcsc, const Order *ord,
                  PASTIX_INT Nrow, PASTIX_INT Ncol, PASTIX_INT Nnzero, PASTIX_INT *colptr, PASTIX_INT *rowind,
                  PASTIX_FLOAT *val, PASTIX_INT forcetrans, const SolverMatrix *solvmtx, PASTIX_INT procnum, PASTIX_INT dof)
{
  pb200_shim_entry_t *e;
  int64_t nnz = 0, *gcol = NULL;
  PASTIX_INT index, iter;
  int trans = 0;
  double t0 = clockGet(), t1, t2;

  if (sizeof(PASTIX_INT) != sizeof(int64_t) || dof != 1 || getenv("PB200_HOST_CSC") != NULL ||
      (e = csc_entry(solvmtx)) == NULL) {
    CscOrdistrib_hostref(thecsc, Type, transcsc, ord, Nrow, Ncol, Nnzero, colptr, rowind, val, forcetrans, solvmtx, procnum, dof);
    return;
  }
  e->csc_fresh = 0;
  e->host_stale = 0;
  if (e->csc == NULL && pb200_csc_create(&e->csc, PB200_FLT, -1) != PB200_SUCCESS) {
    errorPrint("pastix_b200: pb200_csc_create: %s", pb200_last_error());
    EXIT(MOD_SOPALIN, INTERNAL_ERR);
  }
  if (transcsc != NULL) {
    if (Type[1] == 'S' || Type[1] == 'H') trans = (forcetrans == API_YES) ? 2 : 0;
    else trans = 1;
  }
  if (pb200_csc_build(e->csc, Type[1], (int64_t)Ncol, (const int64_t *)colptr, (const int64_t *)rowind, val,
                      (const int64_t *)ord->permtab, trans, &nnz) != PB200_SUCCESS) {
    errorPrint("pastix_b200: pb200_csc_build: %s", pb200_last_error());
    EXIT(MOD_SOPALIN, INTERNAL_ERR);
  }
  t1 = clockGet();

  thecsc->type = Type[1];
  CSC_FNBR(thecsc) = solvmtx->cblknbr;
  MALLOC_INTERN(CSC_FTAB(thecsc), CSC_FNBR(thecsc), CscFormat);
  MALLOC_INTERN(CSC_ROWTAB(thecsc), nnz, PASTIX_INT);
  MALLOC_INTERN(CSC_VALTAB(thecsc), nnz, PASTIX_FLOAT);
  MALLOC_INTERN(gcol, Ncol + 1, int64_t);
  if (trans == 1) { MALLOC_INTERN(*transcsc, nnz, PASTIX_FLOAT); }
  if (getenv("PB200_EAGER_CSC") != NULL) {
    if (pb200_csc_fetch(e->csc, gcol, (int64_t *)CSC_ROWTAB(thecsc), CSC_VALTAB(thecsc), trans == 1 ? *transcsc : NULL) != PB200_SUCCESS) {
      errorPrint("pastix_b200: pb200_csc_fetch: %s", pb200_last_error());
      EXIT(MOD_SOPALIN, INTERNAL_ERR);
    }
  } else {
    /* the numeric phase, the static-pivot threshold and the refinement read the copy in HBM; the host arrays are
     * allocated (CscExit frees them as usual) and filled when a host-side reader of the CscMatrix shows up:
     * Csc2updown (right-hand side generated from the matrix, pastix.c:716), the host statistics of the refinement,
     * the test hooks — pb200_shim_csc_host.  Saves the 16 * nnz byte copy on every API_TASK_NUMFACT. */
    if (pb200_csc_fetch_colptr(e->csc, gcol) != PB200_SUCCESS) {
      errorPrint("pastix_b200: pb200_csc_fetch_colptr: %s", pb200_last_error());
      EXIT(MOD_SOPALIN, INTERNAL_ERR);
    }
    e->lazy_ncol = (int64_t)Ncol; e->lazy_rows = CSC_ROWTAB(thecsc); e->lazy_vals = CSC_VALTAB(thecsc);
    e->lazy_tvals = trans == 1 ? (void *)*transcsc : NULL;
    e->host_stale = 1;
  }
  if (trans == 2) *transcsc = CSC_VALTAB(thecsc);                /* CSC_ALLOC, csc_intern_build.c:163-170 */
  for (index = 0; index < solvmtx->cblknbr; index++) {
    PASTIX_INT fcolnum = solvmtx->cblktab[index].fcolnum;
    PASTIX_INT lcolnum = solvmtx->cblktab[index].lcolnum;
    CSC_COLNBR(thecsc, index) = lcolnum - fcolnum + 1;
    MALLOC_INTERN(CSC_COLTAB(thecsc, index), CSC_COLNBR(thecsc, index) + 1, PASTIX_INT);
    for (iter = 0; iter < CSC_COLNBR(thecsc, index) + 1; iter++)
      CSC_COL(thecsc, index, iter) = (PASTIX_INT)gcol[fcolnum + iter];
  }
  memFree_null(gcol);
  e->csc_fresh = 1;
  t2 = clockGet();
  if (getenv("PB200_SHIM_TIMING") != NULL)
    fprintf(stderr, "[pb200 shim] CscOrdistrib on the device: upload + sort + gather %.1f ms, CscMatrix (host rows / values on demand) %.1f ms (nnz %ld)\n",
            (t1 - t0) * 1e3, (t2 - t1) * 1e3, (long)nnz);
  (void)Nrow; (void)Nnzero; (void)procnum;
}
