/* shim_table.h — side table SolverMatrix* -> device handle, shared by the four factorization
 * variants of one precision build, shim_csc.c, shim_raff.c and shim_hooks.c (which owns it).
 * An entry is keyed by the ADDRESS of the SolverMatrix (&pastix_data->solvmatr, pastixstr.h:48), which survives a
 * re-run of API_TASK_ANALYSE and may be handed out again by malloc after API_TASK_CLEAN: the entry therefore also
 * carries a structural fingerprint of the SolverMatrix its handle was built for (checked at every numeric
 * factorization), and the reference's own release points — CoefMatrix_Free (coefinit.c:479, called from
 * pastix_task_blend pastix.c:2716 and pastix_fillin_csc :3391) and solverExit (solverRealloc.c:217, called from
 * pastix_task_clean pastix.c:4539) — are intercepted (build_dropin.sh, objcopy --redefine-sym) to drop it. */
#ifndef PB200_SHIM_TABLE_H
#define PB200_SHIM_TABLE_H
#include <pthread.h>
#include <stdint.h>
#include "pastix_b200.h"
typedef struct pb200_shim_entry_s {
  const SolverMatrix *m;
  pb200_handle_t     *h;          /* the handle up_down / refinement / read-backs go through (= hs[0]) */
  pb200_handle_t     *hs[8];      /* iparm[IPARM_CUDA_NBR] = ngpu > 1: one handle per device (pb200_attach_local group) */
  int                 ngpu;
  int                 facto;
  int                 schur;      /* handle built for IPARM_SCHUR == API_YES */
  int                 factorized;
  double              critere;
  pb200_csc_t        *csc;        /* internal CSC built on the device by CscOrdistrib (shim_csc.c) */
  int                 csc_fresh;  /* set by CscOrdistrib, consumed by the next numeric factorization */
  uint64_t            fp[2];      /* fingerprint of the SolverMatrix the handle was built for */
  int                 herm;       /* internal CSC type 'H' (conjugated transposed values) */
  /* host copy of the internal CSC filled on demand: CscOrdistrib (shim_csc.c) allocates CSC_ROWTAB / CSC_VALTAB /
   * transcsc and fills CSC_COLTAB, the rows / values stay in HBM until a host-side reader asks (pb200_shim_csc_host) */
  int                 host_stale;
  int64_t             lazy_ncol;
  void               *lazy_rows, *lazy_vals, *lazy_tvals;
} pb200_shim_entry_t;
#define pb200_shim_entry        PASTIX_PREFIX_F(pb200_shim_entry)
#define pb200_shim_entry_drop   PASTIX_PREFIX_F(pb200_shim_entry_drop)
#define pb200_shim_fingerprint  PASTIX_PREFIX_F(pb200_shim_fingerprint)
#define pb200_shim_live_entries PASTIX_PREFIX_F(pb200_shim_live_entries)
#define pb200_shim_csc_host     PASTIX_PREFIX_F(pb200_shim_csc_host)
/* entry of SolverMatrix m (created empty when absent and `create` is set); entries never move */
pb200_shim_entry_t *pb200_shim_entry(const SolverMatrix *m, int create);
/* destroy the handle / device CSC of m's entry and forget it */
void pb200_shim_entry_drop(const SolverMatrix *m);
/* cblknbr, bloknbr, coefnbr and a hash of every cblktab / bloktab field the device schedule is derived from */
void pb200_shim_fingerprint(const SolverMatrix *m, uint64_t fp[2]);
int  pb200_shim_live_entries(void);
/* make the host arrays of m's internal CSC valid (no-op unless CscOrdistrib left them to be filled on demand) */
void pb200_shim_csc_host(const SolverMatrix *m);
#endif
