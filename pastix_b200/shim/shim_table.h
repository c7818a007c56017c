/* shim_table.h — side table SolverMatrix* -> device handle, shared by the four factorization
 * variants of one precision build and by shim_hooks.c */
#ifndef PB200_SHIM_TABLE_H
#define PB200_SHIM_TABLE_H
#include <pthread.h>
#include "pastix_b200.h"
typedef struct pb200_shim_entry_s {
  const SolverMatrix *m;
  pb200_handle_t     *h;
  int                 facto;
  int                 schur;      /* handle built for IPARM_SCHUR == API_YES */
  int                 factorized;
  double              critere;
  pb200_csc_t        *csc;        /* internal CSC built on the device by CscOrdistrib (shim_csc.c) */
  int                 csc_fresh;  /* set by CscOrdistrib, consumed by the next numeric factorization */
} pb200_shim_entry_t;
#define PB200_SHIM_MAX 64
#define shim_table  PASTIX_PREFIX_F(pb200_shim_table)
#define shim_mutex  PASTIX_PREFIX_F(pb200_shim_mutex)
extern pb200_shim_entry_t shim_table[PB200_SHIM_MAX];
extern pthread_mutex_t    shim_mutex;
#endif
