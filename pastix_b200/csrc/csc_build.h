// csc_build.h — device-resident internal CSC (csc_build.cu), shared with engine.cu (pb200_assemble_csc)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PB200_CSC_NPIN 8
struct pb200_csc_s {
  int flt = 0, device = 0;
  size_t esize = 0;
  cudaStream_t stream = nullptr;
  // user CSC as uploaded (1-based) and the ordering
  int64_t *d_ucolptr = nullptr, *d_urows = nullptr, *d_perm = nullptr; void *d_uvals = nullptr;
  size_t cap_ucolptr = 0, cap_urows = 0, cap_perm = 0, cap_uvals = 0;
  // sort work space
  unsigned long long *d_keys0 = nullptr, *d_keys1 = nullptr; unsigned *d_pay0 = nullptr, *d_pay1 = nullptr; void *d_tmp = nullptr;
  size_t cap_keys0 = 0, cap_keys1 = 0, cap_pay0 = 0, cap_pay1 = 0, cap_tmp = 0;
  // result: internal CSC, 0-based, new numbering, rows sorted in every column
  int64_t *d_colptr = nullptr; int *d_rows = nullptr; void *d_vals = nullptr, *d_tvals = nullptr; int64_t *d_extra = nullptr;
  size_t cap_colptr = 0, cap_rows = 0, cap_vals = 0, cap_tvals = 0, cap_extra = 0;
  int64_t n = 0, nnz = 0;
  bool has_t = false, valid = false;
  char type = 'S';   // 'S' / 'H' / 'U' as given to pb200_csc_build (CscMatrix.type, blend/src/csc.h)
  // pinned staging pieces of the parallel device -> host fetch (one per worker thread, created on first use)
  void *pin[PB200_CSC_NPIN] = {}; cudaStream_t pstream[PB200_CSC_NPIN] = {};
};
