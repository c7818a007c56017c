// symbol.cuh — device view of the SolverMatrix and the launch schedule.
// Layout follows blend/src/solver.h:94-117 (SolverCblk/SolverBlok), flattened to
// structure-of-arrays with 32-bit fields where the value fits (n < 2^31) and
// 64-bit panel offsets (coefnbr exceeds 2^31 for the 100^3 27-pt case).
#pragma once
#include <stdint.h>

namespace pb200 {

struct DevSym {
  int cblknbr, bloknbr;
  const int *fcol;      // first column of cblk
  const int *width;     // lcol - fcol + 1
  const int *stride;    // panel leading dimension
  const int *fblok;     // first (diagonal) blok, cblknbr+1 entries
  const int64_t *poff;  // panel offset in the slab, cblknbr+1 entries
  const int *frow;      // first row of blok
  const int *nrow;      // lrow - frow + 1
  const int *fcblk;     // facing cblk
  const int *coefind;   // row offset of blok inside its panel
  const int *col2cblk;  // column -> cblk
};

// one "panel row chunk" task: rows of the off-diagonal part of a cblk
struct RowTask {
  int cblk;
  int tile0;  // first tile index of this task inside its level
};

// one update task: cblk k, off-diagonal blok b1 (sopalin_compute.c:865 compute_1dgemm)
struct UpdTask {
  int cblk;
  int blok;
  int tile0;  // first tile index inside its level
  int ntn;    // tiles along N (rows of b1)
};

// last index i in [lo, hi) with key[i] <= v   (keys ascending; returns lo-1 if none)
__device__ __forceinline__ int upper_le(const int *key, int lo, int hi, int v) {
  int l = lo, h = hi;  // invariant: key[<l] <= v, key[>=h] > v
  while (l < h) {
    int mid = (l + h) >> 1;
    if (key[mid] <= v) l = mid + 1; else h = mid;
  }
  return l - 1;
}

// locate the task owning `tile` (tile0 ascending)
template <class Task>
__device__ __forceinline__ int find_task(const Task *tasks, int ntasks, int tile) {
  int l = 0, h = ntasks;
  while (l < h) {
    int mid = (l + h) >> 1;
    if (tasks[mid].tile0 <= tile) l = mid + 1; else h = mid;
  }
  return l - 1;
}

}  // namespace pb200
