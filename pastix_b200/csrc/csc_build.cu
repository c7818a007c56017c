// csc_build.cu — the internal CSC of PaStiX built on the device.
//
// Reference: CscOrdistrib, src/sopalin/src/csc_intern_build.c:352-570 (macros :100-345).  From the user's CSC
// (1-based colptr/rows, lower triangle for symmetric / Hermitian input) and the ordering permtab it builds the
// matrix in the NEW numbering, column by column: every entry (r, c) lands in column perm[c] with row perm[r];
// for Type 'S' / 'H' every off-diagonal entry is mirrored (conjugated for 'H') into column perm[r]; for 'U' with
// a transposed copy requested the values of A^T are laid out on the pattern of A (transcsc).  Every column is
// then sorted by row (CSC_SORT, :285-345).  On the host this is ~20 ns of cache misses per entry and dominates a
// pastix(API_TASK_NUMFACT) call once the factorization itself runs on the GPU (C2: 38 of 70 ms).
//
// Here: one 64-bit key (kind | new column | new row) + a 32-bit payload (source entry, conj flag) per output
// entry, ONE radix sort over the significant bits (cub::DeviceRadixSort — library plumbing, not the hot path),
// a lower_bound per column for colptr and a gather for rows / values.  The result stays in HBM as the
// assembly input of pb200_assemble_csc (no second upload) and is copied back once for the reference's host-side
// consumers (CscNorm1, the refinement SpMV).
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pastix_b200.h"
#include "csc_build.h"
#include "scalar.cuh"

using namespace pb200;

extern "C" void pb200_set_error(const char *msg);
static int cfail(int code, const std::string &m) { pb200_set_error(m.c_str()); return code; }
#define CCK(call)                                                                        \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) return cfail(PB200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace {

constexpr unsigned CONJ_FLAG = 0x80000000u;

// thread per user column: two key slots per user entry (entry itself; mirror / transposed / unused)
__global__ void k_csc_expand(int64_t n, const int64_t *__restrict__ colptr, const int64_t *__restrict__ rows,
                             const int64_t *__restrict__ perm, int type, int trans, int rb,
                             unsigned long long *keys, unsigned *pay, int *bad) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const unsigned long long nc = (unsigned long long)perm[c];
  const unsigned long long unused = 3ull << (2 * rb), tr = 1ull << (2 * rb);
  for (int64_t it = colptr[c] - 1; it < colptr[c + 1] - 1; ++it) {
    const int64_t r = rows[it] - 1;
    if (r < 0 || r >= n) { *bad = 1; keys[2 * it] = unused; keys[2 * it + 1] = unused; pay[2 * it] = 0; pay[2 * it + 1] = 0; continue; }
    const unsigned long long nr = (unsigned long long)perm[r];
    keys[2 * it] = (nc << rb) | nr; pay[2 * it] = (unsigned)it;
    unsigned long long k2 = unused; unsigned p2 = 0;
    if (type != 'U') {
      if (r != c) { k2 = (nr << rb) | nc; p2 = (unsigned)it | (type == 'H' ? CONJ_FLAG : 0u); }
    } else if (trans == 1) {
      k2 = tr | (nr << rb) | nc; p2 = (unsigned)it;
    }
    keys[2 * it + 1] = k2; pay[2 * it + 1] = p2;
  }
}

__device__ __forceinline__ int64_t lower_bound_u64(const unsigned long long *a, int64_t n, unsigned long long v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// colptr[c] = first sorted key of column c; extra[0] = entries of the transposed copy
__global__ void k_csc_colptr(int64_t n, int rb, const unsigned long long *__restrict__ keys, int64_t N, int64_t *colptr, int64_t *extra) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= n) colptr[c] = lower_bound_u64(keys, N, (unsigned long long)c << rb);
  if (c == n + 1) extra[0] = lower_bound_u64(keys, N, 3ull << (2 * rb)) - lower_bound_u64(keys, N, 1ull << (2 * rb));
}

template <class T>
__global__ void k_csc_gather(int64_t nnz, int rb, const unsigned long long *__restrict__ keys, const unsigned *__restrict__ pay,
                             const T *__restrict__ uvals, int *rows, T *vals, T *tvals, int64_t toff, int *bad) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  rows[k] = (int)(keys[k] & ((1ull << rb) - 1));
  const unsigned p = pay[k];
  T v = uvals[p & ~CONJ_FLAG];
  if (p & CONJ_FLAG) v = ST<T>::conj(v);
  vals[k] = v;
  if (tvals) {
    // the k-th transposed key must sit at the very (column, row) of the k-th main key: a pattern that is not
    // symmetric would pair values by sorted position only (the reference requires a symmetric pattern for LU too)
    const unsigned long long pos = (1ull << (2 * rb)) - 1;
    if ((keys[toff + k] & pos) != (keys[k] & pos)) *bad = 2;
    tvals[k] = uvals[pay[toff + k]];
  }
}

// CscNorm1 (csc_intern_compute.c:120-176): max over columns of the sum of |a_ij|, summed in storage order by one
// thread per column (the same additions in the same order as the reference's loop, so the result is identical
// for real types; positive doubles order like their bit patterns, hence the integer atomicMax)
template <class T>
__global__ void k_csc_norm1(int64_t n, const int64_t *__restrict__ colptr, const T *__restrict__ vals, unsigned long long *out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int64_t k = colptr[c]; k < colptr[c + 1]; ++k) s += (double)ST<T>::abs(vals[k]);
  atomicMax(out, (unsigned long long)__double_as_longlong(s));
}

__global__ void k_csc_widen(int64_t nnz, const int *__restrict__ rows, int64_t *out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) out[k] = rows[k];
}

size_t esize_of(int flt) {
  switch (flt) {
    case PB200_REALSINGLE: return 4;
    case PB200_REALDOUBLE: return 8;
    case PB200_COMPLEXSINGLE: return 8;
    case PB200_COMPLEXDOUBLE: return 16;
  }
  return 0;
}

template <class V>
int ensure(V **p, size_t *cap, size_t need) {
  if (need <= *cap && *p) return 0;
  cudaFree(*p); *p = nullptr; *cap = 0;
  const size_t want = need + need / 8 + 256;
  if (cudaMalloc((void **)p, want) != cudaSuccess) return cfail(PB200_ERR_NOMEM, "cudaMalloc(internal CSC work space) failed");
  *cap = want;
  return 0;
}

}  // namespace

extern "C" int pb200_csc_create(pb200_csc_t **out, int flttype, int device) {
  if (!out) return cfail(PB200_ERR_BADARG, "null argument");
  *out = nullptr;
  if (esize_of(flttype) == 0) return cfail(PB200_ERR_BADARG, "bad flttype");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return cfail(PB200_ERR_CUDA, "no CUDA device: pastix_b200 has no CPU fallback");
  if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
  if (device >= ndev) return cfail(PB200_ERR_BADARG, "bad device");
  CCK(cudaSetDevice(device));
  pb200_csc_t *c = new pb200_csc_t();
  c->flt = flttype; c->esize = esize_of(flttype); c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return cfail(PB200_ERR_CUDA, "cudaStreamCreate failed"); }
  *out = c;
  return PB200_SUCCESS;
}

extern "C" int pb200_csc_destroy(pb200_csc_t *c) {
  if (!c) return PB200_SUCCESS;
  cudaSetDevice(c->device);
  cudaFree(c->d_ucolptr); cudaFree(c->d_urows); cudaFree(c->d_uvals); cudaFree(c->d_perm);
  cudaFree(c->d_keys0); cudaFree(c->d_keys1); cudaFree(c->d_pay0); cudaFree(c->d_pay1); cudaFree(c->d_tmp);
  cudaFree(c->d_colptr); cudaFree(c->d_rows); cudaFree(c->d_vals); cudaFree(c->d_tvals); cudaFree(c->d_extra);
  for (int i = 0; i < PB200_CSC_NPIN; ++i) { if (c->pin[i]) cudaFreeHost(c->pin[i]); if (c->pstream[i]) cudaStreamDestroy(c->pstream[i]); }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return PB200_SUCCESS;
}

template <class T>
static void launch_gather(pb200_csc_t *c, int64_t nnz, int rb, bool want_t) {
  k_csc_gather<T><<<(unsigned)((nnz + 255) / 256), 256, 0, c->stream>>>(nnz, rb, c->d_keys1, c->d_pay1, (const T *)c->d_uvals, c->d_rows,
                                                                       (T *)c->d_vals, want_t ? (T *)c->d_tvals : nullptr, nnz,
                                                                       reinterpret_cast<int *>(c->d_extra + 2));
}

// Pageable host -> device copies of the user's CSC (the arrays pastix() receives: nothing is known about their
// lifetime, so they are not registered).  A plain cudaMemcpy stages them through the driver's pinned buffer with ONE
// thread (~12 GB/s: 2.1 of the 27 ms a pastix(API_TASK_NUMFACT) call takes on the 64^3 problem, 23 ms of 300 on the
// 100^3 one); here the same host threads and pinned pieces as parallel_d2h copy 1 MB chunks of all the arrays into
// pinned memory and hand them to the DMA engine on their own streams, eight chunks in flight per thread.
struct H2DSeg { void *dst; const void *src; size_t bytes; };
static int parallel_h2d(pb200_csc_t *c, const H2DSeg *segs, int nseg);

extern "C" int pb200_csc_build(pb200_csc_t *c, char type, int64_t n, const int64_t *colptr, const int64_t *rows, const void *values,
                               const int64_t *permtab, int trans, int64_t *nnz_out) {
  if (!c || !colptr || !rows || !values || !permtab || !nnz_out) return cfail(PB200_ERR_BADARG, "null argument");
  if (type != 'S' && type != 'H' && type != 'U') return cfail(PB200_ERR_BADARG, "matrix type must be S, H or U");
  if (n <= 0 || n >= (1LL << 31) - 1) return cfail(PB200_ERR_BADARG, "bad n");
  if (colptr[0] != 1) return cfail(PB200_ERR_BADARG, "user CSC must use Fortran (1-based) numbering");
  const int64_t unz = colptr[n] - 1;
  if (unz < 0 || unz >= (1LL << 31)) return cfail(PB200_ERR_BADARG, "nnz of the user CSC out of range");
  if (type != 'U' && trans == 1) trans = 0;
  CCK(cudaSetDevice(c->device));
  c->valid = false;
  int rb = 1; while ((1LL << rb) <= n) ++rb;          // column field must hold the value n
  const int64_t N = 2 * unz;
  { int rc;
    if ((rc = ensure(&c->d_ucolptr, &c->cap_ucolptr, (size_t)(n + 1) * 8))) return rc;
    if ((rc = ensure(&c->d_perm, &c->cap_perm, (size_t)n * 8))) return rc;
    if ((rc = ensure(&c->d_urows, &c->cap_urows, (size_t)std::max<int64_t>(unz, 1) * 8))) return rc;
    if ((rc = ensure(&c->d_uvals, &c->cap_uvals, (size_t)std::max<int64_t>(unz, 1) * c->esize))) return rc;
    if ((rc = ensure(&c->d_keys0, &c->cap_keys0, (size_t)std::max<int64_t>(N, 1) * 8))) return rc;
    if ((rc = ensure(&c->d_keys1, &c->cap_keys1, (size_t)std::max<int64_t>(N, 1) * 8))) return rc;
    if ((rc = ensure(&c->d_pay0, &c->cap_pay0, (size_t)std::max<int64_t>(N, 1) * 4))) return rc;
    if ((rc = ensure(&c->d_pay1, &c->cap_pay1, (size_t)std::max<int64_t>(N, 1) * 4))) return rc;
    if ((rc = ensure(&c->d_colptr, &c->cap_colptr, (size_t)(n + 1) * 8))) return rc;
    if ((rc = ensure(&c->d_extra, &c->cap_extra, 4 * 8))) return rc;
  }
  {
    const H2DSeg segs[4] = {{c->d_ucolptr, colptr, (size_t)(n + 1) * 8}, {c->d_perm, permtab, (size_t)n * 8},
                            {c->d_urows, rows, (size_t)unz * 8}, {c->d_uvals, values, (size_t)unz * c->esize}};
    if (parallel_h2d(c, segs, 4)) return cfail(PB200_ERR_CUDA, "host -> device copy of the user's CSC failed");
  }
  CCK(cudaMemsetAsync(c->d_extra, 0, 4 * 8, c->stream));
  int *d_bad = reinterpret_cast<int *>(c->d_extra + 2);
  k_csc_expand<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(n, c->d_ucolptr, c->d_urows, c->d_perm, (int)type, trans, rb,
                                                                  c->d_keys0, c->d_pay0, d_bad);
  size_t tmp_bytes = 0;
  CCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->d_keys0, c->d_keys1, c->d_pay0, c->d_pay1, N, 0, 2 * rb + 2, c->stream));
  { int rc = ensure(&c->d_tmp, &c->cap_tmp, std::max<size_t>(tmp_bytes, 16)); if (rc) return rc; }
  if (N > 0) CCK(cub::DeviceRadixSort::SortPairs(c->d_tmp, tmp_bytes, c->d_keys0, c->d_keys1, c->d_pay0, c->d_pay1, N, 0, 2 * rb + 2, c->stream));
  k_csc_colptr<<<(unsigned)((n + 2 + 255) / 256), 256, 0, c->stream>>>(n, rb, c->d_keys1, N, c->d_colptr, c->d_extra);
  int64_t tail[3] = {0, 0, 0}, nnz = 0;
  CCK(cudaMemcpyAsync(&nnz, c->d_colptr + n, 8, cudaMemcpyDeviceToHost, c->stream));
  CCK(cudaMemcpyAsync(tail, c->d_extra, 3 * 8, cudaMemcpyDeviceToHost, c->stream));
  CCK(cudaStreamSynchronize(c->stream));
  if (tail[2] != 0) return cfail(PB200_ERR_BADARG, "user CSC: row index out of range");
  if (trans == 1 && tail[0] != nnz) return cfail(PB200_ERR_STRUCT, "unsymmetric pattern: the transposed copy does not fit the pattern of A");
  { int rc;
    if ((rc = ensure(&c->d_rows, &c->cap_rows, (size_t)std::max<int64_t>(nnz, 1) * 4))) return rc;
    if ((rc = ensure(&c->d_vals, &c->cap_vals, (size_t)std::max<int64_t>(nnz, 1) * c->esize))) return rc;
    if (trans) { if ((rc = ensure(&c->d_tvals, &c->cap_tvals, (size_t)std::max<int64_t>(nnz, 1) * c->esize))) return rc; }
  }
  if (nnz > 0) {
    switch (c->flt) {
      case PB200_REALSINGLE: launch_gather<float>(c, nnz, rb, trans == 1); break;
      case PB200_REALDOUBLE: launch_gather<double>(c, nnz, rb, trans == 1); break;
      case PB200_COMPLEXSINGLE: launch_gather<cfloat>(c, nnz, rb, trans == 1); break;
      case PB200_COMPLEXDOUBLE: launch_gather<cdouble>(c, nnz, rb, trans == 1); break;
    }
    if (trans == 2) CCK(cudaMemcpyAsync(c->d_tvals, c->d_vals, (size_t)nnz * c->esize, cudaMemcpyDeviceToDevice, c->stream));
  }
  CCK(cudaGetLastError());
  if (trans == 1) {
    int bad = 0;
    CCK(cudaMemcpyAsync(&bad, c->d_extra + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CCK(cudaStreamSynchronize(c->stream));
    if (bad) return cfail(PB200_ERR_STRUCT, "unsymmetric pattern: the transposed copy does not fit the pattern of A");
  }
  CCK(cudaStreamSynchronize(c->stream));
  c->n = n; c->nnz = nnz; c->has_t = (trans != 0); c->valid = true; c->type = type;
  *nnz_out = nnz;
  return PB200_SUCCESS;
}

// Large device -> pageable host copies: the destination arrays are freshly malloc'ed by the caller (the reference
// frees and re-allocates its CscMatrix on every NUMFACT), so a plain cudaMemcpy is bound by first-touch page faults of
// ONE thread (C3: 424 MB in ~200 ms).  Here several host threads each pull their slice through a small pinned
// buffer on their own stream, so that the DMA, the faults and the memcpy of different slices overlap.
static const size_t kPiece = (size_t)8 << 20;
static int parallel_d2h(pb200_csc_t *c, void *dst, const void *src, size_t bytes) {
  unsigned nt = std::min<unsigned>(PB200_CSC_NPIN, std::max(1u, std::thread::hardware_concurrency()));
  nt = (unsigned)std::min<size_t>(nt, (bytes + kPiece - 1) / kPiece);
  if (bytes < ((size_t)16 << 20) || nt < 2) {
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return 0;
  }
  // pinned pieces and streams are created once per pb200_csc_t (cudaHostAlloc costs milliseconds)
  for (unsigned i = 0; i < nt; ++i)
    if (!c->pin[i]) {
      if (cudaHostAlloc(&c->pin[i], kPiece, cudaHostAllocDefault) != cudaSuccess) { c->pin[i] = nullptr; return 1; }
      if (cudaStreamCreateWithFlags(&c->pstream[i], cudaStreamNonBlocking) != cudaSuccess) return 1;
    }
  {
    // first touch with 2 MB pages where the kernel allows it (transparent huge pages in madvise/always mode):
    // ~500x fewer page faults on a fresh malloc'ed destination; a refusal is harmless
    const uintptr_t lo = ((uintptr_t)dst + 4095) & ~(uintptr_t)4095, hi = ((uintptr_t)dst + bytes) & ~(uintptr_t)4095;
    if (hi > lo) (void)madvise((void *)lo, hi - lo, MADV_HUGEPAGE);
  }
  std::vector<std::thread> th;
  std::vector<int> rc(nt, 0);
  const size_t slice = ((bytes / nt) + 4095) & ~(size_t)4095;
  const int device = c->device;
  for (unsigned i = 0; i < nt; ++i)
    th.emplace_back([&, i]() {
      const size_t lo = std::min(bytes, (size_t)i * slice), hi = (i + 1 == nt) ? bytes : std::min(bytes, lo + slice);
      if (hi <= lo) return;
      if (cudaSetDevice(device) != cudaSuccess) { rc[i] = 1; return; }
      for (size_t off = lo; off < hi; off += kPiece) {
        const size_t n = std::min(kPiece, hi - off);
        if (cudaMemcpyAsync(c->pin[i], (const char *)src + off, n, cudaMemcpyDeviceToHost, c->pstream[i]) != cudaSuccess ||
            cudaStreamSynchronize(c->pstream[i]) != cudaSuccess) { rc[i] = 1; break; }
        memcpy((char *)dst + off, c->pin[i], n);
      }
    });
  for (auto &t : th) t.join();
  for (int r : rc) if (r) return 1;
  return 0;
}

static int parallel_h2d(pb200_csc_t *c, const H2DSeg *segs, int nseg) {
  static const size_t kChunk = (size_t)1 << 20;
  size_t total = 0;
  for (int k = 0; k < nseg; ++k) total += segs[k].bytes;
  unsigned nt = std::min<unsigned>(PB200_CSC_NPIN, std::max(1u, std::thread::hardware_concurrency()));
  nt = (unsigned)std::min<size_t>(nt, (total + 4 * kChunk - 1) / (4 * kChunk));
  if (total < ((size_t)8 << 20) || nt < 2 || getenv("PB200_PLAIN_H2D") != nullptr) {
    for (int k = 0; k < nseg; ++k)
      if (segs[k].bytes && cudaMemcpyAsync(segs[k].dst, segs[k].src, segs[k].bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return 1;
    return 0;
  }
  for (unsigned i = 0; i < nt; ++i)
    if (!c->pin[i]) {
      if (cudaHostAlloc(&c->pin[i], kPiece, cudaHostAllocDefault) != cudaSuccess) { c->pin[i] = nullptr; return 1; }
      if (cudaStreamCreateWithFlags(&c->pstream[i], cudaStreamNonBlocking) != cudaSuccess) return 1;
    }
  // chunk list over all segments; thread i takes chunks i, i + nt, ...
  struct Chunk { char *dst; const char *src; size_t n; };
  std::vector<Chunk> chunks;
  for (int k = 0; k < nseg; ++k)
    for (size_t off = 0; off < segs[k].bytes; off += kChunk)
      chunks.push_back({(char *)segs[k].dst + off, (const char *)segs[k].src + off, std::min(kChunk, segs[k].bytes - off)});
  std::vector<std::thread> th;
  std::vector<int> rc(nt, 0);
  const int device = c->device;
  const size_t slots = kPiece / kChunk;
  for (unsigned i = 0; i < nt; ++i)
    th.emplace_back([&, i]() {
      if (cudaSetDevice(device) != cudaSuccess) { rc[i] = 1; return; }
      size_t used = 0;
      for (size_t q = i; q < chunks.size(); q += nt) {
        if (used == slots) { if (cudaStreamSynchronize(c->pstream[i]) != cudaSuccess) { rc[i] = 1; return; } used = 0; }
        char *p = (char *)c->pin[i] + used * kChunk;
        memcpy(p, chunks[q].src, chunks[q].n);
        if (cudaMemcpyAsync(chunks[q].dst, p, chunks[q].n, cudaMemcpyHostToDevice, c->pstream[i]) != cudaSuccess) { rc[i] = 1; return; }
        ++used;
      }
      if (cudaStreamSynchronize(c->pstream[i]) != cudaSuccess) rc[i] = 1;
    });
  for (auto &t : th) t.join();
  for (int r : rc) if (r) return 1;
  return 0;   // every copy has completed: the kernels on c->stream that follow see the data
}

extern "C" int pb200_csc_fetch(pb200_csc_t *c, int64_t *colptr, int64_t *rows, void *values, void *tvalues) {
  if (!c || !colptr || !rows || !values) return cfail(PB200_ERR_BADARG, "null argument");
  if (!c->valid) return cfail(PB200_ERR_STATE, "no internal CSC built");
  if (tvalues && !c->has_t) return cfail(PB200_ERR_STATE, "no transposed values were built");
  CCK(cudaSetDevice(c->device));
  CCK(cudaMemcpyAsync(colptr, c->d_colptr, (size_t)(c->n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  if (c->nnz > 0) {
    int64_t *wide = reinterpret_cast<int64_t *>(c->d_keys0);   // 2*unz*8 bytes >= nnz*8
    k_csc_widen<<<(unsigned)((c->nnz + 255) / 256), 256, 0, c->stream>>>(c->nnz, c->d_rows, wide);
    CCK(cudaStreamSynchronize(c->stream));
    if (parallel_d2h(c, rows, wide, (size_t)c->nnz * 8) || parallel_d2h(c, values, c->d_vals, (size_t)c->nnz * c->esize) ||
        (tvalues && parallel_d2h(c, tvalues, c->d_tvals, (size_t)c->nnz * c->esize)))
      return cfail(PB200_ERR_CUDA, "device -> host copy of the internal CSC failed");
  }
  CCK(cudaStreamSynchronize(c->stream));
  return PB200_SUCCESS;
}

extern "C" int pb200_csc_fetch_colptr(pb200_csc_t *c, int64_t *colptr) {
  if (!c || !colptr) return cfail(PB200_ERR_BADARG, "null argument");
  if (!c->valid) return cfail(PB200_ERR_STATE, "no internal CSC built");
  CCK(cudaSetDevice(c->device));
  CCK(cudaMemcpyAsync(colptr, c->d_colptr, (size_t)(c->n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
  CCK(cudaStreamSynchronize(c->stream));
  return PB200_SUCCESS;
}

extern "C" int pb200_csc_norm1(pb200_csc_t *c, double *norm) {
  if (!c || !norm) return cfail(PB200_ERR_BADARG, "null argument");
  if (!c->valid) return cfail(PB200_ERR_STATE, "no internal CSC built");
  CCK(cudaSetDevice(c->device));
  unsigned long long *d_out = reinterpret_cast<unsigned long long *>(c->d_extra + 3);
  CCK(cudaMemsetAsync(d_out, 0, 8, c->stream));
  const unsigned grid = (unsigned)((c->n + 127) / 128);
  switch (c->flt) {
    case PB200_REALSINGLE: k_csc_norm1<float><<<grid, 128, 0, c->stream>>>(c->n, c->d_colptr, (const float *)c->d_vals, d_out); break;
    case PB200_REALDOUBLE: k_csc_norm1<double><<<grid, 128, 0, c->stream>>>(c->n, c->d_colptr, (const double *)c->d_vals, d_out); break;
    case PB200_COMPLEXSINGLE: k_csc_norm1<cfloat><<<grid, 128, 0, c->stream>>>(c->n, c->d_colptr, (const cfloat *)c->d_vals, d_out); break;
    case PB200_COMPLEXDOUBLE: k_csc_norm1<cdouble><<<grid, 128, 0, c->stream>>>(c->n, c->d_colptr, (const cdouble *)c->d_vals, d_out); break;
  }
  unsigned long long bits = 0;
  CCK(cudaMemcpyAsync(&bits, d_out, 8, cudaMemcpyDeviceToHost, c->stream));
  CCK(cudaStreamSynchronize(c->stream));
  memcpy(norm, &bits, 8);
  return PB200_SUCCESS;
}
