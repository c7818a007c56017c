// kernels_raff.cuh — the vector back end of the refinement drivers on device-resident vectors.
//
// Reference: the `struct solver` operations that raff_gmres.c / raff_grad.c / raff_bicgstab.c / raff_pivot.c are
// written against (src/sopalin/src/raff_functions.c:100-650), which in the reference run on host vectors through
// CscAx / CscbMAx / CscGradBeta / CscGmresBeta / CscNormFro / CscCopy / CscScal / CscAXPY
// (csc_intern_compute.c:448-1700).  Here the vectors live in HBM (managed allocations, so that the few scalars the
// drivers dereference on the host keep working), the matrix is the internal CSC already resident for the assembly,
// and the preconditioner is the up_down of the same handle — an iteration never crosses PCIe.
//
// SpMV without atomics: the reference's product scatters column by column, r[row] += a * x[col].  The internal CSC
// has a symmetric pattern with every column sorted by row, so row c of A is read off column c:
//   type 'S': A(c, j) = val(j-th entry of column c);  'H': its conjugate;  'U': the transposed values kept for LU.
// One thread per row gathers in increasing column order — the same additions in the same order as the reference's
// sequential scatter, no reductions, reproducible.  A^T x (IPARM_TRANSPOSE_SOLVE) is the column itself.
#pragma once
#include "scalar.cuh"

namespace pb200 {

#define PB200_RAFF_BLOCKS 296   // partial sums of the dot products (2 per SM), summed in index order on the host

template <class T>
__global__ void k_raff_spmv(int64_t n, const int64_t *__restrict__ colptr, const int *__restrict__ rows, const T *__restrict__ vals,
                            int conjv, const T *__restrict__ x, const T *__restrict__ b, T *r) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  T s = ST<T>::zero();
  for (int64_t k = colptr[c]; k < colptr[c + 1]; ++k) {
    T a = vals[k];
    if (conjv) a = ST<T>::conj(a);
    s = s + a * x[rows[k]];
  }
  r[c] = b ? b[c] - s : s;
}

// partial[blockIdx] = sum over this block's strided share of x_i * (conj ? conj(y_i) : y_i)
template <class T>
__global__ void __launch_bounds__(256)
k_raff_dot(int64_t n, const T *__restrict__ x, const T *__restrict__ y, int conjy, T *partial) {
  __shared__ T sh[256];
  T s = ST<T>::zero();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T v = y[i];
    if (conjy) v = ST<T>::conj(v);
    s = s + x[i] * v;
  }
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

template <class T>
__global__ void k_raff_axpy(int64_t n, T alpha, const T *__restrict__ x, T *y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = y[i] + alpha * x[i];
}
template <class T>
__global__ void k_raff_scal(int64_t n, T alpha, T *x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = alpha * x[i];
}

}  // namespace pb200
