// probe.cu — FP64 pipe micro-benchmarks used by bench.py to establish the
// measured FP64 roof the factorization is compared against (MEASURED_PEAKS.json
// carries no FP64 figure).  variant 0: DFMA (vector pipe), 1: DMMA m8n8k4,
// 2: DMMA m16n8k8, 3: DMMA m16n8k16 (sm_90+ shapes).
#include <cuda_runtime.h>
#include "../../include/pastix_b200.h"

namespace {

template <int VARIANT>
__global__ void __launch_bounds__(1024) k_probe(double *out, int iters) {
  double c[8][4];
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; for (int j = 0; j < 4; ++j) c[i][j] = 0.0; }
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 1e-9 * (j + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      if (VARIANT == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) c[ch][j] = fma(a[ch], b[j], c[ch][j]);
      } else if (VARIANT == 1) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[ch][0]), "+d"(c[ch][1]) : "d"(a[ch]), "d"(b[0]));
      } else if (VARIANT == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[ch][0]), "+d"(c[ch][1]), "+d"(c[ch][2]), "+d"(c[ch][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[ch][0]), "+d"(c[ch][1]), "+d"(c[ch][2]), "+d"(c[ch][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456) out[0] = s;
}

// variant 200+: the main loop of k_gemm_scatter in isolation — fragments re-read from shared memory every k-step
// (same k-major layout and padding), 32x32 warp tiles, no global traffic, optional CTA barrier every 2 k-steps.
template <int BARRIER>
__global__ void __launch_bounds__(128) k_probe_loop(double *out, int iters) {
  __shared__ double sA[16 * 68], sB[16 * 68];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 16 * 68; e += 128) { sA[e] = 1.0 + e * 1e-9; sB[e] = 1e-9 * e; }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 1) * 32, wn0 = (warp & 1) * 32;
  double c[2][4][4];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[x][y][q] = 0.0;
  for (int it = 0; it < iters; ++it) {
    if (BARRIER) __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 16; ks += 8) {
      double a[2][4], b[4][2];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int q = 0; q < 4; ++q) a[x][q] = sA[(ks + t + ((q & 2) ? 4 : 0)) * 68 + wm0 + x * 16 + g + ((q & 1) ? 8 : 0)];
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int q = 0; q < 2; ++q) b[y][q] = sB[(ks + t + q * 4) * 68 + wn0 + y * 8 + g];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+d"(c[x][y][0]), "+d"(c[x][y][1]), "+d"(c[x][y][2]), "+d"(c[x][y][3])
                       : "d"(a[x][0]), "d"(a[x][1]), "d"(a[x][2]), "d"(a[x][3]), "d"(b[y][0]), "d"(b[y][1]));
    }
  }
  double s = 0;
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int q = 0; q < 4; ++q) s += c[x][y][q];
  if (s == 123.456) out[0] = s;
}

template <int BARRIER>
double run_loop(int sms, int ctas_per_sm) {
  double *d; cudaMalloc(&d, 8);
  const int iters = 4000, blocks = sms * ctas_per_sm;
  k_probe_loop<BARRIER><<<blocks, 128>>>(d, 10);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_probe_loop<BARRIER><<<blocks, 128>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  const double flops = 2.0 * 64 * 64 * 16 * (double)blocks * iters;
  return flops / (ms * 1e-3) / 1e9;
}

template <int V>
double run(int sms, int bps = 8, int threads = 256) {
  double *d; cudaMalloc(&d, 8);
  const int iters = 20000, blocks = sms * bps;
  k_probe<V><<<blocks, threads>>>(d, 100);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_probe<V><<<blocks, threads>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  // flops per thread-iteration (8 chains)
  double per_warp_iter;
  if (V == 0) per_warp_iter = 8.0 * 4 * 2 * 32;
  else if (V == 1) per_warp_iter = 8.0 * 8 * 8 * 4 * 2;
  else if (V == 2) per_warp_iter = 8.0 * 16 * 8 * 8 * 2;
  else per_warp_iter = 8.0 * 16 * 8 * 16 * 2;
  double flops = per_warp_iter * (threads / 32) * (double)blocks * iters;
  return flops / (ms * 1e-3) / 1e9;
}
}  // namespace

extern "C" double pb200_probe_fp64_gflops(int device, int variant) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1.0;
  if (device >= 0) cudaSetDevice(device); else cudaGetDevice(&device);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, device);
  if (variant >= 300) return run_loop<1>(p.multiProcessorCount, variant - 300);   // with a CTA barrier per chunk
  if (variant >= 200) return run_loop<0>(p.multiProcessorCount, variant - 200);   // (variant-200) CTAs of 4 warps per SM
  if (variant >= 100) {   // DMMA m16n8k8 with (variant-100) warps per SM, 8 independent accumulators per warp
    int warps = variant - 100;
    return run<2>(p.multiProcessorCount, 1, warps * 32);
  }
  switch (variant) {
    case 0: return run<0>(p.multiProcessorCount);
    case 1: return run<1>(p.multiProcessorCount);
    case 2: return run<2>(p.multiProcessorCount);
    case 3: return run<3>(p.multiProcessorCount);
  }
  return -1.0;
}
