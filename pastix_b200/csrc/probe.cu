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
