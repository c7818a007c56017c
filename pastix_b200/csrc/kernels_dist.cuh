// kernels_dist.cuh — multi-GPU fan-in over NVLink peer memory (one process per GPU, slabs mapped into
// every peer with CUDA IPC).
//
// Reference semantics (src/sopalin/src): a contribution to a column block owned by another processor is
// summed into a local fan-in buffer (add_contrib_target, sopalin_compute.c:600-733), shipped once all
// local contributions are in (send_one_fanin, sopalin_sendrecv.c:1219) and added into the owner's panel
// (recv_handle_fanin, sopalin_sendrecv.c:182-300: coeftab[coefind + (fcol-fcolnum)*stride + frow-frownum] += buf).
// Here every GPU keeps the full slab: the region of a cblk it does not own IS its fan-in buffer (zero
// at assembly, filled by the same fused GEMM+scatter kernel, same addresses).  The owner PULLS the
// buffers of its contributors straight out of their HBM with coalesced peer loads and adds them to its
// own panel in one kernel — no staging copy, no message packing.  Ordering between GPUs is carried by
// per-level flags in peer memory (release/acquire at system scope), not by the host.
#pragma once
#include "scalar.cuh"
#include "symbol.cuh"

namespace pb200 {

#define PB200_MAXRANKS 8
struct Peers {
  void *L[PB200_MAXRANKS];
  void *U[PB200_MAXRANKS];
  void *W[PB200_MAXRANKS];      // LDLt / LDLh on the tensor path: the L*D copies beside the panels
  unsigned int *flags[PB200_MAXRANKS];
  int rank, nranks;
};

struct FanTask {      // one owned cblk with remote contributors
  int cblk;
  int tile0;          // first tile of this cblk inside the launch
  unsigned int mask;  // contributing ranks
  int pad;
};
#define PB200_FAN_ELEMS 2048   // elements per CTA (256 threads x 8)

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// "everything this GPU launched before on this stream is done": publish flag[idx] = epoch
__global__ void k_dist_signal(unsigned int *flags, int idx, unsigned int epoch) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(flags + idx, epoch);
  }
}

// wait until every rank in `mask` has published flag[idx] >= epoch (thread p polls rank p over NVLink)
__global__ void k_dist_wait(Peers P, unsigned int mask, int idx, unsigned int epoch, unsigned long long timeout_ns,
                            unsigned int *err) {
  const int p = threadIdx.x;
  if (p >= P.nranks || !((mask >> p) & 1u) || p == P.rank) return;
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(P.flags[p] + idx) - epoch) < 0) {
    __nanosleep(200);
    if (global_ns() - t0 > timeout_ns) { atomicExch(err, 1u + (unsigned)p); return; }
  }
}

// owner's panel += sum over the contributing ranks selected by `filter` of their fan-in buffer (same slab offsets
// everywhere).  The sum is ADDED with an L2 reduction: the panel may be receiving local contributions (and the other
// half of the gather) at the same time.
template <class T>
__global__ void __launch_bounds__(256)
k_fanin_gather(DevSym S, Peers P, T *L, T *U, const FanTask *__restrict__ tasks, int ntasks, unsigned int filter) {
  const int t = find_task(tasks, ntasks, (int)blockIdx.x);
  const FanTask tk = tasks[t];
  const unsigned int mask = tk.mask & filter;
  if (mask == 0u) return;
  const int c = tk.cblk;
  const int64_t base = S.poff[c], len = S.poff[c + 1] - S.poff[c];
  const int64_t e0 = (int64_t)(blockIdx.x - tk.tile0) * PB200_FAN_ELEMS;
#pragma unroll
  for (int q = 0; q < PB200_FAN_ELEMS / 256; ++q) {
    const int64_t e = e0 + q * 256 + threadIdx.x;
    if (e >= len) break;
    T accL = ST<T>::zero(), accU = ST<T>::zero();
    for (int p = 0; p < P.nranks; ++p) {
      if (!((mask >> p) & 1u)) continue;
      accL += reinterpret_cast<const T *>(P.L[p])[base + e];
      if (U != nullptr) accU += reinterpret_cast<const T *>(P.U[p])[base + e];
    }
    atomic_sub(&L[base + e], ST<T>::zero() - accL);
    if (U != nullptr) atomic_sub(&U[base + e], ST<T>::zero() - accU);
  }
}

// after the factorization: copy the factored panels of the other GPUs into the local slab so that
// every GPU can run up_down on its share of the right-hand sides without further exchanges
// (also the fan-out of a freshly factored shared panel to the GPUs that own its targets: then with the L*D copy W)
template <class T>
__global__ void __launch_bounds__(256)
k_pull_panels(DevSym S, Peers P, T *L, T *U, T *W, const int *__restrict__ owner, const FanTask *__restrict__ tasks, int ntasks) {
  const int t = find_task(tasks, ntasks, (int)blockIdx.x);
  const FanTask tk = tasks[t];
  const int c = tk.cblk, o = owner[c];
  const int64_t base = S.poff[c], len = S.poff[c + 1] - S.poff[c];
  const int64_t e0 = (int64_t)(blockIdx.x - tk.tile0) * PB200_FAN_ELEMS;
  const T *pl = reinterpret_cast<const T *>(P.L[o]);
  const T *pu = reinterpret_cast<const T *>(P.U[o]);
  const int ld = S.stride[c];
#pragma unroll
  for (int q = 0; q < PB200_FAN_ELEMS / 256; ++q) {
    const int64_t e = e0 + q * 256 + threadIdx.x;
    if (e >= len) break;
    const T l = pl[base + e];
    L[base + e] = l;
    if (U != nullptr) U[base + e] = pu[base + e];
    // LDLt / LDLh: the L*D copy is rebuilt here from the pulled column and its pivot (one more peer load per element,
    // broadcast over the column) instead of crossing NVLink a second time
    if (W != nullptr) { const int64_t j = e / ld; W[base + e] = l * pl[base + j * (int64_t)(ld + 1)]; }
  }
}

}  // namespace pb200
