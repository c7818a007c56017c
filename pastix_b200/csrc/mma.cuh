// mma.cuh — FP64 tensor-core building blocks for sm_100a.
//
// tcgen05.mma has no f64 kind, so double precision runs on the DMMA path
// (mma.sync.m16n8k8.f64, SASS DMMA) fed from shared-memory tiles staged with
// cp.async (LDGSTS).  Complex arithmetic is carried as four real sub-products
// per tile (re*re - im*im, re*im + im*re) on the same instruction.
//
// Fragment layout of mma.m16n8k8 .f64 (g = lane>>2, t = lane&3):
//   A (16x8, row):  a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)
//   B (8x8,  col):  b0=(k=t,n=g) b1=(k=t+4,n=g)
//   C (16x8):       c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)
#pragma once
#include "scalar.cuh"

namespace pb200 {

__device__ __forceinline__ void dmma(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// cp.async of one element (8 or 16 bytes); `valid == false` zero-fills the destination
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(gmem_src), "n"(BYTES), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on a shared-memory mbarrier
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// real / imaginary parts of a stored element
__device__ __forceinline__ void ld_parts(const double *p, double &re, double &im) { re = *p; im = 0.0; }
__device__ __forceinline__ void ld_parts(const cdouble *p, double &re, double &im) {
  const double2 v = *reinterpret_cast<const double2 *>(p);
  re = v.x; im = v.y;
}
__device__ __forceinline__ double mk(double re, double, double *) { return re; }
__device__ __forceinline__ cdouble mk(double re, double im, cdouble *) { return cdouble(re, im); }

// accumulator tile of one m16n8 MMA in T's arithmetic
template <bool CX> struct Acc;
template <> struct Acc<false> {
  double re[4];
  __device__ __forceinline__ void zero() { re[0] = re[1] = re[2] = re[3] = 0.0; }
};
template <> struct Acc<true> {
  double re[4], im[4];
  __device__ __forceinline__ void zero() {
    re[0] = re[1] = re[2] = re[3] = 0.0; im[0] = im[1] = im[2] = im[3] = 0.0;
  }
};
template <bool CX> struct FragA;
template <> struct FragA<false> { double re[4]; };
template <> struct FragA<true> { double re[4], im[4]; };
template <bool CX> struct FragB;
template <> struct FragB<false> { double re[2]; };
template <> struct FragB<true> { double re[2], im[2], nim[2]; };

__device__ __forceinline__ void mma_acc(Acc<false> &c, const FragA<false> &a, const FragB<false> &b) {
  dmma(c.re, a.re, b.re);
}
__device__ __forceinline__ void mma_acc(Acc<true> &c, const FragA<true> &a, const FragB<true> &b) {
  dmma(c.re, a.re, b.re);
  dmma(c.re, a.im, b.nim);
  dmma(c.im, a.re, b.im);
  dmma(c.im, a.im, b.re);
}

// A fragment from a k-major tile: element (row i, k) at s[k * ld + i]
template <class T>
__device__ __forceinline__ void load_frag_a(FragA<ST<T>::is_complex> &f, const T *s, int ld, int row0, int k0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  double re, im;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = row0 + g + ((q & 1) ? 8 : 0), k = k0 + t + ((q & 2) ? 4 : 0);
    ld_parts(s + (size_t)k * ld + i, re, im);
    f.re[q] = re;
    if constexpr (ST<T>::is_complex) f.im[q] = im;
  }
}
// B fragment from a k-major tile: element (k, n) at s[k * ld + n]; `scale` (optional, per k) and conj applied here
template <class T, bool CONJ, bool SCALE>
__device__ __forceinline__ void load_frag_b(FragB<ST<T>::is_complex> &f, const T *s, int ld, int n0, int k0, int lane,
                                            const T *dk) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int k = k0 + t + q * 4;
    T v = s[(size_t)k * ld + n0 + g];
    if (SCALE) v = v * dk[k];
    if constexpr (ST<T>::is_complex) {
      f.re[q] = v.x; f.im[q] = CONJ ? -v.y : v.y; f.nim[q] = -f.im[q];
    } else {
      f.re[q] = v;
    }
  }
}

}  // namespace pb200
