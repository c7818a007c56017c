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
__device__ __forceinline__ void ld_parts(const float *p, float &re, float &im) { re = *p; im = 0.f; }
__device__ __forceinline__ void ld_parts(const cfloat *p, float &re, float &im) {
  const float2 v = *reinterpret_cast<const float2 *>(p);
  re = v.x; im = v.y;
}

// ---- operand registers and the m16n8k8 product in the two real arithmetics.
// double: mma.sync.m16n8k8.f64 (DMMA).  float: the SAME fragment layout exists for tf32 operands
// (mma.sync.m16n8k8.f32.tf32.tf32.f32), and FP32 accuracy is recovered by the 3xTF32 split: x = hi + lo with
// hi = tf32(x), lo = tf32(x - hi); C += A_lo B_hi + A_hi B_lo + A_hi B_hi (the lo*lo term is below the FP32 ulp).
__device__ __forceinline__ unsigned f2tf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void smma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <class R> struct Mma;
template <> struct Mma<double> {
  struct A { double v[4]; };
  struct B { double v[2]; };
  static __device__ __forceinline__ void set(A &f, int q, double x) { f.v[q] = x; }
  static __device__ __forceinline__ void set(B &f, int q, double x) { f.v[q] = x; }
  static __device__ __forceinline__ void mma(double (&c)[4], const A &a, const B &b) { dmma(c, a.v, b.v); }
};
template <> struct Mma<float> {
  struct A { unsigned hi[4], lo[4]; };
  struct B { unsigned hi[2], lo[2]; };
  static __device__ __forceinline__ void set(A &f, int q, float x) {
    const unsigned h = f2tf32(x); f.hi[q] = h; f.lo[q] = f2tf32(x - __uint_as_float(h));
  }
  static __device__ __forceinline__ void set(B &f, int q, float x) {
    const unsigned h = f2tf32(x); f.hi[q] = h; f.lo[q] = f2tf32(x - __uint_as_float(h));
  }
  static __device__ __forceinline__ void mma(float (&c)[4], const A &a, const B &b) {
    smma(c, a.lo, b.hi); smma(c, a.hi, b.lo); smma(c, a.hi, b.hi);   // small terms first
  }
};

// accumulator tile of one m16n8 MMA in T's arithmetic (R = the real type of T)
template <bool CX, class R> struct Acc;
template <class R> struct Acc<false, R> {
  R re[4];
  __device__ __forceinline__ void zero() { re[0] = re[1] = re[2] = re[3] = R(0); }
};
template <class R> struct Acc<true, R> {
  R re[4], im[4];
  __device__ __forceinline__ void zero() {
    re[0] = re[1] = re[2] = re[3] = R(0); im[0] = im[1] = im[2] = im[3] = R(0);
  }
};
template <bool CX, class R> struct FragA;
template <class R> struct FragA<false, R> { typename Mma<R>::A re; };
template <class R> struct FragA<true, R> { typename Mma<R>::A re, im; };
template <bool CX, class R> struct FragB;
template <class R> struct FragB<false, R> { typename Mma<R>::B re; };
template <class R> struct FragB<true, R> { typename Mma<R>::B re, im, nim; };

template <class R>
__device__ __forceinline__ void mma_acc(Acc<false, R> &c, const FragA<false, R> &a, const FragB<false, R> &b) {
  Mma<R>::mma(c.re, a.re, b.re);
}
template <class R>
__device__ __forceinline__ void mma_acc(Acc<true, R> &c, const FragA<true, R> &a, const FragB<true, R> &b) {
  Mma<R>::mma(c.re, a.re, b.re);
  Mma<R>::mma(c.re, a.im, b.nim);
  Mma<R>::mma(c.im, a.re, b.im);
  Mma<R>::mma(c.im, a.im, b.re);
}

// A fragment from a k-major tile: element (row i, k) at s[k * ld + i]
template <class T>
__device__ __forceinline__ void load_frag_a(FragA<ST<T>::is_complex, typename ST<T>::real> &f, const T *s, int ld, int row0, int k0, int lane) {
  using R = typename ST<T>::real;
  const int g = lane >> 2, t = lane & 3;
  R re, im;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = row0 + g + ((q & 1) ? 8 : 0), k = k0 + t + ((q & 2) ? 4 : 0);
    ld_parts(s + (size_t)k * ld + i, re, im);
    Mma<R>::set(f.re, q, re);
    if constexpr (ST<T>::is_complex) Mma<R>::set(f.im, q, im);
  }
}
// B fragment from a k-major tile: element (k, n) at s[k * ld + n]; `scale` (optional, per k) and conj applied here
template <class T, bool CONJ, bool SCALE>
__device__ __forceinline__ void load_frag_b(FragB<ST<T>::is_complex, typename ST<T>::real> &f, const T *s, int ld, int n0, int k0, int lane,
                                            const T *dk) {
  using R = typename ST<T>::real;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int k = k0 + t + q * 4;
    T v = s[(size_t)k * ld + n0 + g];
    if (SCALE) v = v * dk[k];
    if constexpr (ST<T>::is_complex) {
      const R im = CONJ ? -v.y : v.y;
      Mma<R>::set(f.re, q, v.x); Mma<R>::set(f.im, q, im); Mma<R>::set(f.nim, q, -im);
    } else {
      Mma<R>::set(f.re, q, v);
    }
  }
}

}  // namespace pb200
