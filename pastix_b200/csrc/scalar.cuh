// scalar.cuh — arithmetic traits for the four PASTIX_FLOAT types
// (reference: common/src/common_pastix.h:279-315 — float, double, float complex,
// double complex).  Complex numbers are plain (re, im) pairs so that the device
// code controls exactly which real sub-products are formed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb200 {

template <class R>
struct alignas(2 * sizeof(R)) cx {
  R x, y;
  __host__ __device__ cx() {}
  __host__ __device__ cx(R a) : x(a), y(R(0)) {}
  __host__ __device__ cx(R a, R b) : x(a), y(b) {}
};
template <class R> __host__ __device__ inline cx<R> operator+(cx<R> a, cx<R> b) { return cx<R>(a.x + b.x, a.y + b.y); }
template <class R> __host__ __device__ inline cx<R> operator-(cx<R> a, cx<R> b) { return cx<R>(a.x - b.x, a.y - b.y); }
template <class R> __host__ __device__ inline cx<R> operator-(cx<R> a) { return cx<R>(-a.x, -a.y); }
template <class R> __host__ __device__ inline cx<R> operator*(cx<R> a, cx<R> b) {
  return cx<R>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <class R> __host__ __device__ inline cx<R> operator/(cx<R> a, cx<R> b) {
  // Smith's algorithm (what C99 cdiv does modulo scaling)
  R ar = b.x < 0 ? -b.x : b.x, ai = b.y < 0 ? -b.y : b.y;
  if (ar >= ai) {
    R t = b.y / b.x, d = b.x + b.y * t;
    return cx<R>((a.x + a.y * t) / d, (a.y - a.x * t) / d);
  } else {
    R t = b.x / b.y, d = b.x * t + b.y;
    return cx<R>((a.x * t + a.y) / d, (a.y * t - a.x) / d);
  }
}
template <class R> __host__ __device__ inline cx<R> &operator+=(cx<R> &a, cx<R> b) { a.x += b.x; a.y += b.y; return a; }
template <class R> __host__ __device__ inline cx<R> &operator-=(cx<R> &a, cx<R> b) { a.x -= b.x; a.y -= b.y; return a; }
template <class R> __host__ __device__ inline cx<R> &operator*=(cx<R> &a, cx<R> b) { a = a * b; return a; }

typedef cx<float> cfloat;
typedef cx<double> cdouble;

template <class T> struct ST;
template <> struct ST<float> {
  typedef float real; static const bool is_complex = false;
  static __host__ __device__ inline float conj(float a) { return a; }
  static __host__ __device__ inline float abs(float a) { return fabsf(a); }
  static __host__ __device__ inline float sqrt(float a) { return sqrtf(a); }
  static __host__ __device__ inline float re(float a) { return a; }
  static __host__ __device__ inline float from_real(double r) { return (float)r; }
  static __host__ __device__ inline float zero() { return 0.f; }
};
template <> struct ST<double> {
  typedef double real; static const bool is_complex = false;
  static __host__ __device__ inline double conj(double a) { return a; }
  static __host__ __device__ inline double abs(double a) { return fabs(a); }
  static __host__ __device__ inline double sqrt(double a) { return ::sqrt(a); }
  static __host__ __device__ inline double re(double a) { return a; }
  static __host__ __device__ inline double from_real(double r) { return r; }
  static __host__ __device__ inline double zero() { return 0.0; }
};
template <class R> struct STC {
  typedef R real; static const bool is_complex = true;
  static __host__ __device__ inline cx<R> conj(cx<R> a) { return cx<R>(a.x, -a.y); }
  static __host__ __device__ inline R abs(cx<R> a) { return (R)hypot((double)a.x, (double)a.y); }
  static __host__ __device__ inline cx<R> sqrt(cx<R> a) {
    // principal square root, as csqrt()
    double m = hypot((double)a.x, (double)a.y);
    if (m == 0.0) return cx<R>(R(0), R(0));
    double sr = ::sqrt(0.5 * (m + fabs((double)a.x)));
    double si = 0.5 * (double)a.y / sr;
    if (a.x >= 0) return cx<R>((R)sr, (R)si);
    return cx<R>((R)fabs(si), (R)(a.y >= 0 ? sr : -sr));
  }
  static __host__ __device__ inline R re(cx<R> a) { return a.x; }
  static __host__ __device__ inline cx<R> from_real(double r) { return cx<R>((R)r, R(0)); }
  static __host__ __device__ inline cx<R> zero() { return cx<R>(R(0), R(0)); }
};
template <> struct ST<cfloat> : STC<float> {};
template <> struct ST<cdouble> : STC<double> {};

// fused a += b*c (real: one FMA; complex: four real FMAs)
template <class T> __device__ inline void fma_acc(T &a, T b, T c) { a += b * c; }
template <> __device__ inline void fma_acc<double>(double &a, double b, double c) { a = fma(b, c, a); }
template <> __device__ inline void fma_acc<float>(float &a, float b, float c) { a = fmaf(b, c, a); }
template <> __device__ inline void fma_acc<cdouble>(cdouble &a, cdouble b, cdouble c) {
  a.x = fma(b.x, c.x, a.x); a.x = fma(-b.y, c.y, a.x);
  a.y = fma(b.x, c.y, a.y); a.y = fma(b.y, c.x, a.y);
}
template <> __device__ inline void fma_acc<cfloat>(cfloat &a, cfloat b, cfloat c) {
  a.x = fmaf(b.x, c.x, a.x); a.x = fmaf(-b.y, c.y, a.x);
  a.y = fmaf(b.x, c.y, a.y); a.y = fmaf(b.y, c.x, a.y);
}

// atomic "target -= v" on global memory (the reference serialises these with
// mutex_blok, sopalin_compute.c:563-580; here they are L2 reductions)
__device__ inline void atomic_sub(float *p, float v) { atomicAdd(p, -v); }
__device__ inline void atomic_sub(double *p, double v) { atomicAdd(p, -v); }
__device__ inline void atomic_sub(cfloat *p, cfloat v) { atomicAdd(&p->x, -v.x); atomicAdd(&p->y, -v.y); }
__device__ inline void atomic_sub(cdouble *p, cdouble v) { atomicAdd(&p->x, -v.x); atomicAdd(&p->y, -v.y); }

}  // namespace pb200
