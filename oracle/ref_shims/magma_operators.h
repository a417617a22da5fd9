/* oracle/ref_shims/magma_operators.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE, and NOT a copy of MAGMA.
 *
 * The reference's GPU propagator (Taylor_gpu.cpp, dzgemv_kernels.cu) includes MAGMA's "magma_operators.h" only for
 * the arithmetic operators on cuDoubleComplex and the MAGMA_Z_* constants.  MAGMA is not in this image; this header
 * states those few definitions from their mathematical meaning so that the reference sources compile IN PLACE
 * (oracle/Makefile, target _ref/libref_taylor_gpu.so).  Nothing under dynemol_b200/ includes it. */
#pragma once
#include <cuComplex.h>

#define MAGMA_Z_MAKE(r, i) make_cuDoubleComplex((r), (i))
#define MAGMA_Z_ZERO       make_cuDoubleComplex(0.0, 0.0)
#define MAGMA_Z_ONE        make_cuDoubleComplex(1.0, 0.0)
#define MAGMA_Z_NEG_ONE    make_cuDoubleComplex(-1.0, 0.0)
#define MAGMA_Z_REAL(a)    ((a).x)
#define MAGMA_Z_IMAG(a)    ((a).y)

#ifdef __cplusplus
#define DYB_HD __host__ __device__ static inline
DYB_HD double real(const cuDoubleComplex a) { return a.x; }
DYB_HD double imag(const cuDoubleComplex a) { return a.y; }
DYB_HD cuDoubleComplex conj(const cuDoubleComplex a) { return make_cuDoubleComplex(a.x, -a.y); }
DYB_HD cuDoubleComplex operator-(const cuDoubleComplex a) { return make_cuDoubleComplex(-a.x, -a.y); }
DYB_HD cuDoubleComplex operator+(const cuDoubleComplex a, const cuDoubleComplex b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
DYB_HD cuDoubleComplex operator-(const cuDoubleComplex a, const cuDoubleComplex b) { return make_cuDoubleComplex(a.x - b.x, a.y - b.y); }
DYB_HD cuDoubleComplex operator*(const cuDoubleComplex a, const cuDoubleComplex b) { return make_cuDoubleComplex(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DYB_HD cuDoubleComplex operator*(const cuDoubleComplex a, const double s) { return make_cuDoubleComplex(a.x * s, a.y * s); }
DYB_HD cuDoubleComplex operator*(const double s, const cuDoubleComplex a) { return make_cuDoubleComplex(a.x * s, a.y * s); }
DYB_HD cuDoubleComplex operator/(const cuDoubleComplex a, const double s) { return make_cuDoubleComplex(a.x / s, a.y / s); }
DYB_HD cuDoubleComplex operator/(const cuDoubleComplex a, const cuDoubleComplex b) { return cuCdiv(a, b); }
DYB_HD cuDoubleComplex& operator+=(cuDoubleComplex& a, const cuDoubleComplex b) { a.x += b.x; a.y += b.y; return a; }
DYB_HD cuDoubleComplex& operator-=(cuDoubleComplex& a, const cuDoubleComplex b) { a.x -= b.x; a.y -= b.y; return a; }
DYB_HD cuDoubleComplex& operator*=(cuDoubleComplex& a, const cuDoubleComplex b) { a = a * b; return a; }
DYB_HD cuDoubleComplex& operator*=(cuDoubleComplex& a, const double s) { a.x *= s; a.y *= s; return a; }
DYB_HD bool operator==(const cuDoubleComplex a, const cuDoubleComplex b) { return a.x == b.x && a.y == b.y; }
DYB_HD bool operator!=(const cuDoubleComplex a, const cuDoubleComplex b) { return !(a == b); }
#undef DYB_HD
#endif
