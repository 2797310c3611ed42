// Scalar / 3-vector / rotation helpers shared by the simulator kernels.  Plain C++ under a __host__ __device__ macro so
// the test-only host build can instantiate the same code in fp64.
#ifndef JB_MATH_CUH
#define JB_MATH_CUH

#include <math.h>

#if defined(__CUDACC__)
#define JB_HD __host__ __device__ __forceinline__
#else
#define JB_HD inline
#endif

// ------------------------------------------------------------------------------------------------ scalar helpers
JB_HD float jb_sqrt(float x) { return sqrtf(x); }
JB_HD double jb_sqrt(double x) { return sqrt(x); }
JB_HD float jb_abs(float x) { return fabsf(x); }
JB_HD double jb_abs(double x) { return fabs(x); }
JB_HD void jb_sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
JB_HD void jb_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
// sin / cos for small arguments (leg hinge angles, the per-step quaternion increment): Taylor series to x^11 / x^10 are
// exact to fp32 rounding for |x| <= 0.5; larger arguments fall back to the library call.
JB_HD void jb_sincos_small(float x, float* s, float* c) {
  if (fabsf(x) <= 0.5f) {
    const float x2 = x * x;
    *s = x * (1.f + x2 * (-1.f / 6 + x2 * (1.f / 120 + x2 * (-1.f / 5040 + x2 * (1.f / 362880 + x2 * (-1.f / 39916800))))));
    *c = 1.f + x2 * (-0.5f + x2 * (1.f / 24 + x2 * (-1.f / 720 + x2 * (1.f / 40320 + x2 * (-1.f / 3628800)))));
  } else {
    jb_sincos(x, s, c);
  }
}
JB_HD void jb_sincos_small(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
// sin / cos for the angles of the sub-step (hinge angles, the wrapped motor angle, the half rotation increment): quadrant
// reduction x = q pi/2 + r with a two-constant Cody-Waite split (exact for |q| < 2^10), then the classic single-precision
// kernels on |r| <= pi/4 (errors ~1 ulp).  A few dozen branch-free instructions; the library call and its Payne-Hanek
// slow path (bulky in the instruction stream of a kernel that is instruction-fetch bound) stay only for |x| > 100.
// (the out-of-line fallback returns by value: handing it the caller's output pointers would pin those variables to
// local memory on the fast path as well)
struct JbSinCos { float s, c; };
#if defined(__CUDACC__)
static __host__ __device__ __noinline__ JbSinCos jb_sincos_far(float x) { JbSinCos r; jb_sincos(x, &r.s, &r.c); return r; }
#else
static inline JbSinCos jb_sincos_far(float x) { JbSinCos r; jb_sincos(x, &r.s, &r.c); return r; }
#endif
JB_HD void jb_sincos_fast(float x, float* s, float* c) {
  if (fabsf(x) > 100.f) { const JbSinCos r = jb_sincos_far(x); *s = r.s; *c = r.c; return; }
  const float q = rintf(x * 0.63661977236758134308f);
  float r = fmaf(q, -1.57079625129699707031f, x);
  r = fmaf(q, -7.54978941586159635335e-08f, r);
  const float r2 = r * r;
  const float sp = fmaf(r * r2, fmaf(r2, fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f), -1.6666654611e-1f), r);
  const float cp = fmaf(r2 * r2, fmaf(r2, fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f), 4.166664568298827e-2f),
                        fmaf(r2, -0.5f, 1.0f));
  const int n = (int)q;
  const float ss = (n & 1) ? cp : sp, cc = (n & 1) ? sp : cp;
  *s = (n & 2) ? -ss : ss;
  *c = ((n + 1) & 2) ? -cc : cc;
}
JB_HD void jb_sincos_fast(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
template <typename T> JB_HD T jb_max(T a, T b) { return a > b ? a : b; }
template <typename T> JB_HD T jb_min(T a, T b) { return a < b ? a : b; }
template <typename T> struct JbEps;
template <> struct JbEps<float> {
  static JB_HD float newton_tol() { return 2e-6f; }
#ifndef JBQ_LS_TOL
#define JBQ_LS_TOL 1e-5f
#endif
  static JB_HD float ls_tol() { return JBQ_LS_TOL; }
  static JB_HD int newton_iters() { return 16; }
};
template <> struct JbEps<double> {
  static JB_HD double newton_tol() { return 1e-8; }
  static JB_HD double ls_tol() { return 1e-12; }
  static JB_HD int newton_iters() { return 100; }
};

// ------------------------------------------------------------------------------------------------ 3-vector helpers
template <typename T> JB_HD T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> JB_HD void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> JB_HD void sub3(T* r, const T* a, const T* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <typename T> JB_HD void add3(T* r, const T* a, const T* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
template <typename T> JB_HD void addto3(T* r, const T* a) { r[0] += a[0]; r[1] += a[1]; r[2] += a[2]; }
template <typename T> JB_HD void addscl3(T* r, const T* a, T s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
template <typename T> JB_HD void scl3(T* r, const T* a, T s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
template <typename T, typename U> JB_HD void ld3(T* r, const U* a) { r[0] = (T)a[0]; r[1] = (T)a[1]; r[2] = (T)a[2]; }
// r = M v (row-major 3x3), r = M^T v
template <typename T> JB_HD void mulMV(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
    z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> JB_HD void mulMTV(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
    z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// symmetric 3x3 (xx,yy,zz,xy,xz,yz) times vector
template <typename T> JB_HD void mulSV(T* r, const T* s, const T* v) {
  T x = s[0] * v[0] + s[3] * v[1] + s[4] * v[2], y = s[3] * v[0] + s[1] * v[1] + s[5] * v[2],
    z = s[4] * v[0] + s[5] * v[1] + s[2] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// Rodrigues: rotation by angle (s = sin, c = cos) about unit axis a
template <typename T> JB_HD void rodrigues(T* R, const T* a, T s, T c) {
  T t = (T)1 - c;
  R[0] = c + t * a[0] * a[0];        R[1] = t * a[0] * a[1] - s * a[2]; R[2] = t * a[0] * a[2] + s * a[1];
  R[3] = t * a[0] * a[1] + s * a[2]; R[4] = c + t * a[1] * a[1];        R[5] = t * a[1] * a[2] - s * a[0];
  R[6] = t * a[0] * a[2] - s * a[1]; R[7] = t * a[1] * a[2] + s * a[0]; R[8] = c + t * a[2] * a[2];
}
template <typename T> JB_HD void mulMM(T* r, const T* a, const T* b) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
// I' = R I R^T for symmetric I (6) -> symmetric (6)
template <typename T, typename U> JB_HD void rotSym(T* o, const T* R, const U* Iin) {
  T I[6] = {(T)Iin[0], (T)Iin[1], (T)Iin[2], (T)Iin[3], (T)Iin[4], (T)Iin[5]};
  T t[9];   // t = R * I
#pragma unroll
  for (int i = 0; i < 3; i++) {
    t[3 * i + 0] = R[3 * i] * I[0] + R[3 * i + 1] * I[3] + R[3 * i + 2] * I[4];
    t[3 * i + 1] = R[3 * i] * I[3] + R[3 * i + 1] * I[1] + R[3 * i + 2] * I[5];
    t[3 * i + 2] = R[3 * i] * I[4] + R[3 * i + 1] * I[5] + R[3 * i + 2] * I[2];
  }
  o[0] = t[0] * R[0] + t[1] * R[1] + t[2] * R[2];
  o[1] = t[3] * R[3] + t[4] * R[4] + t[5] * R[5];
  o[2] = t[6] * R[6] + t[7] * R[7] + t[8] * R[8];
  o[3] = t[0] * R[3] + t[1] * R[4] + t[2] * R[5];
  o[4] = t[0] * R[6] + t[1] * R[7] + t[2] * R[8];
  o[5] = t[3] * R[6] + t[4] * R[7] + t[5] * R[8];
}
template <typename T> JB_HD void quat2mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

#endif  // JB_MATH_CUH
