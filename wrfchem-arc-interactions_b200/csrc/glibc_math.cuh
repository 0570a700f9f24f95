// logf / powf / expf that return, bit for bit, what glibc 2.39 (x86-64) returns.
//
// Why: the reference is Fortran built with gfortran, whose LOG / EXP / ** on REAL(4) are calls into glibc's libm.
// setcoef turns log(p) into the table index jp and the interpolation weight fp (module_ra_rrtmg_sw.F:2854-2887,
// module_ra_rrtmg_lw.F:3649-3685), the adapter turns alog / ** into the 14-band aerosol optical depths
// (SW:10996-11021), and reftra_sw amplifies a one-ulp change of any of them into O(1) changes of a layer reflectance
// near its removable singularity k*mu0 = 1 (SW:2629-2660).  CUDA's logf / powf are 1-2 ulp functions, so the device
// path carries glibc's own algorithms instead: table + polynomial in double precision (the "optimized routines"
// logf.c / powf.c / expf.c of glibc's sysdeps/ieee754/flt-32, tables __logf_data, __powf_log2_data, __exp2f_data),
// with the fused multiply-adds of the FMA build that x86-64 CPUs with AVX2 select at run time.  Every operation is an
// IEEE double operation, identical on the GPU and the CPU; tests/test_libm_cpu.py compares the host instantiation with
// the C library over every float in [1e-3, 1200] (the pressures in hPa) and 2e8 powf / expf arguments,
// tests/test_gpu_parity.py compares the device instantiation with the C library.
//
// Arguments outside the fast path of the glibc routines (x <= 0, subnormal, inf, nan, overflow / underflow of the
// result) never occur on the radiation path; they fall back to the CUDA / host library function.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace arc {
namespace glm {

#if defined(__CUDA_ARCH__)
#define GLM_TAB static __device__ const
#define GLM_FN __device__ __forceinline__
#define GLM_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GLM_MUL(a, b) __dmul_rn((a), (b))
#define GLM_ADD(a, b) __dadd_rn((a), (b))
#define GLM_F2U(f) __float_as_uint(f)
#define GLM_U2F(u) __uint_as_float(u)
#define GLM_D2U(d) ((uint64_t)__double_as_longlong(d))
#define GLM_U2D(u) __longlong_as_double((long long)(u))
#else
#define GLM_TAB static const
#define GLM_FN static inline
#define GLM_FMA(a, b, c) ::fma((a), (b), (c))
#define GLM_MUL(a, b) ((a) * (b))
#define GLM_ADD(a, b) ((a) + (b))
static inline uint32_t glm_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float glm_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint64_t glm_d2u(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }
static inline double glm_u2d(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
#define GLM_F2U(f) arc::glm::glm_f2u(f)
#define GLM_U2F(u) arc::glm::glm_u2f(u)
#define GLM_D2U(d) arc::glm::glm_d2u(d)
#define GLM_U2D(u) arc::glm::glm_u2d(u)
#endif

// __logf_data.tab: {1/c, log(c)} for the 16 sub-intervals of [0x1.66p-1, 0x1.66p0)
GLM_TAB double LOGF_T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
// __powf_log2_data.tab: {1/c, log2(c)}
GLM_TAB double POWF_T[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
// __exp2f_data.tab[i] = bits(2^(i/32)) - (i << 47)
GLM_TAB uint64_t EXP2F_T[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

// log(x), glibc logf.c
GLM_FN float logf_(float x) {
  const uint32_t ix = GLM_F2U(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return ::logf(x);      // zero, negative, subnormal, inf, nan
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double invc = LOGF_T[i][0], logc = LOGF_T[i][1];
  const double z = (double)GLM_U2F(iz);
  const double r = GLM_FMA(z, invc, -1.0);
  const double y0 = GLM_FMA((double)k, 0x1.62e42fefa39efp-1, logc);
  const double r2 = GLM_MUL(r, r);
  double y = GLM_FMA(0x1.5575b0be00b6ap-2, r, -0x1.ffffef20a4123p-2);
  y = GLM_FMA(-0x1.00ea348b88334p-2, r2, y);
  y = GLM_FMA(y, r2, GLM_ADD(y0, r));
  return (float)y;
}

// 2^xd for |xd| < 126 (exp2_inline of powf.c, sign_bias = 0)
GLM_FN float exp2_core(double xd) {
  double kd = GLM_ADD(xd, 0x1.8p+47);
  const uint64_t ki = GLM_D2U(kd);
  kd = GLM_ADD(kd, -0x1.8p+47);
  const double r = GLM_ADD(xd, -kd);
  const uint64_t t = EXP2F_T[ki & 31u] + (ki << 47);
  const double s = GLM_U2D(t);
  const double z = GLM_FMA(0x1.c6af84b912394p-5, r, 0x1.ebfce50fac4f3p-3);
  const double r2 = GLM_MUL(r, r);
  double y = GLM_FMA(0x1.62e42ff0c52d6p-1, r, 1.0);
  y = GLM_FMA(z, r2, y);
  return (float)GLM_MUL(y, s);
}

// log2(x) in double as powf.c's log2_inline evaluates it; x must be finite, positive and normal (else NaN is returned and
// powf_with falls back to the library).  Split from powf_ so a caller raising ONE base to many powers (the 14-band
// Angstrom scaling) evaluates it once.
GLM_FN double powf_log2(float x) {
  const uint32_t ix = GLM_F2U(x);
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) return GLM_U2D(0x7ff8000000000000ULL);
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const uint32_t top = tmp & 0xff800000u;
  const uint32_t iz = ix - top;
  const int k = (int32_t)top >> 23;
  const double invc = POWF_T[i][0], logc = POWF_T[i][1];
  const double z = (double)GLM_U2F(iz);
  const double r = GLM_FMA(z, invc, -1.0);
  const double y0 = GLM_ADD(logc, (double)k);
  const double r2 = GLM_MUL(r, r);
  double yy = GLM_FMA(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
  const double p = GLM_FMA(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
  const double r4 = GLM_MUL(r2, r2);
  double q = GLM_FMA(0x1.71547652ab82bp+0, r, y0);
  q = GLM_FMA(p, r2, q);
  return GLM_FMA(yy, r4, q);
}
// x ** y given l2x = powf_log2(x), glibc powf.c
GLM_FN float powf_with(float x, float y, double l2x) {
  const uint32_t iy = GLM_F2U(y);
  const double ylogx = GLM_MUL((double)y, l2x);
  // y zero / inf / nan, a base outside the fast path (l2x NaN), overflow / underflow of the result: the library's special cases
  if ((2u * iy - 1u) >= (2u * 0x7f800000u - 1u) || !(fabs(ylogx) < 126.0)) return ::powf(x, y);
  return exp2_core(ylogx);
}
GLM_FN float powf_(float x, float y) { return powf_with(x, y, powf_log2(x)); }

// exp(x), glibc expf.c
GLM_FN float expf_(float x) {
  if (!(fabsf(x) < 87.0f)) return ::expf(x);                                 // library handles overflow / underflow / nan
  // the FMA build of libm never rounds z = x * N/ln2: both uses are fused (objdump of __expf_fma)
  const double xd = (double)x;
  double kd = GLM_FMA(0x1.71547652b82fep+5, xd, 0x1.8p+52);
  const uint64_t ki = GLM_D2U(kd);
  kd = GLM_ADD(kd, -0x1.8p+52);
  const double r = GLM_FMA(0x1.71547652b82fep+5, xd, -kd);
  const uint64_t t = EXP2F_T[ki & 31u] + (ki << 47);
  const double s = GLM_U2D(t);
  const double zz = GLM_FMA(0x1.c6af84b912394p-20, r, 0x1.ebfce50fac4f3p-13);
  const double r2 = GLM_MUL(r, r);
  double y = GLM_FMA(0x1.62e42ff0c52d6p-6, r, 1.0);
  y = GLM_FMA(zz, r2, y);
  return (float)GLM_MUL(y, s);
}

}  // namespace glm
}  // namespace arc
