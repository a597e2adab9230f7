// mtb_math.cuh — single-precision math layer of the transport kernels.
//
// Everything the collision step needs beyond + - * fma goes through these wrappers so that
// (a) the device build maps them onto the SFU (MUFU.RCP/RSQ/EX2/LG2: one issue slot each) and
// (b) the same source compiles for the host (tests/hostsim.cpp) with libm equivalents, which is
// how the control flow is debugged in a container without a GPU.  Relative accuracy of every
// wrapper is <= ~2.4e-7 (one or two float ulps), far inside the 1e-5 trajectory tolerance.
#ifndef MTB_MATH_CUH
#define MTB_MATH_CUH

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MTB_HD __host__ __device__ __forceinline__
#define MTB_D __device__ __forceinline__
#else
#define MTB_HD inline
#define MTB_D inline
#endif

#if defined(__CUDA_ARCH__)
#define MTB_DEVICE_CODE 1
#else
#define MTB_DEVICE_CODE 0
#endif

namespace mtb
{

MTB_HD float
frcp(float x)
{
#if MTB_DEVICE_CODE
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

MTB_HD float
fdiv(float a, float b)
{
  return a * frcp(b);
}

MTB_HD float
frsqrt(float x)
{
#if MTB_DEVICE_CODE
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}

MTB_HD float
fsqrt(float x)
{
#if MTB_DEVICE_CODE
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}

MTB_HD float
fexp2(float x)
{
#if MTB_DEVICE_CODE
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return exp2f(x);
#endif
}

MTB_HD float
flog2(float x)
{
#if MTB_DEVICE_CODE
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return log2f(x);
#endif
}

// e^x = 2^(x*log2 e).  The product is rounded once, so the relative error grows as
// ~4e-8*|x*log2 e| on top of the 2-ulp MUFU.EX2; every exponential on the hot path has
// |x*log2 e| < ~20 where it matters (larger arguments belong to terms that are e^-14 of the sum).
MTB_HD float
fexp(float x)
{
  return fexp2(x * 1.44269502162933349609375f);
}

MTB_HD float
flog(float x)
{
  return flog2(x) * 0.693147182464599609375f;
}

// x^y for x > 0
MTB_HD float
fpow(float x, float y)
{
  return fexp2(y * flog2(x));
}

MTB_HD float
fmax2(float a, float b)
{
  return fmaxf(a, b);
}
MTB_HD float
fmin2(float a, float b)
{
  return fminf(a, b);
}

// sin/cos of 2*pi*u, u in [0,1]
MTB_HD void
fsincos2pi(float u, float * s, float * c)
{
#if MTB_DEVICE_CODE
  sincospif(2.0f * u, s, c);
#else
  const double a = 6.283185307179586476925 * (double)u;
  *s = (float)sin(a);
  *c = (float)cos(a);
#endif
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), one 128-bit block per collision step.
// ---------------------------------------------------------------------------------------------
MTB_HD uint32_t
mulhi32(uint32_t a, uint32_t b)
{
#if MTB_DEVICE_CODE
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

MTB_HD void
philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
  for (int r = 0; r < 10; ++r)
  {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

// 32 random bits -> uniform in (0,1): the top 23 bits become the mantissa of a float in [1,2),
// minus 1, plus half a grid step.  Pure ALU/FMA-pipe work (an I2F conversion would occupy the SFU
// pipe, the busiest one of the transport kernel), exactly the same on host and device.
MTB_HD float
u01(uint32_t x)
{
  const uint32_t bits = 0x3f800000u | (x >> 9);
  float f;
#if MTB_DEVICE_CODE
  f = __uint_as_float(bits);
#else
  union { uint32_t u; float f; } cvt;
  cvt.u = bits;
  f = cvt.f;
#endif
  return (f - 1.0f) + 0x1p-24f;
}

// Stream id of a recoil: the spare word of the parent's Philox block of the collision that created
// it (already a keyed hash of (parent id, collision index)) in the high half, a mixed copy of the
// parent id and collision index in the low half.
MTB_HD uint64_t
child_uid(uint64_t uid, uint32_t ic, uint32_t w3)
{
  const uint32_t lo = ((uint32_t)uid ^ (uint32_t)(uid >> 32)) * 0x9E3779B9u + ic * 0x85EBCA6Bu;
  return ((uint64_t)w3 << 32) | (uint64_t)lo;
}

} // namespace mtb
#endif
