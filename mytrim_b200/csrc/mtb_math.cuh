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
// Rarely executed, large code of the generic kernels (cluster scan, per-primary table rows, ion log).
// Measured: making these real calls (__noinline__) costs 15-25 % (the call sites spill the lane
// state around them), so they stay inlined; the name only documents what is cold.
#define MTB_HD_COLD __host__ __device__ __forceinline__
#else
#define MTB_HD inline
#define MTB_D inline
#define MTB_HD_COLD inline
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

// Value barrier: the optimiser may not hoist or speculate anything computed from the result above
// the point where this executes (keeps rarely-taken branches out of loop pre-headers).
MTB_HD float
opaque(float x)
{
#if MTB_DEVICE_CODE
  asm volatile("" : "+f"(x));
#endif
  return x;
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

MTB_HD uint32_t
f2u(float f)
{
#if MTB_DEVICE_CODE
  return __float_as_uint(f);
#else
  union { uint32_t u; float f; } cvt;
  cvt.f = f;
  return cvt.u;
#endif
}

MTB_HD float
u2f(uint32_t u)
{
#if MTB_DEVICE_CODE
  return __uint_as_float(u);
#else
  union { uint32_t u; float f; } cvt;
  cvt.u = u;
  return cvt.f;
#endif
}

// (cos, sin) of 2*pi*u01(w) straight from the 32 random bits, on the FMA/ALU pipes only (sincospif
// costs three conversion instructions on the SFU pipe, the busiest one of the transport kernel, and
// ~30 issue slots).  u01(w) = (q + t)/4 with q = the top two bits and t = (the next 21 bits + 1/2)/2^21,
// so the angle is q*90deg + 45deg + 90deg*a with a = t - 1/2 in (-1/2, 1/2): two degree-3 minimax
// polynomials in a^2 (sqrt(1/2)*sin and sqrt(1/2)*cos of 90deg*a; |error| < 8e-8), one add and one subtract for
// the 45deg rotation, and the quadrant as a swap plus sign-bit flips taken from w itself.
MTB_HD void
unit_circle(uint32_t w, float * s, float * c)
{
  const float f = u2f(0x3f800000u | ((w >> 7) & 0x007ffffcu)); // 1 + (21 bits)/2^21
  const float a = f - 1.49999976158142089844f;                 // exact: t - 1/2
  const float z = a * a;
  float ps = -0.003254230599850416f, pc = -0.014430884271860123f;
  ps = fmaf(ps, z, 0.05634243041276932f);
  pc = fmaf(pc, z, 0.1793213188648224f);
  ps = fmaf(ps, z, -0.45676514506340027f);
  pc = fmaf(pc, z, -0.8723555207252502f);
  ps = fmaf(ps, z, 1.1107207536697388f);
  pc = fmaf(pc, z, 0.7071067690849304f);
  ps *= a;
  const float c0 = pc - ps, s0 = pc + ps; // cos, sin of 45deg + 90deg*a
  const bool odd = (w & 0x40000000u) != 0u;
  const uint32_t cb = f2u(odd ? s0 : c0) ^ ((w ^ (w << 1)) & 0x80000000u); // quadrants 1, 2: cos < 0
  const uint32_t sb = f2u(odd ? c0 : s0) ^ (w & 0x80000000u);              // quadrants 2, 3: sin < 0
  *c = u2f(cb);
  *s = u2f(sb);
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

// The key schedule of philox4x32_10 depends only on the key: the host expands it once per launch and
// the kernel reads the round keys as constant-bank operands (saves 18 uniform adds per collision).
MTB_HD void
philox_round_keys(uint32_t k0, uint32_t k1, uint32_t rk[20])
{
  for (int r = 0; r < 10; ++r)
  {
    rk[2 * r] = k0;
    rk[2 * r + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

MTB_HD void
philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t * rk, uint32_t out[4])
{
#pragma unroll
  for (int r = 0; r < 10; ++r)
  {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ rk[2 * r];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c3 = lo0;
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
  return (u2f(0x3f800000u | (x >> 9)) - 1.0f) + 0x1p-24f;
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
