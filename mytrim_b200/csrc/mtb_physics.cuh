// mtb_physics.cuh — the per-collision physics of the transport kernel in single precision.
//
// Same algorithm as the reference's TrimBase::trim() body and MaterialBase::rstop/rpstop, but
// written for the FP32 + SFU pipes of an SM: no libm calls, divisions as MUFU.RCP, powers as
// EX2(y*LG2(x)), and every expression that cancels catastrophically in float is rearranged
// (the closest-approach Newton solve runs in x = r - b, 1 - cos is formed directly, ...).
// Each function cites the reference lines whose result it reproduces.
#ifndef MTB_PHYSICS_CUH
#define MTB_PHYSICS_CUH

#include "mtb_math.cuh"
#include "mtb_types.h"

namespace mtb
{

// .5292 * .8853, the universal screening-length prefactor (material.C:84,104)
#define MTB_SCREEN_K 0.46850076f
#define MTB_PI_F 3.14159265358979323846f

// ZBL proton stopping, e in keV/amu — MaterialBase::rpstop, material.C:133-158.
MTB_HD float
proton_stopping(const DevElement & el, float e)
{
  // below 25 keV/amu: the element's value at 25 (host table, double precision) times (e / 25)^velpwr — the ten
  // MUFU operations of the full expression are the same for every such call
  if (e <= 25.0f)
    return el.sp25 * fpow(e * 0.04f, el.velpwr);
  const float pe = e;
  const float l2 = flog2(pe);
  const float sl = el.pc[0] * fexp2(el.pc[1] * l2) + el.pc[2] * fexp2(el.pc[3] * l2);
  const float sh = el.pc[4] * fexp2(-el.pc[5] * l2) * flog(fdiv(el.pc[6], pe) + el.pc[7] * pe);
  return fdiv(sl * sh, sl + sh);
}

// MaterialBase::rstop outside the tabulated velocity-proportional regime — material.C:187-279.
// e in keV/amu.  The inputs pass through opaque() so that none of this is hoisted out of the
// element loop of material_stopping() into code every collision executes (for Cu->Cu the compiler
// had moved ~30 instructions and 6 MUFU operations of this function in front of the loop).
MTB_HD float
element_stopping_general(const ProjClass & ion, const DevElement & el, float e_in)
{
  const float e = opaque(e_in);
  float se;
  if (ion.Z == 1)
  {
    se = proton_stopping(el, e); // material.C:187-191
  }
  else if (ion.Z == 2)
  {
    // material.C:192-212
    const float he = fmax2(1.0f, e);
    const float b = flog(he);
    const float b2 = b * b;
    const float b4 = b2 * b2;
    float a = 0.2865f + 0.1266f * b - 0.001429f * b2 + 0.02402f * b * b2 - 0.01135f * b4 + 0.001475f * b4 * b;
    float heh = 1.0f - fexp(-fmin2(30.0f, a));
    const float t = 7.6f - b;
    a = 1.0f + (0.007f + 0.00005f * el.fz) * fexp(-(t * t));
    heh *= a * a;
    se = proton_stopping(el, he) * heh * 4.0f;
    if (e <= 1.0f)
      se *= fsqrt(e);
  }
  else
  {
    // material.C:213-279
    const float ifz = opaque(ion.fz);
    const float vfermi = el.vfermi;
    const float v = fsqrt(e * 0.04f) * frcp(vfermi);
    const float v2 = v * v;
    float vr;
    if (v >= 1.0f)
      vr = v * vfermi * (1.0f + fdiv(0.2f, v2));
    else
      vr = (0.75f * vfermi) * (1.0f + v2 * (2.0f / 3.0f) - v2 * v2 * (1.0f / 15.0f));

    const float cb = opaque(ion.cbrt);
    const float icb2 = frcp(cb * cb);
    const float ylow = fmax2(0.13f, icb2); // max(yrmin, vrmin / Z1^(2/3))
    const float yr = fmax2(ylow, vr * icb2);
    const float yr03 = fpow(yr, 0.3f);
    float a = yr03 * (-0.803f + 1.3167f * yr03) + yr * (0.38157f + 0.008983f * yr);

    const float q = fmin2(1.0f, fmax2(0.0f, 1.0f - fexp(-fmin2(a, 50.0f))));
    const float icb = frcp(cb);
    const float b = fmin2(0.43f, fmax2(0.32f, 0.12f + 0.025f * ifz)) * icb;
    const float l0 = (0.8f - q * fmin2(1.2f, 0.6f + ifz * (1.0f / 30.0f))) * icb;
    const float qa = fmax2(0.0f, 0.9f - 0.025f * ifz);
    const float z16 = 0.025f * fmin2(16.0f, ifz);
    float l1;
    if (q < 0.2f)
      l1 = 0.0f;
    else if (q < qa)
      l1 = fdiv(b * (q - 0.2f), fabsf(qa - 0.2000001f));
    else if (q < fmax2(0.0f, 1.0f - z16))
      l1 = b;
    else
      l1 = fdiv(b * (1.0f - q), z16);

    const float l = fmax2(l1, l0 * ion.lfctr);
    const float lx = 4.0f * l * vfermi * (1.0f / 1.919f);
    float zeta = q + el.vf2inv * (1.0f - q) * flog(1.0f + lx * lx);

    // Z1^3 effect: exp(-(7.6 - max(0, ln e))^2).  For e <= 1 the factor is 1 + O(1e-26) == 1.
    if (e > 1.0f)
    {
      const float t = 7.6f - flog(e);
      zeta *= 1.0f + fdiv(0.18f + 0.0015f * el.fz, ifz * ifz) * fexp(-(t * t));
    }

    const float zf = zeta * ifz;
    if (yr <= ylow)
    {
      // velocity-proportional stopping below yrmin
      const float vrmin = fmax2(1.0f, 0.13f * cb * cb);
      const float vmin = 0.5f * (vrmin + fsqrt(fmax2(0.0f, vrmin * vrmin - 0.8f * vfermi * vfermi)));
      const float eee = 25.0f * vmin * vmin;
      const bool p375 = (el.Z == 6) || ((el.Z == 14 || el.Z == 32) && ion.Z <= 19);
      const float ratio = fdiv(e, eee);
      const float scale = p375 ? fpow(ratio, 0.375f) : fsqrt(ratio);
      se = proton_stopping(el, eee) * (zf * zf) * scale;
    }
    else
      se = proton_stopping(el, e) * (zf * zf);
  }
  return se * 10.0f;
}

// Electronic stopping cross-section of one target element — MaterialBase::rstop,
// material.C:160-282.  E in eV; sqrt_e = sqrt(e), e = E * ion.inv_km in keV/amu (the caller has it
// from the free-flight computation).
MTB_HD float
element_stopping(const ProjClass & ion, const LowStop * lowrow, const DevElement & el, float E, float sqrt_e)
{
  const float e = E * ion.inv_km; // keV/amu
  if (ion.Z >= 3)
  {
    // velocity-proportional regime (material.C:259-273): rstop = coef(Z1,Z2) * e^power, with the
    // (Z1, Z2)-only part tabulated in double on the host (mtb_tables.h)
    const LowStop ls = lowrow[el.zslot];
    if (e <= ls.e_max)
    {
      // (a select on purpose: a real branch around the fpow() measured 0.6 % slower than the LG2 the
      // compiler speculates for it)
      return ls.coef * (ls.power == 0.5f ? sqrt_e : fpow(e, ls.power));
    }
  }
  return element_stopping_general(ion, el, e);
}

// MaterialBase::getrstop — material.C:113-122 [eV/Ang]
MTB_HD float
material_stopping(const ProjClass & ion, const LowStop * lowrow, const DevMaterial & M, const DevElement * elements,
                  float E, float sqrt_e)
{
  float se = 0.0f;
  for (int i = 0; i < M.n_elem; ++i)
  {
    const DevElement & el = elements[M.first_elem + i];
    se += element_stopping(ion, lowrow, el, E, sqrt_e) * el.t;
  }
  return se * M.arho;
}

// Scattering result of one binary collision in the centre-of-mass frame.
struct Scatter
{
  float s2; // sin^2(theta/2)
  float c2; // cos^2(theta/2)
};

// Screened potential at reduced distance r: *sum = V(r) r, *dsum = -(V + V' r) r  (trim.C:196-222)
MTB_HD void
screening_sums(int potential, float r, float * sum, float * dsum)
{
  if (potential == MTB_POT_UNIVERSAL)
  {
    // sum c_i exp(-d_i r) and sum c_i d_i exp(-d_i r): the prefactors ride on the FMAs of the sums
    const float ex1 = fexp2(-4.6163355918f * r);
    const float ex2 = fexp2(-1.3594371101f * r);
    const float ex3 = fexp2(-0.58126183197f * r);
    const float ex4 = fexp2(-0.29087617414f * r);
    *sum = fmaf(0.18175f, ex1, fmaf(0.50986f, ex2, fmaf(0.28022f, ex3, 0.028171f * ex4)));
    *dsum = fmaf(0.58156365f, ex1, fmaf(0.4804359794f, ex2, fmaf(0.112900638f, ex3, 0.00567983702f * ex4)));
  }
  else if (potential == MTB_POT_MOLIERE)
  {
    const float ex1 = fexp(-0.3f * r);
    const float ex2 = (ex1 * ex1) * (ex1 * ex1);
    const float e22 = ex2 * ex2;
    const float ex3 = ex2 * (e22 * e22);
    *sum = 0.35f * ex1 + 0.55f * ex2 + 0.1f * ex3;
    *dsum = 0.105f * ex1 + 0.66f * ex2 + 0.6f * ex3;
  }
  else
  {
    const float ex1 = fexp(-0.279f * r);
    const float ex2 = fexp(-0.637f * r);
    const float ex3 = fexp(-1.1919f * r);
    *sum = 0.191f * ex1 + 0.474f * ex2 + 0.335f * ex3;
    *dsum = 0.531865f * ex1 + 0.30181f * ex2 + 0.6437f * ex3;
  }
}

// One Newton step of the closest-approach equation (trim.C:192-233) in x = r - b; returns q = fr/fr1.
//   fr  = b^2/r + v r/eps - r            = sum/eps - x (2b + x)/r
//   fr1 = -b^2/r^2 + (v + v1 r)/eps - 1  = -(b^2/r^2 + 1) - dsum/eps
// Both times r^2: q = r (r sum/eps - x (2b + x)) / -(b^2 + r^2 (1 + dsum/eps)) costs ONE reciprocal; 1/r
// (for v and v1) is formed once, after the iteration.  The SFU pipe bounds this loop: with 1/r inside it,
// 6 of its 34 instructions were MUFU at 8 pipe cycles each (measured: -1 % kernel time without it).
MTB_HD float
newton_step(int potential, float b, float b2, float inv_eps, float x, float * r, float * sum, float * dsum)
{
  *r = b + x;
  screening_sums(potential, *r, sum, dsum);
  const float num = *r * fmaf(*sum * inv_eps, *r, -(x * fmaf(2.0f, b, x)));
  const float den = fmaf(*r * *r, fmaf(*dsum, inv_eps, 1.0f), b2);
  return -fdiv(num, den);
}

// Biersack-Haggmark MAGIC scattering (trim.C:172-272) for reduced energy eps = sqe^2 and reduced
// impact parameter b.  The Newton iteration is the reference's (same start value, same map, same
// stop test |q/r| <= 0.001) but carried in x = r - b so that distant collisions keep full relative
// accuracy in single precision.  Written for the SFU budget of the kernel: per Newton iteration one
// reciprocal and the exponentials of the potential; after it 1/r, one reciprocal for cc, b^cc, one
// square root and ONE reciprocal for everything else (roc, ff, delta and co share a denominator).
MTB_HD Scatter
magic_scatter(int potential, float sqe, float b)
{
  Scatter out;
  const float eps = sqe * sqe;
  if (eps > 10.0f)
  {
    // Rutherford — trim.C:172-179
    const float t = 2.0f * eps * b;
    const float X = (1.0f + b * (1.0f + b)) * (t * t);
    out.s2 = frcp(1.0f + X);
    out.c2 = X * out.s2;
    return out;
  }

  // first guess — trim.C:183-190
  float x = 0.0f;
  {
    float rr = -2.7f * flog(eps * b);
    if (rr >= b)
    {
      rr = -2.7f * flog(eps * rr);
      if (rr >= b)
        x = rr - b;
    }
  }

  const float inv_eps = frcp(eps);
  const float b2 = b * b;
  float r, v, v1, q, sum, dsum; // sum = v*r, dsum = -(v + v1*r)
  int guard = 64; // the reference iterates without a bound; a lane must never spin forever
  do
  {
    q = newton_step(potential, b, b2, inv_eps, x, &r, &sum, &dsum);
    x -= q;
    --guard;
  } while ((fabsf(q) > 0.001f * fabsf(b + x)) & (guard != 0));
  {
    const float inv_r = frcp(r); // r of the last evaluation, as in the reference
    v = sum * inv_r;
    v1 = -(v + dsum) * inv_r;
  }
  r = b + x;

  // trim.C:235-271 (v, v1 are those of the last evaluated r, as in the reference)
  float c_num, c_den, a_k, f_num, f_den;
  if (potential == MTB_POT_UNIVERSAL)
  {
    c_num = 0.011615f; c_den = 0.0071222f; a_k = 0.99229f; f_num = 9.3066f; f_den = 14.813f;
  }
  else if (potential == MTB_POT_MOLIERE)
  {
    c_num = 0.009611f; c_den = 0.005175f; a_k = 0.6743f; f_num = 6.314f; f_den = 10.0f;
  }
  else
  {
    c_num = 0.235809f; c_den = 0.126000f; a_k = 1.0144f; f_num = 6935.0f; f_den = 83550.0f;
  }
  const float cc = fdiv(c_num + sqe, c_den + sqe);
  // aa = 2 eps (1 + a_k/sqrt(eps)) b^cc = 2 (eps + a_k sqrt(eps)) b^cc
  const float aa = 2.0f * fmaf(a_k, sqe, eps) * fpow(b, cc);
  // ff = N/D with N = f_num + eps, D = (f_den + eps)(sqrt(aa^2+1) + aa)   [sqrt(aa^2+1) - aa in its
  // non-cancelling reciprocal form];  delta = (r - b) g,  g = aa ff/(ff + 1) = aa N/(N + D);
  // roc = -2 (eps - v)/v1  =>  r + roc = (r v1 - 2 (eps - v))/v1;
  // 1 - co = (r - b - delta)/(r + roc) = x (N + D - aa N) v1 / ((N + D)(r v1 - 2 (eps - v)))
  const float N = f_num + eps;
  const float D = (f_den + eps) * (fsqrt(fmaf(aa, aa, 1.0f)) + aa);
  const float ND = N + D;
  const float omc = fdiv(x * fmaf(-aa, N, ND) * v1, ND * fmaf(r, v1, -2.0f * (eps - v)));
  const float co = 1.0f - omc;
  out.s2 = omc * (1.0f + co);
  out.c2 = co * co;
  return out;
}

// Free flight (trim.C:88-92) from the tabulated (projectile class, material) constants:
// returns pmax and writes the mean free flight path ls.  sqrtE = sqrt(E) is shared with the
// stopping and the reduced energy of the same step.
MTB_HD float
flight_from_pair(const PairM & pm, float sqrtE, float * ls)
{
  const float eeg = pm.K * sqrtE;
  const float D = eeg + fsqrt(eeg) + 0.125f * fpow(eeg, 0.1f);
  *ls = pm.C2 * (D * D);
  return pm.a * frcp(D);
}

// On-the-fly versions of the pair tables (MaterialBase::average, material.C:77-110) for projectiles
// that have no class (per-primary masses such as fission fragments).
MTB_HD PairM
make_pair_m(const ProjClass & ion, const DevMaterial & M, float tmin)
{
  PairM pm;
  const float a = fdiv(MTB_SCREEN_K, ion.z023 + M.az023);
  const float mu = fdiv(ion.m, M.am);
  const float f = fdiv(a * M.am, M.az * ion.fz * 14.4f * (ion.m + M.am));
  const float opm = 1.0f + mu;
  const float epsdg = fdiv(tmin * f * (opm * opm), 4.0f * mu);
  pm.a = a;
  pm.K = fsqrt(f * epsdg);
  pm.C2 = frcp(MTB_PI_F * M.arho * (a * a));
  pm.sk = fsqrt(ion.inv_km);
  return pm;
}

MTB_HD PairE
make_pair_e(const ProjClass & ion, const DevElement & el)
{
  PairE pe;
  pe.my = fdiv(ion.m, el.m);
  const float opmy = 1.0f + pe.my;
  pe.ec = fdiv(4.0f * pe.my, opmy * opmy);
  const float ai = fdiv(MTB_SCREEN_K, ion.z023 + el.z023);
  pe.inv_ai = frcp(ai);
  pe.sfi = fsqrt(fdiv(ai * el.m, ion.fz * el.fz * 14.4f * (ion.m + el.m)));
  return pe;
}

// ProjClass of an arbitrary (Z, m), computed on the device
MTB_HD ProjClass
make_proj_class(const DevIonZ & iz, int Z, float m)
{
  ProjClass c;
  c.m = (m == 0.0f) ? iz.mm1 : m;
  c.m2 = 2.0f * c.m;
  c.inv_km = fdiv(0.001f, c.m);
  c.fz = (float)Z;
  c.z023 = iz.z023;
  c.cbrt = iz.cbrt;
  c.lfctr = iz.lfctr;
  c.Z = Z;
  return c;
}

} // namespace mtb
#endif
