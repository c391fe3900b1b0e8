// The reference's hair code does not call libm: it goes through the OpenImageIO-derived polynomial approximations
// in src/pbrlab_math.h:135-341 (namespace fast_math), whose error (FastLog2: 7.6e-6 absolute) is *part of the
// reference result*.  These are the same minimax polynomials — the coefficients are the published OIIO fmath
// constants — evaluated in the same order; std::fma there is pbr_fma here (a single-rounding FFMA on the device).
#pragma once
#include "common.cuh"

namespace pbr {
namespace fast_math {

PBR_HD float Madd(float a, float b, float c) { return pbr_fma(a, b, c); }  // src/pbrlab_math.h:101-108

PBR_HD int FastRint(float x) {                                              // src/pbrlab_math.h:111-119
#if defined(__CUDA_ARCH__)
  return __float2int_rn(x);
#else
  return static_cast<int>(rintf(x));
#endif
}

// shared argument reduction of FastSin/FastCos/FastSincos (src/pbrlab_math.h:127-133)
PBR_HD float ReducePi(float x, int* q_out) {
  const int q = FastRint(x * 0.31830988618379067154f);
  const float qf = float(q);
  x = Madd(qf, -0.78515625f * 4, x);
  x = Madd(qf, -0.00024187564849853515625f * 4, x);
  x = Madd(qf, -3.7747668102383613586e-08f * 4, x);
  x = Madd(qf, -1.2816720341285448015e-12f * 4, x);
  x = 1.57079632679489661923f - (1.57079632679489661923f - x);
  *q_out = q;
  return x;
}

PBR_HD float SinPoly(float x, float s) {
  float u = 2.6083159809786593541503e-06f;
  u = Madd(u, s, -0.0001981069071916863322258f);
  u = Madd(u, s, +0.00833307858556509017944336f);
  u = Madd(u, s, -0.166666597127914428710938f);
  u = Madd(s, u * x, x);
  return u;
}
PBR_HD float CosPoly(float s) {
  float u = -2.71811842367242206819355e-07f;
  u = Madd(u, s, +2.47990446951007470488548e-05f);
  u = Madd(u, s, -0.00138888787478208541870117f);
  u = Madd(u, s, +0.0416666641831398010253906f);
  u = Madd(u, s, -0.5f);
  u = Madd(u, s, +1.0f);
  return u;
}

PBR_HD float FastSin(float x) {                                             // src/pbrlab_math.h:121-147
  int q;
  x = ReducePi(x, &q);
  const float s = x * x;
  if ((q & 1) != 0) x = -x;
  float u = SinPoly(x, s);
  if (fabsf(u) > 1.0f) u = 0.0f;
  return u;
}

PBR_HD float FastCos(float x) {                                             // src/pbrlab_math.h:149-171
  int q;
  x = ReducePi(x, &q);
  const float s = x * x;
  float u = CosPoly(s);
  if ((q & 1) != 0) u = -u;
  if (fabsf(u) > 1.0f) u = 0.0f;
  return u;
}

PBR_HD void FastSincos(float x, float* sine, float* cosine) {               // src/pbrlab_math.h:173-201
  int q;
  x = ReducePi(x, &q);
  const float s = x * x;
  if ((q & 1) != 0) x = -x;
  float su = SinPoly(x, s);
  float cu = CosPoly(s);
  if ((q & 1) != 0) cu = -cu;
  if (fabsf(su) > 1.0f) su = 0.0f;
  if (fabsf(cu) > 1.0f) cu = 0.0f;
  *sine = su;
  *cosine = cu;
}

PBR_HD float FastExp2(float xval) {                                          // src/pbrlab_math.h:203-226
  float x = fmaxf_(-126.0f, fminf_(126.0f, xval));
  const int m = int(x);
  x -= float(m);
  x = 1.0f - (1.0f - x);
  float r = 1.33336498402e-3f;
  r = Madd(x, r, 9.810352697968e-3f);
  r = Madd(x, r, 5.551834031939e-2f);
  r = Madd(x, r, 0.2401793301105f);
  r = Madd(x, r, 0.693144857883f);
  r = Madd(x, r, 1.0f);
  return u2f(f2u(r) + (uint32_t(m) << 23));
}

// FastExp2(x * T(1 / kM_LN2)): the double reciprocal is rounded to float before the multiply (:228-233)
PBR_HD float FastExp(float x) { return FastExp2(x * float(1 / 0.69314718055994530942)); }

PBR_HD float FastAtan2(float y, float x) {                                   // src/pbrlab_math.h:235-263
  const float a = fabsf(x);
  const float b = fabsf(y);
  const float k = (b == 0) ? 0.0f : ((a == b) ? 1.0f : (b > a ? a / b : b / a));
  const float s = 1.0f - (1.0f - k);
  const float t = s * s;
  float r = s * Madd(0.430165678f, t, 1.0f) / Madd(Madd(0.0579354987f, t, 0.763007998f), t, 1.0f);
  if (b > a) r = 1.570796326794896557998982f - r;
  if (f2u(x) & 0x80000000u) r = float(kPi) - r;
  return copysignf(r, y);
}

PBR_HD float FastAsin(float x) {                                             // src/pbrlab_math.h:265-278
  const float f = fabsf(x);
  const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
  const float a = 1.57079632679489661923f -
                  sqrtf(1.0f - m) * (1.5707963267f + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
  return copysignf(a, x);
}

PBR_HD float FastLog2(float xval) {                                          // src/pbrlab_math.h:302-330
  const float x = fmaxf_(FLT_MIN, fminf_(FLT_MAX, xval));
  const uint32_t bits = f2u(x);
  const int exponent = int(bits >> 23) - 127;
  const float f = u2f((bits & 0x007FFFFFu) | 0x3f800000u) - 1.0f;
  const float f2 = f * f;
  const float f4 = f2 * f2;
  float hi = Madd(f, -0.00931049621349f, 0.05206469089414f);
  float lo = Madd(f, 0.47868480909345f, -0.72116591947498f);
  hi = Madd(f, hi, -0.13753123777116f);
  hi = Madd(f, hi, 0.24187369696082f);
  hi = Madd(f, hi, -0.34730547155299f);
  lo = Madd(f, lo, 1.442689881667200f);
  return ((f4 * hi) + (f * lo)) + float(exponent);
}

PBR_HD float FastLog(float x) { return FastLog2(x) * 0.69314718055994530942f; }  // src/pbrlab_math.h:332-336

}  // namespace fast_math
}  // namespace pbr
