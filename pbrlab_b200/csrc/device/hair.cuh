// Principled Hair: the energy-conserving 4-lobe (R, TT, TRT, residual) fibre BSDF of
// src/closure/energy‐conserving-hair-bsdf.h:42-572 (pbrt-v3 / Chiang et al. 2016 with the "improved lobe
// evaluation" of the 2018 production-path-tracing course) and the material set-up of
// src/shader/hair-shader.cc:19-151.  Local frame: x = fibre tangent.  Everything transcendental goes through
// fast_math (USE_FAST_MATH = 1 there), and SafeLogI0's small-argument branch is the literal
// log(x^2/4 * P(x^2/4)) + 1 the reference computes.  std::cerr diagnostics have no analogue.
#pragma once
#include "common.cuh"
#include "fastmath.cuh"

namespace pbr {
namespace hair {

struct HairBsdf {   // hair-shader.cc:8-17
  vec3 sigma_a;
  float h;
  float v[4];
  float s;
  float eta;
  float alpha;
  vec3 tints[4];
  float transparent_scale;
};

// Pow<n>: n2 = Pow<n/2>(v); return n2*n2*Pow<n&1>(v)  (pbrlab_math.h:40-55) — same association, unrolled at compile time
template <int n> struct PowT { static PBR_HD float f(float v) { const float n2 = PowT<n / 2>::f(v); return n2 * n2 * PowT<(n & 1)>::f(v); } };
template <> struct PowT<1> { static PBR_HD float f(float v) { return v; } };
template <> struct PowT<0> { static PBR_HD float f(float) { return 1.f; } };

PBR_HD void BetamToV(float beta_m, float* vs) {                 // hair-shader.cc:19-27
  vs[0] = Sqr(0.726f * beta_m + 0.812f * Sqr(beta_m) + 3.7f * PowT<20>::f(beta_m));
  vs[1] = 0.25f * vs[0];
  vs[2] = 4.0f * vs[0];
  vs[3] = vs[2];
}
PBR_HD float CalcS(float beta_n) {                               // hair-shader.cc:29-33
  const float beta_n2 = Sqr(beta_n);
  return sqrtf(kPi / 8.0f) * (0.265f * beta_n + 1.194f * beta_n2 + 5.372f * PowT<11>::f(beta_n2));
}
PBR_HD vec3 CalcSigmaAFromRGB(const vec3& c, float beta_n) {     // hair-shader.cc:35-46 (MY_LOG = FastLog)
  const float den = (5.969f - 0.215f * beta_n + 2.532f * Sqr(beta_n) - 10.73f * PowT<3>::f(beta_n) +
                     5.574f * PowT<4>::f(beta_n) + 0.245f * PowT<5>::f(beta_n));
  return vec3(Sqr(fast_math::FastLog(c.x) / den), Sqr(fast_math::FastLog(c.y) / den),
              Sqr(fast_math::FastLog(c.z) / den));
}
PBR_HD vec3 CalcSigmaAUsingMelanin(float melanin, float redness) {   // hair-shader.cc:48-64, random_value = 0.5
  const float factor_random_value = 1.f + 2.f * (0.5f - 0.5f);
  melanin = Clampf(melanin, 0.0f, 1.0f) * factor_random_value;
  redness = Clampf(redness, 0.0f, 1.0f);
  melanin = -fast_math::FastLog(fmaxf_(1.0f - melanin, 0.0001f));
  const float eumelanin = melanin * (1.0f - redness);
  const float pheomelanin = melanin * redness;
  return vec3(fmaxf_(0.0f, eumelanin * 0.506f + pheomelanin * 0.343f),
              fmaxf_(0.0f, eumelanin * 0.841f + pheomelanin * 0.733f),
              fmaxf_(0.0f, eumelanin * 1.653f + pheomelanin * 1.924f));
}

// p: pbrgpu_material.p for type 1 (hair-shader.cc:100-151)
PBR_HD HairBsdf ParamToBsdf(const float* p, float geom_v) {
  HairBsdf b;
  if (p[0] == 0.f) b.sigma_a = CalcSigmaAFromRGB(vec3(p[1], p[2], p[3]), p[8]);
  else b.sigma_a = CalcSigmaAUsingMelanin(p[4], p[5]);
  b.h = geom_v;
  BetamToV(p[7], b.v);
  b.s = CalcS(p[8]);
  b.eta = p[9];
  b.alpha = p[10] * kPi / 180.f;
  b.tints[0] = vec3(p[11], p[12], p[13]);
  b.tints[1] = vec3(p[17], p[18], p[19]);   // transmission tint
  b.tints[2] = vec3(p[14], p[15], p[16]);   // second specular tint
  b.tints[3] = vec3(1.f);
  b.transparent_scale = 1.f;
  return b;
}

#define PBR_LOG(x) fast_math::FastLog((x))
#define PBR_EXP(x) fast_math::FastExp((x))

PBR_HD float SafeASin(float x) {                                  // :42-49 (FastAsin never returns NaN for NaN-free x)
  const float r = fast_math::FastAsin(x);
  if (r != r) return fast_math::FastAsin(Clampf(x, -1.0f, 1.0f));
  return r;
}

PBR_HD float Horner(float x, const float* a, int n) {             // :82-90
  float f = a[n];
  for (int i = n - 1; i >= 0; i--) f = f * x + a[i];
  return f;
}

PBR_HD float SafeLogI0(float x) {                                 // :92-170, USE_IMPROVED_ROBE_EVALUATION
  x = fabsf(x);
  if (x < 7.5f) {
    const float P[9] = {1.00000003928615375e+00f, 2.49999576572179639e-01f, 2.77785268558399407e-02f,
                        1.73560257755821695e-03f, 6.96166518788906424e-05f, 1.89645733877137904e-06f,
                        4.29455004657565361e-08f, 3.90565476357034480e-10f, 1.48095934745267240e-11f};
    const float x22 = x * x / 4.0f;
    return PBR_LOG(x22 * Horner(x22, P, 8)) + 1.0f;
  }
  const float P[5] = {3.98942651588301770e-01f, 4.98327234176892844e-02f, 2.91866904423115499e-02f,
                      1.35614940793742178e-02f, 1.31409251787866793e-01f};
  const float inv_x = 1.0f / x;
  const float Px = Horner(inv_x, P, 4);
  return x + 0.5f * PBR_LOG(Px * Px * inv_x);
}

PBR_HD float Mp(float sin_theta_i, float cos_theta_i, float sin_theta_o, float cos_theta_o, float v) {   // :172-202
  const float ccv = cos_theta_i * cos_theta_o / v;
  const float ssv = sin_theta_i * sin_theta_o / v;
  v = Clampf(v, 1e-5f, 1e4f);
  return PBR_EXP(SafeLogI0(ccv) - ssv - 1.0f / v + PBR_LOG(1.0f / v) - PBR_LOG(1.0f - PBR_EXP(-2.0f / v)));
}

PBR_HD float FrDielectric(float cos_theta_i, float eta_i, float eta_t) {   // :205-229
  cos_theta_i = Clampf(cos_theta_i, -1.0f, 1.0f);
  const bool entering = cos_theta_i > 0.0f;
  if (!entering) {
    const float a = eta_i;
    eta_i = eta_t;
    eta_t = a;
    cos_theta_i = fabsf(cos_theta_i);
  }
  const float sin_theta_i = sqrtf(fmaxf_(0.0f, 1.0f - cos_theta_i * cos_theta_i));
  const float sin_theta_t = eta_i / eta_t * sin_theta_i;
  if (sin_theta_t >= 1.0f) return 1.0f;
  const float cos_theta_t = sqrtf(fmaxf_(0.0f, 1.0f - sin_theta_t * sin_theta_t));
  const float r_parl = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
  const float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
  return (r_parl * r_parl + r_perp * r_perp) * 0.5f;
}

PBR_HD void Ap(float cos_theta_o, float eta, float h, const vec3& T, vec3* ap) {   // :231-255
  const float cos_gamma_o = SafeSqrtf(1.0f - h * h);
  const float cos_theta = cos_theta_o * cos_gamma_o;
  const float f = FrDielectric(cos_theta, 1.0f, eta);
  ap[0] = vec3(f);
  ap[1] = Sqr(1.0f - f) * T;
  ap[2] = ap[1] * T * f;
  ap[3] = ap[2] * f * T / (vec3(1.0f) - T * f);
  if (!finitef_(ap[3].x) || !finitef_(ap[3].y) || !finitef_(ap[3].z)) ap[3] = vec3(0.0f);
}

PBR_HD float Logistic(float x, float s) {                         // :257-262
  x = fabsf(x);
  const float numerator = PBR_EXP(-x / s);
  return numerator / (s * Sqr(1.0f + numerator));
}
PBR_HD float LogisticCDF(float x, float s) { return 1.0f / (1.0f + PBR_EXP(-x / s)); }   // :264-266
PBR_HD float TrimmedLogistic(float x, float s, float a, float b) {   // :268-271
  return Logistic(x, s) / (LogisticCDF(b, s) - LogisticCDF(a, s));
}
PBR_HD float Phi(int p, float gamma_o, float gamma_t) {           // :273-275
  return 2.0f * float(p) * gamma_t - 2.0f * gamma_o + float(p) * kPi;
}
PBR_HD float Fmod(float a, float b) { return a - floorf(a / b) * b; }   // :277-279
PBR_HD float Np(float phi, int p, float s, float gamma_o, float gamma_t) {   // :281-289
  float dphi = Fmod(phi - Phi(p, gamma_o, gamma_t), 2.0f * kPi);
  if (dphi >= kPi) dphi -= 2.0f * kPi;
  return TrimmedLogistic(dphi, s, -kPi, kPi);
}

// per-(omega_out, material) precomputation shared by eval and sample (:300-355 / :425-476)
struct HairCommon {
  float sin_theta_o, cos_theta_o;
  float sin_o[4], cos_o[4];   // tilt-corrected (sin,cos) theta_o per lobe
  float phi_o, gamma_o, gamma_t;
  vec3 ap[4];
  float apPdf[4];
};

PBR_HD HairCommon HairSetup(const vec3& omega_out, const HairBsdf& b) {
  HairCommon c;
  c.sin_theta_o = omega_out.x;
  c.cos_theta_o = SafeSqrtf(1.0f - Sqr(c.sin_theta_o));
  float s2k[3], c2k[3];
  fast_math::FastSincos(b.alpha, &s2k[0], &c2k[0]);
  for (int i = 1; i < 3; i++) {
    s2k[i] = 2.0f * s2k[i - 1] * c2k[i - 1];
    c2k[i] = Sqr(c2k[i - 1]) - Sqr(s2k[i - 1]);
  }
  c.sin_o[0] = c.sin_theta_o * c2k[1] - c.cos_theta_o * s2k[1];
  c.cos_o[0] = c.cos_theta_o * c2k[1] + c.sin_theta_o * s2k[1];
  c.sin_o[1] = c.sin_theta_o * c2k[0] + c.cos_theta_o * s2k[0];
  c.cos_o[1] = c.cos_theta_o * c2k[0] - c.sin_theta_o * s2k[0];
  c.sin_o[2] = c.sin_theta_o * c2k[2] + c.cos_theta_o * s2k[2];
  c.cos_o[2] = c.cos_theta_o * c2k[2] - c.sin_theta_o * s2k[2];
  c.sin_o[3] = c.sin_theta_o;
  c.cos_o[3] = c.cos_theta_o;
  c.phi_o = fast_math::FastAtan2(omega_out.z, omega_out.y);

  const float sin_theta_t = c.sin_theta_o / b.eta;
  const float cos_theta_t = SafeSqrtf(1.f - Sqr(sin_theta_t));
  const float etap = sqrtf(b.eta * b.eta - Sqr(c.sin_theta_o)) / c.cos_theta_o;
  const float sin_gamma_t = b.h / etap;
  const float cos_gamma_t = SafeSqrtf(1.0f - Sqr(sin_gamma_t));
  c.gamma_t = SafeASin(sin_gamma_t);
  const float l = b.transparent_scale * 2.0f * cos_gamma_t / cos_theta_t;
  const vec3 T(PBR_EXP(-b.sigma_a.x * l), PBR_EXP(-b.sigma_a.y * l), PBR_EXP(-b.sigma_a.z * l));
  c.gamma_o = SafeASin(b.h);
  Ap(c.cos_theta_o, b.eta, b.h, T, c.ap);
  float sum = 0.0f;
  for (int i = 0; i < 4; i++) sum = sum + RgbToY(c.ap[i]);
  for (int i = 0; i < 4; i++) c.apPdf[i] = RgbToY(c.ap[i]) / sum;
  return c;
}

// shared tail of eval and sample: sum the four lobes (:369-404 / :540-571)
PBR_HD vec3 HairLobes(const HairCommon& c, const HairBsdf& b, float sin_theta_i, float cos_theta_i, float phi,
                      float* pdf) {
  *pdf = 0.0f;
  float pdfs[4];
  vec3 ret(0.0f);
  for (int p = 0; p < 3; p++) {
    const float mpnp = Mp(sin_theta_i, cos_theta_i, c.sin_o[p], c.cos_o[p], b.v[p]) *
                       Np(phi, p, b.s, c.gamma_o, c.gamma_t);
    pdfs[p] = mpnp * c.apPdf[p];
    ret = ret + mpnp * c.ap[p] * b.tints[p];
  }
  const float mpnp = Mp(sin_theta_i, cos_theta_i, c.sin_theta_o, c.cos_theta_o, b.v[3]) * (1.0f / (2.0f * kPi));
  pdfs[3] = mpnp * c.apPdf[3];
  ret = ret + mpnp * c.ap[3] * b.tints[3];
  if (!finitef_(ret.x) || !finitef_(ret.y) || !finitef_(ret.z)) return vec3(0.0f);
  float sum = 0.0f;
  for (int i = 0; i < 4; i++) sum = sum + pdfs[i];
  *pdf = sum;
  if (!finitef_(*pdf)) {
    *pdf = 0.0f;
    return vec3(0.0f);
  }
  return ret;
}

// returns f * cos, writes pdf (:295-405)
PBR_HD vec3 EnergyConservingHairBsdfCosPdf(const vec3& omega_in, const vec3& omega_out, const HairBsdf& b,
                                           float* pdf) {
  const HairCommon c = HairSetup(omega_out, b);
  const float sin_theta_i = omega_in.x;
  const float cos_theta_i = SafeSqrtf(1.0f - Sqr(sin_theta_i));
  const float phi_i = fast_math::FastAtan2(omega_in.z, omega_in.y);
  const float phi = phi_i - c.phi_o;
  return HairLobes(c, b, sin_theta_i, cos_theta_i, phi, pdf);
}

PBR_HD float SampleTrimmedLogistic(float s, float a, float b, float u) {   // :407-417 (the inf guard there is dead)
  const float T = LogisticCDF(b, s) - LogisticCDF(a, s);
  return -s * PBR_LOG(1.0f / (u * T + 1.0f / (1.0f + PBR_EXP(-a / s))) - 1.0f);
}

// 4 randoms: lobe, longitudinal (2), azimuthal.  Returns f * cos, writes omega_in and pdf (:419-572)
PBR_HD vec3 EnergyConservingHairSample(const vec3& omega_out, const HairBsdf& b, const float* us, vec3* omega_in,
                                       float* pdf) {
  const HairCommon c = HairSetup(omega_out, b);
  int p = 0;
  float u0 = us[0];
  for (p = 0; p < 4 - 1; p++) {
    if (u0 < c.apPdf[p]) break;
    u0 -= c.apPdf[p];
  }
  float sin_theta_i, cos_theta_i;
  {
    const float u1 = us[1], u2 = us[2];
    const float u = 1.0f + b.v[p] * PBR_LOG(u1 + (1.0f - u1) * PBR_EXP(-2.0f / b.v[p]));
    sin_theta_i = -u * c.sin_o[p] + SafeSqrtf(1.0f - Sqr(u)) * fast_math::FastCos(2.0f * kPi * u2) * c.cos_o[p];
    cos_theta_i = SafeSqrtf(1.0f - Sqr(sin_theta_i));
  }
  float dphi;
  if (p < 3) dphi = Phi(p, c.gamma_o, c.gamma_t) + SampleTrimmedLogistic(b.s, -kPi, kPi, us[3]);
  else dphi = 2.0f * kPi * us[3];
  const float phi_i = c.phi_o + dphi;
  *omega_in = vec3(sin_theta_i, cos_theta_i * fast_math::FastCos(phi_i), cos_theta_i * fast_math::FastSin(phi_i));
  return HairLobes(c, b, sin_theta_i, cos_theta_i, dphi, pdf);
}

#undef PBR_LOG
#undef PBR_EXP

}  // namespace hair
}  // namespace pbr
