// pbr_oracle — CPU restatement of pbrlab's path-tracing hot path.  TEST INFRASTRUCTURE ONLY (see pbr_oracle.h):
// never linked into, imported by or called from the product path; pinned against the golden vectors generated from
// the compiled unmodified reference (tests/test_oracle_port.py).
//
// Plain scalar C++, one function per reference function, each citing the file:line under /root/reference it
// follows.  Arithmetic is float32 in the reference's evaluation order; build with -ffp-contract=off (the reference
// is built without FMA contraction; its explicit std::fma calls are std::fma here too).
// Ray queries restate what Embree 4.2 does for pbrlab's two geometry types over a plain binary
// BVH of our own — the acceleration structure does not influence which hit is the closest one.
#include "pbr_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

// ------------------------------------------------------------------ float3 = nanort::real3<float> (src/nanort.h:314-404)
struct F3 {
  float x, y, z;
  F3() : x(0), y(0), z(0) {}
  explicit F3(float s) : x(s), y(s), z(s) {}
  F3(float a, float b, float c) : x(a), y(b), z(c) {}
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  float& at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline F3 operator+(F3 a, F3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline F3 operator-(F3 a, F3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline F3 operator*(F3 a, F3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline F3 operator/(F3 a, F3 b) { return F3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline F3 operator*(F3 a, float f) { return F3(a.x * f, a.y * f, a.z * f); }
inline F3 operator*(float f, F3 a) { return F3(a.x * f, a.y * f, a.z * f); }
inline F3 operator/(F3 a, float f) { return F3(a.x / f, a.y / f, a.z / f); }
inline F3 operator-(F3 a) { return F3(-a.x, -a.y, -a.z); }
inline float Dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 Cross(F3 a, F3 b) { return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float Length(F3 v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
const float kFltEps = std::numeric_limits<float>::epsilon();
inline F3 Normalized(F3 v) {   // nanort::vnormalize: untouched when the length is <= FLT_EPSILON (nanort.h:380-390)
  const float len = Length(v);
  if (std::fabs(len) > kFltEps) {
    const float inv = 1.0f / len;
    return F3(v.x * inv, v.y * inv, v.z * inv);
  }
  return v;
}

// ------------------------------------------------------------------ src/pbrlab_math.h, src/pbrlab-util.h
const float kPi = 3.141592653589793f, kPiInv = 0.318309886183f, kEps = 1e-3f, kInf = 1.844E18f;   // pbrlab_math.h:7-11
inline float Sqr(float v) { return v * v; }
inline float SafeSqrtf(float f) { return std::sqrt(std::max(f, 0.0f)); }                            // :17
inline F3 Lerp(F3 a, F3 b, float u) { return (1.0f - u) * a + u * b; }                               // :29-32
inline F3 Lerp3(F3 a, F3 b, F3 c, float u, float v) { return (1.0f - u - v) * a + u * b + v * c; }  // :34-38
inline float Clampf(float x, float a, float b) { return std::max(a, std::min(b, x)); }              // pbrlab-util.h:9-12
inline float Saturate(float x) { return Clampf(x, 0.f, 1.f); }
inline float Average(F3 c) { return (c.x + c.y + c.z) / 3.f; }                                      // :19
inline float SpectrumNorm(F3 c) { return std::max({c.x, c.y, c.z}); }                               // :21-23
inline F3 SafeDivideSpectrum(F3 a, F3 b) {                                                           // :25-46
  return F3(std::fabs(b.x) < kFltEps ? 0.f : a.x / b.x, std::fabs(b.y) < kFltEps ? 0.f : a.y / b.y,
            std::fabs(b.z) < kFltEps ? 0.f : a.z / b.z);
}
inline float RgbToY(F3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }           // :48-51
inline bool IsBlack(F3 v) { return (std::fabs(v.x) + std::fabs(v.y) + std::fabs(v.z)) < kFltEps; }  // :53-56
inline bool IsFinite(F3 v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); }

// ------------------------------------------------------------------ fast_math (src/pbrlab_math.h:101-336, OIIO fmath)
namespace fm {
inline float Madd(float a, float b, float c) { return std::fma(a, b, c); }
inline int FastRint(float x) { return static_cast<int>(std::rint(x)); }
inline unsigned Bits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float FromBits(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
const float kPi2 = float(1.57079632679489661923);
inline float Reduce(float x, int* q) {   // argument reduction shared by FastSin / FastCos / FastSincos (:126-133)
  *q = FastRint(x * float(0.31830988618379067154));
  const float qf = float(*q);
  x = Madd(qf, -0.78515625f * 4, x);
  x = Madd(qf, -0.00024187564849853515625f * 4, x);
  x = Madd(qf, -3.7747668102383613586e-08f * 4, x);
  x = Madd(qf, -1.2816720341285448015e-12f * 4, x);
  return kPi2 - (kPi2 - x);
}
inline float SinTail(float x, float s) {
  float u = 2.6083159809786593541503e-06f;
  u = Madd(u, s, -0.0001981069071916863322258f);
  u = Madd(u, s, +0.00833307858556509017944336f);
  u = Madd(u, s, -0.166666597127914428710938f);
  return Madd(s, u * x, x);
}
inline float CosTail(float s) {
  float u = -2.71811842367242206819355e-07f;
  u = Madd(u, s, +2.47990446951007470488548e-05f);
  u = Madd(u, s, -0.00138888787478208541870117f);
  u = Madd(u, s, +0.0416666641831398010253906f);
  u = Madd(u, s, -0.5f);
  return Madd(u, s, +1.0f);
}
inline float FastSin(float x) {    // :121-147
  int q;
  x = Reduce(x, &q);
  const float s = x * x;
  if (q & 1) x = -x;
  float u = SinTail(x, s);
  if (std::fabs(u) > 1.0f) u = 0.0f;
  return u;
}
inline float FastCos(float x) {    // :149-171
  int q;
  x = Reduce(x, &q);
  float u = CosTail(x * x);
  if (q & 1) u = -u;
  if (std::fabs(u) > 1.0f) u = 0.0f;
  return u;
}
inline void FastSincos(float x, float* sine, float* cosine) {   // :173-201
  int q;
  x = Reduce(x, &q);
  const float s = x * x;
  if (q & 1) x = -x;
  float su = SinTail(x, s), cu = CosTail(s);
  if (q & 1) cu = -cu;
  if (std::fabs(su) > 1.0f) su = 0.0f;
  if (std::fabs(cu) > 1.0f) cu = 0.0f;
  *sine = su;
  *cosine = cu;
}
inline float FastExp2(float xval) {   // :203-226
  float x = Clampf(xval, -126.0f, 126.0f);
  const int m = int(x);
  x -= float(m);
  x = 1.0f - (1.0f - x);
  float r = 1.33336498402e-3f;
  r = Madd(x, r, 9.810352697968e-3f);
  r = Madd(x, r, 5.551834031939e-2f);
  r = Madd(x, r, 0.2401793301105f);
  r = Madd(x, r, 0.693144857883f);
  r = Madd(x, r, 1.0f);
  return FromBits(Bits(r) + (unsigned(m) << 23));
}
inline float FastExp(float x) { return FastExp2(x * float(1 / 0.69314718055994530942)); }   // :228-233
inline float FastAtan2(float y, float x) {   // :235-263
  const float a = std::fabs(x), b = std::fabs(y);
  const float k = (b == 0) ? 0.0f : ((a == b) ? 1.0f : (b > a ? a / b : b / a));
  const float s = 1.0f - (1.0f - k);
  const float t = s * s;
  float r = s * Madd(0.430165678f, t, 1.0f) / Madd(Madd(0.0579354987f, t, 0.763007998f), t, 1.0f);
  if (b > a) r = 1.570796326794896557998982f - r;
  if (Bits(x) & 0x80000000u) r = float(kPi) - r;
  return copysignf(r, y);
}
inline float FastAsin(float x) {   // :265-278
  const float f = std::fabs(x);
  const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
  const float a = kPi2 - sqrtf(1.0f - m) * (1.5707963267f + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
  return std::copysign(a, x);
}
inline float FastLog2(float xval) {   // :302-330
  const float x = Clampf(xval, std::numeric_limits<float>::min(), std::numeric_limits<float>::max());
  const unsigned bits = Bits(x);
  const int exponent = int(bits >> 23) - 127;
  const float f = FromBits((bits & 0x007FFFFF) | 0x3f800000) - 1.0f;
  const float f2 = f * f, f4 = f2 * f2;
  float hi = Madd(f, -0.00931049621349f, 0.05206469089414f);
  float lo = Madd(f, 0.47868480909345f, -0.72116591947498f);
  hi = Madd(f, hi, -0.13753123777116f);
  hi = Madd(f, hi, 0.24187369696082f);
  hi = Madd(f, hi, -0.34730547155299f);
  lo = Madd(f, lo, 1.442689881667200f);
  return ((f4 * hi) + (f * lo)) + float(exponent);
}
inline float FastLog(float x) { return FastLog2(x) * float(0.69314718055994530942); }   // :332-336
}  // namespace fm

// ------------------------------------------------------------------ PCG32 (src/random/rng.h:17-65)
struct Rng {
  uint64_t state, inc;
  Rng(uint64_t initstate, uint64_t initseq) {
    state = 0U;
    inc = (initseq << 1U) | 1U;
    Next();
    state += initstate;
    Next();
  }
  uint32_t Next() {
    const uint64_t old = state;
    state = old * uint64_t(6364136223846793005) + inc;
    const uint32_t xorshifted = uint32_t(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = uint32_t(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((-static_cast<int>(rot)) & 31));
  }
  float Draw() {
    const uint32_t u = (Next() >> 9) | 0x3f800000u;
    float f;
    memcpy(&f, &u, 4);
    return f - 1.0f;
  }
};

// ------------------------------------------------------------------ src/sampler/sampling-utils.h
inline F3 CosineSampleHemisphere(float u1, float u2) {   // :10-14
  const float a = u1 * 2.0f * kPi, r = std::sqrt(u2);
  return F3(std::cos(a) * r, std::sin(a) * r, std::sqrt(std::max(1.0f - u2, 0.0f)));
}
inline F3 UniformSampleSphere(float u1, float u2) {      // :16-23
  const float u = 2.0f * u2 - 1.0f;
  const float norm = std::sqrt(std::max(0.0f, 1.0f - u * u));
  const float theta = 2.0f * kPi * u1;
  return F3(norm * std::cos(theta), u, norm * std::sin(theta));
}
inline float PowerHeuristicWeight(float sampled_pdf, float other_pdf) {   // :27-57
  if (sampled_pdf > other_pdf) {
    const float r = other_pdf / sampled_pdf;
    return 1 / (1 + r * r);
  }
  if (sampled_pdf < other_pdf) {
    const float r = sampled_pdf / other_pdf;
    return 1 - 1 / (1 + r * r);
  }
  return 0.5f;
}

// ------------------------------------------------------------------ closures
inline float FresnelDielectricCos(float cos_, float eta) {   // src/closure/closure-util.h:10-29
  if (std::fabs(eta) < kFltEps) return 1.0f;
  if (cos_ < 0.0f) eta = 1.0f / eta;
  const float c = std::fabs(cos_);
  float g = eta * eta - 1 + c * c;
  if (g > 0) {
    g = std::sqrt(g);
    const float A = (g - c) / (g + c);
    const float B = (c * (g + c) - 1) / (c * (g - c) + 1);
    return 0.5f * A * A * (1 + B * B);
  }
  return 1.0f;
}

// src/closure/microfacet-ggx.h
inline float D_GTR1(F3 h, float alpha) {     // :48-53
  if (alpha >= 1.0f) return 1.0f / kPi;
  const float a2 = alpha * alpha;
  const float t = 1.0f + (a2 - 1.0f) * h.z * h.z;
  return (a2 - 1.0f) / (kPi * logf(a2) * t);
}
inline float D_GTR2(F3 h, float alpha2) {    // :55-63
  const float c2 = h.z * h.z, c4 = c2 * c2;
  const float t2 = (1.0f - c2) / c2;
  return alpha2 / (kPi * c4 * (alpha2 + t2) * (alpha2 + t2));
}
float GgxEvalPdf(F3 wi, F3 wo, float ax, float ay, int distrib, float* pdf) {   // MicrofacetGGXBsdfPdf :164-245
  const float cos_o = wo.z, cos_i = wi.z;
  if (cos_o > 0 && cos_i > 0) {
    const F3 m = Normalized(wi + wo);
    float alpha2 = ax * ay;
    float D, G1o, G1i;
    if (std::fabs(ax - ay) < kFltEps) {
      if (distrib == 1) {
        D = D_GTR1(m, ax);
        alpha2 = 0.0625f;
      } else {
        D = D_GTR2(m, alpha2);
      }
      G1o = 2 / (1 + SafeSqrtf(1 + alpha2 * (1 - cos_o * cos_o) / (cos_o * cos_o)));
      G1i = 2 / (1 + SafeSqrtf(1 + alpha2 * (1 - cos_i * cos_i) / (cos_i * cos_i)));
    } else {
      const float sx = -m.x / (m.z * ax), sy = -m.y / (m.z * ay);
      const float slope_len = 1 + sx * sx + sy * sy;
      const float c2 = m.z * m.z, c4 = c2 * c2;
      D = 1.f / ((slope_len * slope_len) * kPi * alpha2 * c4);
      const float tan_o2 = (1.f - cos_o * cos_o) / (cos_o * cos_o);
      float alphaO2 = (wo.x * wo.x) * (ax * ax) + (wo.y * wo.y) * (ay * ay);
      alphaO2 /= wo.x * wo.x + wo.y * wo.y;
      G1o = 2 / (1 + SafeSqrtf(1 + alphaO2 * tan_o2));
      const float tan_i2 = (1 - cos_i * cos_i) / (cos_i * cos_i);
      float alphaI2 = (wi.x * wi.x) * (ax * ax) + (wi.y * wi.y) * (ay * ay);
      alphaI2 /= wi.x * wi.x + wi.y * wi.y;
      G1i = 2 / (1 + SafeSqrtf(1 + alphaI2 * tan_i2));
    }
    const float G = G1o * G1i;
    const float common = D * 0.25f / cos_o / cos_i;
    float f = G * common;
    if (distrib == 1) f = 0.25f * f;
    *pdf = G1o * common;
    return f;
  }
  *pdf = 0.f;
  return 0.f;
}
void GgxSampleSlopes(float cos_i, float sin_i, float randu, float randv, float* slope_x, float* slope_y) {   // :65-118
  if (cos_i >= 0.99999f) {
    const float r = sqrtf(randu / (1.0f - randu));
    const float phi = 2.0f * kPi * randv;
    *slope_x = r * cosf(phi);
    *slope_y = r * sinf(phi);
    return;
  }
  const float tan_i = sin_i / cos_i;
  const float G1_inv = 0.5f * (1.0f + SafeSqrtf(1.0f + tan_i * tan_i));
  const float A = 2.0f * randu * G1_inv - 1.0f;
  const float AA = A * A;
  const float tmp = 1.0f / (AA - 1.0f);
  const float B = tan_i, BB = B * B;
  const float D = SafeSqrtf(BB * (tmp * tmp) - (AA - BB) * tmp);
  const float s1 = B * tmp - D, s2 = B * tmp + D;
  *slope_x = (A < 0.0f || s2 * tan_i > 1.0f) ? s1 : s2;
  float S;
  if (randv > 0.5f) {
    S = 1.0f;
    randv = 2.0f * (randv - 0.5f);
  } else {
    S = -1.0f;
    randv = 2.0f * (0.5f - randv);
  }
  const float z = (randv * (randv * (randv * 0.27385f - 0.73369f) + 0.46341f)) /
                  (randv * (randv * (randv * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
  *slope_y = S * z * SafeSqrtf(1.0f + (*slope_x) * (*slope_x));
}
F3 GgxSampleHalfVector(F3 wo, float ax, float ay, float randu, float randv) {   // MicrofacetSampleStretched :121-162
  const F3 s = Normalized(F3(ax * wo.x, ay * wo.y, wo.z));
  float costheta = 1.0f, sintheta = 0.0f, cosphi = 1.0f, sinphi = 0.0f;
  if (s.z < 0.99999f) {
    costheta = s.z;
    sintheta = SafeSqrtf(1.0f - costheta * costheta);
    const float invlen = 1.0f / sintheta;
    cosphi = s.x * invlen;
    sinphi = s.y * invlen;
  }
  float slope_x = 0.f, slope_y = 0.f;
  GgxSampleSlopes(costheta, sintheta, randu, randv, &slope_x, &slope_y);
  const float tmp = cosphi * slope_x - sinphi * slope_y;
  slope_y = sinphi * slope_x + cosphi * slope_y;
  slope_x = tmp;
  slope_x = ax * slope_x;
  slope_y = ay * slope_y;
  return Normalized(F3(-slope_x, -slope_y, 1.0f));
}
void GgxSample(F3 wo, float ax, float ay, float u0, float u1, int distrib, F3* wi) {   // MicrofacetGGXSample :247-286
  if (wo.z > 0.f) {
    const F3 m = GgxSampleHalfVector(wo, ax, ay, u0, u1);
    const float cos_m_o = Dot(m, wo);
    if (cos_m_o > 0) *wi = 2 * cos_m_o * m - wo;   // otherwise wi keeps the caller's value (0)
  }
  (void)distrib;
}

// ------------------------------------------------------------------ Principled BSDF (src/shader/cycles-principled-shader.cc)
struct Principled {   // CyclesPrincipledBsdf :20-45
  bool diffuse = false, subsurface = false, specular = false, clearcoat = false;
  F3 diffuse_w, ss_weight, ss_albedo, ss_radius, spec_w, spec_color, cc_w, cc_color;
  float ax = 1.f, ay = 1.f, ior = 1.5f, cc_ax = 1.f, cc_ay = 1.f, cc_ior = 1.5f;
};
struct Weights { float d, ss, sp, cc; };

F3 SpecularColor(F3 wi, F3 wo, F3 color, float ior) {   // :54-61
  const F3 h = Normalized(wi + wo);
  const float f0 = FresnelDielectricCos(1.0f, ior);
  const float fh = (FresnelDielectricCos(Dot(h, wo), ior) - f0) / (1.0f - f0);
  return color * (1.f - fh) + F3(fh);
}
Weights SampleWeights(F3 wo, const Principled& b) {      // FetchClosureSampleWeight :63-112
  Weights w;
  const F3 mirror(-wo.x, -wo.y, wo.z);
  w.d = b.diffuse ? RgbToY(b.diffuse_w) : 0.f;
  w.ss = b.subsurface ? RgbToY(b.ss_weight) : 0.f;
  w.sp = b.specular ? RgbToY(b.spec_w * SpecularColor(mirror, wo, b.spec_color, b.ior)) : 0.f;
  w.cc = b.clearcoat ? RgbToY(b.cc_w * SpecularColor(mirror, wo, b.cc_color, b.cc_ior)) : 0.f;
  float sum = 0.0f;
  sum += w.d; sum += w.ss; sum += w.sp; sum += w.cc;
  w.d /= sum; w.ss /= sum; w.sp /= sum; w.cc /= sum;
  if (!std::isfinite(w.d)) w.d = 0.f;
  if (!std::isfinite(w.ss)) w.ss = 0.f;
  if (!std::isfinite(w.sp)) w.sp = 0.f;
  if (!std::isfinite(w.cc)) w.cc = 0.f;
  return w;
}
void EvalBsdf(F3 wi, F3 wo, const Principled& b, F3* f, float* pdf) {   // :114-155
  const Weights w = SampleWeights(wo, b);
  *f = F3(0.f);
  *pdf = 0.f;
  if (b.diffuse) {   // LambertBrdfPdf (src/closure/lambert.h:11-20)
    *f = *f + b.diffuse_w * kPiInv;
    *pdf += w.d * (wi.z * kPiInv);
  }
  if (b.specular) {
    float p = 0.f;
    const float v = GgxEvalPdf(wi, wo, b.ax, b.ay, 2, &p);
    *f = *f + b.spec_w * SpecularColor(wi, wo, b.spec_color, b.ior) * v;
    *pdf += w.sp * p;
  }
  if (b.clearcoat) {
    float p = 0.f;
    const float v = GgxEvalPdf(wi, wo, b.cc_ax, b.cc_ay, 1, &p);
    *f = *f + b.cc_w * SpecularColor(wi, wo, b.cc_color, b.cc_ior) * v;
    *pdf += w.cc * p;
  }
}

// random-walk-sss.h:35-104
void BssrdfSetup(F3* weight, F3* albedo, F3* radius, F3* diffuse_weight) {   // (burley=true, scale_mfp=true, eq5)
  *diffuse_weight = F3(0.f);
  F3 kd(0.f);
  int channels = 3;
  for (int i = 0; i < 3; ++i) {
    if ((*radius)[i] < 1e-8f) {
      kd.at(i) = (*weight)[i];
      weight->at(i) = 0.f;
      radius->at(i) = 0.f;
      channels--;
    }
  }
  if (channels < 3) *diffuse_weight = kd;
  if (channels > 0) {
    const F3 l = 0.25f * (1.0f / kPi) * (*radius);   // BssrdfBurleyCompatibleMfp :48-50
    const F3 A = *albedo;
    auto fit5 = [](float a) { return 1.85f - a + 7.0f * std::fabs((a - 0.8f) * (a - 0.8f) * (a - 0.8f)); };   // :40-43
    *radius = l / F3(fit5(A.x), fit5(A.y), fit5(A.z));   // `bool use_eq5` arrives as `int mode` != 0 -> equation 5
  }
}

// Texture::FetchFloat3 -> BilinearFilter, clamp addressing (src/texture.cc:43-72, src/image-utils.cc:99-167)
struct TextureRef { const float* px; uint32_t w, h, c; };
F3 TextureFetch3(const TextureRef& t, float u, float v) {
  const float uu = std::min(std::max(u, 0.0f), 1.0f), vv = std::min(std::max(v, 0.0f), 1.0f);
  const float px = float(t.w) * uu, py = float(t.h) * vv;
  const int W = int(t.w), H = int(t.h);
  const int x0 = std::max(0, std::min(W - 1, int(px))), y0 = std::max(0, std::min(H - 1, int(py)));
  const int x1 = (x0 + 1 >= W) ? W - 1 : x0 + 1, y1 = (y0 + 1 >= H) ? H - 1 : y0 + 1;
  const float dx = px - float(x0), dy = py - float(y0);
  const float w0 = (1.0f - dx) * (1.0f - dy), w1 = (1.0f - dx) * dy, w2 = dx * (1.0f - dy), w3 = dx * dy;
  const int st = int(t.c);
  const int i00 = st * (y0 * W + x0), i01 = st * (y0 * W + x1), i10 = st * (y1 * W + x0), i11 = st * (y1 * W + x1);
  float o[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < 3 && k < st; ++k)
    o[k] = t.px[i00 + k] * w0 + t.px[i10 + k] * w1 + t.px[i01 + k] * w2 + t.px[i11 + k] * w3;
  return F3(o[0], o[1], o[2]);
}

// :244-412; base / ss_color arrive texture-resolved (:281-301)
Principled ParamToBsdf(const float* p, const F3& base, const F3& ss_color) {
  const float subsurface = p[3];
  const F3 ss_radius(p[4], p[5], p[6]);
  const float metallic = p[10], specular = p[11], specular_tint = p[12], roughness = p[13], anisotropic = p[14];
  const float clearcoat = p[18], clearcoat_roughness = p[19], transmission = p[21];
  const float cut = kEps;
  Principled b;
  const float diffuse_weight = (1.0f - Saturate(metallic)) * (1.0f - Saturate(transmission));
  const float final_transmission = Saturate(transmission) * (1.0f - Saturate(metallic));
  const float specular_weight = 1.0f - final_transmission;
  {
    const F3 mixed = ss_color * subsurface + base * (1.0f - subsurface);
    if (Average(mixed) > cut) {
      if (subsurface < cut && diffuse_weight > cut) {
        b.diffuse = true;
        b.diffuse_w = F3(1.f) * base * diffuse_weight;
      } else if (subsurface > cut) {
        b.subsurface = true;
        b.ss_weight = F3(1.f) * mixed * diffuse_weight;
        b.ss_albedo = mixed;
        b.ss_radius = ss_radius * subsurface;
        F3 add(0.f);
        BssrdfSetup(&b.ss_weight, &b.ss_albedo, &b.ss_radius, &add);
        if (!IsBlack(add)) {
          b.diffuse = true;
          b.diffuse_w = b.diffuse_w + add;
        }
      }
    }
  }
  if (specular_weight > cut && (specular > cut || metallic > cut)) {
    b.specular = true;
    b.spec_w = F3(1.f) * specular_weight;
    b.ior = (2.0f / (1.0f - SafeSqrtf(0.08f * specular))) - 1.0f;
    const float aspect = SafeSqrtf(1.0f - anisotropic * 0.9f);
    const float r2 = roughness * roughness;
    b.ax = r2 / aspect;
    b.ay = r2 * aspect;
    const float y = RgbToY(base);
    const F3 rho_tint = y > 0.0f ? base / y : F3(0.0f);
    const F3 rho_spec = Lerp(F3(1.0f), rho_tint, specular_tint);
    b.spec_color = Lerp(0.08f * specular * rho_spec, base, metallic);
  }
  if (clearcoat > cut) {
    b.clearcoat = true;
    b.cc_w = F3(0.25f * clearcoat);
    b.cc_ax = b.cc_ay = clearcoat_roughness * clearcoat_roughness;
    b.cc_color = F3(0.04f);
    b.cc_ior = 1.5f;
  }
  return b;
}

// ------------------------------------------------------------------ hair closure (src/closure/energy‐conserving-hair-bsdf.h)
namespace hair {
inline float SafeASin(float x) { return fm::FastAsin(x); }   // :42-49 (FastAsin never yields NaN for finite x)
inline float Horner(float x, const float* a, int n) {        // :82-90
  float f = a[n];
  for (int i = n - 1; i >= 0; i--) f = f * x + a[i];
  return f;
}
float SafeLogI0(float x) {                                    // :92-170, USE_IMPROVED_ROBE_EVALUATION
  x = std::fabs(x);
  if (x < 7.5f) {
    static const float P[] = {1.00000003928615375e+00f, 2.49999576572179639e-01f, 2.77785268558399407e-02f,
                              1.73560257755821695e-03f, 6.96166518788906424e-05f, 1.89645733877137904e-06f,
                              4.29455004657565361e-08f, 3.90565476357034480e-10f, 1.48095934745267240e-11f};
    const float x22 = x * x / 4.0f;
    return fm::FastLog(x22 * Horner(x22, P, 8)) + 1.0f;   // the live code: log(x22 * P(x22)) + 1
  }
  static const float P[] = {3.98942651588301770e-01f, 4.98327234176892844e-02f, 2.91866904423115499e-02f,
                            1.35614940793742178e-02f, 1.31409251787866793e-01f};
  const float inv_x = 1.0f / x;
  const float Px = Horner(inv_x, P, 4);
  return x + 0.5f * fm::FastLog(Px * Px * inv_x);
}
float Mp(float sin_i, float cos_i, float sin_o, float cos_o, float v) {   // :172-202
  const float ccv = cos_i * cos_o / v;
  const float ssv = sin_i * sin_o / v;
  v = Clampf(v, 1e-5f, 1e4f);
  return fm::FastExp(SafeLogI0(ccv) - ssv - 1.0f / v + fm::FastLog(1.0f / v) - fm::FastLog(1.0f - fm::FastExp(-2.0f / v)));
}
float FrDielectric(float cos_i, float eta_i, float eta_t) {   // :205-229
  cos_i = Clampf(cos_i, -1.0f, 1.0f);
  if (!(cos_i > 0.0f)) {
    std::swap(eta_i, eta_t);
    cos_i = std::fabs(cos_i);
  }
  const float sin_i = std::sqrt(std::max(0.0f, 1.0f - cos_i * cos_i));
  const float sin_t = eta_i / eta_t * sin_i;
  if (sin_t >= 1.0f) return 1.0f;
  const float cos_t = std::sqrt(std::max(0.0f, 1.0f - sin_t * sin_t));
  const float r_parl = ((eta_t * cos_i) - (eta_i * cos_t)) / ((eta_t * cos_i) + (eta_i * cos_t));
  const float r_perp = ((eta_i * cos_i) - (eta_t * cos_t)) / ((eta_i * cos_i) + (eta_t * cos_t));
  return (r_parl * r_parl + r_perp * r_perp) * 0.5f;
}
void Ap(float cos_theta_o, float eta, float h, F3 T, F3* ap) {   // :231-255
  const float cos_gamma_o = SafeSqrtf(1.0f - h * h);
  const float f = FrDielectric(cos_theta_o * cos_gamma_o, 1.0f, eta);
  ap[0] = F3(f);
  ap[1] = Sqr(1.0f - f) * T;
  ap[2] = ap[1] * T * f;
  ap[3] = ap[2] * f * T / (F3(1.0f) - T * f);
  if (!IsFinite(ap[3])) ap[3] = F3(0.0f);
}
inline float Logistic(float x, float s) {                      // :257-262
  x = std::fabs(x);
  const float n = fm::FastExp(-x / s);
  return n / (s * Sqr(1.0f + n));
}
inline float LogisticCDF(float x, float s) { return 1.0f / (1.0f + fm::FastExp(-x / s)); }   // :264-266
inline float TrimmedLogistic(float x, float s, float a, float b) {                              // :268-271
  return Logistic(x, s) / (LogisticCDF(b, s) - LogisticCDF(a, s));
}
inline float Phi(int p, float gamma_o, float gamma_t) { return 2.0f * float(p) * gamma_t - 2.0f * gamma_o + float(p) * kPi; }
inline float Fmod(float a, float b) { return a - std::floor(a / b) * b; }
float Np(float phi, int p, float s, float gamma_o, float gamma_t) {   // :281-289
  float dphi = Fmod(phi - Phi(p, gamma_o, gamma_t), 2.0f * kPi);
  if (dphi >= kPi) dphi -= 2.0f * kPi;
  return TrimmedLogistic(dphi, s, -kPi, kPi);
}

struct Bsdf {   // HairBsdf (src/shader/hair-shader.cc:8-17)
  F3 sigma_a;
  float h, v[4], s, eta, alpha;
  F3 tints[4];
  float transparent_scale;
};
struct Common {   // the part of eval and sample that depends on omega_out only (:302-357 == :427-489)
  float sin_o, cos_o, sin_crt[4], cos_crt[4], phi_o, gamma_o, gamma_t;
  F3 ap[4];
  float ap_pdf[4];
};
Common Setup(F3 wo, const Bsdf& b) {
  Common c;
  c.sin_o = wo.x;
  c.cos_o = SafeSqrtf(1.0f - Sqr(c.sin_o));
  float s2k[3], c2k[3];
  fm::FastSincos(b.alpha, &s2k[0], &c2k[0]);
  for (int i = 1; i < 3; i++) {
    s2k[i] = 2.0f * s2k[i - 1] * c2k[i - 1];
    c2k[i] = Sqr(c2k[i - 1]) - Sqr(s2k[i - 1]);
  }
  c.sin_crt[0] = c.sin_o * c2k[1] - c.cos_o * s2k[1];
  c.cos_crt[0] = c.cos_o * c2k[1] + c.sin_o * s2k[1];
  c.sin_crt[1] = c.sin_o * c2k[0] + c.cos_o * s2k[0];
  c.cos_crt[1] = c.cos_o * c2k[0] - c.sin_o * s2k[0];
  c.sin_crt[2] = c.sin_o * c2k[2] + c.cos_o * s2k[2];
  c.cos_crt[2] = c.cos_o * c2k[2] - c.sin_o * s2k[2];
  c.sin_crt[3] = c.sin_o;
  c.cos_crt[3] = c.cos_o;
  c.phi_o = fm::FastAtan2(wo.z, wo.y);
  const float sin_t = c.sin_o / b.eta;
  const float cos_t = SafeSqrtf(1.f - Sqr(sin_t));
  const float etap = std::sqrt(b.eta * b.eta - Sqr(c.sin_o)) / c.cos_o;
  const float sin_gamma_t = b.h / etap;
  const float cos_gamma_t = SafeSqrtf(1.0f - Sqr(sin_gamma_t));
  c.gamma_t = SafeASin(sin_gamma_t);
  const float l = b.transparent_scale * 2.0f * cos_gamma_t / cos_t;
  const F3 T(fm::FastExp(-b.sigma_a.x * l), fm::FastExp(-b.sigma_a.y * l), fm::FastExp(-b.sigma_a.z * l));
  c.gamma_o = SafeASin(b.h);
  Ap(c.cos_o, b.eta, b.h, T, c.ap);
  float sum = 0.0f;
  for (int i = 0; i < 4; i++) sum = sum + RgbToY(c.ap[i]);
  for (int i = 0; i < 4; i++) c.ap_pdf[i] = RgbToY(c.ap[i]) / sum;
  return c;
}
F3 Lobes(const Common& c, const Bsdf& b, float sin_i, float cos_i, float phi, float* pdf) {   // :366-404 == :540-571
  float pdfs[4];
  F3 ret(0.0f);
  for (int p = 0; p < 3; p++) {
    const float mpnp = Mp(sin_i, cos_i, c.sin_crt[p], c.cos_crt[p], b.v[p]) * Np(phi, p, b.s, c.gamma_o, c.gamma_t);
    pdfs[p] = mpnp * c.ap_pdf[p];
    ret = ret + mpnp * c.ap[p] * b.tints[p];
  }
  const float mpnp = Mp(sin_i, cos_i, c.sin_o, c.cos_o, b.v[3]) * (1.0f / (2.0f * kPi));
  pdfs[3] = mpnp * c.ap_pdf[3];
  ret = ret + mpnp * c.ap[3] * b.tints[3];
  *pdf = 0.f;
  if (!IsFinite(ret)) return F3(0.0f);
  float sum = 0.0f;
  for (int i = 0; i < 4; i++) sum = sum + pdfs[i];
  if (!std::isfinite(sum)) return F3(0.0f);
  *pdf = sum;
  return ret;
}
F3 EvalCosPdf(F3 wi, F3 wo, const Bsdf& b, float* pdf) {   // EnergyConservingHairBsdfCosPdf :295-405
  const Common c = Setup(wo, b);
  const float sin_i = wi.x;
  const float cos_i = SafeSqrtf(1.0f - Sqr(sin_i));
  const float phi_i = fm::FastAtan2(wi.z, wi.y);
  return Lobes(c, b, sin_i, cos_i, phi_i - c.phi_o, pdf);
}
float SampleTrimmedLogistic(float s, float a, float b, float u) {   // :407-417
  const float T = LogisticCDF(b, s) - LogisticCDF(a, s);
  return -s * fm::FastLog(1.0f / (u * T + 1.0f / (1.0f + fm::FastExp(-a / s))) - 1.0f);
}
F3 Sample(F3 wo, const Bsdf& b, const float* us, F3* wi, float* pdf) {   // EnergyConservingHairSample :419-572
  const Common c = Setup(wo, b);
  int p = 0;
  float u0 = us[0];
  for (p = 0; p < 3; p++) {
    if (u0 < c.ap_pdf[p]) break;
    u0 -= c.ap_pdf[p];
  }
  const float u1 = us[1], u2 = us[2];
  const float u = 1.0f + b.v[p] * fm::FastLog(u1 + (1.0f - u1) * fm::FastExp(-2.0f / b.v[p]));
  const float sin_i = -u * c.sin_crt[p] + SafeSqrtf(1.0f - Sqr(u)) * fm::FastCos(2.0f * kPi * u2) * c.cos_crt[p];
  const float cos_i = SafeSqrtf(1.0f - Sqr(sin_i));
  float dphi;
  if (p < 3) dphi = Phi(p, c.gamma_o, c.gamma_t) + SampleTrimmedLogistic(b.s, -kPi, kPi, us[3]);
  else dphi = 2.0f * kPi * us[3];
  const float phi_i = c.phi_o + dphi;
  *wi = F3(sin_i, cos_i * fm::FastCos(phi_i), cos_i * fm::FastSin(phi_i));
  return Lobes(c, b, sin_i, cos_i, dphi, pdf);
}
template <int N> float PowN(float v) { const float h = PowN<N / 2>(v); return h * h * PowN<(N & 1)>(v); }
template <> float PowN<1>(float v) { return v; }
template <> float PowN<0>(float) { return 1.f; }
Bsdf FromParam(const float* p, float geom_v) {   // ParamToBsdf (src/shader/hair-shader.cc:100-151)
  Bsdf b;
  const float beta_m = p[7], beta_n = p[8];
  if (p[0] == 0.f) {   // kRGB: CalcSigmaAFromRGB :35-46
    const float den = 5.969f - 0.215f * beta_n + 2.532f * Sqr(beta_n) - 10.73f * PowN<3>(beta_n) +
                      5.574f * PowN<4>(beta_n) + 0.245f * PowN<5>(beta_n);
    b.sigma_a = F3(Sqr(fm::FastLog(p[1]) / den), Sqr(fm::FastLog(p[2]) / den), Sqr(fm::FastLog(p[3]) / den));
  } else {             // kMelanin: CalcSigmaAUsingMelaninParameter :48-64 (random_value = 0.5)
    const float factor = 1.f + 2.f * (0.5f - 0.5f);
    float melanin = Clampf(p[4], 0.0f, 1.0f) * factor;
    const float redness = Clampf(p[5], 0.0f, 1.0f);
    melanin = -fm::FastLog(std::max(1.0f - melanin, 0.0001f));
    const float eu = melanin * (1.0f - redness), pheo = melanin * redness;
    b.sigma_a = F3(std::max(0.0f, eu * 0.506f + pheo * 0.343f), std::max(0.0f, eu * 0.841f + pheo * 0.733f),
                   std::max(0.0f, eu * 1.653f + pheo * 1.924f));
  }
  b.h = geom_v;
  b.v[0] = Sqr(0.726f * beta_m + 0.812f * Sqr(beta_m) + 3.7f * PowN<20>(beta_m));   // BetamToV :19-27
  b.v[1] = 0.25f * b.v[0];
  b.v[2] = 4.0f * b.v[0];
  b.v[3] = b.v[2];
  const float bn2 = Sqr(beta_n);
  b.s = std::sqrt(kPi / 8.0f) * (0.265f * beta_n + 1.194f * bn2 + 5.372f * PowN<11>(bn2));   // CalcS :29-33
  b.eta = p[9];
  b.alpha = p[10] * kPi / 180.f;
  b.tints[0] = F3(p[11], p[12], p[13]);
  b.tints[1] = F3(p[17], p[18], p[19]);
  b.tints[2] = F3(p[14], p[15], p[16]);
  b.tints[3] = F3(1.f);
  b.transparent_scale = 1.f;
  return b;
}
}  // namespace hair

// ------------------------------------------------------------------ scene + ray queries
struct Ray { F3 o, d; float tmin, tmax; };
struct Hit {   // TraceResult (src/raytracer/raytracer.h:9-17)
  F3 ng;
  float t = 0, u = 0, v = 0;
  uint32_t inst = 0xFFFFFFFFu, geom = 0xFFFFFFFFu, prim = 0xFFFFFFFFu;
  uint32_t flat_prim = 0xFFFFFFFFu;   // index in pbo_set_* order
  bool curve = false;
  bool valid() const { return inst != 0xFFFFFFFFu; }
};
struct Material { uint32_t type, tex[2], reserved; float p[24]; };

struct Box { float lo[3], hi[3]; };
struct Node { Box box; uint32_t left, right, first, count; };   // count > 0: leaf

struct Bvh {
  std::vector<Node> nodes;
  std::vector<uint32_t> order;
  void Build(const std::vector<Box>& b) {
    nodes.clear();
    order.resize(b.size());
    for (size_t i = 0; i < b.size(); ++i) order[i] = uint32_t(i);
    if (b.empty()) return;
    nodes.reserve(b.size() / 2 + 16);
    Split(b, 0, uint32_t(b.size()));
  }
  uint32_t Split(const std::vector<Box>& b, uint32_t first, uint32_t last) {
    const uint32_t id = uint32_t(nodes.size());
    nodes.push_back(Node());
    Box bx, cb;
    for (int k = 0; k < 3; ++k) { bx.lo[k] = cb.lo[k] = 3e38f; bx.hi[k] = cb.hi[k] = -3e38f; }
    for (uint32_t i = first; i < last; ++i) {
      const Box& p = b[order[i]];
      for (int k = 0; k < 3; ++k) {
        bx.lo[k] = std::min(bx.lo[k], p.lo[k]); bx.hi[k] = std::max(bx.hi[k], p.hi[k]);
        const float c = 0.5f * (p.lo[k] + p.hi[k]);
        cb.lo[k] = std::min(cb.lo[k], c); cb.hi[k] = std::max(cb.hi[k], c);
      }
    }
    nodes[id].box = bx;
    int axis = 0;
    for (int k = 1; k < 3; ++k) if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
    if (last - first <= 4 || !(cb.hi[axis] > cb.lo[axis])) {
      nodes[id].first = first; nodes[id].count = last - first; nodes[id].left = nodes[id].right = 0;
      return id;
    }
    const uint32_t mid = (first + last) / 2;
    std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + last, [&](uint32_t x, uint32_t y) {
      return b[x].lo[axis] + b[x].hi[axis] < b[y].lo[axis] + b[y].hi[axis];
    });
    nodes[id].count = 0;
    const uint32_t l = Split(b, first, mid);
    const uint32_t r = Split(b, mid, last);
    nodes[id].left = l; nodes[id].right = r;
    return id;
  }
};

inline bool RayBox(const Ray& r, const F3& inv, const Box& b, float tfar) {
  float t0 = r.tmin, t1 = tfar;
  const float o[3] = {r.o.x, r.o.y, r.o.z}, id[3] = {inv.x, inv.y, inv.z};
  for (int k = 0; k < 3; ++k) {
    float a = (b.lo[k] - o[k]) * id[k], c = (b.hi[k] - o[k]) * id[k];
    if (a > c) std::swap(a, c);
    // NaN (0 * inf) must not cull: widen instead
    if (!(a == a)) a = -std::numeric_limits<float>::infinity();
    if (!(c == c)) c = std::numeric_limits<float>::infinity();
    t0 = std::max(t0, a);
    t1 = std::min(t1, c);
  }
  return t0 <= t1 * 1.000001f + 1e-6f;
}

// Embree's Moeller-Trumbore (kernels/geometry/triangle_intersector_moeller.h:69-110, triangle.h:40-41):
// v0, e1 = v0 - v1, e2 = v2 - v0, Ng = e2 x e1; inclusive edges, no culling, tnear < t <= tfar.
// Embree's SSE dot/cross association: a.x*b.x + (a.y*b.y + a.z*b.z)  (common/math/vec3.h:205-209)
inline float EDot(F3 a, F3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline bool IntersectTriangle(const Ray& r, float tfar, F3 v0, F3 v1, F3 v2, float* t, float* u, float* v, F3* ng) {
  const F3 e1 = v0 - v1, e2 = v2 - v0;
  const F3 Ng = Cross(e2, e1);
  const F3 C = v0 - r.o;
  const F3 R = Cross(C, r.d);
  const float den = EDot(Ng, r.d);
  const float absden = std::fabs(den);
  const float sgn = den < 0.f || (den == 0.f && std::signbit(den)) ? -1.f : 1.f;
  const float U = EDot(R, e2) * sgn, V = EDot(R, e1) * sgn;
  if (!(den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= absden)) return false;
  const float T = EDot(Ng, C) * sgn;
  if (!(absden * r.tmin < T && T <= absden * tfar)) return false;
  const float rcp = 1.0f / absden;
  *t = T * rcp; *u = U * rcp; *v = V * rcp;
  *ng = Ng;
  return true;
}

// Flat cubic Bezier curves (RTC_GEOMETRY_TYPE_FLAT_BEZIER_CURVE), Embree's ribbon intersector with N = 4 sub-segments
// (kernels/geometry/curve_intersector_ribbon.h:72-177, quad_intersector.h:15-74,
//  curve_intersector_precalculations.h:20-26, common/math/linearspace3.h:117-124, kernels/subdiv/bezier_curve.h:12-51).
struct P4 { float x, y, z, w; };
inline void BezierW(float t1, float* b) { const float t0 = 1.0f - t1; b[0] = t0 * t0 * t0; b[1] = 3.0f * t1 * (t0 * t0); b[2] = 3.0f * (t1 * t1) * t0; b[3] = t1 * t1 * t1; }
inline void BezierD(float t1, float* b) { const float t0 = 1.0f - t1; b[0] = 3.0f * (-(t0 * t0)); b[1] = 3.0f * (-2.0f * (t0 * t1) + t0 * t0); b[2] = 3.0f * (2.0f * (t0 * t1) - t1 * t1); b[3] = 3.0f * (t1 * t1); }
inline float Mix(const float* b, float a0, float a1, float a2, float a3) { return b[0] * a0 + (b[1] * a1 + (b[2] * a2 + b[3] * a3)); }
inline F3 ENormalize(F3 v) { return v * (1.0f / std::sqrt(EDot(v, v))); }
bool IntersectCurve(const Ray& r, float tfar, const P4* cp, float* t_out, float* u_out, float* v_out) {
  const float depth_scale = 1.0f / std::sqrt(EDot(r.d, r.d));
  const F3 N = depth_scale * r.d;
  const F3 dx0(0.f, N.z, -N.y), dx1(-N.z, 0.f, N.x);
  const F3 dx = ENormalize(EDot(dx0, dx0) > EDot(dx1, dx1) ? dx0 : dx1);
  const F3 dy = ENormalize(Cross(N, dx));
  const F3 dz = N * depth_scale;
  F3 q[4];
  float m = 0.f;
  for (int i = 0; i < 4; ++i) {
    const F3 p = F3(cp[i].x, cp[i].y, cp[i].z) - r.o;
    q[i] = F3(p.x * dx.x + (p.y * dx.y + p.z * dx.z), p.x * dy.x + (p.y * dy.y + p.z * dy.z),
              p.x * dz.x + (p.y * dz.y + p.z * dz.z));
    m = std::max(m, std::max(std::max(std::fabs(q[i].x), std::fabs(q[i].y)), std::fabs(q[i].z)));
  }
  const float eps = 4.0f * kFltEps * m;
  bool found = false;
  float bt = 0, bu = 0, bv = 0;
  for (int i = 0; i < 4; ++i) {
    float w0[4], w1[4], d0[4], d1[4];
    BezierW(float(i) / 4.0f, w0); BezierW(float(i + 1) / 4.0f, w1);
    const P4 p0 = {Mix(w0, q[0].x, q[1].x, q[2].x, q[3].x), Mix(w0, q[0].y, q[1].y, q[2].y, q[3].y),
                   Mix(w0, q[0].z, q[1].z, q[2].z, q[3].z), Mix(w0, cp[0].w, cp[1].w, cp[2].w, cp[3].w)};
    const P4 p1 = {Mix(w1, q[0].x, q[1].x, q[2].x, q[3].x), Mix(w1, q[0].y, q[1].y, q[2].y, q[3].y),
                   Mix(w1, q[0].z, q[1].z, q[2].z, q[3].z), Mix(w1, cp[0].w, cp[1].w, cp[2].w, cp[3].w)};
    {   // cylinder_culling_test
      const float ax = p1.x - p0.x, ay = p1.y - p0.y;
      const float num = ax * p0.y - ay * p0.x, den2 = ax * ax + ay * ay;
      const float rr = std::max(p0.w, p1.w);
      if (!(num * num <= rr * rr * den2)) continue;
    }
    BezierD(float(i) / 4.0f, d0); BezierD(float(i + 1) / 4.0f, d1);
    F3 t0(Mix(d0, q[0].x, q[1].x, q[2].x, q[3].x), Mix(d0, q[0].y, q[1].y, q[2].y, q[3].y), Mix(d0, q[0].z, q[1].z, q[2].z, q[3].z));
    F3 t1(Mix(d1, q[0].x, q[1].x, q[2].x, q[3].x), Mix(d1, q[0].y, q[1].y, q[2].y, q[3].y), Mix(d1, q[0].z, q[1].z, q[2].z, q[3].z));
    const F3 chord(p1.x - p0.x, p1.y - p0.y, p1.z - p0.z);
    if (std::max(std::max(std::fabs(t0.x), std::fabs(t0.y)), std::fabs(t0.z)) < eps) t0 = chord;
    if (std::max(std::max(std::fabs(t1.x), std::fabs(t1.y)), std::fabs(t1.z)) < eps) t1 = chord;
    const F3 n0 = ENormalize(F3(t0.y, -t0.x, 0.0f)), n1 = ENormalize(F3(t1.y, -t1.x, 0.0f));
    const F3 P0(p0.x, p0.y, p0.z), P1(p1.x, p1.y, p1.z);
    const F3 lp0 = p0.w * n0 + P0, lp1 = p1.w * n1 + P1, up0 = P0 - p0.w * n0, up1 = P1 - p1.w * n1;
    // intersect_quad_backface_culling with O = 0, D = (0,0,1), quad (lp0, lp1, up1, up0)
    const F3 edb = lp1 - up0;
    const float WW = Cross(up0, edb).z;
    const bool first = WW <= 0.0f;
    const F3 a = first ? lp0 : up1, b = first ? lp1 : up0, c = first ? up0 : lp1;
    const F3 e0 = c - a, e1 = a - b;
    const float U = Cross(a, e0).z, V = Cross(b, e1).z;
    if (!(std::max(U, V) <= 0.0f)) continue;
    const F3 Ng = Cross(e1, e0);
    const float den = Ng.z;
    const float rcp = 1.0f / den;
    const float t = rcp * EDot(a, Ng);
    if (!(r.tmin <= t && t <= tfar)) continue;
    if (!(den != 0.0f)) continue;
    float u = U * rcp, v = V * rcp;
    if (!first) { u = 1.0f - u; v = 1.0f - v; }
    const float rad = u * (p1.w - p0.w) + p0.w;
    if (!(t > 2.0f * rad * depth_scale)) continue;   // EMBREE_CURVE_SELF_INTERSECTION_AVOIDANCE_FACTOR = 2
    if (!found || t < bt) {
      found = true;
      bt = t;
      bu = (float(i) + u + 0.0f) * (1.0f / 4.0f);
      bv = 2.0f * v + -1.0f;
    }
  }
  if (found) { *t_out = bt; *u_out = bu; *v_out = bv; }
  return found;
}

}  // namespace

struct pbo_scene {
  std::vector<float> verts, normals, uvs, curve_cp;           // xyzw / xyzw / uv / xyzr
  std::vector<uint32_t> vidx, nidx, tidx, tri_mat, tri_inst, tri_geom, tri_prim;
  std::vector<uint32_t> seg_first, seg_mat, seg_inst, seg_geom, seg_prim;
  std::vector<Material> materials;
  std::vector<float> tex_pixels;
  std::vector<TextureRef> textures;
  // LightManager tables (src/light-manager.h:172-193)
  std::vector<float> light_prob, light_cdf, prim_prob, prim_cdf, prim_area_pdf, prim_emission;
  std::vector<uint32_t> light_off, prim_emissive, prim_tri;
  std::vector<int32_t> tri_light_entry;   // triangle -> entry in the per-light-primitive tables, -1 = not in a light mesh
  std::vector<uint32_t> entry_light;      // per light primitive: its light
  Bvh tri_bvh, seg_bvh;
  float bmin[3] = {0, 0, 0}, bmax[3] = {0, 0, 0};
  bool committed = false;
  std::string error;

  F3 Vert(uint32_t i) const { return F3(verts[4 * i], verts[4 * i + 1], verts[4 * i + 2]); }
  F3 Norm(uint32_t i) const { return F3(normals[4 * i], normals[4 * i + 1], normals[4 * i + 2]); }
  size_t ntris() const { return tri_prim.size(); }
  size_t nsegs() const { return seg_prim.size(); }

  // Scene::TraceFirstHit1 (src/scene.cc:261-264, raytracer_impl.cc:268-278): closest hit over both geometry types
  Hit Trace(const Ray& r) const {
    Hit h;
    float tfar = r.tmax;
    const F3 inv(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    uint32_t stack[128];
    if (!tri_bvh.nodes.empty()) {
      int sp = 0;
      stack[sp++] = 0;
      while (sp) {
        const Node& n = tri_bvh.nodes[stack[--sp]];
        if (!RayBox(r, inv, n.box, tfar)) continue;
        if (n.count) {
          for (uint32_t k = 0; k < n.count; ++k) {
            const uint32_t f = tri_bvh.order[n.first + k];
            float t, u, v;
            F3 ng;
            if (IntersectTriangle(r, tfar, Vert(vidx[3 * f]), Vert(vidx[3 * f + 1]), Vert(vidx[3 * f + 2]), &t, &u, &v, &ng)) {
              tfar = t;
              h.t = t; h.u = u; h.v = v; h.ng = ng;
              h.inst = tri_inst[f]; h.geom = tri_geom[f]; h.prim = tri_prim[f]; h.flat_prim = f; h.curve = false;
            }
          }
        } else {
          stack[sp++] = n.left; stack[sp++] = n.right;
        }
      }
    }
    if (!seg_bvh.nodes.empty()) {
      int sp = 0;
      stack[sp++] = 0;
      while (sp) {
        const Node& n = seg_bvh.nodes[stack[--sp]];
        if (!RayBox(r, inv, n.box, tfar)) continue;
        if (n.count) {
          for (uint32_t k = 0; k < n.count; ++k) {
            const uint32_t sgm = seg_bvh.order[n.first + k];
            const P4* cp = reinterpret_cast<const P4*>(&curve_cp[4 * size_t(seg_first[sgm])]);
            float t, u, v;
            if (IntersectCurve(r, tfar, cp, &t, &u, &v)) {
              tfar = t;
              h.t = t; h.u = u; h.v = v;
              float d[4];
              BezierD(u, d);   // Ng = dB/du (RibbonHit::Ng, curve_intersector_ribbon.h:35)
              h.ng = F3(Mix(d, cp[0].x, cp[1].x, cp[2].x, cp[3].x), Mix(d, cp[0].y, cp[1].y, cp[2].y, cp[3].y),
                        Mix(d, cp[0].z, cp[1].z, cp[2].z, cp[3].z));
              h.inst = seg_inst[sgm]; h.geom = seg_geom[sgm]; h.prim = seg_prim[sgm]; h.flat_prim = sgm; h.curve = true;
            }
          }
        } else {
          stack[sp++] = n.left; stack[sp++] = n.right;
        }
      }
    }
    if (h.valid()) {   // EmbreeRayToTraceResult normalises Ng without a length check (raytracer_impl.cc:213-232)
      const float inv_norm = 1.0f / std::sqrt(h.ng.x * h.ng.x + h.ng.y * h.ng.y + h.ng.z * h.ng.z);
      h.ng = F3(h.ng.x * inv_norm, h.ng.y * inv_norm, h.ng.z * inv_norm);
    }
    return h;
  }
  // Scene::AnyHit1 (src/scene.cc:266-268): any intersection in (tmin, tmax]
  bool Occluded(const Ray& r) const { return Trace(r).valid(); }
};

namespace {

enum Face { kFront = 0, kBack = 1, kAmbiguous = 2 };
struct Surface {   // SurfaceInfo (src/shader/shader-utils.h:18-41)
  F3 P, Ns, Ng;
  float u, v;
  float tex_u = 0.f, tex_v = 0.f;   // Scene::FetchMeshTexcoord (scene.cc:226-249)
  uint32_t inst, mat;
  int light_entry;
  int face;
};

// TraceResultToSufaceInfo (shader-utils.h:131-164) with Scene::FetchMesh* (scene.cc:186-249, triangle-mesh.cc:62-101)
Surface MakeSurface(const pbo_scene& s, const Ray& r, const Hit& h) {
  Surface si;
  si.u = h.u; si.v = h.v;
  si.inst = h.inst;
  si.P = r.o + h.t * r.d;
  si.Ng = h.ng;
  si.light_entry = -1;
  if (h.curve) {
    si.Ns = h.ng;
    si.mat = s.seg_mat[h.flat_prim];
  } else {
    const uint32_t f = h.flat_prim;
    si.mat = s.tri_mat[f];
    si.light_entry = s.tri_light_entry.empty() ? -1 : s.tri_light_entry[f];
    const uint32_t n0 = s.nidx[3 * f], n1 = s.nidx[3 * f + 1], n2 = s.nidx[3 * f + 2];
    if (n0 == 0xFFFFFFFFu || n1 == 0xFFFFFFFFu || n2 == 0xFFFFFFFFu) {
      const F3 p0 = s.Vert(s.vidx[3 * f]), p1 = s.Vert(s.vidx[3 * f + 1]), p2 = s.Vert(s.vidx[3 * f + 2]);
      si.Ns = Normalized(Cross(p1 - p0, p2 - p1));   // CalcGeometryNormal (triangle-mesh.cc:181-184)
    } else {
      si.Ns = Normalized(Lerp3(s.Norm(n0), s.Norm(n1), s.Norm(n2), h.u, h.v));
    }
    // TriangleMesh::FetchTexcoord (triangle-mesh.cc:126-155): barycentrics when a corner has no texcoord
    const uint32_t t0 = s.tidx[3 * f], t1 = s.tidx[3 * f + 1], t2 = s.tidx[3 * f + 2];
    if (t0 == 0xFFFFFFFFu || t1 == 0xFFFFFFFFu || t2 == 0xFFFFFFFFu) {
      si.tex_u = h.u; si.tex_v = h.v;
    } else {
      const float w = 1.0f - h.u - h.v;   // Lerp3 (pbrlab_math.h:35-38)
      si.tex_u = w * s.uvs[2 * t0] + h.u * s.uvs[2 * t1] + h.v * s.uvs[2 * t2];
      si.tex_v = w * s.uvs[2 * t0 + 1] + h.u * s.uvs[2 * t1 + 1] + h.v * s.uvs[2 * t2 + 1];
    }
  }
  const float dg = Dot(r.d, si.Ng), ds = Dot(r.d, si.Ns);
  if (dg < 0.0f && ds < 0.0f) si.face = kFront;
  else if (dg > 0.0f && ds > 0.0f) si.face = kBack;
  else si.face = kAmbiguous;
  return si;
}

struct Frame {   // rows of Rgl: v_local = (ex.v, ey.v, ez.v) (shader-utils.h:66-89, matrix.cc:218-222)
  F3 ex, ey, ez;
  F3 ToLocal(F3 v) const {
    return F3(ex.x * v.x + ex.y * v.y + ex.z * v.z + 0.0f, ey.x * v.x + ey.y * v.y + ey.z * v.z + 0.0f,
              ez.x * v.x + ez.y * v.y + ez.z * v.z + 0.0f);
  }
  F3 ToWorld(F3 v) const {
    return F3(ex.x * v.x + ey.x * v.y + ez.x * v.z + 0.0f, ex.y * v.x + ey.y * v.y + ez.y * v.z + 0.0f,
              ex.z * v.x + ey.z * v.y + ez.z * v.z + 0.0f);
  }
};
void BranchlessONB(F3 n, F3* x, F3* y) {   // shader-utils.h:44-50
  const float sign = copysignf(1.0f, n.z);
  const float a = -1.0f / (sign + n.z);
  const float b = n.x * n.y * a;
  *x = F3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
  *y = F3(b, sign + n.y * n.y * a, -n.y);
}

struct Counters { uint64_t closest = 0, shadow = 0, sss = 0; };

uint32_t LowerBound(const float* cdf, uint32_t n, float u) {   // std::lower_bound, clamped
  const uint32_t i = uint32_t(std::lower_bound(cdf, cdf + n, u) - cdf);
  return i < n ? i : n - 1;
}

// DirectIllumination (shader-utils.h:166-212) with LightManager::SampleAllLight (light-manager.h:79-170)
template <class Eval>
F3 DirectIllumination(const pbo_scene& s, const Surface& si, const Frame& Rgl, F3 normal, Rng& rng, bool hemisphere,
                      const Eval& eval, Counters* cnt) {
  if (s.light_cdf.empty()) return F3(0.f);
  const float u0 = rng.Draw();
  const uint32_t li = LowerBound(s.light_cdf.data(), uint32_t(s.light_cdf.size()), u0);
  const uint32_t off = s.light_off[li], cntp = s.light_off[li + 1] - off;
  const float u1 = rng.Draw();
  const uint32_t pi = off + LowerBound(s.prim_cdf.data() + off, cntp, u1);
  const float u2 = rng.Draw(), u3 = rng.Draw();
  const bool flag = u2 > u3;   // TriangleUniformSampler (sampling-utils.h:59-66)
  const float M = flag ? u2 : u3, m = (!flag) ? u2 : u3;
  const float bu = 1.0f - M, bv = M - m;
  const uint32_t f = s.prim_tri[pi];
  const F3 p0 = s.Vert(s.vidx[3 * f]), p1 = s.Vert(s.vidx[3 * f + 1]), p2 = s.Vert(s.vidx[3 * f + 2]);
  const F3 lpos = Lerp3(p0, p1, p2, bu, bv);
  const F3 lnormal = Normalized(Cross(p1 - p0, p2 - p1));
  const F3 emission(s.prim_emission[3 * pi], s.prim_emission[3 * pi + 1], s.prim_emission[3 * pi + 2]);
  const float pdf_area = s.light_prob[li] * s.prim_prob[pi] * s.prim_area_pdf[pi];

  const F3 dir = Normalized(lpos - si.P);
  const float dist = Length(si.P - lpos);
  const float wl_nl = -Dot(dir, lnormal), wl_np = Dot(dir, normal);
  const float pdf_sigma = std::fabs(pdf_area * dist * dist / (wl_nl * wl_np));
  if ((!hemisphere) || (wl_nl > 0.0f && wl_np > 0.0f)) {
    Ray sr;   // ShadowRay (shader-utils.h:116-129)
    sr.o = si.P; sr.d = dir; sr.tmin = kEps; sr.tmax = std::max(kEps, dist - kEps);
    if (cnt) cnt->shadow++;
    if (!s.Occluded(sr)) {
      F3 f3(0.f);
      float pdf = 0.f;
      eval(Rgl.ToLocal(dir), &f3, &pdf);
      const float w = PowerHeuristicWeight(pdf_sigma, pdf);
      return f3 * emission * w / pdf_sigma;
    }
  }
  return F3(0.f);
}

struct Vertex { F3 wi, throughput, contribute; float pdf; };

// RandomWalkSubsurface (src/shader/random-walk-sss.h:227-405).  On success *si is the exit point, *Rgl the exit frame.
bool RandomWalk(const pbo_scene& s, const Principled& b, Rng& rng, Surface* si, Frame* Rgl, F3* new_wo, F3* thr_out,
                Counters* cnt) {
  if (si->face != kFront) return false;
  F3 dir;
  {
    const float u0 = rng.Draw(), u1 = rng.Draw();
    const F3 tmp = -CosineSampleHemisphere(u0, u1);
    dir = Rgl->ToWorld(tmp);
    if (Dot(-si->Ng, dir) <= 0.0f) return false;
  }
  F3 sigma_t, sigma_s;
  for (int c = 0; c < 3; ++c) {   // ComputeScatteringCoefficientFromAlbedo :111-121
    const float A = b.ss_albedo[c], d = b.ss_radius[c];
    const float a = 1.0f - std::exp(A * (-5.09406f + A * (2.61188f - A * 4.31805f)));
    const float sfit = 1.9f - A + 3.5f * Sqr(A - 0.8f);
    sigma_t.at(c) = 1.0f / std::max(d * sfit, 1e-16f);
    sigma_s.at(c) = sigma_t[c] * a;
  }
  F3 throughput = SafeDivideSpectrum(b.ss_weight, b.ss_albedo);
  Ray ray;
  ray.o = si->P; ray.d = dir; ray.tmin = 1e-3f; ray.tmax = kInf;
  Hit hit;
  bool is_hit = false;
  for (uint32_t bounce = 0; bounce <= 8192; ++bounce) {
    if (bounce > 0) {
      // UniformSampleSphere(rng.Draw(), rng.Draw()) (:296): g++ evaluates the arguments right to left, so the FIRST
      // draw is u2 (asserted against the compiled reference by the golden path vectors)
      const float u2 = rng.Draw(), u1 = rng.Draw();
      ray.d = Normalized(UniformSampleSphere(u1, u2));
      ray.tmin = 0.f;
    }
    F3 cpdf;
    const float ua = rng.Draw(), ub = rng.Draw();
    {   // SampleScatterDistance / SampleChannel :141-187
      const F3 albedo = SafeDivideSpectrum(sigma_s, sigma_t);
      const float w0 = std::fabs(throughput.x * albedo.x), w1 = std::fabs(throughput.y * albedo.y), w2 = std::fabs(throughput.z * albedo.z);
      const float sum = w0 + w1 + w2;
      cpdf = sum > 0.0f ? F3(w0 / sum, w1 / sum, w2 / sum) : F3(1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f);
    }
    const float st = ua < cpdf.x ? sigma_t.x : (ua < cpdf.x + cpdf.y ? sigma_t.y : sigma_t.z);
    const float t_scatter = -logf(1.0f - ub) / st;
    ray.tmax = t_scatter;
    hit = s.Trace(ray);
    if (cnt) cnt->sss++;
    is_hit = hit.valid();
    const float t = is_hit ? hit.t : t_scatter;
    const F3 tr(std::exp(-sigma_t.x * t), std::exp(-sigma_t.y * t), std::exp(-sigma_t.z * t));
    if (is_hit) {
      throughput = throughput * tr / Dot(cpdf, tr);
      break;
    }
    throughput = throughput * (sigma_s * tr) / Dot(cpdf, sigma_t * tr);
    const float p = Saturate(SpectrumNorm(throughput));
    if (rng.Draw() >= p) break;
    throughput = throughput / p;
    ray.o = ray.o + t * ray.d;
  }
  if (!is_hit) return false;
  const uint32_t prev_inst = si->inst;
  *si = MakeSurface(s, ray, hit);
  if (si->inst != prev_inst) return false;
  if (si->face != kBack) return false;
  Rgl->ez = si->Ns;
  BranchlessONB(Rgl->ez, &Rgl->ex, &Rgl->ey);
  *new_wo = Rgl->ToLocal(ray.d);
  *thr_out = throughput;
  return true;
}

// SampleBsdf (cycles-principled-shader.cc:169-242)
void SampleBsdf(const pbo_scene& s, F3 wo, const Principled& b, Rng& rng, Surface* si, Frame* Rgl, F3* wi, F3* f,
                F3* contribute, float* pdf, Counters* cnt) {
  *contribute = F3(0.f);
  const Weights w = SampleWeights(wo, b);
  const float sel = rng.Draw();
  if (sel < w.d) {
    const float u0 = rng.Draw(), u1 = rng.Draw();
    *wi = CosineSampleHemisphere(u0, u1);
  } else if (sel < w.d + w.ss) {
    F3 new_wo, thr;
    if (RandomWalk(s, b, rng, si, Rgl, &new_wo, &thr, cnt)) {
      Principled nb;
      nb.diffuse = true;
      nb.diffuse_w = thr;
      *contribute = DirectIllumination(s, *si, *Rgl, si->Ns, rng, true,
                                       [&](F3 wl, F3* ff, float* pp) { EvalBsdf(wl, new_wo, nb, ff, pp); }, cnt);
      F3 dummy;
      SampleBsdf(s, new_wo, nb, rng, si, Rgl, wi, f, &dummy, pdf, cnt);
      return;
    }
    *wi = F3(0.f); *f = F3(0.f); *pdf = 0.f;
    return;
  } else if (sel < w.d + w.ss + w.sp) {
    const float u0 = rng.Draw(), u1 = rng.Draw();
    GgxSample(wo, b.ax, b.ay, u0, u1, 2, wi);
  } else {
    const float u0 = rng.Draw(), u1 = rng.Draw();
    GgxSample(wo, b.cc_ax, b.cc_ay, u0, u1, 1, wi);
  }
  EvalBsdf(*wi, wo, b, f, pdf);
}

// CyclesPrincipledShader (cycles-principled-shader.cc:414-484)
void PrincipledShader(const pbo_scene& s, F3 wo_world, Rng& rng, Surface* si, Vertex* out, Counters* cnt) {
  if (si->face == kAmbiguous) {
    out->wi = wo_world; out->throughput = F3(0.f); out->contribute = F3(0.f); out->pdf = 0.f;
    return;
  }
  Frame entry;
  entry.ez = (si->face == kFront) ? si->Ns : -si->Ns;
  BranchlessONB(entry.ez, &entry.ex, &entry.ey);
  Frame Rgl = entry;
  const F3 wo = Rgl.ToLocal(wo_world);
  const Material& mp = s.materials[si->mat];
  F3 base(mp.p[0], mp.p[1], mp.p[2]), ss_color(mp.p[7], mp.p[8], mp.p[9]);
  if (mp.tex[0] < s.textures.size()) base = TextureFetch3(s.textures[mp.tex[0]], si->tex_u, si->tex_v);
  if (mp.tex[1] < s.textures.size()) ss_color = TextureFetch3(s.textures[mp.tex[1]], si->tex_u, si->tex_v);
  const Principled b = ParamToBsdf(mp.p, base, ss_color);
  F3 contribute = DirectIllumination(s, *si, Rgl, entry.ez, rng, true,
                                     [&](F3 wl, F3* ff, float* pp) { EvalBsdf(wl, wo, b, ff, pp); }, cnt);
  F3 wi(0.f), f(0.f), c2(0.f);
  float pdf = 0.f;
  SampleBsdf(s, wo, b, rng, si, &Rgl, &wi, &f, &c2, &pdf, cnt);
  contribute = contribute + c2;
  out->wi = entry.ToWorld(wi);   // the ENTRY frame even after the walk moved si (cycles-principled-shader.cc:468-470)
  out->throughput = f * std::fabs(wi.z) / pdf;
  out->pdf = pdf;
  out->contribute = contribute;
  if (!IsFinite(out->throughput) || !std::isfinite(out->pdf)) {
    out->throughput = F3(0.f);
    out->pdf = 0.f;
  }
}

// HairShader (src/shader/hair-shader.cc:153-229)
void HairShader(const pbo_scene& s, F3 wo_world, Rng& rng, Surface* si, Vertex* out, Counters* cnt) {
  if (si->face == kAmbiguous) {
    out->wi = wo_world; out->throughput = F3(0.f); out->contribute = F3(0.f); out->pdf = 0.f;
    return;
  }
  Frame fr;
  fr.ex = si->Ns;
  fr.ey = Normalized(Cross(Cross(wo_world, fr.ex), fr.ex));
  fr.ez = Cross(fr.ex, fr.ey);
  const F3 wo = fr.ToLocal(wo_world);
  const hair::Bsdf b = hair::FromParam(s.materials[si->mat].p, si->v);
  out->contribute = DirectIllumination(s, *si, fr, fr.ex, rng, false,
                                       [&](F3 wl, F3* ff, float* pp) {
                                         const F3 fc = hair::EvalCosPdf(wl, wo, b, pp);
                                         *ff = fc / std::fabs(wl.x);
                                       }, cnt);
  float us[4];
  us[0] = rng.Draw(); us[1] = rng.Draw(); us[2] = rng.Draw(); us[3] = rng.Draw();
  F3 wi(0.f);
  float pdf = 0.f;
  const F3 fc = hair::Sample(wo, b, us, &wi, &pdf);
  out->wi = fr.ToWorld(wi);
  out->throughput = fc / pdf;
  out->pdf = pdf;
  if (!IsFinite(out->throughput) || !std::isfinite(out->pdf)) {
    out->throughput = F3(0.f);
    out->pdf = 0.f;
  }
}

// GetRadiance (src/render.cc:24-90)
F3 GetRadiance(const pbo_scene& s, Ray ray, Rng& rng, Counters* cnt) {
  F3 L(0.0f), thr(1.0f);
  float bsdf_pdf = 0.f;
  for (uint32_t depth = 0;; depth++) {
    if (IsBlack(thr)) break;
    const Hit h = s.Trace(ray);
    if (cnt) cnt->closest++;
    if (!h.valid()) break;
    Surface si = MakeSurface(s, ray, h);
    if (si.face == kFront && si.light_entry >= 0 && s.prim_emissive[si.light_entry]) {   // ImplicitAreaLight (light-manager.h:37-74)
      const int e = si.light_entry;
      const F3 emission(s.prim_emission[3 * e], s.prim_emission[3 * e + 1], s.prim_emission[3 * e + 2]);
      const float pdf_area = s.light_prob[s.entry_light[e]] * s.prim_prob[e] * s.prim_area_pdf[e];
      const float a2s = std::fabs((h.t * h.t) / Dot(si.Ns, ray.d));
      const float w = (depth == 0) ? 1.0f : PowerHeuristicWeight(bsdf_pdf, pdf_area * a2s);
      L = L + w * emission * thr;
    }
    const float rr = SpectrumNorm(thr);
    if (rr < rng.Draw()) break;
    thr = thr * F3(1.0f / rr);
    Vertex v;
    v.contribute = F3(0.f);
    const uint32_t mat = si.mat;
    if (mat == 0xFFFFFFFFu || mat >= s.materials.size()) {   // Shader(): no material -> absorbed (shader.cc:11-17)
      v.wi = -ray.d; v.throughput = F3(0.f); v.pdf = 0.f;
    } else if (s.materials[mat].type == 0) {
      PrincipledShader(s, -ray.d, rng, &si, &v, cnt);
    } else {
      HairShader(s, -ray.d, rng, &si, &v, cnt);
    }
    L = L + thr * v.contribute;
    thr = v.throughput * thr;
    bsdf_pdf = v.pdf;
    ray.o = si.P;   // after a successful walk this is the exit point
    ray.d = v.wi;
    ray.tmin = 1e-3f;
    ray.tmax = kInf;
  }
  return L;
}

Box TriBox(const pbo_scene& s, uint32_t f) {
  Box b;
  for (int k = 0; k < 3; ++k) { b.lo[k] = 3e38f; b.hi[k] = -3e38f; }
  for (int c = 0; c < 3; ++c) {
    const uint32_t v = s.vidx[3 * f + c];
    for (int k = 0; k < 3; ++k) {
      b.lo[k] = std::min(b.lo[k], s.verts[4 * v + k]);
      b.hi[k] = std::max(b.hi[k], s.verts[4 * v + k]);
    }
  }
  return b;
}
// conservative box of a flat curve segment: convex hull of the control points grown by the largest radius
Box SegBox(const pbo_scene& s, uint32_t g) {
  Box b;
  for (int k = 0; k < 3; ++k) { b.lo[k] = 3e38f; b.hi[k] = -3e38f; }
  float r = 0.f;
  for (int c = 0; c < 4; ++c) {
    const float* p = &s.curve_cp[4 * (size_t(s.seg_first[g]) + c)];
    r = std::max(r, std::fabs(p[3]));
    for (int k = 0; k < 3; ++k) { b.lo[k] = std::min(b.lo[k], p[k]); b.hi[k] = std::max(b.hi[k], p[k]); }
  }
  for (int k = 0; k < 3; ++k) { b.lo[k] -= r; b.hi[k] += r; }
  return b;
}

}  // namespace

static Ray RayFrom(const float* r) {
  Ray ray;
  ray.o = F3(r[0], r[1], r[2]); ray.tmin = r[3];
  ray.d = F3(r[4], r[5], r[6]); ray.tmax = r[7];
  return ray;
}

template <class Fn>
static void ParallelFor(uint64_t n, int threads, const Fn& fn) {
  int nt = threads > 0 ? threads : int(std::max(1u, std::thread::hardware_concurrency()));
  if (n < 4096) nt = 1;
  std::atomic<uint64_t> next(0);
  auto work = [&](int tid) {
    for (;;) {
      const uint64_t b = next.fetch_add(1024);
      if (b >= n) break;
      fn(b, std::min<uint64_t>(n, b + 1024), tid);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
}

extern "C" {

pbo_scene* pbo_create(void) { return new pbo_scene(); }
void pbo_destroy(pbo_scene* s) { delete s; }
const char* pbo_last_error(const pbo_scene* s) { return s ? s->error.c_str() : "null scene"; }

int pbo_set_triangles(pbo_scene* s, const float* xyzw, uint32_t nverts, const uint32_t* vidx, const float* nxyzw,
                      uint32_t nnormals, const uint32_t* nidx, const float* uv, uint32_t nuv, const uint32_t* tidx,
                      const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                      const uint32_t* prim_id, uint64_t ntris) {
  if (!s) return 1;
  s->committed = false;
  s->verts.assign(xyzw, xyzw + size_t(nverts) * 4);
  s->normals.clear();
  if (nxyzw) s->normals.assign(nxyzw, nxyzw + size_t(nnormals) * 4);
  s->uvs.clear();
  if (uv) s->uvs.assign(uv, uv + size_t(nuv) * 2);
  s->vidx.assign(vidx, vidx + ntris * 3);
  s->nidx.assign(ntris * 3, 0xFFFFFFFFu);
  if (nidx) s->nidx.assign(nidx, nidx + ntris * 3);
  s->tidx.assign(ntris * 3, 0xFFFFFFFFu);
  if (tidx) s->tidx.assign(tidx, tidx + ntris * 3);
  s->tri_mat.assign(material_id, material_id + ntris);
  s->tri_inst.assign(instance_id, instance_id + ntris);
  s->tri_geom.assign(geom_id, geom_id + ntris);
  s->tri_prim.assign(prim_id, prim_id + ntris);
  for (uint32_t v : s->vidx) if (v >= nverts) { s->error = "vertex index out of range"; return 1; }
  for (uint32_t v : s->nidx) if (v != 0xFFFFFFFFu && v >= nnormals) { s->error = "normal index out of range"; return 1; }
  return 0;
}

int pbo_set_curves(pbo_scene* s, const float* xyzr, uint32_t nverts, const uint32_t* first_cp,
                   const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                   const uint32_t* prim_id, uint64_t nsegs) {
  if (!s) return 1;
  s->committed = false;
  s->curve_cp.clear();
  if (xyzr) s->curve_cp.assign(xyzr, xyzr + size_t(nverts) * 4);
  s->seg_first.assign(first_cp, first_cp + nsegs);
  s->seg_mat.assign(material_id, material_id + nsegs);
  s->seg_inst.assign(instance_id, instance_id + nsegs);
  s->seg_geom.assign(geom_id, geom_id + nsegs);
  s->seg_prim.assign(prim_id, prim_id + nsegs);
  for (uint32_t f : s->seg_first) if (uint64_t(f) + 4 > nverts) { s->error = "control point index out of range"; return 1; }
  return 0;
}

int pbo_set_materials(pbo_scene* s, const void* materials, uint32_t n) {
  if (!s) return 1;
  s->materials.resize(n);
  if (n) memcpy(s->materials.data(), materials, sizeof(Material) * n);
  return 0;
}

int pbo_set_textures(pbo_scene* s, const pbo_texture* t, uint32_t n) {
  if (!s || (n && !t)) return 1;
  size_t total = 0;
  for (uint32_t i = 0; i < n; ++i) total += size_t(t[i].width) * t[i].height * t[i].channels;
  s->tex_pixels.resize(total);
  s->textures.resize(n);
  size_t off = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const size_t cnt = size_t(t[i].width) * t[i].height * t[i].channels;
    memcpy(s->tex_pixels.data() + off, t[i].pixels, cnt * sizeof(float));
    s->textures[i] = {s->tex_pixels.data() + off, t[i].width, t[i].height, t[i].channels};
    off += cnt;
  }
  return 0;
}

int pbo_set_lights(pbo_scene* s, const pbo_light_tables* t) {
  if (!s || !t) return 1;
  s->committed = false;
  const uint32_t nl = t->num_lights, np = t->num_light_prims;
  s->light_prob.assign(t->light_probability, t->light_probability + nl);
  s->light_cdf.assign(t->light_cdf, t->light_cdf + nl);
  s->light_off.clear();
  if (nl) s->light_off.assign(t->light_prim_offset, t->light_prim_offset + nl + 1);
  s->prim_prob.assign(t->prim_probability, t->prim_probability + np);
  s->prim_cdf.assign(t->prim_cdf, t->prim_cdf + np);
  s->prim_area_pdf.assign(t->prim_area_pdf, t->prim_area_pdf + np);
  s->prim_emission.assign(t->prim_emission, t->prim_emission + size_t(np) * 3);
  s->prim_emissive.assign(t->prim_is_emissive, t->prim_is_emissive + np);
  s->prim_tri.assign(t->prim_triangle, t->prim_triangle + np);
  return 0;
}

int pbo_commit(pbo_scene* s, const float* bmin, const float* bmax) {
  if (!s) return 1;
  std::vector<Box> tb(s->ntris()), sb(s->nsegs());
  for (size_t i = 0; i < tb.size(); ++i) tb[i] = TriBox(*s, uint32_t(i));
  for (size_t i = 0; i < sb.size(); ++i) sb[i] = SegBox(*s, uint32_t(i));
  s->tri_bvh.Build(tb);
  s->seg_bvh.Build(sb);
  s->tri_light_entry.assign(s->ntris(), -1);
  s->entry_light.assign(s->prim_tri.size(), 0);
  for (size_t l = 0; l + 1 < s->light_off.size(); ++l)
    for (uint32_t e = s->light_off[l]; e < s->light_off[l + 1]; ++e) {
      s->entry_light[e] = uint32_t(l);
      if (s->prim_tri[e] >= s->ntris()) { s->error = "light primitive out of range"; return 1; }
      s->tri_light_entry[s->prim_tri[e]] = int32_t(e);
    }
  if (bmin && bmax) {
    for (int k = 0; k < 3; ++k) { s->bmin[k] = bmin[k]; s->bmax[k] = bmax[k]; }
  } else {
    for (int k = 0; k < 3; ++k) { s->bmin[k] = 3e38f; s->bmax[k] = -3e38f; }
    for (const Box& b : tb) for (int k = 0; k < 3; ++k) { s->bmin[k] = std::min(s->bmin[k], b.lo[k]); s->bmax[k] = std::max(s->bmax[k], b.hi[k]); }
    for (const Box& b : sb) for (int k = 0; k < 3; ++k) { s->bmin[k] = std::min(s->bmin[k], b.lo[k]); s->bmax[k] = std::max(s->bmax[k], b.hi[k]); }
  }
  s->committed = true;
  return 0;
}

int pbo_scene_bounds(const pbo_scene* s, float* bmin, float* bmax) {
  if (!s || !s->committed) return 1;
  for (int k = 0; k < 3; ++k) { bmin[k] = s->bmin[k]; bmax[k] = s->bmax[k]; }
  return 0;
}

int pbo_trace(pbo_scene* s, const float* rays, uint64_t n, void* hits) {
  if (!s || !s->committed) return 1;
  struct Out { float ng[3], t, u, v; uint32_t inst, geom, prim; };
  Out* out = static_cast<Out*>(hits);
  ParallelFor(n, 0, [&](uint64_t b, uint64_t e, int) {
    for (uint64_t i = b; i < e; ++i) {
      const Hit h = s->Trace(RayFrom(rays + 8 * i));
      Out o = {{1.f, 0.f, 0.f}, 1.0f, 0.f, 0.f, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
      if (h.valid()) { o.ng[0] = h.ng.x; o.ng[1] = h.ng.y; o.ng[2] = h.ng.z; o.t = h.t; o.u = h.u; o.v = h.v; o.inst = h.inst; o.geom = h.geom; o.prim = h.prim; }
      out[i] = o;
    }
  });
  return 0;
}

int pbo_occluded(pbo_scene* s, const float* rays, uint64_t n, uint8_t* occluded) {
  if (!s || !s->committed) return 1;
  ParallelFor(n, 0, [&](uint64_t b, uint64_t e, int) {
    for (uint64_t i = b; i < e; ++i) occluded[i] = s->Occluded(RayFrom(rays + 8 * i)) ? 1 : 0;
  });
  return 0;
}

int pbo_radiance(pbo_scene* s, const float* rays, const uint64_t* seeds, uint64_t n, float* radiance_out,
                 uint64_t* counts3) {
  if (!s || !s->committed) return 1;
  std::atomic<uint64_t> c0(0), c1(0), c2(0);
  ParallelFor(n, 0, [&](uint64_t b, uint64_t e, int) {
    Counters cnt;
    for (uint64_t i = b; i < e; ++i) {
      Rng rng(seeds[2 * i], seeds[2 * i + 1]);
      const F3 L = GetRadiance(*s, RayFrom(rays + 8 * i), rng, &cnt);
      radiance_out[3 * i] = L.x; radiance_out[3 * i + 1] = L.y; radiance_out[3 * i + 2] = L.z;
    }
    c0 += cnt.closest; c1 += cnt.shadow; c2 += cnt.sss;
  });
  if (counts3) { counts3[0] = c0; counts3[1] = c1; counts3[2] = c2; }
  return 0;
}

double pbo_render(pbo_scene* s, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed, float* rgba,
                  uint32_t* count, int threads, uint64_t* counts3) {
  if (!s || !s->committed || !width || !height) return -1.0;
  const auto t0 = std::chrono::steady_clock::now();
  // camera of RenderingTile (src/render.cc:132-158)
  const float* bmin = s->bmin;
  const float* bmax = s->bmax;
  float hs, vs;
  if (bmax[0] - bmin[0] > bmax[1] - bmin[1]) {
    hs = bmax[0] - bmin[0];
    vs = hs * float(height) / float(width);
  } else {
    vs = bmax[1] - bmin[1];
    hs = vs * float(width) / float(height);
  }
  const F3 eye((bmax[0] + bmin[0]) * 0.5f, (bmax[1] + bmin[1]) * 0.5f, bmax[2] + hs * 0.5f * sqrtf(3.f));
  const float x_corner = (bmax[0] + bmin[0]) * 0.5f - hs * 0.5f;
  const float y_corner = (bmax[1] + bmin[1]) * 0.5f + vs * 0.5f;
  const float z_corner = bmax[2];
  const float dx = hs / float(width), dy = vs / float(height);
  const uint64_t npix = uint64_t(width) * height;
  std::atomic<uint64_t> c0(0), c1(0), c2(0);
  ParallelFor(npix, threads, [&](uint64_t b, uint64_t e, int) {
    Counters cnt;
    for (uint64_t p = b; p < e; ++p) {
      const uint32_t x = uint32_t(p % width), y = uint32_t(p / width);
      float acc[3] = {0, 0, 0};
      for (uint32_t k = 0; k < spp; ++k) {
        Rng rng(seed + k, p);
        const float jx = rng.Draw(), jy = rng.Draw();   // render.cc:160-171
        float d[3] = {x_corner + dx * (float(x) + jx) - eye.x, y_corner - dy * (float(y) + jy) - eye.y, z_corner - eye.z};
        const float inv_norm = 1.0f / std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);   // Normalize :243-249
        Ray ray;
        ray.o = eye; ray.d = F3(d[0] * inv_norm, d[1] * inv_norm, d[2] * inv_norm); ray.tmin = 0.f; ray.tmax = kInf;
        const F3 L = GetRadiance(*s, ray, rng, &cnt);
        acc[0] += L.x; acc[1] += L.y; acc[2] += L.z;   // render.cc:175-183
      }
      rgba[4 * p] = acc[0]; rgba[4 * p + 1] = acc[1]; rgba[4 * p + 2] = acc[2]; rgba[4 * p + 3] = float(spp);
      count[p] = spp;
    }
    c0 += cnt.closest; c1 += cnt.shadow; c2 += cnt.sss;
  });
  if (counts3) { counts3[0] = c0; counts3[1] = c1; counts3[2] = c2; }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
