// Sampling warps and MIS weight (src/sampler/sampling-utils.h:10-66), Lambert (src/closure/lambert.h:11-27),
// dielectric Fresnel (src/closure/closure-util.h:10-29) and the local shading frame helpers
// (src/shader/shader-utils.h:44-114, src/matrix.cc:218-222).
#pragma once
#include "common.cuh"

namespace pbr {

// full-precision sinf/cosf/sqrtf: the reference uses std::cos/std::sin/std::sqrt (sampling-utils.h:10-14)
PBR_HD vec3 CosineSampleHemisphere(float u1, float u2) {
  const float u1_ = u1 * 2.0f * kPi, u3 = sqrtf(u2);
  return vec3(cosf(u1_) * u3, sinf(u1_) * u3, sqrtf(fmaxf_(1.0f - u2, 0.0f)));
}

PBR_HD vec3 UniformSampleSphere(float u1, float u2) {   // sampling-utils.h:16-23 (Y up)
  const float u = 2.0f * u2 - 1.0f;
  const float norm = sqrtf(fmaxf_(0.0f, 1.0f - u * u));
  const float theta = 2.0f * kPi * u1;
  return vec3(norm * cosf(theta), u, norm * sinf(theta));
}

PBR_HD float PowerHeuristicWeight(float sampled_pdf, float other_pdf) {   // sampling-utils.h:27-57
  float r, mis;
  if (sampled_pdf > other_pdf) {
    r = other_pdf / sampled_pdf;
    mis = 1 / (1 + r * r);
  } else if (sampled_pdf < other_pdf) {
    r = sampled_pdf / other_pdf;
    mis = 1 - 1 / (1 + r * r);
  } else {
    mis = 0.5f;  // equal or unordered (NaN)
  }
  return mis;
}

// returns (u, v) = (1 - max, max - min) (sampling-utils.h:59-66)
PBR_HD void TriangleUniformSampler(float u1, float u2, float* u, float* v) {
  const bool flag = (u1 > u2);
  const float M = flag ? u1 : u2;
  const float m = (!flag) ? u1 : u2;
  *u = 1.0f - M;
  *v = M - m;
}

// ---- Lambert: f = 1/pi, pdf = wi.z/pi, deliberately not clamped (closure/lambert.h:11-27)
PBR_HD float LambertPdf(const vec3& omega_in) { return omega_in.z * kPiInv; }

// ---- FresnelDielectricCos (closure/closure-util.h:10-29)
PBR_HD float FresnelDielectricCos(float cos_, float eta) {
  if (fabsf(eta) < kFltEps) return 1.0f;
  if (cos_ < 0.0f) eta = 1.0f / eta;
  const float c = fabsf(cos_);
  float g = eta * eta - 1 + c * c;
  if (g > 0) {
    g = sqrtf(g);
    const float A = (g - c) / (g + c);
    const float B = (c * (g + c) - 1) / (c * (g - c) + 1);
    return 0.5f * A * A * (1 + B * B);
  }
  return 1.0f;
}

// ---- Pixar branchless ONB (shader-utils.h:44-50)
PBR_HD void BranchlessONB(const vec3& n, vec3* x, vec3* y) {
  const float sign = copysignf(1.0f, n.z);
  const float a = -1.0f / (sign + n.z);
  const float b = n.x * n.y * a;
  *x = vec3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
  *y = vec3(b, sign + n.y * n.y * a, -n.y);
}

// A shading frame.  ToLocal is Matrix::MultV with GrobalToShadingLocal's matrix (rows = world axes, columns =
// ex,ey,ez; the translation row is zero): dst_j = e_j.x*v.x + e_j.y*v.y + e_j.z*v.z + 0, summed in that order.
// ToWorld is Matrix::MultV with ShadingLocalToGlobal's matrix: dst = ex*v.x + ey*v.y + ez*v.z + 0 per component.
struct Frame {
  vec3 ex, ey, ez;
  PBR_HD vec3 ToLocal(const vec3& v) const {
    return vec3(ex.x * v.x + ex.y * v.y + ex.z * v.z + 0.0f, ey.x * v.x + ey.y * v.y + ey.z * v.z + 0.0f,
                ez.x * v.x + ez.y * v.y + ez.z * v.z + 0.0f);
  }
  PBR_HD vec3 ToWorld(const vec3& v) const {
    return vec3(ex.x * v.x + ey.x * v.y + ez.x * v.z + 0.0f, ex.y * v.x + ey.y * v.y + ez.y * v.z + 0.0f,
                ex.z * v.x + ey.z * v.y + ez.z * v.z + 0.0f);
  }
};

}  // namespace pbr
