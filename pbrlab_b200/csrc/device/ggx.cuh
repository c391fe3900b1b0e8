// GGX (GTR2) reflection and GTR1 clearcoat: eval, pdf and visible-normal slope sampling, following
// src/closure/microfacet-ggx.h:48-286 (itself the Cycles/OSL flavour of Heitz & d'Eon 2014).
// Quirks kept on purpose: pdf = D*G1o/(4 cos_o cos_i) without the |m.o| Jacobian term;
// clearcoat (distrib == 1) evaluates G with a fixed alpha^2 = 0.0625 and carries an extra 0.25; the sampler leaves
// omega_in untouched when m.o <= 0.
#pragma once
#include "common.cuh"

namespace pbr {

PBR_HD float D_GTR1(const vec3& h, float alpha) {                       // microfacet-ggx.h:48-53
  if (alpha >= 1.0f) return 1.0f / kPi;
  const float alpha2 = alpha * alpha;
  const float t = 1.0f + (alpha2 - 1.0f) * h.z * h.z;
  return (alpha2 - 1.0f) / (kPi * logf(alpha2) * t);
}

PBR_HD float D_GTR2(const vec3& h, float alpha2) {                      // microfacet-ggx.h:55-63
  const float cos_theta_m = h.z;
  const float cos_theta_m2 = cos_theta_m * cos_theta_m;
  const float cos_theta_m4 = cos_theta_m2 * cos_theta_m2;
  const float tan_theta_m2 = (1.0f - cos_theta_m2) / cos_theta_m2;
  return alpha2 / (kPi * cos_theta_m4 * (alpha2 + tan_theta_m2) * (alpha2 + tan_theta_m2));
}

PBR_HD void MicrofacetGgxSampleSlopes(float cos_theta_i, float sin_theta_i, float randu, float randv, float* slope_x,
                                      float* slope_y, float* G1i) {       // microfacet-ggx.h:65-118
  const float k2PI = 2.0f * kPi;
  if (cos_theta_i >= 0.99999f) {
    const float r = sqrtf(randu / (1.0f - randu));
    const float phi = k2PI * randv;
    *slope_x = r * cosf(phi);
    *slope_y = r * sinf(phi);
    *G1i = 1.0f;
    return;
  }
  const float tan_theta_i = sin_theta_i / cos_theta_i;
  const float G1_inv = 0.5f * (1.0f + SafeSqrtf(1.0f + tan_theta_i * tan_theta_i));
  *G1i = 1.0f / G1_inv;

  const float A = 2.0f * randu * G1_inv - 1.0f;
  const float AA = A * A;
  const float tmp = 1.0f / (AA - 1.0f);
  const float B = tan_theta_i;
  const float BB = B * B;
  const float D = SafeSqrtf(BB * (tmp * tmp) - (AA - BB) * tmp);
  const float slope_x_1 = B * tmp - D;
  const float slope_x_2 = B * tmp + D;
  *slope_x = (A < 0.0f || slope_x_2 * tan_theta_i > 1.0f) ? slope_x_1 : slope_x_2;

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

PBR_HD vec3 MicrofacetSampleStretched(const vec3& omega_i, float alpha_x, float alpha_y, float randu, float randv,
                                      float* G1i) {                          // microfacet-ggx.h:121-162
  const vec3 omega_i_ = vnormalized(vec3(alpha_x * omega_i.x, alpha_y * omega_i.y, omega_i.z));
  float costheta_ = 1.0f, sintheta_ = 0.0f, cosphi_ = 1.0f, sinphi_ = 0.0f;
  if (omega_i_.z < 0.99999f) {
    costheta_ = omega_i_.z;
    sintheta_ = SafeSqrtf(1.0f - costheta_ * costheta_);
    const float invlen = 1.0f / sintheta_;
    cosphi_ = omega_i_.x * invlen;
    sinphi_ = omega_i_.y * invlen;
  }
  float slope_x = 0.f, slope_y = 0.f;
  MicrofacetGgxSampleSlopes(costheta_, sintheta_, randu, randv, &slope_x, &slope_y, G1i);
  const float tmp = cosphi_ * slope_x - sinphi_ * slope_y;
  slope_y = sinphi_ * slope_x + cosphi_ * slope_y;
  slope_x = tmp;
  slope_x = alpha_x * slope_x;
  slope_y = alpha_y * slope_y;
  return vnormalized(vec3(-slope_x, -slope_y, 1.0f));
}

// returns f, writes pdf (microfacet-ggx.h:164-245). distrib: 2 = GTR2 (specular), 1 = GTR1 (clearcoat)
PBR_HD float MicrofacetGGXBsdfPdf(const vec3& omega_in, const vec3& omega_out, float alpha_x, float alpha_y,
                                  int distrib, float* pdf) {
  const float cos_n_o = omega_out.z;
  const float cos_n_i = omega_in.z;
  if (cos_n_o > 0 && cos_n_i > 0) {
    const vec3 m = vnormalized(omega_in + omega_out);
    float alpha2 = alpha_x * alpha_y;
    float D, G1o, G1i;
    if (fabsf(alpha_x - alpha_y) < kFltEps) {
      if (distrib == 1) {
        D = D_GTR1(m, alpha_x);
        alpha2 = 0.0625f;
      } else {
        D = D_GTR2(m, alpha2);
      }
      G1o = 2 / (1 + SafeSqrtf(1 + alpha2 * (1 - cos_n_o * cos_n_o) / (cos_n_o * cos_n_o)));
      G1i = 2 / (1 + SafeSqrtf(1 + alpha2 * (1 - cos_n_i * cos_n_i) / (cos_n_i * cos_n_i)));
    } else {
      const float slope_x = -m.x / (m.z * alpha_x);
      const float slope_y = -m.y / (m.z * alpha_y);
      const float slope_len = 1 + slope_x * slope_x + slope_y * slope_y;
      const float cosThetaM = m.z;
      const float cosThetaM2 = cosThetaM * cosThetaM;
      const float cosThetaM4 = cosThetaM2 * cosThetaM2;
      D = 1.f / ((slope_len * slope_len) * kPi * alpha2 * cosThetaM4);

      const float tanThetaO2 = (1.f - cos_n_o * cos_n_o) / (cos_n_o * cos_n_o);
      const float cosPhiO = omega_out.x;
      const float sinPhiO = omega_out.y;
      float alphaO2 = (cosPhiO * cosPhiO) * (alpha_x * alpha_x) + (sinPhiO * sinPhiO) * (alpha_y * alpha_y);
      alphaO2 /= cosPhiO * cosPhiO + sinPhiO * sinPhiO;
      G1o = 2 / (1 + SafeSqrtf(1 + alphaO2 * tanThetaO2));

      const float tanThetaI2 = (1 - cos_n_i * cos_n_i) / (cos_n_i * cos_n_i);
      const float cosPhiI = omega_in.x;
      const float sinPhiI = omega_in.y;
      float alphaI2 = (cosPhiI * cosPhiI) * (alpha_x * alpha_x) + (sinPhiI * sinPhiI) * (alpha_y * alpha_y);
      alphaI2 /= cosPhiI * cosPhiI + sinPhiI * sinPhiI;
      G1i = 2 / (1 + SafeSqrtf(1 + alphaI2 * tanThetaI2));
    }
    const float G = G1o * G1i;
    const float common = D * 0.25f / cos_n_o / cos_n_i;
    float bsdf_f = G * common;
    if (distrib == 1) bsdf_f = 0.25f * bsdf_f;
    *pdf = G1o * common;
    return bsdf_f;
  }
  *pdf = 0.f;
  return 0.f;
}

// omega_in is in/out: left untouched (caller passes 0) when the sampled normal faces away (microfacet-ggx.h:247-286)
PBR_HD float MicrofacetGGXSample(const vec3& omega_out, float alpha_x, float alpha_y, float u0, float u1, int distrib,
                                 vec3* omega_in, float* pdf) {
  const float cos_n_o = omega_out.z;
  float ret = 0.f;
  if (cos_n_o > 0.f) {
    float G1o = 0.f;
    const vec3 m = MicrofacetSampleStretched(omega_out, alpha_x, alpha_y, u0, u1, &G1o);
    const float cos_m_o = vdot(m, omega_out);
    if (cos_m_o > 0) {
      *omega_in = 2 * cos_m_o * m - omega_out;
      ret = MicrofacetGGXBsdfPdf(*omega_in, omega_out, alpha_x, alpha_y, distrib, pdf);
    }
  }
  return ret;
}

}  // namespace pbr
