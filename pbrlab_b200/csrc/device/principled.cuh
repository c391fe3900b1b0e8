// Cycles-style PrincipledBSDF as pbrlab implements it (src/shader/cycles-principled-shader.cc:20-412):
// material parameters -> closure set (diffuse | subsurface, GGX specular, GTR1 clearcoat), closure sample weights,
// eval + pdf.  Sheen, transmission, anisotropic rotation and diffuse roughness are read but unused there, hence
// absent here.  The random-walk set-up it calls (BssrdfSetup & co., src/shader/random-walk-sss.h:35-104) is included.
#pragma once
#include "common.cuh"
#include "ggx.cuh"
#include "sampling.cuh"

namespace pbr {

// POD mirror of CyclesPrincipledBsdfParameter (src/material-param.h:24-49); layout = include/pbrgpu.h
struct PrincipledParams {
  float base_color[3];
  float subsurface;
  float subsurface_radius[3];
  float subsurface_color[3];
  float metallic, specular, specular_tint, roughness, anisotropic, anisotropic_rotation, sheen, sheen_tint;
  float clearcoat, clearcoat_roughness, ior, transmission, transmission_roughness;
  uint32_t base_color_tex_id, subsurface_color_tex_id;
};

struct PrincipledBsdf {   // cycles-principled-shader.cc:20-45
  bool enable_diffuse, enable_subsurface, enable_specular, enable_clearcoat;
  vec3 diffuse_weight;
  vec3 subsurface_weight, subsurface_albedo, subsurface_radius;
  vec3 specular_weight;
  float alpha_x, alpha_y, ior;
  vec3 specular_color;
  vec3 clearcoat_weight;
  float clearcoat_alpha_x, clearcoat_alpha_y, clearcoat_ior;
  vec3 clearcoat_color;
};

PBR_HD void InitBsdf(PrincipledBsdf* b) {
  b->enable_diffuse = b->enable_subsurface = b->enable_specular = b->enable_clearcoat = false;
  b->diffuse_weight = b->subsurface_weight = b->subsurface_albedo = b->subsurface_radius = vec3(0.f);
  b->specular_weight = b->specular_color = b->clearcoat_weight = b->clearcoat_color = vec3(0.f);
  b->alpha_x = b->alpha_y = 1.f; b->ior = 1.5f;
  b->clearcoat_alpha_x = b->clearcoat_alpha_y = 1.f; b->clearcoat_ior = 1.5f;
}

// ---- random-walk-sss.h:40-104
PBR_HD float BssrdfBurleyFitting(float A) { return 1.9f - A + 3.5f * (A - 0.8f) * (A - 0.8f); }
PBR_HD float BssrdfBurleyFitting5(float A) { return 1.85f - A + 7.0f * fabsf((A - 0.8f) * (A - 0.8f) * (A - 0.8f)); }

PBR_HD void BssrdfSetup(bool burley_radius, bool scale_mfp, bool use_eq5, vec3* weight, vec3* albedo, vec3* radius,
                        vec3* diffuse_weight) {
  *diffuse_weight = vec3(0.f);
  const float kBssrdfMinRadius = 1e-8f;
  float kd[3] = {0.f, 0.f, 0.f};
  float w[3] = {weight->x, weight->y, weight->z};
  float r[3] = {radius->x, radius->y, radius->z};
  int channels = 3;
  for (int i = 0; i < 3; ++i) {
    if (r[i] < kBssrdfMinRadius) {
      kd[i] = w[i];
      w[i] = 0.f;
      r[i] = 0.f;
      channels--;
    }
  }
  *weight = vec3(w[0], w[1], w[2]);
  *radius = vec3(r[0], r[1], r[2]);
  if (channels < 3) *diffuse_weight = vec3(kd[0], kd[1], kd[2]);
  if (channels > 0 && burley_radius) {
    // BssrdfBurleySetup(albedo, radius, scale_mfp, mode = int(use_eq5)): mode != 0 -> equation (5) fit
    const vec3 l = scale_mfp ? (0.25f * (1.0f / kPi)) * (*radius) : *radius;
    const vec3 A = *albedo;
    vec3 s;
    if (!use_eq5) s = vec3(BssrdfBurleyFitting(A.x), BssrdfBurleyFitting(A.y), BssrdfBurleyFitting(A.z));
    else s = vec3(BssrdfBurleyFitting5(A.x), BssrdfBurleyFitting5(A.y), BssrdfBurleyFitting5(A.z));
    *radius = l / s;
  }
}

// ---- ParamToBsdf (cycles-principled-shader.cc:244-412); base/subsurface colours already texture-resolved
PBR_HD PrincipledBsdf ParamToBsdf(const PrincipledParams& m, const vec3& base_color, const vec3& subsurface_color) {
  const vec3 weight(1.f);
  const float kCut = kEps;
  PrincipledBsdf bsdf;
  InitBsdf(&bsdf);

  const float diffuse_w = (1.0f - Saturatef(m.metallic)) * (1.0f - Saturatef(m.transmission));
  const float final_transmission = Saturatef(m.transmission) * (1.0f - Saturatef(m.metallic));
  const float specular_w = (1.0f - final_transmission);
  const float subsurface = m.subsurface;

  {
    const vec3 mixed = subsurface_color * subsurface + base_color * (1.0f - subsurface);
    if (Average(mixed) > kCut) {
      if (subsurface < kCut && diffuse_w > kCut) {
        bsdf.enable_diffuse = true;
        bsdf.diffuse_weight = weight * base_color * diffuse_w;
      } else if (subsurface > kCut) {
        bsdf.enable_subsurface = true;
        bsdf.subsurface_weight = weight * mixed * diffuse_w;
        bsdf.subsurface_albedo = mixed;
        bsdf.subsurface_radius =
            vec3(m.subsurface_radius[0], m.subsurface_radius[1], m.subsurface_radius[2]) * subsurface;
        vec3 add_diffuse(0.f);
        BssrdfSetup(true, true, true, &bsdf.subsurface_weight, &bsdf.subsurface_albedo, &bsdf.subsurface_radius,
                    &add_diffuse);
        if (!IsBlack(add_diffuse)) {
          bsdf.enable_diffuse = true;
          bsdf.diffuse_weight = bsdf.diffuse_weight + add_diffuse;
        }
      }
    }
  }

  if (specular_w > kCut && (m.specular > kCut || m.metallic > kCut)) {
    bsdf.enable_specular = true;
    bsdf.specular_weight = weight * specular_w;
    bsdf.ior = (2.0f / (1.0f - SafeSqrtf(0.08f * m.specular))) - 1.0f;
    const float aspect = SafeSqrtf(1.0f - m.anisotropic * 0.9f);
    const float roughness2 = m.roughness * m.roughness;
    bsdf.alpha_x = roughness2 / aspect;
    bsdf.alpha_y = roughness2 * aspect;
    const float y_base = RgbToY(base_color);
    const vec3 rho_tint = y_base > 0.0f ? base_color / y_base : vec3(0.0f);
    const vec3 rho_specular = Lerp3v(vec3(1.0f), rho_tint, m.specular_tint);
    bsdf.specular_color = Lerp3v(0.08f * m.specular * rho_specular, base_color, m.metallic);
  }

  if (m.clearcoat > kCut) {
    bsdf.enable_clearcoat = true;
    bsdf.clearcoat_weight = vec3(0.25f * m.clearcoat);
    bsdf.clearcoat_alpha_x = m.clearcoat_roughness * m.clearcoat_roughness;
    bsdf.clearcoat_alpha_y = m.clearcoat_roughness * m.clearcoat_roughness;
    bsdf.clearcoat_color = vec3(0.04f);
    bsdf.clearcoat_ior = 1.5f;
  }
  return bsdf;
}

PBR_HD vec3 SpecularColor(const vec3& omega_in, const vec3& omega_out, const vec3& specular_color, float ior) {
  const vec3 h = vnormalized(omega_in + omega_out);                 // cycles-principled-shader.cc:54-61
  const float f0 = FresnelDielectricCos(1.0f, ior);
  const float fh = (FresnelDielectricCos(vdot(h, omega_out), ior) - f0) / (1.0f - f0);
  return (specular_color) * (1.f - fh) + vec3(fh);
}

struct SampleWeight { float diffuse, subsurface, specular, clearcoat; };

// luma-normalised closure selection weights, NaN/inf -> 0 (cycles-principled-shader.cc:63-112)
PBR_HD SampleWeight FetchClosureSampleWeight(const vec3& omega_out, const PrincipledBsdf& bsdf) {
  SampleWeight w;
  const vec3 mirror(-omega_out.x, -omega_out.y, omega_out.z);
  w.diffuse = bsdf.enable_diffuse ? RgbToY(bsdf.diffuse_weight) : 0.f;
  w.subsurface = bsdf.enable_subsurface ? RgbToY(bsdf.subsurface_weight) : 0.f;
  w.specular = bsdf.enable_specular
                   ? RgbToY(bsdf.specular_weight * SpecularColor(mirror, omega_out, bsdf.specular_color, bsdf.ior))
                   : 0.f;
  w.clearcoat = bsdf.enable_clearcoat ? RgbToY(bsdf.clearcoat_weight * SpecularColor(mirror, omega_out,
                                                                                     bsdf.clearcoat_color,
                                                                                     bsdf.clearcoat_ior))
                                      : 0.f;
  float sum = 0.0f;
  sum += w.diffuse;
  sum += w.subsurface;
  sum += w.specular;
  sum += w.clearcoat;
  w.diffuse /= sum;
  w.subsurface /= sum;
  w.specular /= sum;
  w.clearcoat /= sum;
  if (!finitef_(w.diffuse)) w.diffuse = 0.f;
  if (!finitef_(w.subsurface)) w.subsurface = 0.f;
  if (!finitef_(w.specular)) w.specular = 0.f;
  if (!finitef_(w.clearcoat)) w.clearcoat = 0.f;
  return w;
}

// f = sum over enabled closures, pdf = sum w_i pdf_i (cycles-principled-shader.cc:114-155)
PBR_HD void EvalBsdf(const vec3& omega_in, const vec3& omega_out, const PrincipledBsdf& bsdf, vec3* bsdf_f,
                     float* pdf) {
  const SampleWeight w = FetchClosureSampleWeight(omega_out, bsdf);
  *bsdf_f = vec3(0.0f);
  *pdf = 0.0f;
  if (bsdf.enable_diffuse) {
    const float p = LambertPdf(omega_in);
    *bsdf_f = *bsdf_f + bsdf.diffuse_weight * kPiInv;
    *pdf += w.diffuse * p;
  }
  if (bsdf.enable_specular) {
    float p = 0.f;
    const float f = MicrofacetGGXBsdfPdf(omega_in, omega_out, bsdf.alpha_x, bsdf.alpha_y, 2, &p);
    *bsdf_f = *bsdf_f + bsdf.specular_weight * SpecularColor(omega_in, omega_out, bsdf.specular_color, bsdf.ior) * f;
    *pdf += w.specular * p;
  }
  if (bsdf.enable_clearcoat) {
    float p = 0.f;
    const float f = MicrofacetGGXBsdfPdf(omega_in, omega_out, bsdf.clearcoat_alpha_x, bsdf.clearcoat_alpha_y, 1, &p);
    *bsdf_f = *bsdf_f +
              bsdf.clearcoat_weight * SpecularColor(omega_in, omega_out, bsdf.clearcoat_color, bsdf.clearcoat_ior) * f;
    *pdf += w.clearcoat * p;
  }
}

}  // namespace pbr
