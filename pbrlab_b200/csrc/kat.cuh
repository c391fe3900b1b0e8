// Closure-level known-answer hook behind pbrgpu_eval_closure(): one record in, one record out, evaluated by the very
// device functions the shading kernels call.  The same switch is compiled by g++ for the CPU tests.
//   op  params                       in (per record)                  out (per record)
//    0  -                            4 x u32 bits: initstate lo,hi, initseq lo,hi      out_stride PCG32 draws
//    1  [0] = fast_math function id  x, y                             1   (0 sin 1 cos 2 exp2 3 exp 4 log2 5 log
//                                                                          6 atan2(y=x-arg? no: FastAtan2(in0,in1)) 7 asin
//                                                                          8 sincos.sin 9 sincos.cos)
//    2  -                            u0, u1                           3   CosineSampleHemisphere
//    3  -                            u0, u1                           3   UniformSampleSphere
//    4  -                            sampled_pdf, other_pdf           1   PowerHeuristicWeight
//    5  -                            cos, eta                         1   FresnelDielectricCos
//    6  ax, ay, distrib              wi3, wo3                         2   MicrofacetGGXBsdfPdf -> f, pdf
//    7  ax, ay, distrib              wo3, u0, u1                      5   MicrofacetGGXSample -> wi3, f, pdf
//    8  principled p[23]             wi3, wo3                         4   EvalBsdf -> f3, pdf
//    9  principled p[23]             wo3                              4   FetchClosureSampleWeight
//   10  principled p[23]             -                                36  ParamToBsdf fields (tests/refbind.py order)
//   11  hair p[20]                   h, wi3, wo3                      4   EnergyConservingHairBsdfCosPdf -> f*cos 3, pdf
//   12  hair p[20]                   h, wo3, us4                      7   EnergyConservingHairSample -> wi3, f*cos 3, pdf
//   13  hair p[20]                   -                                9   sigma_a3, v[4], s, alpha
//   14  -                            albedo3, radius3, weight3        9   sigma_t3, sigma_s3, throughput3
//   15  -                            thr3, sigma_s3, sigma_t3, u0,u1  4   SampleScatterDistance -> distance, channel_pdf3
#pragma once
#include "device/shade.cuh"

namespace pbr {

PBR_HD PrincipledBsdf KatBsdf(const float* p23) {
  PrincipledParams pp;
  memcpy(&pp, p23, 23 * sizeof(float));
  pp.base_color_tex_id = kInvalid;
  pp.subsurface_color_tex_id = kInvalid;
  return ParamToBsdf(pp, vec3(p23[0], p23[1], p23[2]), vec3(p23[7], p23[8], p23[9]));
}

PBR_HD void KatEval(int op, const float* prm, const float* in, float* out, uint32_t out_stride) {
  switch (op) {
    case 0: {
      Pcg32 rng;
      const uint64_t st = uint64_t(f2u(in[0])) | (uint64_t(f2u(in[1])) << 32);
      const uint64_t sq = uint64_t(f2u(in[2])) | (uint64_t(f2u(in[3])) << 32);
      pcg32_srandom(&rng, st, sq);
      for (uint32_t i = 0; i < out_stride; ++i) out[i] = Draw(&rng);
    } break;
    case 1: {
      const int f = int(prm[0]);
      float s, c;
      switch (f) {
        case 0: out[0] = fast_math::FastSin(in[0]); break;
        case 1: out[0] = fast_math::FastCos(in[0]); break;
        case 2: out[0] = fast_math::FastExp2(in[0]); break;
        case 3: out[0] = fast_math::FastExp(in[0]); break;
        case 4: out[0] = fast_math::FastLog2(in[0]); break;
        case 5: out[0] = fast_math::FastLog(in[0]); break;
        case 6: out[0] = fast_math::FastAtan2(in[0], in[1]); break;
        case 7: out[0] = fast_math::FastAsin(in[0]); break;
        case 8: fast_math::FastSincos(in[0], &s, &c); out[0] = s; break;
        default: fast_math::FastSincos(in[0], &s, &c); out[0] = c; break;
      }
    } break;
    case 2: { const vec3 w = CosineSampleHemisphere(in[0], in[1]); out[0] = w.x; out[1] = w.y; out[2] = w.z; } break;
    case 3: { const vec3 w = UniformSampleSphere(in[0], in[1]); out[0] = w.x; out[1] = w.y; out[2] = w.z; } break;
    case 4: out[0] = PowerHeuristicWeight(in[0], in[1]); break;
    case 5: out[0] = FresnelDielectricCos(in[0], in[1]); break;
    case 6: {
      float pdf = 0.f;
      out[0] = MicrofacetGGXBsdfPdf(vec3(in[0], in[1], in[2]), vec3(in[3], in[4], in[5]), prm[0], prm[1], int(prm[2]), &pdf);
      out[1] = pdf;
    } break;
    case 7: {
      vec3 wi(0.f);
      float pdf = 0.f;
      const float f = MicrofacetGGXSample(vec3(in[0], in[1], in[2]), prm[0], prm[1], in[3], in[4], int(prm[2]), &wi, &pdf);
      out[0] = wi.x; out[1] = wi.y; out[2] = wi.z; out[3] = f; out[4] = pdf;
    } break;
    case 8: {
      const PrincipledBsdf b = KatBsdf(prm);
      vec3 f(0.f);
      float pdf = 0.f;
      EvalBsdf(vec3(in[0], in[1], in[2]), vec3(in[3], in[4], in[5]), b, &f, &pdf);
      out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = pdf;
    } break;
    case 9: {
      const PrincipledBsdf b = KatBsdf(prm);
      const SampleWeight w = FetchClosureSampleWeight(vec3(in[0], in[1], in[2]), b);
      out[0] = w.diffuse; out[1] = w.subsurface; out[2] = w.specular; out[3] = w.clearcoat;
    } break;
    case 10: {
      const PrincipledBsdf b = KatBsdf(prm);
      float* o = out;
      *o++ = b.enable_diffuse; *o++ = b.diffuse_weight.x; *o++ = b.diffuse_weight.y; *o++ = b.diffuse_weight.z;
      *o++ = b.enable_subsurface;
      *o++ = b.subsurface_weight.x; *o++ = b.subsurface_weight.y; *o++ = b.subsurface_weight.z;
      *o++ = b.subsurface_albedo.x; *o++ = b.subsurface_albedo.y; *o++ = b.subsurface_albedo.z;
      *o++ = b.subsurface_radius.x; *o++ = b.subsurface_radius.y; *o++ = b.subsurface_radius.z;
      *o++ = b.enable_specular; *o++ = b.specular_weight.x; *o++ = b.specular_weight.y; *o++ = b.specular_weight.z;
      *o++ = b.alpha_x; *o++ = b.alpha_y; *o++ = b.ior;
      *o++ = b.specular_color.x; *o++ = b.specular_color.y; *o++ = b.specular_color.z;
      *o++ = b.enable_clearcoat;
      *o++ = b.clearcoat_weight.x; *o++ = b.clearcoat_weight.y; *o++ = b.clearcoat_weight.z;
      *o++ = b.clearcoat_alpha_x; *o++ = b.clearcoat_alpha_y; *o++ = b.clearcoat_ior;
      *o++ = b.clearcoat_color.x; *o++ = b.clearcoat_color.y; *o++ = b.clearcoat_color.z;
      *o++ = 0.f; *o++ = 0.f;
    } break;
    case 11: {
      const hair::HairBsdf b = hair::ParamToBsdf(prm, in[0]);
      float pdf = 0.f;
      const vec3 f = hair::EnergyConservingHairBsdfCosPdf(vec3(in[1], in[2], in[3]), vec3(in[4], in[5], in[6]), b, &pdf);
      out[0] = f.x; out[1] = f.y; out[2] = f.z; out[3] = pdf;
    } break;
    case 12: {
      const hair::HairBsdf b = hair::ParamToBsdf(prm, in[0]);
      float pdf = 0.f;
      vec3 wi(0.f);
      const vec3 f = hair::EnergyConservingHairSample(vec3(in[1], in[2], in[3]), b, in + 4, &wi, &pdf);
      out[0] = wi.x; out[1] = wi.y; out[2] = wi.z; out[3] = f.x; out[4] = f.y; out[5] = f.z; out[6] = pdf;
    } break;
    case 13: {
      const hair::HairBsdf b = hair::ParamToBsdf(prm, 0.f);
      out[0] = b.sigma_a.x; out[1] = b.sigma_a.y; out[2] = b.sigma_a.z;
      out[3] = b.v[0]; out[4] = b.v[1]; out[5] = b.v[2]; out[6] = b.v[3];
      out[7] = b.s; out[8] = b.alpha;
    } break;
    case 14: {
      float st[3], ss[3];
      for (int k = 0; k < 3; ++k) ComputeScatteringCoefficientFromAlbedo(in[k], in[3 + k], &st[k], &ss[k]);
      const vec3 thr = SafeDivideSpectrum(vec3(in[6], in[7], in[8]), vec3(in[0], in[1], in[2]));
      for (int k = 0; k < 3; ++k) { out[k] = st[k]; out[3 + k] = ss[k]; }
      out[6] = thr.x; out[7] = thr.y; out[8] = thr.z;
    } break;
    case 15: {
      vec3 cp;
      out[0] = SampleScatterDistance(vec3(in[0], in[1], in[2]), vec3(in[3], in[4], in[5]), vec3(in[6], in[7], in[8]),
                                     in[9], in[10], &cp);
      out[1] = cp.x; out[2] = cp.y; out[3] = cp.z;
    } break;
    default: break;
  }
}

// Reference-order megakernel form of GetRadiance (src/render.cc:24-90): one thread walks one path to the end,
// tracing its own shadow rays.  It is the cross-check for the wavefront scheduling (same per-vertex functions, so
// the two must agree to the last bit up to the order of the two NEE additions) and what the CPU tests run.
// `L`, `throughput`, `pdf_prev`, `depth`: the state a path carries from one vertex to the next (render.cc:79-86), so a
// path can be taken up in the middle — the wavefront hands the last few paths of a frame to FinishPathsKernel.
PBR_HD vec3 PathRadianceFrom(const SceneView& s, RayT ray, Pcg32* rng, vec3 L, vec3 throughput, float pdf_prev,
                             uint32_t depth, uint64_t* ray_counts /* closest, shadow, sss */) {
  for (;; ++depth) {
    if (IsBlack(throughput)) break;
    HitT hit;
    const bool found = TraceClosest<false>(s, ray, &hit, nullptr);
    if (ray_counts) ray_counts[0]++;
    if (!found) break;
    const Surface si = MakeSurface(s, ray, hit);
    if (!EmissionAndRoulette(s, ray, hit, si, depth, pdf_prev, rng, &L, &throughput)) break;
    VertexResult vr;
    const vec3 wo = -ray.d;
    const int kind = MaterialKind(s, si);
    if (kind == 1) {
      if (PrincipledVertex(s, si, wo, rng, &vr)) SubsurfaceVertex(s, si, rng, &vr, ray_counts ? &ray_counts[2] : nullptr);
    } else if (kind == 2) {
      HairVertex(s, si, wo, rng, &vr);
    } else {
      AbsorbVertex(wo, si.P, &vr);
    }
    vec3 direct(0.f);
    for (int k = 0; k < 2; ++k) {
      if (vr.shadow[k].active) {
        if (ray_counts) ray_counts[1]++;
        if (!TraceAny<false>(s, vr.shadow[k].ray, nullptr)) direct = direct + vr.shadow[k].contribute;
      }
    }
    L = L + throughput * direct;
    throughput = vr.throughput * throughput;
    pdf_prev = vr.pdf;
    ray.o = vr.P;
    ray.d = vr.wi;
    ray.tmin = 1e-3f;
    ray.tmax = kInf;
  }
  return L;
}

PBR_HD vec3 PathRadiance(const SceneView& s, RayT ray, Pcg32* rng, uint64_t* ray_counts /* closest, shadow, sss */) {
  return PathRadianceFrom(s, ray, rng, vec3(0.f), vec3(1.f), 0.f, 0u, ray_counts);
}

}  // namespace pbr
