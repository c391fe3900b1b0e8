// Per-vertex shading of one path, restating the reference's integrator pieces so that, given the same PCG32
// stream, a path takes the same decisions in the same order:
//   hit -> SurfaceInfo            src/shader/shader-utils.h:131-164, src/scene.cc:186-249, src/mesh/triangle-mesh.cc:62-184
//   emission + MIS, roulette      src/render.cc:39-69
//   area-light sampling / NEE     src/light-manager.h:37-170, src/shader/shader-utils.h:116-129,166-212
//   Principled vertex             src/shader/cycles-principled-shader.cc:169-242,414-484
//   random-walk subsurface        src/shader/random-walk-sss.h:111-405
//   hair vertex                   src/shader/hair-shader.cc:153-229
// Difference by design (wavefront): the NEE shadow ray is not traced here; the vertex returns it as a request with
// the contribution it would add, and the any-hit kernel adds it when unoccluded.  No random number depends on the
// outcome, so the draw order is unchanged.
#pragma once
#include "common.cuh"
#include "hair.cuh"
#include "principled.cuh"
#include "rng.cuh"
#include "sampling.cuh"
#include "scene_view.cuh"
#include "traverse.cuh"

// g++ evaluates the two rng.Draw() arguments of UniformSampleSphere(rng.Draw(), rng.Draw())
// (random-walk-sss.h:296) right to left: the FIRST draw becomes u2 (cos theta).  Checked against the compiled
// reference by tests/test_gpu_parity.py::test_shading_vertices (GPU) and tests/test_emul_parity.py::test_shading_vertices (g++ emulation).
#ifndef PBR_SSS_SPHERE_DRAW_RIGHT_TO_LEFT
#define PBR_SSS_SPHERE_DRAW_RIGHT_TO_LEFT 1
#endif

namespace pbr {

enum FaceDirection { kFront = 0, kBack = 1, kAmbiguous = 2 };

struct Surface {                 // SurfaceInfo (shader-utils.h:18-41)
  vec3 P, Ns, Ng;
  float u, v, tex_u, tex_v;
  uint32_t instance_id, geom_id, prim_id, material_id, light_entry;
  int face;
  bool is_curve;
};

PBR_HD vec3 NormalizeNoCheck(const vec3& v) {      // raytracer_impl.cc:213-219
  const float inv_norm = 1.0f / sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  return vec3(v.x * inv_norm, v.y * inv_norm, v.z * inv_norm);
}

// geometric normal Embree reports, normalised as EmbreeRayToTraceResult does (raytracer_impl.cc:221-243)
PBR_HD vec3 HitGeometricNormal(const SceneView& s, const HitT& hit) {
  if (hit.prim & kCurveFlag) {
    const uint32_t i = hit.prim & ~kCurveFlag;
    return NormalizeNoCheck(CurveTangent(s.curve_data[i * 4 + 0], s.curve_data[i * 4 + 1], s.curve_data[i * 4 + 2],
                                         s.curve_data[i * 4 + 3], hit.u));
  }
  const vec3 e1 = from4(s.tri_data[hit.prim * 3 + 1]), e2 = from4(s.tri_data[hit.prim * 3 + 2]);
  return NormalizeNoCheck(ecross(e2, e1));
}

PBR_HD Surface MakeSurface(const SceneView& s, const RayT& ray, const HitT& hit) {
  Surface si;
  si.u = hit.u;
  si.v = hit.v;
  si.P = ray.o + hit.t * ray.d;
  si.Ng = HitGeometricNormal(s, hit);
  si.light_entry = kInvalid;
  if (hit.prim & kCurveFlag) {
    const uint32_t prim = s.curve_prim[hit.prim & ~kCurveFlag];
    const uint4 ids = s.curve_ids[prim];
    si.instance_id = ids.x; si.geom_id = ids.y; si.prim_id = ids.z; si.material_id = ids.w;
    si.is_curve = true;
    si.Ns = si.Ng;                              // scene.cc:222-223
    si.tex_u = 0.f; si.tex_v = 0.f;             // scene.cc:243-245
  } else {
    const uint32_t prim = f2u(s.tri_data[hit.prim * 3 + 0].w);
    const uint4 ids = s.tri_ids[prim];
    si.instance_id = ids.x; si.geom_id = ids.y; si.prim_id = ids.z; si.material_id = ids.w;
    si.is_curve = false;
    const uint4 ni = s.tri_nidx[prim];
    si.light_entry = ni.w;
    if (ni.x == kInvalid || ni.y == kInvalid || ni.z == kInvalid) {
      // FetchGeometryNormal: normalize((p1-p0) x (p2-p1)) from the mesh vertices (triangle-mesh.cc:62-75,181-184)
      const uint4 vi = s.tri_vidx[prim];
      const vec3 p0 = from4(s.verts[vi.x]), p1 = from4(s.verts[vi.y]), p2 = from4(s.verts[vi.z]);
      si.Ns = vnormalized(vcross(p1 - p0, p2 - p1));
    } else {
      si.Ns = vnormalized(Lerp3(from4(s.normals[ni.x]), from4(s.normals[ni.y]), from4(s.normals[ni.z]), hit.u, hit.v));
    }
    const uint4 ti = s.tri_tidx[prim];
    if (ti.x == kInvalid || ti.y == kInvalid || ti.z == kInvalid) {
      si.tex_u = hit.u; si.tex_v = hit.v;       // triangle-mesh.cc:131-135
    } else {
      const float2 t0 = s.texcoords[ti.x], t1 = s.texcoords[ti.y], t2 = s.texcoords[ti.z];
      const float w = 1.0f - hit.u - hit.v;
      si.tex_u = (w * t0.x + hit.u * t1.x) + hit.v * t2.x;
      si.tex_v = (w * t0.y + hit.u * t1.y) + hit.v * t2.y;
    }
  }
  const float dg = vdot(ray.d, si.Ng), ds = vdot(ray.d, si.Ns);
  if (dg < 0.0f && ds < 0.0f) si.face = kFront;
  else if (dg > 0.0f && ds > 0.0f) si.face = kBack;
  else si.face = kAmbiguous;
  return si;
}

// ---------------------------------------------------------------- lights
struct LightSample {
  bool valid;
  vec3 pos, normal, emission;
  float pdf;
};

// std::lower_bound(cdf, cdf+n, u) - cdf, clamped to n-1 (release builds of the reference index past the end when
// rounding leaves the last CDF entry below u)
PBR_HD uint32_t LowerBound(const float* cdf, uint32_t n, float u) {
  uint32_t lo = 0, len = n;
  while (len > 0) {
    const uint32_t half = len >> 1;
    if (cdf[lo + half] < u) { lo += half + 1; len -= half + 1; }
    else len = half;
  }
  return lo < n ? lo : n - 1;
}

// LightManager::SampleAllLight (light-manager.h:79-170): 4 draws (light, primitive, u, v); none without lights
PBR_HD LightSample SampleAllLight(const SceneView& s, Pcg32* rng) {
  LightSample r;
  r.valid = false;
  r.pos = r.normal = r.emission = vec3(0.f);
  r.pdf = 0.f;
  if (s.num_lights == 0) return r;
  const float u0 = Draw(rng);
  const uint32_t li = LowerBound(s.light_cdf, s.num_lights, u0);
  const LightRec L = s.lights[li];
  const float u1 = Draw(rng);
  const uint32_t pi = L.prim_offset + LowerBound(s.lprim_cdf + L.prim_offset, L.prim_count, u1);
  const float4 info = s.lprim_info[pi];
  const float u2 = Draw(rng), u3 = Draw(rng);
  float bu, bv;
  TriangleUniformSampler(u2, u3, &bu, &bv);
  const uint4 vi = s.tri_vidx[s.lprim_tri[pi]];
  const vec3 p0 = from4(s.verts[vi.x]), p1 = from4(s.verts[vi.y]), p2 = from4(s.verts[vi.z]);
  r.pos = Lerp3(p0, p1, p2, bu, bv);                         // FetchLocalPosition, not transformed
  r.normal = vnormalized(vcross(p1 - p0, p2 - p1));          // FetchGeometryNormal
  r.emission = vec3(info.x, info.y, info.z);
  r.pdf = info.w;                                            // P(light) * P(prim) * 1/area
  r.valid = true;
  return r;
}

struct ShadowRequest {
  bool active;
  RayT ray;
  vec3 contribute;   // already includes f * Le * w / pdf, excludes the path throughput
};

// DirectIllumination (shader-utils.h:166-212) up to, but not including, the occlusion test.
// eval(omega_l_local, &f, &pdf) is the closure's EvalFunc with omega_out bound.
template <class Eval>
PBR_HD void DirectIllumination(const SceneView& s, const vec3& P, const Frame& Rgl, const vec3& global_normal,
                               Pcg32* rng, const Eval& eval, bool hemisphere, ShadowRequest* req) {
  req->active = false;
  const LightSample ls = SampleAllLight(s, rng);
  if (!ls.valid) return;
  const vec3 dir_to_light = vnormalized(ls.pos - P);
  const float dist = vlength(P - ls.pos);
  const float wl_dot_nl = -vdot(dir_to_light, ls.normal);
  const float wl_dot_np = vdot(dir_to_light, global_normal);
  const float pdf_sigma = fabsf(ls.pdf * dist * dist / (wl_dot_nl * wl_dot_np));
  if ((!hemisphere) || (wl_dot_nl > 0.0f && wl_dot_np > 0.0f)) {
    const vec3 omega_l = Rgl.ToLocal(dir_to_light);
    vec3 f(0.0f);
    float pdf = 0.f;
    eval(omega_l, &f, &pdf);
    const float weight = PowerHeuristicWeight(pdf_sigma, pdf);
    req->contribute = f * ls.emission * weight / pdf_sigma;
    req->ray.o = P;
    req->ray.d = dir_to_light;
    req->ray.tmin = kEps;                               // ShadowRay (shader-utils.h:116-129)
    req->ray.tmax = fmaxf_(kEps, dist - kEps);
    req->active = true;
  }
}

// ---------------------------------------------------------------- result of one vertex
struct VertexResult {
  vec3 wi;          // next direction (world)
  vec3 throughput;  // f * |cos| / pdf
  float pdf;        // pdf of the BSDF sample (for emission MIS at the next vertex)
  vec3 P;           // origin of the next ray (the SSS exit point after a walk)
  ShadowRequest shadow[2];
};

PBR_HD void AbsorbVertex(const vec3& wo, const vec3& P, VertexResult* out) {
  out->wi = wo;
  out->throughput = vec3(0.f);
  out->pdf = 0.f;
  out->P = P;
  out->shadow[0].active = false;
  out->shadow[1].active = false;
}

struct PrincipledEval {
  const PrincipledBsdf* bsdf;
  vec3 wo;
  PBR_HD void operator()(const vec3& wi, vec3* f, float* pdf) const { EvalBsdf(wi, wo, *bsdf, f, pdf); }
};

// SampleBsdf without the subsurface branch (cycles-principled-shader.cc:169-242); `select` already drawn.
// Falls through to the clearcoat lobe when no weight catches the selector (quirk 9).
PBR_HD void SampleBsdfLobes(const PrincipledBsdf& bsdf, const SampleWeight& w, float select, const vec3& wo,
                            Pcg32* rng, vec3* wi, vec3* f, float* pdf) {
  if (select < w.diffuse) {
    const float u0 = Draw(rng), u1 = Draw(rng);
    *wi = CosineSampleHemisphere(u0, u1);
  } else if (select < w.diffuse + w.subsurface + w.specular) {
    const float u0 = Draw(rng), u1 = Draw(rng);
    float p = 0.f;
    MicrofacetGGXSample(wo, bsdf.alpha_x, bsdf.alpha_y, u0, u1, 2, wi, &p);
  } else {
    const float u0 = Draw(rng), u1 = Draw(rng);
    float p = 0.f;
    MicrofacetGGXSample(wo, bsdf.clearcoat_alpha_x, bsdf.clearcoat_alpha_y, u0, u1, 1, wi, &p);
  }
  EvalBsdf(*wi, wo, bsdf, f, pdf);
}

PBR_HD void FinishPrincipled(const Frame& entry, const vec3& omega_in, const vec3& bsdf_f, float ret_pdf,
                             VertexResult* out) {
  out->wi = entry.ToWorld(omega_in);                  // entry-point frame even after SSS (quirk 13)
  const float cos_i = fabsf(omega_in.z);
  out->throughput = bsdf_f * cos_i / ret_pdf;
  out->pdf = ret_pdf;
  if (!IsFinite3(out->throughput) || !finitef_(out->pdf)) {
    out->throughput = vec3(0.f);
    out->pdf = 0.f;
  }
}

// Texture::FetchFloat3 (texture.cc:43-72) -> BilinearFilter with clamp addressing (image-utils.cc:99-167): the
// texel grid is addressed by px = width * clamp(u, 0, 1) (no half-texel offset), neighbours clamp at the border,
// channels beyond the texture's own read as 0.
PBR_HD vec3 TextureFetch3(const SceneView& s, uint32_t tex_id, float u, float v) {
  const TexDesc td = s.tex_desc[tex_id];
  const float uu = fminf_(fmaxf_(u, 0.0f), 1.0f), vv = fminf_(fmaxf_(v, 0.0f), 1.0f);
  const float px = float(td.width) * uu, py = float(td.height) * vv;
  const int w = int(td.width), h = int(td.height);
  int x0 = int(px), y0 = int(py);
  x0 = x0 < 0 ? 0 : (x0 > w - 1 ? w - 1 : x0);
  y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0);
  const int x1 = (x0 + 1 >= w) ? w - 1 : x0 + 1, y1 = (y0 + 1 >= h) ? h - 1 : y0 + 1;
  const float dx = px - float(x0), dy = py - float(y0);
  const float w0 = (1.0f - dx) * (1.0f - dy), w1 = (1.0f - dx) * dy, w2 = dx * (1.0f - dy), w3 = dx * dy;
  const int st = int(td.channels);
  const float* img = s.tex_pixels + td.offset;
  const int i00 = st * (y0 * w + x0), i01 = st * (y0 * w + x1), i10 = st * (y1 * w + x0), i11 = st * (y1 * w + x1);
  float out[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < 3 && c < st; ++c)
    out[c] = ((img[i00 + c] * w0 + img[i10 + c] * w1) + img[i01 + c] * w2) + img[i11 + c] * w3;
  return vec3(out[0], out[1], out[2]);
}

PBR_HD PrincipledBsdf SurfaceBsdf(const SceneView& s, const Surface& si) {
  const DeviceMaterial& m = s.materials[si.material_id];
  PrincipledParams pp;
  memcpy(&pp, m.p, 23 * sizeof(float));
  pp.base_color_tex_id = m.tex_id[0];
  pp.subsurface_color_tex_id = m.tex_id[1];
  // cycles-principled-shader.cc:281-301: a texture id other than -1 replaces the constant colour (ids are validated
  // against the texture table at pbrgpu_commit / pbrgpu_set_materials)
  vec3 base(pp.base_color[0], pp.base_color[1], pp.base_color[2]);
  vec3 sub(pp.subsurface_color[0], pp.subsurface_color[1], pp.subsurface_color[2]);
  if (pp.base_color_tex_id < s.num_textures) base = TextureFetch3(s, pp.base_color_tex_id, si.tex_u, si.tex_v);
  if (pp.subsurface_color_tex_id < s.num_textures) sub = TextureFetch3(s, pp.subsurface_color_tex_id, si.tex_u, si.tex_v);
  return ParamToBsdf(pp, base, sub);
}

PBR_HD Frame PrincipledFrame(const Surface& si) {
  Frame f;
  f.ez = (si.face == kFront) ? si.Ns : -si.Ns;
  BranchlessONB(f.ez, &f.ex, &f.ey);
  return f;
}

// CyclesPrincipledShader (cycles-principled-shader.cc:414-484).  Returns true when the sampled closure is the
// random walk: the caller then runs the walk (SubsurfaceVertex, or SssBegin + the wavefront's walk kernels); rng is
// left right after the selector draw in that case.  fr_out / bsdf_out: the shading frame and closure set of the
// vertex, for the caller that continues with the walk.
// DIFFUSE_ONLY: the material is of class kClassDiffuse (scene_host.cc: ClassifyMaterial) — ParamToBsdf enables the
// Lambert closure and nothing else, and its selection weight is exactly 1, so `select < w.diffuse` always holds.  The
// same statements run in the same order; the other closures are compiled out.
template <bool DIFFUSE_ONLY>
PBR_HD bool PrincipledVertexT(const SceneView& s, const Surface& si, const vec3& wo_world, Pcg32* rng,
                              VertexResult* out, Frame* fr_out, PrincipledBsdf* bsdf_out) {
  if (si.face == kAmbiguous) {
    AbsorbVertex(wo_world, si.P, out);
    return false;
  }
  const Frame fr = PrincipledFrame(si);
  const vec3 wo = fr.ToLocal(wo_world);
  PrincipledBsdf bsdf = SurfaceBsdf(s, si);
  if (DIFFUSE_ONLY) {
    bsdf.enable_diffuse = true;
    bsdf.enable_subsurface = bsdf.enable_specular = bsdf.enable_clearcoat = false;
  }
  out->P = si.P;
  out->shadow[1].active = false;
  PrincipledEval ev = {&bsdf, wo};
  DirectIllumination(s, si.P, fr, fr.ez, rng, ev, true, &out->shadow[0]);

  const SampleWeight w = FetchClosureSampleWeight(wo, bsdf);
  const float select = Draw(rng);
  vec3 wi(0.f), f(0.f);
  float pdf = 0.f;
  if (DIFFUSE_ONLY) {
    const float u0 = Draw(rng), u1 = Draw(rng);
    wi = CosineSampleHemisphere(u0, u1);
    EvalBsdf(wi, wo, bsdf, &f, &pdf);
  } else {
    if (!(select < w.diffuse) && select < w.diffuse + w.subsurface) {
      *fr_out = fr;
      *bsdf_out = bsdf;
      return true;
    }
    SampleBsdfLobes(bsdf, w, select, wo, rng, &wi, &f, &pdf);
  }
  FinishPrincipled(fr, wi, f, pdf, out);
  return false;
}
PBR_HD bool PrincipledVertex(const SceneView& s, const Surface& si, const vec3& wo_world, Pcg32* rng,
                             VertexResult* out, Frame* fr_out, PrincipledBsdf* bsdf_out) {
  return PrincipledVertexT<false>(s, si, wo_world, rng, out, fr_out, bsdf_out);
}
PBR_HD bool PrincipledVertex(const SceneView& s, const Surface& si, const vec3& wo_world, Pcg32* rng,
                             VertexResult* out) {
  Frame fr;
  PrincipledBsdf bsdf;
  return PrincipledVertex(s, si, wo_world, rng, out, &fr, &bsdf);
}

// ---------------------------------------------------------------- random-walk SSS (random-walk-sss.h)
PBR_HD void ComputeScatteringCoefficientFromAlbedo(float A, float d, float* sigma_t, float* sigma_s) {   // :111-121
  const float a = 1.0f - expf(A * (-5.09406f + A * (2.61188f - A * 4.31805f)));
  const float sfit = 1.9f - A + 3.5f * Sqr(A - 0.8f);
  *sigma_t = 1.0f / fmaxf_(d * sfit, 1e-16f);
  *sigma_s = *sigma_t * a;
}

// SampleChannel's pdf (:141-172): proportional to |throughput * albedo|, uniform when that is black
PBR_HD vec3 SssChannelPdf(const vec3& throughput, const vec3& sigma_s, const vec3& sigma_t) {
  const vec3 albedo = SafeDivideSpectrum(sigma_s, sigma_t);
  const float w0 = fabsf(throughput.x * albedo.x), w1 = fabsf(throughput.y * albedo.y),
              w2 = fabsf(throughput.z * albedo.z);
  const float sum = w0 + w1 + w2;
  if (sum > 0.0f) return vec3(w0 / sum, w1 / sum, w2 / sum);
  return vec3(1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f);
}

PBR_HD float SampleScatterDistance(const vec3& throughput, const vec3& sigma_s, const vec3& sigma_t, float u0,
                                   float u1, vec3* channel_pdf) {                                          // :141-187
  *channel_pdf = SssChannelPdf(throughput, sigma_s, sigma_t);
  float sample_sigma_t;
  if (u0 < channel_pdf->x) sample_sigma_t = sigma_t.x;
  else if (u0 < channel_pdf->x + channel_pdf->y) sample_sigma_t = sigma_t.y;
  else sample_sigma_t = sigma_t.z;
  return -logf(1.0f - u1) / sample_sigma_t;
}

struct SssWalkState {     // what one in-flight random walk carries between bounces
  vec3 sigma_t, sigma_s, throughput;
  RayT ray;
  uint32_t bounce;
};

// Steps 1-2 of RandomWalkSubsurface (:227-279): entry direction + coefficients.  false = walk rejected.
PBR_HD bool SssBegin(const Surface& si, const Frame& entry, const PrincipledBsdf& bsdf, Pcg32* rng,
                     SssWalkState* w) {
  if (si.face != kFront) return false;
  const float u0 = Draw(rng), u1 = Draw(rng);
  const vec3 tmp = -CosineSampleHemisphere(u0, u1);
  const vec3 global_dir = entry.ToWorld(tmp);
  if (vdot(-si.Ng, global_dir) <= 0.0f) return false;
  float st[3], ss[3];
  ComputeScatteringCoefficientFromAlbedo(bsdf.subsurface_albedo.x, bsdf.subsurface_radius.x, &st[0], &ss[0]);
  ComputeScatteringCoefficientFromAlbedo(bsdf.subsurface_albedo.y, bsdf.subsurface_radius.y, &st[1], &ss[1]);
  ComputeScatteringCoefficientFromAlbedo(bsdf.subsurface_albedo.z, bsdf.subsurface_radius.z, &st[2], &ss[2]);
  w->sigma_t = vec3(st[0], st[1], st[2]);
  w->sigma_s = vec3(ss[0], ss[1], ss[2]);
  w->throughput = SafeDivideSpectrum(bsdf.subsurface_weight, bsdf.subsurface_albedo);
  w->ray.o = si.P;
  w->ray.d = global_dir;
  w->ray.tmin = 1e-3f;
  w->ray.tmax = kInf;
  w->bounce = 0;
  return true;
}

enum SssStep { kSssContinue = 0, kSssHit = 1, kSssAbsorbed = 2 };

// Can a walk segment (origin o, unit direction d, length len) be declared free of intersections without tracing it?
// Most segments of a random walk stay away from the surface, yet each one costs a traversal, and a deep one: a point
// inside a closed mesh lies inside the boxes of many BVH nodes (measured: 3x the node steps of a camera ray).  The
// clearance field answers conservatively: the byte of the cell holding a point p is a lower bound of the distance
// from any point of that cell to any primitive of the scene, so the stretch of the segment from p up to that
// distance is free; the march continues from its end (sphere tracing along the segment) until the segment is
// exhausted (clear), or a cell next to the surface is reached, or the step budget runs out (trace it).  A segment
// that is skipped is one the query would have reported "no hit" for, so the walk takes exactly the same decisions
// (same random numbers, same result).
PBR_HD bool SegmentIsClear(const SceneView& s, const vec3& o, const vec3& d, float len) {
  if (!s.clear_dist) return false;
  const float need = len * 1.02f;   // 2 % of margin: |d| = 1 up to rounding, rounding of the hit point
  const float slack = 0.0625f * s.clear_quantum;   // a point a few ulps across a cell face
  float t = 0.f;
  for (uint32_t i = 0; i < s.clear_march_steps; ++i) {
    const float gx = ((o.x + t * d.x) - s.clear_org[0]) * s.clear_inv_cell;
    const float gy = ((o.y + t * d.y) - s.clear_org[1]) * s.clear_inv_cell;
    const float gz = ((o.z + t * d.z) - s.clear_org[2]) * s.clear_inv_cell;
    if (!(gx >= 0.f && gy >= 0.f && gz >= 0.f && gx < float(s.clear_dims[0]) && gy < float(s.clear_dims[1]) &&
          gz < float(s.clear_dims[2])))
      return false;
    const uint32_t idx = (uint32_t(gz) * s.clear_dims[1] + uint32_t(gy)) * s.clear_dims[0] + uint32_t(gx);
    const float bound = float(s.clear_dist[idx]) * s.clear_quantum - slack;
    if (!(bound > 0.f)) return false;        // next to the surface
    t += bound * 0.98f;                       // the stretch [t, t + bound) is free; restart a little before its end
    if (t >= need) return true;
  }
  return false;
}

// One iteration of the walk loop (:281-383), split at the ray query so that the wavefront's walk kernel can run the
// query in its warp traversal engine:
//   SssPrepareSegment: new direction (bounces > 0) + scatter distance -> w->ray (tmax = scatter distance)
//   SssFinishSegment:  transmittance / throughput update, roulette, advance.  On kSssHit the caller holds the exit
//                      intersection along w->ray.
// `channel_pdf`: the channel probabilities the distance was drawn with; the caller hands them back to
// SssFinishSegment (throughput is unchanged in between), which saves recomputing six IEEE divisions per segment.
PBR_HD void SssPrepareSegment(Pcg32* rng, SssWalkState* w, vec3* channel_pdf_out) {
  if (w->bounce > 0) {
#if PBR_SSS_SPHERE_DRAW_RIGHT_TO_LEFT
    const float u2 = Draw(rng), u1 = Draw(rng);
#else
    const float u1 = Draw(rng), u2 = Draw(rng);
#endif
    w->ray.d = vnormalized(UniformSampleSphere(u1, u2));
    w->ray.tmin = 0.f;
  }
  const float ua = Draw(rng), ub = Draw(rng);
  w->ray.tmax = SampleScatterDistance(w->throughput, w->sigma_s, w->sigma_t, ua, ub, channel_pdf_out);
}

PBR_HD SssStep SssFinishSegment(bool is_hit, float hit_t, Pcg32* rng, SssWalkState* w, const vec3& channel_pdf) {
  const float t = is_hit ? hit_t : w->ray.tmax;
  const vec3 tr(expf(-w->sigma_t.x * t), expf(-w->sigma_t.y * t), expf(-w->sigma_t.z * t));
  if (is_hit) {
    const float pdf = vdot(channel_pdf, tr);
    w->throughput = w->throughput * tr / pdf;
    return kSssHit;
  }
  const float pdf = vdot(channel_pdf, w->sigma_t * tr);
  w->throughput = w->throughput * (w->sigma_s * tr) / pdf;
  const float p = Saturatef(SpectrumNorm(w->throughput));
  const float q = Draw(rng);
  if (q >= p) return kSssAbsorbed;
  w->throughput = w->throughput / p;
  w->ray.o = w->ray.o + t * w->ray.d;
  w->bounce++;
  if (w->bounce > 8192u) return kSssAbsorbed;   // for (bounce = 0; bounce <= 8192; ++bounce)
  return kSssContinue;
}

PBR_HD SssStep SssBounce(const SceneView& s, Pcg32* rng, SssWalkState* w, HitT* hit, uint64_t* rays) {
  vec3 channel_pdf;
  SssPrepareSegment(rng, w, &channel_pdf);
  const bool is_hit = TraceClosest<false>(s, w->ray, hit, nullptr);
#ifdef PBR_CLEARANCE_PROBE   // tests/host_emul only: every segment the field would skip must be a miss
  PBR_CLEARANCE_PROBE(SegmentIsClear(s, w->ray.o, w->ray.d, w->ray.tmax * 1.001f), is_hit);
#endif
#ifdef PBR_WALK_SEGMENT_PROBE   // tests/host_emul only: the walks' segments, for offline traversal statistics
  PBR_WALK_SEGMENT_PROBE(w->ray, SegmentIsClear(s, w->ray.o, w->ray.d, w->ray.tmax * 1.001f));
#endif
  if (rays) ++*rays;
  return SssFinishSegment(is_hit, hit->t, rng, w, channel_pdf);
}

// After a surface hit (:385-404) + the success branch of SampleBsdf (cycles-principled-shader.cc:187-216).
PBR_HD void SssFinish(const SceneView& s, const Surface& entry_si, const Frame& entry, const SssWalkState& w,
                      const HitT& hit, Pcg32* rng, VertexResult* out) {
  const Surface ex = MakeSurface(s, w.ray, hit);
  if (ex.instance_id != entry_si.instance_id || ex.face != kBack) {
    FinishPrincipled(entry, vec3(0.f), vec3(0.f), 0.f, out);   // omega_in = f = pdf = 0 -> throughput 0
    return;
  }
  Frame xf;
  xf.ez = ex.Ns;                                   // exit frame, normal not flipped
  BranchlessONB(xf.ez, &xf.ex, &xf.ey);
  const vec3 new_wo = xf.ToLocal(w.ray.d);
  PrincipledBsdf nb;
  InitBsdf(&nb);
  nb.enable_diffuse = true;
  nb.diffuse_weight = w.throughput;
  PrincipledEval ev = {&nb, new_wo};
  out->P = ex.P;
  DirectIllumination(s, ex.P, xf, ex.Ns, rng, ev, true, &out->shadow[1]);
  const SampleWeight sw = FetchClosureSampleWeight(new_wo, nb);
  const float select = Draw(rng);
  vec3 wi(0.f), f(0.f);
  float pdf = 0.f;
  SampleBsdfLobes(nb, sw, select, new_wo, rng, &wi, &f, &pdf);
  FinishPrincipled(entry, wi, f, pdf, out);
}

// Whole subsurface branch for one path, run to completion by one thread (the wavefront's sss kernel refills idle
// lanes between bounces instead; this form serves the host emulation and small test hooks).
PBR_HD void SubsurfaceVertex(const SceneView& s, const Surface& si, Pcg32* rng, VertexResult* out, uint64_t* rays) {
  const Frame fr = PrincipledFrame(si);
  const PrincipledBsdf bsdf = SurfaceBsdf(s, si);
  SssWalkState w;
  if (!SssBegin(si, fr, bsdf, rng, &w)) {
    FinishPrincipled(fr, vec3(0.f), vec3(0.f), 0.f, out);
    return;
  }
  HitT hit;
  for (;;) {
    const SssStep st = SssBounce(s, rng, &w, &hit, rays);
    if (st == kSssHit) { SssFinish(s, si, fr, w, hit, rng, out); return; }
    if (st == kSssAbsorbed) { FinishPrincipled(fr, vec3(0.f), vec3(0.f), 0.f, out); return; }
  }
}

// ---------------------------------------------------------------- hair vertex (hair-shader.cc:153-229)
struct HairEval {
  const hair::HairBsdf* bsdf;
  vec3 wo;
  PBR_HD void operator()(const vec3& wi, vec3* f, float* pdf) const {
    const vec3 fc = hair::EnergyConservingHairBsdfCosPdf(wi, wo, *bsdf, pdf);
    *f = fc / fabsf(wi.x);
  }
};

PBR_HD void HairVertex(const SceneView& s, const Surface& si, const vec3& wo_world, Pcg32* rng, VertexResult* out) {
  if (si.face == kAmbiguous) {
    AbsorbVertex(wo_world, si.P, out);
    return;
  }
  Frame fr;
  fr.ex = si.Ns;
  fr.ey = vnormalized(vcross(vcross(wo_world, fr.ex), fr.ex));
  fr.ez = vcross(fr.ex, fr.ey);
  const vec3 wo = fr.ToLocal(wo_world);
  const hair::HairBsdf bsdf = hair::ParamToBsdf(s.materials[si.material_id].p, si.v);
  out->P = si.P;
  out->shadow[1].active = false;
  HairEval ev = {&bsdf, wo};
  DirectIllumination(s, si.P, fr, fr.ex, rng, ev, false, &out->shadow[0]);
  float us[4];
  us[0] = Draw(rng); us[1] = Draw(rng); us[2] = Draw(rng); us[3] = Draw(rng);
  vec3 wi(0.f);
  float pdf = 0.f;
  const vec3 fc = hair::EnergyConservingHairSample(wo, bsdf, us, &wi, &pdf);
  out->wi = fr.ToWorld(wi);
  out->throughput = fc / pdf;
  out->pdf = pdf;
  if (!IsFinite3(out->throughput) || !finitef_(out->pdf)) {
    out->throughput = vec3(0.f);
    out->pdf = 0.f;
  }
}

// ---------------------------------------------------------------- emission + roulette (render.cc:39-69)
// Returns false when the path ends at this vertex.  L and throughput are updated in place.
PBR_HD bool EmissionAndRoulette(const SceneView& s, const RayT& ray, const HitT& hit, const Surface& si,
                                uint32_t depth, float bsdf_pdf_prev, Pcg32* rng, vec3* L, vec3* throughput) {
  if (si.face == kFront && si.light_entry != kInvalid) {
    const float4 e = s.emissive[si.light_entry];
    const float area_to_solid_angle = fabsf((hit.t * hit.t) / vdot(si.Ns, ray.d));
    const float weight = (depth == 0) ? 1.0f : PowerHeuristicWeight(bsdf_pdf_prev, e.w * area_to_solid_angle);
    *L = *L + weight * vec3(e.x, e.y, e.z) * (*throughput);
  }
  const float p = SpectrumNorm(*throughput);
  if (p < Draw(rng)) return false;
  *throughput = (*throughput) * vec3(1.0f / p);
  return true;
}

// material dispatch of Shader() (shader.cc:8-34): 0 = absorbed here, 1 = principled, 2 = hair
PBR_HD int MaterialKind(const SceneView& s, const Surface& si) {
  if (si.material_id == kInvalid || si.material_id >= s.num_materials) return 0;
  return s.materials[si.material_id].type == 0 ? 1 : 2;
}

}  // namespace pbr
