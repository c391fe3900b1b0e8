// TEST INFRASTRUCTURE ONLY.  extern "C" door into the UNMODIFIED reference (pbrlab + vendored Embree)
// compiled from /root/reference by oracle/Makefile into oracle/_ref/libpbrlab_ref.so.
//
// This file contains no reference code: it only *calls* the reference through its own headers so the
// Python tests (tests/) and bench.py's cpu_baseline / --impl reference legs can
//   * pin the restated oracle (oracle/pbr_oracle.cc) against the real thing,
//   * generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   * time pbrlab::Render() on the host cores.
// The product path (pbrlab_b200/) never links or loads it.
//
// The two shader translation units are #included (not linked) so their file-static helpers
// (ParamToBsdf / EvalBsdf / SampleBsdf / FetchClosureSampleWeight, src/shader/cycles-principled-shader.cc:54-412;
// BetamToV / CalcS / CalcSigmaA* / ParamToBsdf, src/shader/hair-shader.cc:19-151) are reachable for
// closure-level known-answer vectors.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "shader/cycles-principled-shader.cc"  // NOLINT(build/include)
#include "shader/hair-shader.cc"               // NOLINT(build/include)

#include "image-utils.h"
#include "io/curve-mesh-io.h"
#include "io/image-io.h"
#include "io/triangle-mesh-io.h"
#include "pc-common.h"
#include "render.h"
#include "scene.h"

namespace pbrlab {
// external linkage, not declared in any header (src/render.cc:22-24)
float3 GetRadiance(const Ray& input_ray, const Scene& scene, const RNG& rng);
}  // namespace pbrlab

using namespace pbrlab;  // NOLINT

namespace {

template <typename F>
void ParallelFor(uint64_t n, F f) {
  const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<uint64_t> next(0);
  const uint64_t chunk = 4096;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&]() {
      for (;;) {
        const uint64_t b = next.fetch_add(chunk);
        if (b >= n) break;
        const uint64_t e = std::min(n, b + chunk);
        for (uint64_t i = b; i < e; ++i) f(i);
      }
    });
  }
  for (auto& t : th) t.join();
}

Ray MakeRay(const float* r) {
  Ray ray;
  ray.ray_org = float3(r[0], r[1], r[2]);
  ray.min_t   = r[3];
  ray.ray_dir = float3(r[4], r[5], r[6]);
  ray.max_t   = r[7];
  return ray;
}

struct ObjData {
  std::vector<TriangleMesh> meshes;
  std::vector<MaterialParameter> materials;
  std::vector<Texture> textures;
};

CyclesPrincipledBsdfParameter UnpackPrincipled(const float* p) {
  CyclesPrincipledBsdfParameter m;
  m.base_color             = float3(p[0], p[1], p[2]);
  m.subsurface             = p[3];
  m.subsurface_radius      = float3(p[4], p[5], p[6]);
  m.subsurface_color       = float3(p[7], p[8], p[9]);
  m.metallic               = p[10];
  m.specular               = p[11];
  m.specular_tint          = p[12];
  m.roughness              = p[13];
  m.anisotropic            = p[14];
  m.anisotropic_rotation   = p[15];
  m.sheen                  = p[16];
  m.sheen_tint             = p[17];
  m.clearcoat              = p[18];
  m.clearcoat_roughness    = p[19];
  m.ior                    = p[20];
  m.transmission           = p[21];
  m.transmission_roughness = p[22];
  return m;
}

void PackPrincipled(const CyclesPrincipledBsdfParameter& m, float* p) {
  p[0] = m.base_color[0]; p[1] = m.base_color[1]; p[2] = m.base_color[2];
  p[3] = m.subsurface;
  p[4] = m.subsurface_radius[0]; p[5] = m.subsurface_radius[1]; p[6] = m.subsurface_radius[2];
  p[7] = m.subsurface_color[0]; p[8] = m.subsurface_color[1]; p[9] = m.subsurface_color[2];
  p[10] = m.metallic; p[11] = m.specular; p[12] = m.specular_tint; p[13] = m.roughness;
  p[14] = m.anisotropic; p[15] = m.anisotropic_rotation; p[16] = m.sheen; p[17] = m.sheen_tint;
  p[18] = m.clearcoat; p[19] = m.clearcoat_roughness; p[20] = m.ior; p[21] = m.transmission;
  p[22] = m.transmission_roughness;
}

HairBsdfParameter UnpackHair(const float* p) {
  HairBsdfParameter m;
  m.coloring_hair        = (p[0] != 0.f) ? HairBsdfParameter::kMelanin : HairBsdfParameter::kRGB;
  m.base_color           = float3(p[1], p[2], p[3]);
  m.melanin              = p[4];
  m.melanin_redness      = p[5];
  m.melanin_randomize    = p[6];
  m.roughness            = p[7];
  m.azimuthal_roughness  = p[8];
  m.ior                  = p[9];
  m.shift                = p[10];
  m.specular_tint        = float3(p[11], p[12], p[13]);
  m.second_specular_tint = float3(p[14], p[15], p[16]);
  m.transmission_tint    = float3(p[17], p[18], p[19]);
  return m;
}

}  // namespace

extern "C" {

#define REF_API __attribute__((visibility("default")))

// ---------------------------------------------------------------- scene (pc/pc-common.cc:239 CreateScene)
REF_API void* ref_scene_create(int nfiles, const char** files) {
  std::vector<std::string> store;
  store.emplace_back("ref");
  for (int i = 0; i < nfiles; ++i) store.emplace_back(files[i]);
  std::vector<char*> argv;
  for (auto& s : store) argv.push_back(const_cast<char*>(s.c_str()));
  Scene* scene = new Scene();
  if (!CreateScene(int(argv.size()), argv.data(), scene)) {
    delete scene;
    return nullptr;
  }
  return scene;
}
REF_API void ref_scene_destroy(void* s) { delete static_cast<Scene*>(s); }
REF_API void ref_scene_aabb(void* s, float* bmin, float* bmax) {
  static_cast<Scene*>(s)->FetchSceneAABB(bmin, bmax);
}
REF_API int ref_num_threads() { return int(std::max(1u, std::thread::hardware_concurrency())); }

// ---------------------------------------------------------------- ray queries (src/scene.cc:261-268)
// rays: n x 8 floats (org.xyz, tmin, dir.xyz, tmax); out_f: n x 6 (t,u,v,Ng.xyz); out_id: n x 3
REF_API void ref_trace(void* s, const float* rays, uint64_t n, float* out_f, uint32_t* out_id) {
  const Scene* scene = static_cast<Scene*>(s);
  ParallelFor(n, [&](uint64_t i) {
    const TraceResult tr = scene->TraceFirstHit1(MakeRay(rays + 8 * i));
    float* f = out_f + 6 * i;
    f[0] = tr.t; f[1] = tr.u; f[2] = tr.v;
    f[3] = tr.normal_g[0]; f[4] = tr.normal_g[1]; f[5] = tr.normal_g[2];
    out_id[3 * i + 0] = tr.instance_id;
    out_id[3 * i + 1] = tr.geom_id;
    out_id[3 * i + 2] = tr.prim_id;
  });
}
REF_API void ref_occluded(void* s, const float* rays, uint64_t n, uint8_t* out) {
  const Scene* scene = static_cast<Scene*>(s);
  ParallelFor(n, [&](uint64_t i) { out[i] = scene->AnyHit1(MakeRay(rays + 8 * i)) ? 1 : 0; });
}

// ---------------------------------------------------------------- per-path radiance (src/render.cc:24-90)
// seeds: n x 2 u64 (initstate, initseq) for RNG (src/random/rng.h:29-36)
REF_API void ref_radiance(void* s, const float* rays, const uint64_t* seeds, uint64_t n, float* out) {
  const Scene* scene = static_cast<Scene*>(s);
  ParallelFor(n, [&](uint64_t i) {
    RNG rng(seeds[2 * i], seeds[2 * i + 1]);
    Ray ray = MakeRay(rays + 8 * i);
    const float3 L = GetRadiance(ray, *scene, rng);
    out[3 * i + 0] = L[0]; out[3 * i + 1] = L[1]; out[3 * i + 2] = L[2];
  });
}

// Full pixel sample exactly as RenderingTile does it (src/render.cc:132-171): jitter draws come from the
// same RNG as the path.  cam: 8 floats (eye.xyz, x_corner, y_corner, z_corner, dx, dy) as computed there.
REF_API void ref_camera(void* s, uint32_t width, uint32_t height, float* cam) {
  float bmax[3], bmin[3];
  static_cast<Scene*>(s)->FetchSceneAABB(bmin, bmax);
  float hs, vs;
  if (bmax[0] - bmin[0] > bmax[1] - bmin[1]) {
    hs = bmax[0] - bmin[0];
    vs = hs * float(height) / float(width);
  } else {
    vs = bmax[1] - bmin[1];
    hs = vs * float(width) / float(height);
  }
  cam[0] = (bmax[0] + bmin[0]) * 0.5f;
  cam[1] = (bmax[1] + bmin[1]) * 0.5f;
  cam[2] = bmax[2] + hs * 0.5f * sqrtf(3.f);
  cam[3] = (bmax[0] + bmin[0]) * 0.5f - hs * 0.5f;
  cam[4] = (bmax[1] + bmin[1]) * 0.5f + vs * 0.5f;
  cam[5] = bmax[2];
  cam[6] = hs / float(width);
  cam[7] = vs / float(height);
}

// ---------------------------------------------------------------- one shading vertex (src/shader/shader.cc:8-34)
// For each ray: closest hit -> SurfaceInfo -> Shader().  out: n x 16 floats
//   [0] hit(0/1) [1..3] wi [4..6] throughput [7..9] contribute [10] pdf [11..13] position after shading
//   [14] face_direction [15] t
REF_API void ref_shade(void* s, const float* rays, const uint64_t* seeds, uint64_t n, float* out) {
  const Scene* scene = static_cast<Scene*>(s);
  ParallelFor(n, [&](uint64_t i) {
    float* o = out + 16 * i;
    for (int k = 0; k < 16; ++k) o[k] = 0.f;
    const Ray ray = MakeRay(rays + 8 * i);
    const TraceResult tr = scene->TraceFirstHit1(ray);
    if (tr.instance_id == uint32_t(-1)) return;
    SurfaceInfo si = TraceResultToSufaceInfo(ray, *scene, tr);
    o[14] = float(int(si.face_direction));
    o[15] = tr.t;
    RNG rng(seeds[2 * i], seeds[2 * i + 1]);
    float3 wi, thr, contrib;
    float pdf = 0.f;
    Shader(*scene, -ray.ray_dir, rng, &si, &wi, &thr, &contrib, &pdf);
    o[0] = 1.f;
    o[1] = wi[0]; o[2] = wi[1]; o[3] = wi[2];
    o[4] = thr[0]; o[5] = thr[1]; o[6] = thr[2];
    o[7] = contrib[0]; o[8] = contrib[1]; o[9] = contrib[2];
    o[10] = pdf;
    o[11] = si.global_position[0]; o[12] = si.global_position[1]; o[13] = si.global_position[2];
  });
}

// hit -> shading inputs (src/shader/shader-utils.h:131-164). out: n x 12: P(3) Ns(3) Ng(3) uv(2) face
REF_API void ref_surface(void* s, const float* rays, uint64_t n, float* out) {
  const Scene* scene = static_cast<Scene*>(s);
  ParallelFor(n, [&](uint64_t i) {
    float* o = out + 12 * i;
    for (int k = 0; k < 12; ++k) o[k] = 0.f;
    const Ray ray = MakeRay(rays + 8 * i);
    const TraceResult tr = scene->TraceFirstHit1(ray);
    if (tr.instance_id == uint32_t(-1)) { o[11] = -1.f; return; }
    const SurfaceInfo si = TraceResultToSufaceInfo(ray, *scene, tr);
    for (int k = 0; k < 3; ++k) {
      o[k] = si.global_position[k]; o[3 + k] = si.normal_s[k]; o[6 + k] = si.normal_g[k];
    }
    o[9] = si.texcoord[0]; o[10] = si.texcoord[1];
    o[11] = float(int(si.face_direction));
  });
}

// ---------------------------------------------------------------- lights (src/light-manager.h:37-170)
// out: n x 10 (pos3, normal3, emission3, pdf); returns light_type of the last sample
REF_API int ref_sample_light(void* s, const uint64_t* seeds, uint64_t n, float* out) {
  const LightManager* lm = static_cast<Scene*>(s)->GetLightManager();
  int type = 0;
  for (uint64_t i = 0; i < n; ++i) {
    RNG rng(seeds[2 * i], seeds[2 * i + 1]);
    const auto r = lm->SampleAllLight(rng);
    float* o = out + 10 * i;
    for (int k = 0; k < 3; ++k) { o[k] = r.v1[k]; o[3 + k] = r.v2[k]; o[6 + k] = r.emission[k]; }
    o[9] = r.pdf;
    type = int(r.light_type);
  }
  return type;
}
REF_API int ref_implicit_light(void* s, uint32_t inst, uint32_t geom, uint32_t prim, float* out4) {
  float3 e(0.f);
  float pdf = 0.f;
  const bool has = static_cast<Scene*>(s)->GetLightManager()->ImplicitAreaLight(inst, geom, prim, &e, &pdf);
  out4[0] = e[0]; out4[1] = e[1]; out4[2] = e[2]; out4[3] = pdf;
  return has ? 1 : 0;
}

// ---------------------------------------------------------------- Render (src/render.h:14-17)
// returns seconds spent inside pbrlab::Render only
REF_API double ref_render(void* s, uint32_t w, uint32_t h, uint32_t spp, float* rgba, uint32_t* count) {
  const Scene* scene = static_cast<Scene*>(s);
  std::atomic_bool cancel(false);
  std::atomic_size_t finish_pass(0);
  RenderLayer layer;
  const auto t0 = std::chrono::steady_clock::now();
  Render(*scene, w, h, spp, cancel, &layer, &finish_pass);
  const auto t1 = std::chrono::steady_clock::now();
  if (rgba) memcpy(rgba, layer.rgba.data(), sizeof(float) * layer.rgba.size());
  if (count) memcpy(count, layer.count.data(), sizeof(uint32_t) * layer.count.size());
  return std::chrono::duration<double>(t1 - t0).count();
}

// ---------------------------------------------------------------- loaders
REF_API void* ref_obj_load(const char* path) {
  ObjData* d = new ObjData();
  if (!io::LoadTriangleMeshFromObj(path, &d->meshes, &d->materials, &d->textures)) {
    delete d;
    return nullptr;
  }
  return d;
}
REF_API void ref_obj_free(void* h) { delete static_cast<ObjData*>(h); }
REF_API int ref_obj_num_shapes(void* h) { return int(static_cast<ObjData*>(h)->meshes.size()); }
REF_API int ref_obj_num_materials(void* h) { return int(static_cast<ObjData*>(h)->materials.size()); }
REF_API uint32_t ref_obj_shape_faces(void* h, int i, char* name, int name_cap) {
  const TriangleMesh& m = static_cast<ObjData*>(h)->meshes[size_t(i)];
  if (name) { strncpy(name, m.GetName().c_str(), size_t(name_cap - 1)); name[name_cap - 1] = 0; }
  return m.GetNumFaces();
}
REF_API uint32_t ref_obj_num_vertices(void* h) {
  return static_cast<ObjData*>(h)->meshes.empty() ? 0 : static_cast<ObjData*>(h)->meshes[0].GetNumVertices();
}
REF_API void ref_obj_vertices(void* h, float* xyzw) {
  const auto& v = static_cast<ObjData*>(h)->meshes[0].GetVertices();
  memcpy(xyzw, v.data(), sizeof(float) * v.size());
}
REF_API void ref_obj_shape_ids(void* h, int i, uint32_t* vertex_ids, uint32_t* material_ids) {
  const TriangleMesh& m = static_cast<ObjData*>(h)->meshes[size_t(i)];
  memcpy(vertex_ids, m.GetVertexIds().data(), sizeof(uint32_t) * m.GetVertexIds().size());
  memcpy(material_ids, m.GetMaterials().data(), sizeof(uint32_t) * m.GetNumFaces());
}
// normals are private in TriangleMesh; expose them through the public fetch (src/mesh/triangle-mesh.cc:77-101)
REF_API void ref_obj_shading_normal(void* h, int i, const uint32_t* prim, const float* uv, uint64_t n, float* out) {
  const TriangleMesh& m = static_cast<ObjData*>(h)->meshes[size_t(i)];
  for (uint64_t k = 0; k < n; ++k) {
    const float3 ns = m.FetchShadingNormal(prim[k], uv[2 * k], uv[2 * k + 1]);
    out[3 * k] = ns[0]; out[3 * k + 1] = ns[1]; out[3 * k + 2] = ns[2];
  }
}
REF_API int ref_obj_material(void* h, int i, float* p23, uint32_t* tex2, char* name, int name_cap) {
  const MaterialParameter& mp = static_cast<ObjData*>(h)->materials[size_t(i)];
  if (mp.index() != kCyclesPrincipledBsdfParameter) return int(mp.index());
  const auto& m = mpark::get<kCyclesPrincipledBsdfParameter>(mp);
  PackPrincipled(m, p23);
  tex2[0] = m.base_color_tex_id; tex2[1] = m.subsurface_color_tex_id;
  if (name) { strncpy(name, m.name.c_str(), size_t(name_cap - 1)); name[name_cap - 1] = 0; }
  return 0;
}
// textures the OBJ loader produced (src/io/triangle-mesh-io.cc:80-141): call with pixels == nullptr for the sizes
REF_API int ref_obj_num_textures(void* h) { return int(static_cast<ObjData*>(h)->textures.size()); }
REF_API void ref_obj_texture(void* h, int i, uint32_t* whc, float* pixels) {
  const Texture& t = static_cast<ObjData*>(h)->textures[size_t(i)];
  whc[0] = t.GetWidth(); whc[1] = t.GetHeight(); whc[2] = t.GetChannels();
  if (!pixels) return;
  // Texture keeps its pixels private; read them back through FetchFloatN at the texel origins, where the bilinear
  // weights are exactly (1, 0, 0, 0): px = width * u must land on the integer x, so u = x / width is checked
  for (uint32_t y = 0; y < whc[1]; ++y)
    for (uint32_t x = 0; x < whc[0]; ++x) {
      float tmp[4] = {0.f, 0.f, 0.f, 0.f};
      float u = float(x) / float(whc[0]), v = float(y) / float(whc[1]);
      if (float(whc[0]) * u != float(x)) u = std::nextafter(u, float(whc[0]) * u < float(x) ? 2.0f : -1.0f);
      if (float(whc[1]) * v != float(y)) v = std::nextafter(v, float(whc[1]) * v < float(y) ? 2.0f : -1.0f);
      t.FetchFloatN(u, v, whc[2], tmp);
      for (uint32_t c = 0; c < whc[2]; ++c) pixels[(size_t(y) * whc[0] + x) * whc[2] + c] = tmp[c];
    }
}
// Texture::FetchFloat3 at arbitrary coordinates (src/texture.cc:43-72)
REF_API void ref_obj_texture_fetch3(void* h, int i, const float* uv, uint64_t n, float* out) {
  const Texture& t = static_cast<ObjData*>(h)->textures[size_t(i)];
  for (uint64_t k = 0; k < n; ++k) t.FetchFloat3(uv[2 * k], uv[2 * k + 1], out + 3 * k);
}
// CyHair -> cubic Bezier (src/io/curve-mesh-io.cc:32-121), memory_saving_mode=false as the CLI uses.
// call with vt==nullptr to query sizes.
REF_API int ref_hair_load(const char* path, float* vt, uint64_t* nfloats, uint32_t* idx, uint64_t* nidx) {
  std::vector<float> v;
  std::vector<uint32_t> ind;
  const bool ok = io::LoadCurveMeshAsCubicBezierCurve(path, false, &v, &ind);
  *nfloats = v.size();
  *nidx = ind.size();
  if (vt) memcpy(vt, v.data(), sizeof(float) * v.size());
  if (idx) memcpy(idx, ind.data(), sizeof(uint32_t) * ind.size());
  return ok ? 1 : 0;
}

// ---------------------------------------------------------------- RNG (src/random/rng.h)
REF_API void ref_rng_draws(uint64_t initstate, uint64_t initseq, uint64_t n, float* out) {
  RNG rng(initstate, initseq);
  for (uint64_t i = 0; i < n; ++i) out[i] = rng.Draw();
}

// ---------------------------------------------------------------- fast_math (src/pbrlab_math.h:135-341)
// op: 0 sin 1 cos 2 exp2 3 exp 4 log2 5 log 6 atan2(x=y-arg, y=x-arg) 7 asin 8 sincos(sin) 9 sincos(cos)
REF_API void ref_fastmath(int op, const float* x, const float* y, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {
    float s, c;
    switch (op) {
      case 0: out[i] = fast_math::FastSin(x[i]); break;
      case 1: out[i] = fast_math::FastCos(x[i]); break;
      case 2: out[i] = fast_math::FastExp2(x[i]); break;
      case 3: out[i] = fast_math::FastExp(x[i]); break;
      case 4: out[i] = fast_math::FastLog2(x[i]); break;
      case 5: out[i] = fast_math::FastLog(x[i]); break;
      case 6: out[i] = fast_math::FastAtan2(x[i], y[i]); break;
      case 7: out[i] = fast_math::FastAsin(x[i]); break;
      case 8: fast_math::FastSincos(x[i], &s, &c); out[i] = s; break;
      case 9: fast_math::FastSincos(x[i], &s, &c); out[i] = c; break;
      default: out[i] = 0.f;
    }
  }
}

// ---------------------------------------------------------------- sampling utils / closures
REF_API void ref_cosine_hemisphere(const float* u, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {
    const float3 w = CosineSampleHemisphere(u[2 * i], u[2 * i + 1]);
    out[3 * i] = w[0]; out[3 * i + 1] = w[1]; out[3 * i + 2] = w[2];
  }
}
REF_API void ref_uniform_sphere(const float* u, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {
    const float3 w = UniformSampleSphere(u[2 * i], u[2 * i + 1]);
    out[3 * i] = w[0]; out[3 * i + 1] = w[1]; out[3 * i + 2] = w[2];
  }
}
REF_API void ref_power_heuristic(const float* a, const float* b, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = PowerHeuristicWeight(a[i], b[i]);
}
REF_API void ref_fresnel_dielectric_cos(const float* c, const float* eta, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = FresnelDielectricCos(c[i], eta[i]);
}
// out n x 2 (f, pdf)
REF_API void ref_ggx_eval(const float* wi, const float* wo, float ax, float ay, int distrib, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {
    float pdf = 0.f;
    const float f = MicrofacetGGXBsdfPdf(float3(wi + 3 * i), float3(wo + 3 * i), ax, ay, distrib, &pdf);
    out[2 * i] = f; out[2 * i + 1] = pdf;
  }
}
// out n x 5 (wi3, f, pdf); wi pre-set to 0 as the caller does (cycles-principled-shader.cc:463)
REF_API void ref_ggx_sample(const float* wo, float ax, float ay, const float* u, int distrib, uint64_t n, float* out) {
  for (uint64_t i = 0; i < n; ++i) {
    float3 wi(0.f);
    float pdf = 0.f;
    const std::array<float, 2> uu = {u[2 * i], u[2 * i + 1]};
    const float f = MicrofacetGGXSample(float3(wo + 3 * i), ax, ay, uu, false, distrib, &wi, &pdf);
    out[5 * i] = wi[0]; out[5 * i + 1] = wi[1]; out[5 * i + 2] = wi[2]; out[5 * i + 3] = f; out[5 * i + 4] = pdf;
  }
}

// Principled: params (23 floats, see UnpackPrincipled) -> closure set (ParamToBsdf) -> EvalBsdf.
// out n x 4 (f3, pdf).  bsdf_out (optional, 40 floats): the CyclesPrincipledBsdf struct fields in order.
REF_API void ref_principled_eval(const float* p23, const float* wi, const float* wo, uint64_t n, float* out,
                                 float* bsdf_out) {
  static Scene dummy_scene;  // only consulted for textures; ids are -1 here
  MaterialParameter mp = UnpackPrincipled(p23);
  SurfaceInfo si = {};
  si.material_param = &mp;
  si.texcoord = float2(0.f, 0.f);
  const CyclesPrincipledBsdf bsdf = ParamToBsdf(dummy_scene, si);
  if (bsdf_out) {
    float* b = bsdf_out;
    *b++ = bsdf.enable_diffuse; for (int k = 0; k < 3; ++k) *b++ = bsdf.diffuse_weight[k];
    *b++ = bsdf.enable_subsurface;
    for (int k = 0; k < 3; ++k) *b++ = bsdf.subsurface_weight[k];
    for (int k = 0; k < 3; ++k) *b++ = bsdf.subsurface_albedo[k];
    for (int k = 0; k < 3; ++k) *b++ = bsdf.subsurface_radius[k];
    *b++ = bsdf.enable_specular; for (int k = 0; k < 3; ++k) *b++ = bsdf.specular_weight[k];
    *b++ = bsdf.alpha_x; *b++ = bsdf.alpha_y; *b++ = bsdf.ior;
    for (int k = 0; k < 3; ++k) *b++ = bsdf.specular_color[k];
    *b++ = bsdf.enable_clearcoat; for (int k = 0; k < 3; ++k) *b++ = bsdf.clearcoat_weight[k];
    *b++ = bsdf.clearcoat_alpha_x; *b++ = bsdf.clearcoat_alpha_y; *b++ = bsdf.clearcoat_ior;
    for (int k = 0; k < 3; ++k) *b++ = bsdf.clearcoat_color[k];
  }
  for (uint64_t i = 0; i < n; ++i) {
    float3 f(0.f);
    float pdf = 0.f;
    EvalBsdf(float3(wi + 3 * i), float3(wo + 3 * i), bsdf, &f, &pdf);
    out[4 * i] = f[0]; out[4 * i + 1] = f[1]; out[4 * i + 2] = f[2]; out[4 * i + 3] = pdf;
  }
}
// closure sample weights (cycles-principled-shader.cc:63-112): out n x 4
REF_API void ref_principled_weights(const float* p23, const float* wo, uint64_t n, float* out) {
  static Scene dummy_scene;
  MaterialParameter mp = UnpackPrincipled(p23);
  SurfaceInfo si = {};
  si.material_param = &mp;
  si.texcoord = float2(0.f, 0.f);
  const CyclesPrincipledBsdf bsdf = ParamToBsdf(dummy_scene, si);
  for (uint64_t i = 0; i < n; ++i) {
    const CyclesSampleWeight w = FetchClosureSampleWeight(float3(wo + 3 * i), bsdf);
    out[4 * i] = w.diffuse_sample_weight; out[4 * i + 1] = w.subsurface_sample_weight;
    out[4 * i + 2] = w.specular_sample_weight; out[4 * i + 3] = w.clearcoat_sample_weight;
  }
}

// Hair: params (20 floats, see UnpackHair), h in [-1,1].  out n x 4 (f*cos 3, pdf)
REF_API void ref_hair_eval(const float* p20, const float* h, const float* wi, const float* wo, uint64_t n, float* out) {
  const HairBsdfParameter mp = UnpackHair(p20);
  for (uint64_t i = 0; i < n; ++i) {
    const HairBsdf b = ParamToBsdf(mp, h[i]);
    float pdf = 0.f;
    const float3 f = hair_bsdf::EnergyConservingHairBsdfCosPdf(float3(wi + 3 * i), float3(wo + 3 * i), b.h, b.v, b.s,
                                                               b.sigma_a, b.eta, b.alpha, b.tints,
                                                               b.transparent_scale, &pdf);
    out[4 * i] = f[0]; out[4 * i + 1] = f[1]; out[4 * i + 2] = f[2]; out[4 * i + 3] = pdf;
  }
}
// out n x 7 (wi3, f*cos 3, pdf)
REF_API void ref_hair_sample(const float* p20, const float* h, const float* wo, const float* us, uint64_t n, float* out) {
  const HairBsdfParameter mp = UnpackHair(p20);
  for (uint64_t i = 0; i < n; ++i) {
    const HairBsdf b = ParamToBsdf(mp, h[i]);
    float pdf = 0.f;
    float3 wi(0.f);
    const std::array<float, 4> u = {us[4 * i], us[4 * i + 1], us[4 * i + 2], us[4 * i + 3]};
    const float3 f = hair_bsdf::EnergyConservingHairSample(float3(wo + 3 * i), b.h, b.v, b.s, b.sigma_a, b.eta,
                                                           b.alpha, b.tints, b.transparent_scale, u, &wi, &pdf);
    float* o = out + 7 * i;
    o[0] = wi[0]; o[1] = wi[1]; o[2] = wi[2]; o[3] = f[0]; o[4] = f[1]; o[5] = f[2]; o[6] = pdf;
  }
}
// derived hair constants (hair-shader.cc:100-151): out 9 floats: sigma_a3, v[4], s, alpha
REF_API void ref_hair_setup(const float* p20, float* out) {
  const HairBsdf b = ParamToBsdf(UnpackHair(p20), 0.f);
  out[0] = b.sigma_a[0]; out[1] = b.sigma_a[1]; out[2] = b.sigma_a[2];
  out[3] = b.v[0]; out[4] = b.v[1]; out[5] = b.v[2]; out[6] = b.v[3];
  out[7] = b.s; out[8] = b.alpha;
}

// SSS helpers (src/shader/random-walk-sss.h:111-197). in: albedo3 radius3 weight3; out: sigma_t3 sigma_s3 thr3
REF_API void ref_sss_coefficients(const float* in9, float* out9) {
  float3 st, ss, thr;
  random_walk_sss::ComputeScatteringCoefficient(float3(in9 + 6), float3(in9), float3(in9 + 3), &st, &ss, &thr);
  for (int k = 0; k < 3; ++k) { out9[k] = st[k]; out9[3 + k] = ss[k]; out9[6 + k] = thr[k]; }
}
// in: thr3 sigma_s3 sigma_t3 u2 ; out: distance, channel_pdf3
REF_API void ref_sss_sample_distance(const float* in11, float* out4) {
  float3 cp;
  const std::array<float, 2> u = {in11[9], in11[10]};
  out4[0] = random_walk_sss::SampleScatterDistance(float3(in11), float3(in11 + 3), float3(in11 + 6), u, &cp);
  out4[1] = cp[0]; out4[2] = cp[1]; out4[3] = cp[2];
}

// ---------------------------------------------------------------- output stage of the CLI (pc/pbrlab-cli.cc:47-57)
// sums / count -> pbrlab::LinerToSrgb -> pbrlab::io::WritePNG (quantisation src/io/image-io.cc:200-206): the statements
// of the reference's main(), on caller-supplied RenderLayer contents.  Writes dir/name; returns WritePNG's result.
REF_API int ref_output_stage(const float* rgba_sums, const uint32_t* count, uint32_t width, uint32_t height,
                             const char* name, const char* dir) {
  std::vector<float> color(size_t(width) * height * 4);
  for (size_t i = 0; i < size_t(width) * height; ++i) {
    color[i * 4 + 0] = rgba_sums[i * 4 + 0] / float(count[i]);
    color[i * 4 + 1] = rgba_sums[i * 4 + 1] / float(count[i]);
    color[i * 4 + 2] = rgba_sums[i * 4 + 2] / float(count[i]);
    color[i * 4 + 3] = rgba_sums[i * 4 + 3] / float(count[i]);
  }
  pbrlab::LinerToSrgb(color, width, height, 4, &color);
  return pbrlab::io::WritePNG(name, dir, color, width, height, 4) ? 1 : 0;
}

}  // extern "C"
