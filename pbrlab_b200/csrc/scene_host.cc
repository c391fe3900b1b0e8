// See scene_host.h.
#include "scene_host.h"
#include "bvh_ploc.h"

#include "device/principled.cuh"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>

namespace pbrhost {

bool HostScene::SetTriangles(const float* xyzw, uint32_t nverts, const uint32_t* vidx, const float* nxyzw,
                             uint32_t nnormals, const uint32_t* nidx, const float* uv, uint32_t nuv,
                             const uint32_t* tidx, const uint32_t* material_id, const uint32_t* instance_id,
                             const uint32_t* geom_id, const uint32_t* prim_id, uint64_t ntris) {
  committed = false;
  if (ntris > 0x7fffffffull) { error = "pbrgpu_set_triangles: too many triangles"; return false; }
  if (ntris && (!xyzw || !vidx || !instance_id || !geom_id || !prim_id)) {
    error = "pbrgpu_set_triangles: null array";
    return false;
  }
  verts.resize(nverts);
  if (nverts) memcpy(verts.data(), xyzw, sizeof(F4) * nverts);
  normals.resize(nxyzw ? nnormals : 0);
  if (nxyzw && nnormals) memcpy(normals.data(), nxyzw, sizeof(F4) * nnormals);
  texcoords.resize(uv ? nuv : 0);
  if (uv && nuv) memcpy(texcoords.data(), uv, sizeof(F2) * nuv);
  tri_vidx.resize(ntris); tri_nidx.resize(ntris); tri_tidx.resize(ntris); tri_ids.resize(ntris);
  for (uint64_t i = 0; i < ntris; ++i) {
    const uint32_t a = vidx[3 * i], b = vidx[3 * i + 1], c = vidx[3 * i + 2];
    if (a >= nverts || b >= nverts || c >= nverts) { error = "pbrgpu_set_triangles: vertex index out of range"; return false; }
    tri_vidx[i] = {a, b, c, 0u};
    U4 n = {PBRGPU_INVALID_ID, PBRGPU_INVALID_ID, PBRGPU_INVALID_ID, PBRGPU_INVALID_ID};
    if (nidx && !normals.empty()) {
      n.x = nidx[3 * i]; n.y = nidx[3 * i + 1]; n.z = nidx[3 * i + 2];
      if ((n.x != PBRGPU_INVALID_ID && n.x >= normals.size()) || (n.y != PBRGPU_INVALID_ID && n.y >= normals.size()) ||
          (n.z != PBRGPU_INVALID_ID && n.z >= normals.size())) {
        error = "pbrgpu_set_triangles: normal index out of range";
        return false;
      }
    }
    tri_nidx[i] = n;
    U4 t = {PBRGPU_INVALID_ID, PBRGPU_INVALID_ID, PBRGPU_INVALID_ID, 0u};
    if (tidx && !texcoords.empty()) {
      t.x = tidx[3 * i]; t.y = tidx[3 * i + 1]; t.z = tidx[3 * i + 2];
      if ((t.x != PBRGPU_INVALID_ID && t.x >= texcoords.size()) || (t.y != PBRGPU_INVALID_ID && t.y >= texcoords.size()) ||
          (t.z != PBRGPU_INVALID_ID && t.z >= texcoords.size())) {
        error = "pbrgpu_set_triangles: texcoord index out of range";
        return false;
      }
    }
    tri_tidx[i] = t;
    tri_ids[i] = {instance_id[i], geom_id[i], prim_id[i], material_id ? material_id[i] : PBRGPU_INVALID_ID};
  }
  return true;
}

bool HostScene::SetCurves(const float* xyzr, uint32_t nverts, const uint32_t* first_cp, const uint32_t* material_id,
                          const uint32_t* instance_id, const uint32_t* geom_id, const uint32_t* prim_id,
                          uint64_t nsegs) {
  committed = false;
  if (nsegs > 0x7fffffffull) { error = "pbrgpu_set_curves: too many segments"; return false; }
  if (nsegs && (!xyzr || !first_cp || !instance_id || !geom_id || !prim_id)) {
    error = "pbrgpu_set_curves: null array";
    return false;
  }
  curve_cps.resize(4 * nsegs);
  curve_ids.resize(nsegs);
  for (uint64_t i = 0; i < nsegs; ++i) {
    const uint32_t f = first_cp[i];
    if (uint64_t(f) + 3 >= nverts) { error = "pbrgpu_set_curves: control point index out of range"; return false; }
    memcpy(&curve_cps[4 * i], xyzr + 4 * size_t(f), sizeof(F4) * 4);
    curve_ids[i] = {instance_id[i], geom_id[i], prim_id[i], material_id ? material_id[i] : PBRGPU_INVALID_ID};
  }
  return true;
}

// Which shading queue a material's vertices go to (device/scene_view.cuh: MaterialClass).  A Principled material is
// "diffuse only" when ParamToBsdf (cycles-principled-shader.cc:244-412) enables the Lambert closure and nothing
// else for EVERY hit, i.e. no texture can change the closure set, and its selection weight is exactly 1
// (luma > 0 and finite, so w = x / x).  The specialised kernel runs the same code with the other closures compiled out.
static uint32_t ClassifyMaterial(const pbrgpu_material& m) {
  if (m.type == 1) return pbr::kClassHair;
  if (m.tex_id[0] != PBRGPU_INVALID_ID || m.tex_id[1] != PBRGPU_INVALID_ID) return pbr::kClassGeneral;
  pbr::PrincipledParams pp;
  memcpy(&pp, m.p, 23 * sizeof(float));
  pp.base_color_tex_id = pp.subsurface_color_tex_id = PBRGPU_INVALID_ID;
  const pbr::PrincipledBsdf b =
      pbr::ParamToBsdf(pp, pbr::vec3(pp.base_color[0], pp.base_color[1], pp.base_color[2]),
                       pbr::vec3(pp.subsurface_color[0], pp.subsurface_color[1], pp.subsurface_color[2]));
  if (!b.enable_diffuse || b.enable_subsurface || b.enable_specular || b.enable_clearcoat) return pbr::kClassGeneral;
  const float y = pbr::RgbToY(b.diffuse_weight);
  if (!(y > 0.0f) || !std::isfinite(y) || !std::isfinite(y / y)) return pbr::kClassGeneral;
  return pbr::kClassDiffuse;
}

bool HostScene::SetMaterials(const pbrgpu_material* m, uint32_t n) {
  if (n && !m) { error = "pbrgpu_set_materials: null array"; return false; }
  for (uint32_t i = 0; i < n; ++i) {
    if (m[i].type > 1) { error = "pbrgpu_set_materials: unknown material type"; return false; }
  }
  if (committed) {   // live edit: texture ids must stay inside the committed texture table
    for (uint32_t i = 0; i < n; ++i)
      for (int k = 0; k < 2; ++k)
        if (m[i].type == 0 && m[i].tex_id[k] != PBRGPU_INVALID_ID && m[i].tex_id[k] >= tex_desc.size()) {
          error = "pbrgpu_set_materials: texture id out of range";
          return false;
        }
  }
  materials.assign(m, m + n);
  material_class.resize(n);
  for (uint32_t i = 0; i < n; ++i) material_class[i] = ClassifyMaterial(m[i]);
  return true;
}

// Scene::AddTexture (scene.h:45-51): pixels are copied; channel count 1..4 as Texture stores them (texture.cc:10-41)
bool HostScene::SetTextures(const pbrgpu_texture* t, uint32_t n) {
  committed = false;
  if (n && !t) { error = "pbrgpu_set_textures: null array"; return false; }
  size_t total = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (!t[i].pixels || t[i].width == 0 || t[i].height == 0 || t[i].channels == 0 || t[i].channels > 4) {
      error = "pbrgpu_set_textures: empty texture or bad channel count";
      return false;
    }
    total += size_t(t[i].width) * t[i].height * t[i].channels;
  }
  if (total > 0xffffffffull) { error = "pbrgpu_set_textures: more than 2^32 texels in total"; return false; }
  tex_pixels.resize(total);
  tex_desc.resize(n);
  size_t off = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const size_t cnt = size_t(t[i].width) * t[i].height * t[i].channels;
    memcpy(tex_pixels.data() + off, t[i].pixels, cnt * sizeof(float));
    tex_desc[i] = {uint32_t(off), t[i].width, t[i].height, t[i].channels};
    off += cnt;
  }
  return true;
}

bool HostScene::CheckTextureIds() {
  for (const auto& m : materials)
    for (int k = 0; k < 2; ++k)
      if (m.type == 0 && m.tex_id[k] != PBRGPU_INVALID_ID && m.tex_id[k] >= tex_desc.size()) {
        error = "pbrgpu_commit: texture id out of range (call pbrgpu_set_textures first)";
        return false;
      }
  return true;
}

bool HostScene::SetLights(const pbrgpu_light_tables* t) {
  committed = false;
  light_cdf.clear(); lights.clear(); lprim_cdf.clear(); lprim_info.clear(); lprim_tri.clear();
  if (!t || t->num_lights == 0) return true;
  light_cdf.assign(t->light_cdf, t->light_cdf + t->num_lights);
  lights.resize(t->num_lights);
  for (uint32_t l = 0; l < t->num_lights; ++l) {
    lights[l].choose_probability = t->light_probability[l];
    lights[l].prim_offset = t->light_prim_offset[l];
    lights[l].prim_count = t->light_prim_offset[l + 1] - t->light_prim_offset[l];
    lights[l].pad = 0;
    if (lights[l].prim_count == 0) { error = "pbrgpu_set_lights: light without primitives"; return false; }
  }
  const uint32_t np = t->num_light_prims;
  lprim_cdf.assign(t->prim_cdf, t->prim_cdf + np);
  lprim_tri.assign(t->prim_triangle, t->prim_triangle + np);
  lprim_info.resize(np);
  for (uint32_t l = 0; l < t->num_lights; ++l) {
    for (uint32_t k = lights[l].prim_offset; k < lights[l].prim_offset + lights[l].prim_count; ++k) {
      // pdf = choose_light * choose_prim * 1/area, multiplied in the reference's order (light-manager.h:59-62,151-152)
      const float pdf = t->light_probability[l] * t->prim_probability[k] * t->prim_area_pdf[k];
      lprim_info[k] = {t->prim_emission[3 * k], t->prim_emission[3 * k + 1], t->prim_emission[3 * k + 2], pdf};
    }
  }
  // emissive entries are resolved against the triangles in Commit()
  pending_emissive_.assign(t->prim_is_emissive, t->prim_is_emissive + np);
  return true;
}

void MakeCamera(const float* bmin, const float* bmax, uint32_t width, uint32_t height, float* cam) {
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

namespace {
inline void BezierBasisQuarter(int i, float* b) {   // Embree BezierBasis::eval(float(i)/4) (bezier_curve.h:16-26)
  const float t1 = float(i) / 4.0f, t0 = 1.0f - t1;
  b[0] = t0 * t0 * t0;
  b[1] = 3.0f * t1 * (t0 * t0);
  b[2] = 3.0f * (t1 * t1) * t0;
  b[3] = t1 * t1 * t1;
}
}  // namespace

namespace {
template <class F>
void ParallelFor(int n, F fn) {
  const int nt = std::max(1, std::min(int(std::thread::hardware_concurrency()), n));
  std::vector<std::thread> th;
  std::atomic<int> next{0};
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < n;) fn(i); });
  for (auto& t : th) t.join();
}
// fn(begin, end) over [0, n) in chunks on all host threads (the per-primitive passes of Commit at 20 M triangles)
template <class F>
void ParallelRanges(uint64_t n, F fn) {
  const uint64_t chunk = 1u << 16;
  if (n <= chunk) { fn(uint64_t(0), n); return; }
  const int nchunks = int((n + chunk - 1) / chunk);
  ParallelFor(nchunks, [&](int c) { fn(uint64_t(c) * chunk, std::min<uint64_t>(n, uint64_t(c + 1) * chunk)); });
}
}  // namespace

bool HostScene::BuildBvh(const pbrbvh::Aabb* boxes, uint32_t n, const pbrbvh::BuildParams& prm, pbrbvh::Bvh8* out,
                         std::string* which, bool curves) {
  // PBRGPU_BVH_TRIS / PBRGPU_BVH_CURVES (or PBRGPU_BVH for both) = sah | ploc | auto.  auto: triangles -> PLOC when a
  // device builder is installed (better trees AND a 100x faster build, measured: DESIGN.md §6), curves -> the host SAH
  // builder (PLOC trees over hair are no better and the curve BVHs are small)
  const char* mode_env = getenv(curves ? "PBRGPU_BVH_CURVES" : "PBRGPU_BVH_TRIS");
  if (!mode_env) mode_env = getenv("PBRGPU_BVH");
  const std::string mode = mode_env ? mode_env : "auto";
  uint32_t min_prims = 0u;
  if (const char* e = getenv("PBRGPU_BVH_DEVICE_MIN")) min_prims = uint32_t(std::max(0, atoi(e)));
  uint32_t radius = 8;
  if (const char* e = getenv("PBRGPU_PLOC_RADIUS")) radius = uint32_t(std::min(128, std::max(1, atoi(e))));
  const bool ploc = mode == "ploc" || (mode == "auto" && !curves && n >= min_prims && device_builder != nullptr);
  const char* err = nullptr;
  bool ok;
  if (ploc && device_builder) {
    *which = "ploc-device";
    ok = device_builder(device_builder_user, boxes, n, prm, radius, out, &err);
    if (!ok && getenv("PBRGPU_VERBOSE_COMMIT")) fprintf(stderr, "commit: device builder failed (%s), host SAH builder takes over\n", err ? err : "?");
    if (!ok) { *which = "sah"; ok = pbrbvh::BuildBvh8(boxes, n, prm, out, &err); }   // e.g. a tree too deep to encode
  } else if (ploc) {
    *which = "ploc-host";
    ok = pbrploc::BuildBvh8PlocHost(boxes, n, prm, radius, out, &err);
  } else {
    *which = "sah";
    ok = pbrbvh::BuildBvh8(boxes, n, prm, out, &err);
  }
  if (!ok) error = err ? err : "BVH build failed";
  return ok;
}

bool HostScene::Commit(const float* bmin_in, const float* bmax_in) {
  const auto t0 = std::chrono::steady_clock::now();
  bvh_seconds = 0.0;
  clearance_seconds = 0.0;
  const uint32_t nt = num_tris(), nc = num_curves();
  if (nt == 0 && nc == 0) { error = "pbrgpu_commit: empty scene"; return false; }
  if (!CheckTextureIds()) return false;
  max_material_id = 0;
  any_material_id = false;
  for (auto& id : tri_ids) {
    if (id.w == PBRGPU_INVALID_ID) continue;
    if (id.w >= materials.size()) { error = "pbrgpu_commit: material id out of range"; return false; }
    max_material_id = std::max(max_material_id, id.w); any_material_id = true;
  }
  for (auto& id : curve_ids) {
    if (id.w == PBRGPU_INVALID_ID) continue;
    if (id.w >= materials.size()) { error = "pbrgpu_commit: material id out of range"; return false; }
    max_material_id = std::max(max_material_id, id.w); any_material_id = true;
  }
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  pbrbvh::BuildParams prm;


  // ---- triangles: Embree's TriangleM stores v0, e1 = v0 - v1, e2 = v2 - v0 (kernels/geometry/triangle.h:40-41)
  tri_data.clear();
  if (nt) {
    std::vector<pbrbvh::Aabb> boxes(nt);
    std::mutex lohi_mutex;
    ParallelRanges(nt, [&](uint64_t b0, uint64_t e0) {
      float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
      for (uint64_t i = b0; i < e0; ++i) {
        const F4 &a = verts[tri_vidx[i].x], &b = verts[tri_vidx[i].y], &c = verts[tri_vidx[i].z];
        pbrbvh::Aabb& bx = boxes[i];
        bx.lo[0] = std::min(a.x, std::min(b.x, c.x)); bx.hi[0] = std::max(a.x, std::max(b.x, c.x));
        bx.lo[1] = std::min(a.y, std::min(b.y, c.y)); bx.hi[1] = std::max(a.y, std::max(b.y, c.y));
        bx.lo[2] = std::min(a.z, std::min(b.z, c.z)); bx.hi[2] = std::max(a.z, std::max(b.z, c.z));
        for (int k = 0; k < 3; ++k) { llo[k] = std::min(llo[k], bx.lo[k]); lhi[k] = std::max(lhi[k], bx.hi[k]); }
      }
      std::lock_guard<std::mutex> lock(lohi_mutex);
      for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], llo[k]); hi[k] = std::max(hi[k], lhi[k]); }
    });
    const auto tb0 = std::chrono::steady_clock::now();
    if (!BuildBvh(boxes.data(), nt, prm, &tri_bvh, &last_builder, false)) return false;
    bvh_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count();
    if (getenv("PBRGPU_VERBOSE_COMMIT"))
      fprintf(stderr, "commit: triangle BVH (%u prims) %.3f s\n", nt, std::chrono::duration<double>(std::chrono::steady_clock::now() - tb0).count());
    tri_data.resize(size_t(3) * nt);
    ParallelRanges(nt, [&](uint64_t b0, uint64_t e0) {
      for (uint64_t k = b0; k < e0; ++k) {
        const uint32_t i = tri_bvh.prim_order[k];
        const F4 &a = verts[tri_vidx[i].x], &b = verts[tri_vidx[i].y], &c = verts[tri_vidx[i].z];
        float idbits;
        memcpy(&idbits, &i, 4);
        tri_data[3 * k + 0] = {a.x, a.y, a.z, idbits};
        float matbits;   // material id rides in the spare lane of e1: the closest-hit kernel routes by material class
        memcpy(&matbits, &tri_ids[i].w, 4);
        tri_data[3 * k + 1] = {a.x - b.x, a.y - b.y, a.z - b.z, matbits};
        tri_data[3 * k + 2] = {c.x - a.x, c.y - a.y, c.z - a.z, 0.f};
      }
    });
  } else {
    tri_bvh = pbrbvh::Bvh8();
  }

  // ---- curves.  Embree's ribbon intersector cuts a segment at u = 0, 1/4, .. 1 into four ray-facing quads; each
  // quad lies inside the capsule of radius max(r_a, r_b) around the chord between consecutive cut points.  The BVH is
  // built over PARTS of segments (curve_split = 1, 2 or 4 parts of 4, 2 or 1 quads): boxes of thin diagonal segments
  // are mostly empty and hair is nothing but such segments; four parts have a quarter of the surface area.  A part
  // is tested like the whole segment, restricted to its quads, so the hits are those of the unsplit segment.
  curve_data.clear(); curve_prim.clear(); curve_sub.clear();
  if (nc) {
    int split = 4;
    if (const char* e = getenv("PBRGPU_CURVE_SPLIT")) split = atoi(e);
    if (split != 1 && split != 2 && split != 4) split = 4;
    if (uint64_t(nc) * uint64_t(split) > 0x3fffffffull) split = 1;
    if (nc > 0x1fffffffu) { error = "pbrgpu_commit: too many curve segments"; return false; }
    const uint32_t quads_per_part = uint32_t(4 / split);
    curve_part_quads = quads_per_part;
    std::vector<pbrbvh::Aabb> boxes(size_t(nc) * split);
    for (uint32_t i = 0; i < nc; ++i) {
      const F4* cp = &curve_cps[4 * size_t(i)];
      // Embree accurateFlatBounds(4): bbox of B(0), B(1/4), B(1/2), B(3/4) and v3, enlarged by the largest |r| of
      // those points (kernels/subdiv/bezier_curve.h:631-640) — this is what feeds rtcGetSceneBounds, hence the camera
      float ebl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ebh[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, er = 0.f;
      float pts[5][4];
      float mag = 0.f;
      for (int j = 0; j <= 4; ++j) {
        float* p = pts[j];
        if (j < 4) {
          float b[4];
          BezierBasisQuarter(j, b);
          const float* c0 = &cp[0].x; const float* c1 = &cp[1].x; const float* c2 = &cp[2].x; const float* c3 = &cp[3].x;
          for (int k = 0; k < 4; ++k) p[k] = b[0] * c0[k] + (b[1] * c1[k] + (b[2] * c2[k] + b[3] * c3[k]));
        } else {
          p[0] = cp[3].x; p[1] = cp[3].y; p[2] = cp[3].z; p[3] = cp[3].w;
        }
        er = std::max(er, std::fabs(p[3]));
        for (int k = 0; k < 3; ++k) {
          ebl[k] = std::min(ebl[k], p[k]); ebh[k] = std::max(ebh[k], p[k]);
          mag = std::max(mag, std::fabs(p[k]));
        }
      }
      // BVH boxes: capsule bound per quad + a margin for the float evaluation of the cut points in ray space
      const float slack = 16.0f * FLT_EPSILON * mag;
      for (int part = 0; part < split; ++part) {
        pbrbvh::Aabb& bx = boxes[size_t(i) * split + part];
        for (int k = 0; k < 3; ++k) { bx.lo[k] = FLT_MAX; bx.hi[k] = -FLT_MAX; }
        for (uint32_t q = part * quads_per_part; q < (part + 1) * quads_per_part; ++q) {
          const float r = std::max(std::fabs(pts[q][3]), std::fabs(pts[q + 1][3])) * 1.0001f + slack;
          for (int k = 0; k < 3; ++k) {
            bx.lo[k] = std::min(bx.lo[k], std::min(pts[q][k], pts[q + 1][k]) - r);
            bx.hi[k] = std::max(bx.hi[k], std::max(pts[q][k], pts[q + 1][k]) + r);
          }
        }
      }
      float size = 0.f;   // enlarge_bounds: + 4 ulp of the largest coordinate (kernels/common/scene_curves.cpp:377-381)
      for (int k = 0; k < 3; ++k) {
        ebl[k] = ebl[k] - er; ebh[k] = ebh[k] + er;
        size = std::max(size, std::max(std::fabs(ebl[k]), std::fabs(ebh[k])));
      }
      const float pad = 4.0f * FLT_EPSILON * size;
      for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], ebl[k] - pad); hi[k] = std::max(hi[k], ebh[k] + pad); }
    }
    const uint32_t nparts = nc * uint32_t(split);
    {
      std::string which;
      const auto tcb = std::chrono::steady_clock::now();
      // a curve candidate costs more than a triangle (rejection test, then now and again the ribbon test): smaller
      // leaves.  Sweep (profiles/r3i_tune_curve_sah.log): prim_cost 0.6 -> 2.5 is +2.5 % on C4, neutral on C3.
      pbrbvh::BuildParams cprm = prm;
      cprm.prim_cost = 2.5f;
      if (const char* e = getenv("PBRGPU_CURVE_PRIM_COST")) cprm.prim_cost = std::max(0.01f, float(atof(e)));
      if (const char* e = getenv("PBRGPU_CURVE_LEAF")) cprm.max_leaf_prims = std::min(3, std::max(1, atoi(e)));
      if (!BuildBvh(boxes.data(), nparts, cprm, &curve_bvh, &which, true)) return false;
      bvh_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tcb).count();
      if (!nt) last_builder = which;
    }
    // segment storage ("slots") in order of first appearance in the leaves: neighbours in space are neighbours in memory
    std::vector<uint32_t> slot_of(nc, 0xffffffffu);
    curve_prim.reserve(nc);
    curve_sub.resize(nparts);
    for (uint32_t k = 0; k < nparts; ++k) {
      const uint32_t part = curve_bvh.prim_order[k];
      const uint32_t i = part / uint32_t(split), j = part % uint32_t(split);
      if (slot_of[i] == 0xffffffffu) { slot_of[i] = uint32_t(curve_prim.size()); curve_prim.push_back(i); }
      curve_sub[k] = (slot_of[i] << 2) | (j * quads_per_part);
    }
    curve_data.resize(size_t(4) * nc);
    for (uint32_t k = 0; k < nc; ++k)
      memcpy(&curve_data[4 * size_t(k)], &curve_cps[4 * size_t(curve_prim[k])], sizeof(F4) * 4);
    // CurveMayHit: capsule radius around the line c0c3 = largest |r| + largest distance of c1, c2 from that line
    // (double arithmetic, then 0.1 % + 1e-6 |coordinates| of slack for the float evaluation on the device)
    curve_cull.assign(size_t(8) * nc, 0.f);
    if (!getenv("PBRGPU_NO_CURVE_CULL")) {
      for (uint32_t k = 0; k < nc; ++k) {
        const F4* cp = &curve_data[4 * size_t(k)];
        const double e[3] = {double(cp[3].x) - cp[0].x, double(cp[3].y) - cp[0].y, double(cp[3].z) - cp[0].z};
        const double ee = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
        double rmax = 0.0, dev = 0.0, mag = 0.0;
        for (int j = 0; j < 4; ++j) {
          rmax = std::max(rmax, std::fabs(double(cp[j].w)));
          mag = std::max(mag, std::max(std::fabs(double(cp[j].x)), std::max(std::fabs(double(cp[j].y)), std::fabs(double(cp[j].z)))));
        }
        float R = std::numeric_limits<float>::infinity();   // degenerate chord: never reject
        if (ee > 1e-24 && ee > 1e-10 * mag * mag) {
          for (int j = 1; j <= 2; ++j) {
            const double a[3] = {double(cp[j].x) - cp[0].x, double(cp[j].y) - cp[0].y, double(cp[j].z) - cp[0].z};
            const double c[3] = {a[1] * e[2] - a[2] * e[1], a[2] * e[0] - a[0] * e[2], a[0] * e[1] - a[1] * e[0]};
            dev = std::max(dev, std::sqrt((c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / ee));
          }
          R = float((rmax + dev) * 1.001 + 1e-6 * mag);
        }
        // the device subtracts in float from these rounded values: (c0, e) only has to describe SOME line within the
        // slack of R, which the 0.1 % + 1e-6 |coordinates| above covers
        float* rec = &curve_cull[8 * size_t(k)];
        rec[0] = cp[0].x; rec[1] = cp[0].y; rec[2] = cp[0].z; rec[3] = R;
        rec[4] = float(e[0]); rec[5] = float(e[1]); rec[6] = float(e[2]); rec[7] = float(std::sqrt(ee) * 1.0001);
      }
    } else {
      curve_cull.clear();
    }
  } else {
    curve_bvh = pbrbvh::Bvh8();
    curve_cull.clear();
    curve_part_quads = 4;
  }

  // ---- emissive triangle entries (LightManager::ImplicitAreaLight)
  emissive.clear();
  for (auto& n : tri_nidx) n.w = PBRGPU_INVALID_ID;
  for (size_t k = 0; k < lprim_tri.size(); ++k) {
    if (lprim_tri[k] >= nt) { error = "pbrgpu_commit: light primitive refers to a missing triangle"; return false; }
    if (k < pending_emissive_.size() && pending_emissive_[k]) {
      tri_nidx[lprim_tri[k]].w = uint32_t(emissive.size());
      emissive.push_back(lprim_info[k]);
    }
  }

  for (int k = 0; k < 3; ++k) {
    bmin[k] = bmin_in ? bmin_in[k] : lo[k];
    bmax[k] = bmax_in ? bmax_in[k] : hi[k];
  }
  const auto tc0 = std::chrono::steady_clock::now();
  BuildClearance();
  clearance_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count();
  if (getenv("PBRGPU_VERBOSE_COMMIT"))
    fprintf(stderr, "commit: clearance field %.3f s, total %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count(),
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  committed = true;
  build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return true;
}

// Clearance field of the region where random walks happen (the bounds of every triangle whose material has
// subsurface enabled): one byte per cell of an isotropic grid = a lower bound, in quarter cells, of the distance from
// any point of the cell to any primitive.  ALL primitives are rasterised — a walk segment is intersected with the
// whole scene (random-walk-sss.h:310) — by their boxes, refined for large triangles by the triangle's plane.  The
// occupancy is dilated by one cell, so the exact Euclidean distance transform between cell centres of the dilated
// set equals the box-to-box distance between a cell and the nearest occupied cell; the outermost layer of the grid
// counts as occupied (primitives outside the grid are not rasterised).
namespace {
// 1-D squared distance transform (lower envelope of parabolas, Felzenszwalb & Huttenlocher 2012) of f[0..n), in place
inline void Edt1D(float* f, int n, float* z, int* v, float* out) {
  const float kBig = 1e20f;
  int k = 0;
  v[0] = 0; z[0] = -kBig; z[1] = kBig;
  for (int q = 1; q < n; ++q) {
    float sx;
    for (;;) {   // z[0] = -kBig is below every finite intersection, so k never drops under 0
      const int p = v[k];
      sx = ((f[q] + float(q) * float(q)) - (f[p] + float(p) * float(p))) / (2.0f * float(q - p));
      if (sx <= z[k]) --k; else break;
    }
    ++k; v[k] = q; z[k] = sx; z[k + 1] = kBig;
  }
  k = 0;
  for (int q = 0; q < n; ++q) {
    while (z[k + 1] < float(q)) ++k;
    const float dq = float(q - v[k]);
    out[q] = dq * dq + f[v[k]];
  }
  for (int q = 0; q < n; ++q) f[q] = out[q];
}

}  // namespace

void HostScene::BuildClearance() {
  clear_dist.clear();
  clear_dims[0] = clear_dims[1] = clear_dims[2] = 0;
  const uint32_t nt = num_tris();
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  bool any = false;
  for (uint32_t i = 0; i < nt; ++i) {
    const uint32_t m = tri_ids[i].w;
    if (m == PBRGPU_INVALID_ID || m >= materials.size()) continue;
    if (materials[m].type != 0 || !(materials[m].p[3] > 1e-3f)) continue;   // subsurface > kClosureWeightCutOff
    any = true;
    const F4* v[3] = {&verts[tri_vidx[i].x], &verts[tri_vidx[i].y], &verts[tri_vidx[i].z]};
    for (int c = 0; c < 3; ++c) {
      lo[0] = std::min(lo[0], v[c]->x); hi[0] = std::max(hi[0], v[c]->x);
      lo[1] = std::min(lo[1], v[c]->y); hi[1] = std::max(hi[1], v[c]->y);
      lo[2] = std::min(lo[2], v[c]->z); hi[2] = std::max(hi[2], v[c]->z);
    }
  }
  if (!any) return;
  // isotropic cells: as fine as the cell budget allows (PBRGPU_CLEAR_CELLS, default 24 Mi cells = 24 MiB in L2)
  double budget = 24.0 * 1048576.0;
  if (const char* e = getenv("PBRGPU_CLEAR_CELLS")) budget = std::max(4096.0, atof(e));
  double ext[3];
  for (int k = 0; k < 3; ++k) ext[k] = std::max(double(hi[k]) - double(lo[k]), 1e-6);
  const double longest = std::max(ext[0], std::max(ext[1], ext[2]));
  for (int k = 0; k < 3; ++k) ext[k] = std::max(ext[k], longest * 0.01);   // a flat region still gets a few layers
  double h = std::cbrt(ext[0] * ext[1] * ext[2] / budget);
  for (;;) {
    double cells = 1.0;
    for (int k = 0; k < 3; ++k) cells *= std::ceil(ext[k] / h) + 4.0;
    if (cells <= budget) break;
    h *= 1.02;
  }
  int D[3];
  for (int k = 0; k < 3; ++k) {
    D[k] = int(std::ceil(ext[k] / h)) + 4;                       // two spare layers on each side
    clear_org[k] = float(0.5 * (double(lo[k]) + double(hi[k])) - 0.5 * double(D[k]) * h);
    clear_dims[k] = uint32_t(D[k]);
  }
  const float cell = float(h);
  clear_inv_cell = 1.0f / cell;
  clear_quantum = 0.25f * cell;
  const size_t ncell = size_t(D[0]) * D[1] * D[2];
  std::vector<uint8_t> occ(ncell, 0);
  auto mark_box = [&](const float* blo, const float* bhi, const float* plane /* n.xyz, d or null */) {
    int a[3], b[3];
    for (int k = 0; k < 3; ++k) {
      const float fa = (blo[k] - clear_org[k]) * clear_inv_cell - 0.01f, fb = (bhi[k] - clear_org[k]) * clear_inv_cell + 0.01f;
      if (fb < 0.f || fa >= float(D[k])) return;
      a[k] = std::max(0, int(std::floor(fa)));
      b[k] = std::min(D[k] - 1, int(std::floor(fb)));
    }
    const bool refine = plane && (size_t(b[0] - a[0] + 1) * (b[1] - a[1] + 1) * (b[2] - a[2] + 1) > 64);
    const float rad = refine ? 0.51f * cell * (std::fabs(plane[0]) + std::fabs(plane[1]) + std::fabs(plane[2])) : 0.f;
    for (int z = a[2]; z <= b[2]; ++z)
      for (int y = a[1]; y <= b[1]; ++y) {
        uint8_t* row = &occ[(size_t(z) * D[1] + y) * D[0]];
        if (!refine) {
          for (int x = a[0]; x <= b[0]; ++x) row[x] = 1;
        } else {
          const float cy = clear_org[1] + (y + 0.5f) * cell, cz = clear_org[2] + (z + 0.5f) * cell;
          for (int x = a[0]; x <= b[0]; ++x) {
            const float cx = clear_org[0] + (x + 0.5f) * cell;
            if (std::fabs(plane[0] * cx + plane[1] * cy + plane[2] * cz + plane[3]) <= rad) row[x] = 1;
          }
        }
      }
  };
  // (all host threads: cells are only ever set to 1, so concurrent writers cannot disagree)
  ParallelRanges(nt, [&](uint64_t tb, uint64_t te) {
  for (uint64_t i = tb; i < te; ++i) {
    const F4 &a = verts[tri_vidx[i].x], &b = verts[tri_vidx[i].y], &c = verts[tri_vidx[i].z];
    const float blo[3] = {std::min(a.x, std::min(b.x, c.x)), std::min(a.y, std::min(b.y, c.y)), std::min(a.z, std::min(b.z, c.z))};
    const float bhi[3] = {std::max(a.x, std::max(b.x, c.x)), std::max(a.y, std::max(b.y, c.y)), std::max(a.z, std::max(b.z, c.z))};
    const float e1[3] = {b.x - a.x, b.y - a.y, b.z - a.z}, e2[3] = {c.x - a.x, c.y - a.y, c.z - a.z};
    float n[4] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0], 0.f};
    const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (len > 0.f) {
      n[0] /= len; n[1] /= len; n[2] /= len;
      n[3] = -(n[0] * a.x + n[1] * a.y + n[2] * a.z);
      mark_box(blo, bhi, n);
    } else {
      mark_box(blo, bhi, nullptr);
    }
  }
  });
  for (uint32_t i = 0; i < num_curves(); ++i) {
    const F4* cp = &curve_cps[4 * size_t(i)];
    float blo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, r = 0.f;
    for (int c = 0; c < 4; ++c) {
      r = std::max(r, std::fabs(cp[c].w));
      blo[0] = std::min(blo[0], cp[c].x); bhi[0] = std::max(bhi[0], cp[c].x);
      blo[1] = std::min(blo[1], cp[c].y); bhi[1] = std::max(bhi[1], cp[c].y);
      blo[2] = std::min(blo[2], cp[c].z); bhi[2] = std::max(bhi[2], cp[c].z);
    }
    for (int k = 0; k < 3; ++k) { blo[k] -= r; bhi[k] += r; }   // convex hull of the control points grown by the radius
    mark_box(blo, bhi, nullptr);
  }
  // the outermost layer vouches for nothing: whatever lies outside the grid was not rasterised
  for (int z = 0; z < D[2]; ++z)
    for (int y = 0; y < D[1]; ++y) {
      uint8_t* row = &occ[(size_t(z) * D[1] + y) * D[0]];
      if (z == 0 || z == D[2] - 1 || y == 0 || y == D[1] - 1) memset(row, 1, size_t(D[0]));
      else row[0] = row[D[0] - 1] = 1;
    }
  // dilate by one cell (separable 3-wide maximum), then the squared distance transform along x, y, z
  const float kFar = 1e12f;
  std::vector<float> d2(ncell);
  const int maxd = std::max(D[0], std::max(D[1], D[2]));
  // pass x: dilation along x folded into the seed (a cell is a seed if it or an x-neighbour is occupied after the
  // y/z dilation below) -> do the y/z dilation first
  {
    std::vector<uint8_t> tmp(ncell);
    ParallelFor(D[2], [&](int z) {
      for (int y = 0; y < D[1]; ++y)
        for (int x = 0; x < D[0]; ++x) {
          const size_t i = (size_t(z) * D[1] + y) * D[0] + x;
          uint8_t v = occ[i];
          if (y > 0) v |= occ[i - D[0]];
          if (y + 1 < D[1]) v |= occ[i + D[0]];
          tmp[i] = v;
        }
    });
    const size_t sz = size_t(D[0]) * D[1];
    ParallelFor(D[2], [&](int z) {
      for (size_t j = 0; j < sz; ++j) {
        const size_t i = size_t(z) * sz + j;
        uint8_t v = tmp[i];
        if (z > 0) v |= tmp[i - sz];
        if (z + 1 < D[2]) v |= tmp[i + sz];
        occ[i] = v;
      }
    });
    ParallelFor(D[2], [&](int z) {
      for (int y = 0; y < D[1]; ++y) {
        const size_t r = (size_t(z) * D[1] + y) * D[0];
        for (int x = 0; x < D[0]; ++x) {
          uint8_t v = occ[r + x];
          if (x > 0) v |= occ[r + x - 1];
          if (x + 1 < D[0]) v |= occ[r + x + 1];
          d2[r + x] = v ? 0.f : kFar;
        }
      }
    });
  }
  ParallelFor(D[2], [&](int z) {   // along x
    std::vector<float> zb(maxd + 1), ob(maxd); std::vector<int> vb(maxd);
    for (int y = 0; y < D[1]; ++y) Edt1D(&d2[(size_t(z) * D[1] + y) * D[0]], D[0], zb.data(), vb.data(), ob.data());
  });
  ParallelFor(D[2], [&](int z) {   // along y
    std::vector<float> zb(maxd + 1), ob(maxd), col(maxd); std::vector<int> vb(maxd);
    for (int x = 0; x < D[0]; ++x) {
      float* base = &d2[size_t(z) * D[1] * D[0] + x];
      for (int y = 0; y < D[1]; ++y) col[y] = base[size_t(y) * D[0]];
      Edt1D(col.data(), D[1], zb.data(), vb.data(), ob.data());
      for (int y = 0; y < D[1]; ++y) base[size_t(y) * D[0]] = col[y];
    }
  });
  ParallelFor(D[1], [&](int y) {   // along z
    std::vector<float> zb(maxd + 1), ob(maxd), col(maxd); std::vector<int> vb(maxd);
    const size_t sz = size_t(D[0]) * D[1];
    for (int x = 0; x < D[0]; ++x) {
      float* base = &d2[size_t(y) * D[0] + x];
      for (int z = 0; z < D[2]; ++z) col[z] = base[size_t(z) * sz];
      Edt1D(col.data(), D[2], zb.data(), vb.data(), ob.data());
      for (int z = 0; z < D[2]; ++z) base[size_t(z) * sz] = col[z];
    }
  });
  clear_dist.assign((ncell + 3) / 4, 0u);
  uint8_t* q = reinterpret_cast<uint8_t*>(clear_dist.data());
  ParallelFor(D[2], [&](int z) {
    const size_t sz = size_t(D[0]) * D[1];
    for (size_t i = size_t(z) * sz; i < size_t(z + 1) * sz; ++i) {
      // floor(4 d) with a hair of slack for the float arithmetic of the transform
      const float d = std::sqrt(std::min(d2[i], 1e8f)) * 4.0f * 0.9999f;
      q[i] = uint8_t(std::min(255.0f, std::floor(d)));
    }
  });
  const char* exact = getenv("PBRGPU_CLEAR_EXACT");
  if (!exact || atoi(exact) != 0) RefineClearanceNearSurface(d2.data());
}

// Near the surface the box-to-box bound above gives away one to three cells: the primitive may lie anywhere in its
// (dilated) cell.  The walk segments that reach the ray engine are half a cell long and 90 % of them hit nothing
// (profiles/r7_walk_segment_stats_cpu.log), so this pass replaces the bound of every cell whose centre lies within
// kRefineCells cells of a primitive by  (exact distance from the cell's centre to the nearest primitive) - (half the
// cell's diagonal): any point of the cell is at most half a diagonal from the centre, so it is still a lower bound
// for every point of the cell.  Triangles by their exact point-triangle distance, curve segments by the distance to
// their (radius-grown) bounding box.  A cell no primitive comes within kRefineCells cells of keeps its old bound or
// gets the bound of that radius, whichever is larger.  `scratch`: ncell floats (the squared distances of the
// transform, no longer needed).  The radius is 2.5 cells (3.5 answers 0.3 % more segments for twice the time), 1.5 if
// that does not fit a budget of point-
// primitive distance evaluations, and the pass is skipped when neither does: meshes whose triangles are much smaller than the cells
// (the 20 M-triangle case) keep the box-to-box bound.  PBRGPU_CLEAR_EXACT=0 switches the pass off.  On the Cornell
// scene the field then answers 69 % instead of 46 % of the walk segments (CPU emulation, identical radiance:
// tests/test_emul_parity.py::test_clearance_field_only_skips_segments_that_miss).
namespace {
// squared distance from p to triangle abc (Ericson, Real-Time Collision Detection 5.1.5), double precision
inline double PointTriangleDist2(const double p[3], const double a[3], const double b[3], const double c[3]) {
  double ab[3], ac[3], ap[3];
  for (int k = 0; k < 3; ++k) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; ap[k] = p[k] - a[k]; }
  auto dot = [](const double* x, const double* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
  auto dist2_to = [&](double u, double v) {   // a + u ab + v ac
    double s = 0;
    for (int k = 0; k < 3; ++k) { const double q = a[k] + u * ab[k] + v * ac[k] - p[k]; s += q * q; }
    return s;
  };
  const double d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= 0 && d2 <= 0) return dist2_to(0, 0);
  double bp[3]; for (int k = 0; k < 3; ++k) bp[k] = p[k] - b[k];
  const double d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= 0 && d4 <= d3) return dist2_to(1, 0);
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) return dist2_to(d1 / (d1 - d3), 0);
  double cp[3]; for (int k = 0; k < 3; ++k) cp[k] = p[k] - c[k];
  const double d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= 0 && d5 <= d6) return dist2_to(0, 1);
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) return dist2_to(0, d2 / (d2 - d6));
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); return dist2_to(1 - w, w); }
  const double denom = va + vb + vc;
  if (!(denom > 0)) {   // degenerate triangle: the nearest of its edges' end points and the projections above
    return std::min(dist2_to(0, 0), std::min(dist2_to(1, 0), dist2_to(0, 1)));
  }
  return dist2_to(vb / denom, vc / denom);
}
inline void AtomicMinFloat(std::atomic<uint32_t>* cell, float v) {   // v >= 0: the bit pattern orders like the value
  uint32_t bits;
  memcpy(&bits, &v, 4);
  uint32_t cur = cell->load(std::memory_order_relaxed);
  while (bits < cur && !cell->compare_exchange_weak(cur, bits, std::memory_order_relaxed)) {}
}
}  // namespace

void HostScene::RefineClearanceNearSurface(float* scratch) {
  const int D[3] = {int(clear_dims[0]), int(clear_dims[1]), int(clear_dims[2])};
  const size_t ncell = size_t(D[0]) * D[1] * D[2];
  if (ncell == 0 || clear_dist.empty()) return;
  // work of the pass for a radius: (cell, primitive) pairs it visits
  double budget = 3e8;
  if (const char* e = getenv("PBRGPU_CLEAR_EXACT_BUDGET")) budget = std::max(0.0, atof(e));
  const double inv = double(clear_inv_cell);
  // cells of the grid whose centre can be within `radius` cells of a box [lo, hi] (world units)
  auto cell_range_r = [&](const double* blo, const double* bhi, double radius, int* a, int* b) {
    for (int k = 0; k < 3; ++k) {
      const double fa = (blo[k] - double(clear_org[k])) * inv - radius - 0.5, fb = (bhi[k] - double(clear_org[k])) * inv + radius - 0.5;
      if (fb < 0.0 || fa > double(D[k] - 1)) return false;
      a[k] = std::max(0, int(std::ceil(fa)));
      b[k] = std::min(D[k] - 1, int(std::floor(fb)));
      if (a[k] > b[k]) return false;
    }
    return true;
  };
  auto cells_of = [&](const double* blo, const double* bhi, double radius) -> uint64_t {
    int a[3], b[3];
    if (!cell_range_r(blo, bhi, radius, a, b)) return 0;
    return uint64_t(b[0] - a[0] + 1) * uint64_t(b[1] - a[1] + 1) * uint64_t(b[2] - a[2] + 1);
  };
  auto work_for = [&](double radius) {
    std::atomic<uint64_t> total(0);
    ParallelRanges(num_tris(), [&](uint64_t tb, uint64_t te) {
      uint64_t local = 0;
      for (uint64_t i = tb; i < te; ++i) {
        const F4 &A = verts[tri_vidx[i].x], &B = verts[tri_vidx[i].y], &C = verts[tri_vidx[i].z];
        const double lo3[3] = {std::min(A.x, std::min(B.x, C.x)), std::min(A.y, std::min(B.y, C.y)), std::min(A.z, std::min(B.z, C.z))};
        const double hi3[3] = {std::max(A.x, std::max(B.x, C.x)), std::max(A.y, std::max(B.y, C.y)), std::max(A.z, std::max(B.z, C.z))};
        local += cells_of(lo3, hi3, radius);
      }
      total += local;
    });
    ParallelRanges(num_curves(), [&](uint64_t cb, uint64_t ce) {
      uint64_t local = 0;
      for (uint64_t i = cb; i < ce; ++i) {
        const F4* cp = &curve_cps[4 * size_t(i)];
        double lo3[3] = {1e300, 1e300, 1e300}, hi3[3] = {-1e300, -1e300, -1e300}, r = 0.0;
        for (int c = 0; c < 4; ++c) {
          r = std::max(r, std::fabs(double(cp[c].w)));
          const double v[3] = {cp[c].x, cp[c].y, cp[c].z};
          for (int k = 0; k < 3; ++k) { lo3[k] = std::min(lo3[k], v[k]); hi3[k] = std::max(hi3[k], v[k]); }
        }
        for (int k = 0; k < 3; ++k) { lo3[k] -= r * 1.0001; hi3[k] += r * 1.0001; }
        local += cells_of(lo3, hi3, radius);
      }
      total += local;
    });
    return double(total.load());
  };
  double radius = 0.0;
  for (double r : {2.5, 1.5}) {
    const double wk = work_for(r);
    if (getenv("PBRGPU_VERBOSE_COMMIT")) fprintf(stderr, "commit: clearance exact pass: radius %.1f cells = %.0f M cell visits\n", r, wk * 1e-6);
    if (wk <= budget) { radius = r; break; }
  }
  if (radius == 0.0) {
    if (getenv("PBRGPU_VERBOSE_COMMIT")) fprintf(stderr, "commit: clearance field keeps the box-to-box bound (exact pass over budget)\n");
    return;
  }
  const double kRefineCells = radius;
  const auto tr0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (getenv("PBRGPU_VERBOSE_COMMIT"))
      fprintf(stderr, "commit: clearance exact pass: %s at %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - tr0).count());
  };
  static_assert(sizeof(std::atomic<uint32_t>) == sizeof(float), "atomic<uint32_t> must overlay a float");
  std::atomic<uint32_t>* dmin2 = reinterpret_cast<std::atomic<uint32_t>*>(scratch);   // squared distance in CELLS
  const float kInfCells2 = 1e30f;
  ParallelRanges(ncell, [&](uint64_t b, uint64_t e) {
    uint32_t bits;
    memcpy(&bits, &kInfCells2, 4);
    for (uint64_t i = b; i < e; ++i) dmin2[i].store(bits, std::memory_order_relaxed);
  });
  lap("cleared");
  auto cell_range = [&](const double* blo, const double* bhi, int* a, int* b) { return cell_range_r(blo, bhi, kRefineCells, a, b); };
  const uint32_t nt = num_tris();
  ParallelRanges(nt, [&](uint64_t tb, uint64_t te) {
    for (uint64_t i = tb; i < te; ++i) {
      const F4 &A = verts[tri_vidx[i].x], &B = verts[tri_vidx[i].y], &C = verts[tri_vidx[i].z];
      // (in cell units, relative to the grid origin)
      const double a[3] = {(double(A.x) - clear_org[0]) * inv, (double(A.y) - clear_org[1]) * inv, (double(A.z) - clear_org[2]) * inv};
      const double b[3] = {(double(B.x) - clear_org[0]) * inv, (double(B.y) - clear_org[1]) * inv, (double(B.z) - clear_org[2]) * inv};
      const double c[3] = {(double(C.x) - clear_org[0]) * inv, (double(C.y) - clear_org[1]) * inv, (double(C.z) - clear_org[2]) * inv};
      const double blo[3] = {std::min(double(A.x), std::min(double(B.x), double(C.x))), std::min(double(A.y), std::min(double(B.y), double(C.y))), std::min(double(A.z), std::min(double(B.z), double(C.z)))};
      const double bhi[3] = {std::max(double(A.x), std::max(double(B.x), double(C.x))), std::max(double(A.y), std::max(double(B.y), double(C.y))), std::max(double(A.z), std::max(double(B.z), double(C.z)))};
      int lo[3], hi[3];
      if (!cell_range(blo, bhi, lo, hi)) continue;
      double tlo[3], thi[3];   // the triangle's box in cell units
      for (int k = 0; k < 3; ++k) { tlo[k] = std::min(a[k], std::min(b[k], c[k])); thi[k] = std::max(a[k], std::max(b[k], c[k])); }
      for (int z = lo[2]; z <= hi[2]; ++z)
        for (int y = lo[1]; y <= hi[1]; ++y)
          for (int x = lo[0]; x <= hi[0]; ++x) {
            const double p[3] = {x + 0.5, y + 0.5, z + 0.5};
            // the triangle's box is a lower bound of its distance: most (cell, triangle) pairs end here, either beyond the
            // radius or no nearer than what the cell already holds
            double box2 = 0.0;
            for (int k = 0; k < 3; ++k) {
              const double q = p[k] < tlo[k] ? tlo[k] - p[k] : (p[k] > thi[k] ? p[k] - thi[k] : 0.0);
              box2 += q * q;
            }
            std::atomic<uint32_t>* cellp = &dmin2[(size_t(z) * D[1] + y) * D[0] + x];
            if (box2 > kRefineCells * kRefineCells) continue;
            const uint32_t cur_bits = cellp->load(std::memory_order_relaxed);
            float cur;
            memcpy(&cur, &cur_bits, 4);
            if (box2 >= double(cur)) continue;
            const double d2v = PointTriangleDist2(p, a, b, c);
            if (d2v <= kRefineCells * kRefineCells) AtomicMinFloat(cellp, float(d2v * (1.0 - 1e-6)));
          }
    }
  });
  lap("triangles");
  ParallelRanges(num_curves(), [&](uint64_t cb, uint64_t ce) {
  for (uint64_t i = cb; i < ce; ++i) {   // curve segments: distance to the radius-grown box of the control points
    const F4* cp = &curve_cps[4 * size_t(i)];
    double blo[3] = {1e300, 1e300, 1e300}, bhi[3] = {-1e300, -1e300, -1e300}, r = 0.0;
    for (int c = 0; c < 4; ++c) {
      r = std::max(r, std::fabs(double(cp[c].w)));
      const double v[3] = {cp[c].x, cp[c].y, cp[c].z};
      for (int k = 0; k < 3; ++k) { blo[k] = std::min(blo[k], v[k]); bhi[k] = std::max(bhi[k], v[k]); }
    }
    for (int k = 0; k < 3; ++k) { blo[k] -= r * 1.0001; bhi[k] += r * 1.0001; }
    int lo[3], hi[3];
    if (!cell_range(blo, bhi, lo, hi)) continue;
    for (int z = lo[2]; z <= hi[2]; ++z)
      for (int y = lo[1]; y <= hi[1]; ++y)
        for (int x = lo[0]; x <= hi[0]; ++x) {
          const double p[3] = {x + 0.5, y + 0.5, z + 0.5};
          double d2v = 0.0;
          for (int k = 0; k < 3; ++k) {
            const double a = (blo[k] - double(clear_org[k])) * inv, b = (bhi[k] - double(clear_org[k])) * inv;
            const double q = p[k] < a ? a - p[k] : (p[k] > b ? p[k] - b : 0.0);
            d2v += q * q;
          }
          if (d2v <= kRefineCells * kRefineCells) AtomicMinFloat(&dmin2[(size_t(z) * D[1] + y) * D[0] + x], float(d2v * (1.0 - 1e-6)));
        }
  }
  });
  lap("curves");
  uint8_t* q = reinterpret_cast<uint8_t*>(clear_dist.data());
  ParallelRanges(ncell, [&](uint64_t b, uint64_t e) {
    for (uint64_t i = b; i < e; ++i) {
      const float v2 = scratch[i];                                           // (relaxed stores are complete: the threads were joined)
      const double dc = v2 >= 1e29f ? kRefineCells : std::min(kRefineCells, std::sqrt(double(v2)));   // (nothing within the radius: at least the radius)
      // any point of the cell is within half a diagonal (0.8660254 cells) of the centre; 0.05 % + a hair for the rounding
      const double bound = (dc - 0.8660255) * 0.9995 - 1e-4;
      if (bound <= 0.0) continue;
      const uint8_t nv = uint8_t(std::min(255.0, std::floor(4.0 * bound)));
      if (nv > q[i]) q[i] = nv;
    }
  });
}

pbr::SceneView HostScene::HostView() const {
  pbr::SceneView v;
  memset(&v, 0, sizeof(v));
  v.tri_nodes = reinterpret_cast<const float4*>(tri_bvh.nodes.data());
  v.tri_data = reinterpret_cast<const float4*>(tri_data.data());
  v.curve_nodes = reinterpret_cast<const float4*>(curve_bvh.nodes.data());
  v.curve_data = reinterpret_cast<const float4*>(curve_data.data());
  v.curve_prim = curve_prim.data();
  v.curve_sub = curve_sub.data();
  v.curve_part_quads = curve_part_quads;
  v.curve_cull = curve_cull.empty() ? nullptr : reinterpret_cast<const float4*>(curve_cull.data());
  v.num_tris = num_tris();
  v.num_curves = num_curves();
  v.tri_ids = reinterpret_cast<const uint4*>(tri_ids.data());
  v.tri_nidx = reinterpret_cast<const uint4*>(tri_nidx.data());
  v.tri_vidx = reinterpret_cast<const uint4*>(tri_vidx.data());
  v.tri_tidx = reinterpret_cast<const uint4*>(tri_tidx.data());
  v.verts = reinterpret_cast<const float4*>(verts.data());
  v.normals = reinterpret_cast<const float4*>(normals.data());
  v.texcoords = reinterpret_cast<const float2*>(texcoords.data());
  v.curve_ids = reinterpret_cast<const uint4*>(curve_ids.data());
  v.materials = reinterpret_cast<const pbr::DeviceMaterial*>(materials.data());
  v.num_materials = uint32_t(materials.size());
  v.material_class = material_class.data();
  for (const auto& m : materials) v.num_hair_materials += (m.type == 1u) ? 1u : 0u;
  v.tex_pixels = tex_pixels.data();
  v.tex_desc = tex_desc.data();
  v.num_textures = uint32_t(tex_desc.size());
  v.emissive = reinterpret_cast<const float4*>(emissive.data());
  v.light_cdf = light_cdf.data();
  v.lights = lights.data();
  v.num_lights = uint32_t(lights.size());
  v.lprim_cdf = lprim_cdf.data();
  v.lprim_info = reinterpret_cast<const float4*>(lprim_info.data());
  v.lprim_tri = lprim_tri.data();
  v.clear_dist = clear_dist.empty() ? nullptr : reinterpret_cast<const uint8_t*>(clear_dist.data());
  for (int k = 0; k < 3; ++k) { v.clear_org[k] = clear_org[k]; v.clear_dims[k] = clear_dims[k]; }
  v.clear_inv_cell = clear_inv_cell;
  v.clear_quantum = clear_quantum;
  v.clear_march_steps = 4;
  return v;
}

}  // namespace pbrhost
