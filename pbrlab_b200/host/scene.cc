#include "scene.h"

#include <array>
#include <thread>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>

#include "../../include/pbrgpu.h"

namespace pbrlab {

Scene::Scene(void) : light_manager_(new LightManager()) {
  for (int k = 0; k < 3; ++k) { bmin_[k] = 0.f; bmax_[k] = 0.f; }
}
Scene::~Scene(void) {
  if (ctx_) pbrgpu_destroy(ctx_);
}

uint32_t Scene::CreateLocalScene(void) {
  local_scenes_.emplace_back(new LocalScene);
  return uint32_t(local_scenes_.size() - 1);
}

uint32_t Scene::AddMeshToLocalScene(const uint32_t local_scene_id, const MeshPtr& mesh_ptr) {
  LocalScene* ls = local_scenes_.at(local_scene_id).get();
  ls->meshes.push_back(mesh_ptr);
  return uint32_t(ls->meshes.size() - 1);   // local geom id, as Embree's attach-by-id counter would give
}

// reference src/scene.cc:106-158
uint32_t Scene::CreateInstance(const uint32_t local_scene_id, const float transform[4][4]) {
  MeshInstance inst;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      inst.transform_lg[i][j] = transform ? transform[i][j] : (i == j ? 1.f : 0.f);
      inst.transform_gl[i][j] = inst.transform_lg[i][j];   // identity in every front-end; inverse not needed here
    }
  inst.local_scene = local_scenes_.at(local_scene_id);
  for (const MeshPtr& m : inst.local_scene->meshes) {
    if (m.index() == kTriangleMesh) inst.material_ids.emplace_back(std::get<kTriangleMesh>(m)->GetMaterials());
    else inst.material_ids.emplace_back(std::get<kCubicBezierCurveMesh>(m)->GetMaterials());
    inst.light_param_ids.emplace_back();
  }
  instances_.push_back(inst);
  return uint32_t(instances_.size() - 1);
}

void Scene::AttachLightParamIdsToInstance(const uint32_t instance_id,
                                          const std::vector<std::vector<uint32_t>>& light_param_ids) {
  MeshInstance& inst = instances_.at(instance_id);
  if (inst.light_param_ids.size() != light_param_ids.size()) throw std::runtime_error("light param error");
  inst.light_param_ids = light_param_ids;
}

void Scene::AttachMaterialParamIdsToInstance(const uint32_t instance_id,
                                             const std::vector<std::vector<uint32_t>>& material_ids) {
  MeshInstance& inst = instances_.at(instance_id);
  if (inst.local_scene->meshes.size() != material_ids.size()) throw std::runtime_error("material param error");
  for (size_t g = 0; g < material_ids.size(); ++g) {
    if (material_ids[g].size() != GetNumPrimitive(inst.local_scene->meshes[g]))
      throw std::runtime_error("material param error");
  }
  inst.material_ids = material_ids;
}

void Scene::PackMaterials(std::vector<float>* out) const {
  out->assign(material_params_.size() * 28, 0.f);
  for (size_t i = 0; i < material_params_.size(); ++i) {
    pbrgpu_material m;
    memset(&m, 0, sizeof(m));
    m.tex_id[0] = m.tex_id[1] = PBRGPU_INVALID_ID;
    if (material_params_[i].index() == kCyclesPrincipledBsdfParameter) {
      const auto& p = std::get<kCyclesPrincipledBsdfParameter>(material_params_[i]);
      m.type = 0;
      float* q = m.p;
      q[0] = p.base_color[0]; q[1] = p.base_color[1]; q[2] = p.base_color[2];
      q[3] = p.subsurface;
      q[4] = p.subsurface_radius[0]; q[5] = p.subsurface_radius[1]; q[6] = p.subsurface_radius[2];
      q[7] = p.subsurface_color[0]; q[8] = p.subsurface_color[1]; q[9] = p.subsurface_color[2];
      q[10] = p.metallic; q[11] = p.specular; q[12] = p.specular_tint; q[13] = p.roughness;
      q[14] = p.anisotropic; q[15] = p.anisotropic_rotation; q[16] = p.sheen; q[17] = p.sheen_tint;
      q[18] = p.clearcoat; q[19] = p.clearcoat_roughness; q[20] = p.ior; q[21] = p.transmission;
      q[22] = p.transmission_roughness;
      m.tex_id[0] = p.base_color_tex_id;
      m.tex_id[1] = p.subsurface_color_tex_id;
    } else {
      const auto& p = std::get<kHairBsdfParameter>(material_params_[i]);
      m.type = 1;
      float* q = m.p;
      q[0] = (p.coloring_hair == HairBsdfParameter::kMelanin) ? 1.f : 0.f;
      q[1] = p.base_color[0]; q[2] = p.base_color[1]; q[3] = p.base_color[2];
      q[4] = p.melanin; q[5] = p.melanin_redness; q[6] = p.melanin_randomize;
      q[7] = p.roughness; q[8] = p.azimuthal_roughness; q[9] = p.ior; q[10] = p.shift;
      q[11] = p.specular_tint[0]; q[12] = p.specular_tint[1]; q[13] = p.specular_tint[2];
      q[14] = p.second_specular_tint[0]; q[15] = p.second_specular_tint[1]; q[16] = p.second_specular_tint[2];
      q[17] = p.transmission_tint[0]; q[18] = p.transmission_tint[1]; q[19] = p.transmission_tint[2];
    }
    memcpy(out->data() + i * 28, &m, sizeof(m));
  }
}

namespace {
bool IsIdentity(const float m[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      if (m[i][j] != (i == j ? 1.f : 0.f)) return false;
  return true;
}
}  // namespace

void Scene::CommitHostOnly(void) {
  // light tables (reference src/scene.cc:96-104)
  for (uint32_t i = 0; i < instances_.size(); ++i) light_manager_->RegisterInstanceMesh(instances_[i], i);
  light_manager_->Commit();

  FlatScene& f = flat_;
  f = FlatScene();
  std::map<const Attribute*, std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> pools;   // -> vbase,(nbase,tbase)
  std::vector<std::vector<uint32_t>> tri_offset(instances_.size());
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};

  for (uint32_t i = 0; i < instances_.size(); ++i) {
    const MeshInstance& inst = instances_[i];
    const bool identity = IsIdentity(inst.transform_lg);
    tri_offset[i].assign(inst.local_scene->meshes.size(), 0u);
    for (uint32_t g = 0; g < inst.local_scene->meshes.size(); ++g) {
      const MeshPtr& mp = inst.local_scene->meshes[g];
      if (mp.index() == kTriangleMesh) {
        const TriangleMesh& mesh = *std::get<kTriangleMesh>(mp);
        const Attribute* attr = mesh.GetAttribute().get();
        uint32_t vbase, nbase, tbase;
        auto it = identity ? pools.find(attr) : pools.end();
        if (it == pools.end()) {
          vbase = uint32_t(f.verts.size() / 4);
          nbase = uint32_t(f.normals.size() / 4);
          tbase = uint32_t(f.texcoords.size() / 2);
          f.verts.insert(f.verts.end(), attr->vertices.begin(), attr->vertices.end());
          f.normals.insert(f.normals.end(), attr->normals.begin(), attr->normals.end());
          f.texcoords.insert(f.texcoords.end(), attr->texcoords.begin(), attr->texcoords.end());
          if (identity) {
            pools[attr] = {vbase, {nbase, tbase}};
          } else {
            // Embree transforms instanced geometry (row-vector convention, v' = v * M); the reference transforms
            // neither shading normals nor light samples (// TODO transform at src/scene.cc:219).  Positions are baked
            // here, which the device's light sampling (shade.cuh: SampleAllLight) then reads as well: for a transformed
            // EMISSIVE instance NEE samples the transformed surface, the reference the untransformed one — a known
            // deviation; neither front-end creates non-identity transforms.
            for (size_t v = size_t(vbase) * 4; v < f.verts.size(); v += 4) {
              const float x = f.verts[v], y = f.verts[v + 1], z = f.verts[v + 2];
              for (int j = 0; j < 3; ++j)
                f.verts[v + j] = x * inst.transform_lg[0][j] + y * inst.transform_lg[1][j] +
                                 z * inst.transform_lg[2][j] + inst.transform_lg[3][j];
            }
          }
        } else {
          vbase = it->second.first; nbase = it->second.second.first; tbase = it->second.second.second;
        }
        tri_offset[i][g] = uint32_t(f.tri_prim.size());
        const uint32_t nf = mesh.GetNumFaces();
        const auto& vid = mesh.GetVertexIds();
        const auto& nid = mesh.GetNormalIds();
        const auto& tid = mesh.GetTexcoordIds();
        const auto& mat = inst.material_ids[g];
        // sized fill on all host threads (20 M triangles: 260 M element appends otherwise)
        const size_t t0 = f.tri_prim.size();
        f.vidx.resize((t0 + nf) * 3); f.nidx.resize((t0 + nf) * 3); f.tidx.resize((t0 + nf) * 3);
        f.tri_material.resize(t0 + nf); f.tri_instance.resize(t0 + nf); f.tri_geom.resize(t0 + nf);
        f.tri_prim.resize(t0 + nf);
        const unsigned nth = std::max(1u, std::min(nf / 65536u + 1u, std::thread::hardware_concurrency()));
        std::vector<std::array<float, 6>> part(nth);
        std::vector<std::thread> th;
        auto fill = [&](unsigned t) {
          float llo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, lhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
          const uint32_t p0 = uint32_t(uint64_t(nf) * t / nth), p1 = uint32_t(uint64_t(nf) * (t + 1) / nth);
          for (uint32_t p = p0; p < p1; ++p) {
            for (int k = 0; k < 3; ++k) {
              const uint32_t v = vid[p * 3 + k] + vbase;
              f.vidx[(t0 + p) * 3 + k] = v;
              const uint32_t n = nid[p * 3 + k], tt = tid[p * 3 + k];
              f.nidx[(t0 + p) * 3 + k] = (n == uint32_t(-1) ? n : n + nbase);
              f.tidx[(t0 + p) * 3 + k] = (tt == uint32_t(-1) ? tt : tt + tbase);
              for (int c = 0; c < 3; ++c) {
                llo[c] = std::min(llo[c], f.verts[size_t(v) * 4 + c]);
                lhi[c] = std::max(lhi[c], f.verts[size_t(v) * 4 + c]);
              }
            }
            f.tri_material[t0 + p] = (p < mat.size() ? mat[p] : uint32_t(-1));
            f.tri_instance[t0 + p] = i;
            f.tri_geom[t0 + p] = g;
            f.tri_prim[t0 + p] = p;
          }
          for (int c = 0; c < 3; ++c) { part[t][c] = llo[c]; part[t][3 + c] = lhi[c]; }
        };
        for (unsigned t = 1; t < nth; ++t) th.emplace_back(fill, t);
        fill(0);
        for (auto& x : th) x.join();
        for (unsigned t = 0; t < nth; ++t)
          for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], part[t][c]); hi[c] = std::max(hi[c], part[t][3 + c]); }
      } else {
        const CubicBezierCurveMesh& mesh = *std::get<kCubicBezierCurveMesh>(mp);
        const uint32_t base = uint32_t(f.curve_verts.size() / 4);
        f.curve_verts.insert(f.curve_verts.end(), mesh.GetVertices().begin(), mesh.GetVertices().end());
        if (!identity) {
          for (size_t v = size_t(base) * 4; v < f.curve_verts.size(); v += 4) {
            const float x = f.curve_verts[v], y = f.curve_verts[v + 1], z = f.curve_verts[v + 2];
            for (int j = 0; j < 3; ++j)
              f.curve_verts[v + j] = x * inst.transform_lg[0][j] + y * inst.transform_lg[1][j] +
                                     z * inst.transform_lg[2][j] + inst.transform_lg[3][j];
          }
        }
        const auto& idx = mesh.GetIndices();
        const auto& mat = inst.material_ids[g];
        for (uint32_t s = 0; s < idx.size(); ++s) {
          const uint32_t first = idx[s] + base;
          f.curve_first.push_back(first);
          f.curve_material.push_back(s < mat.size() ? mat[s] : uint32_t(-1));
          f.curve_instance.push_back(i);
          f.curve_geom.push_back(g);
          f.curve_prim.push_back(s);
          // scene bounds as Embree computes them for flat curves: bbox of B(0), B(1/4), B(1/2), B(3/4), B(1) enlarged
          // by the largest |radius| among those points (kernels/subdiv/bezier_curve.h:631-640)
          const float* cp = &f.curve_verts[size_t(first) * 4];
          float pl[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, ph[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, rmax = 0.f;
          for (int j = 0; j <= 4; ++j) {
            float q[4];
            if (j < 4) {
              const float t1 = float(j) / 4.0f, t0 = 1.0f - t1;
              const float b0 = t0 * t0 * t0, b1 = 3.0f * t1 * (t0 * t0), b2 = 3.0f * (t1 * t1) * t0, b3 = t1 * t1 * t1;
              for (int c = 0; c < 4; ++c) q[c] = b0 * cp[c] + (b1 * cp[4 + c] + (b2 * cp[8 + c] + b3 * cp[12 + c]));
            } else {
              for (int c = 0; c < 4; ++c) q[c] = cp[12 + c];
            }
            rmax = std::max(rmax, std::fabs(q[3]));
            for (int c = 0; c < 3; ++c) { pl[c] = std::min(pl[c], q[c]); ph[c] = std::max(ph[c], q[c]); }
          }
          // ... then padded by 4 ulp of the largest coordinate (enlarge_bounds, kernels/common/scene_curves.cpp:377-381)
          float size = 0.f;
          for (int c = 0; c < 3; ++c) {
            pl[c] = pl[c] - rmax; ph[c] = ph[c] + rmax;
            size = std::max(size, std::max(std::fabs(pl[c]), std::fabs(ph[c])));
          }
          const float pad = 4.0f * std::numeric_limits<float>::epsilon() * size;
          for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], pl[c] - pad); hi[c] = std::max(hi[c], ph[c] + pad); }
        }
      }
    }
  }

  f.lights = light_manager_->ExportTables();
  f.light_prim_triangle.clear();
  for (size_t l = 0; l < f.lights.light_probability.size(); ++l) {
    const uint32_t base = tri_offset[f.lights.light_instance[l]][f.lights.light_geom[l]];
    const uint32_t n = f.lights.light_prim_offset[l + 1] - f.lights.light_prim_offset[l];
    for (uint32_t p = 0; p < n; ++p) f.light_prim_triangle.push_back(base + p);
  }
  PackMaterials(&f.materials);
  for (const auto& tex : textures_) {
    f.tex_desc.push_back(uint32_t(f.tex_pixels.size()));
    f.tex_desc.push_back(tex->GetWidth());
    f.tex_desc.push_back(tex->GetHeight());
    f.tex_desc.push_back(tex->GetChannels());
    f.tex_pixels.insert(f.tex_pixels.end(), tex->GetPixels().begin(), tex->GetPixels().end());
  }
  for (int c = 0; c < 3; ++c) { f.bmin[c] = bmin_[c] = lo[c]; f.bmax[c] = bmax_[c] = hi[c]; }
}

void Scene::CommitScene(void) {
  CommitHostOnly();
  if (!ctx_) {
    ctx_ = pbrgpu_create(device_ids_.empty() ? nullptr : device_ids_.data(), int(device_ids_.size()));
    if (!ctx_) throw std::runtime_error(std::string("pbrlab: no B200 backend: ") + pbrgpu_last_error(nullptr));
  }
  const FlatScene& f = flat_;
  auto check = [this](int rc) {
    if (rc != PBRGPU_OK) throw std::runtime_error(std::string("pbrlab: device upload failed: ") + pbrgpu_last_error(ctx_));
  };
  std::vector<pbrgpu_texture> tex(f.tex_desc.size() / 4);
  for (size_t i = 0; i < tex.size(); ++i) {
    tex[i].pixels = f.tex_pixels.data() + f.tex_desc[4 * i];
    tex[i].width = f.tex_desc[4 * i + 1]; tex[i].height = f.tex_desc[4 * i + 2]; tex[i].channels = f.tex_desc[4 * i + 3];
    tex[i].reserved = 0;
  }
  check(pbrgpu_set_textures(ctx_, tex.data(), uint32_t(tex.size())));
  check(pbrgpu_set_materials(ctx_, reinterpret_cast<const pbrgpu_material*>(f.materials.data()),
                             uint32_t(f.materials.size() / 28)));
  check(pbrgpu_set_triangles(ctx_, f.verts.data(), uint32_t(f.verts.size() / 4), f.vidx.data(), f.normals.data(),
                             uint32_t(f.normals.size() / 4), f.nidx.data(), f.texcoords.data(),
                             uint32_t(f.texcoords.size() / 2), f.tidx.data(), f.tri_material.data(),
                             f.tri_instance.data(), f.tri_geom.data(), f.tri_prim.data(), f.tri_prim.size()));
  check(pbrgpu_set_curves(ctx_, f.curve_verts.data(), uint32_t(f.curve_verts.size() / 4), f.curve_first.data(),
                          f.curve_material.data(), f.curve_instance.data(), f.curve_geom.data(), f.curve_prim.data(),
                          f.curve_prim.size()));
  pbrgpu_light_tables lt;
  memset(&lt, 0, sizeof(lt));
  lt.num_lights = uint32_t(f.lights.light_probability.size());
  lt.light_probability = f.lights.light_probability.data();
  lt.light_cdf = f.lights.light_cdf.data();
  lt.light_prim_offset = f.lights.light_prim_offset.data();
  lt.num_light_prims = uint32_t(f.lights.prim_probability.size());
  lt.prim_probability = f.lights.prim_probability.data();
  lt.prim_cdf = f.lights.prim_cdf.data();
  lt.prim_area_pdf = f.lights.prim_area_pdf.data();
  lt.prim_emission = f.lights.prim_emission.data();
  lt.prim_is_emissive = f.lights.prim_is_emissive.data();
  lt.prim_triangle = f.light_prim_triangle.data();
  check(pbrgpu_set_lights(ctx_, &lt));
  check(pbrgpu_commit(ctx_, f.bmin, f.bmax));
}

void Scene::SyncMaterialsToDevice(void) const {
  if (!ctx_) throw std::runtime_error("pbrlab: SyncMaterialsToDevice before CommitScene");
  std::vector<float> packed;
  PackMaterials(&packed);
  if (pbrgpu_set_materials(ctx_, reinterpret_cast<const pbrgpu_material*>(packed.data()),
                           uint32_t(packed.size() / 28)) != PBRGPU_OK)
    throw std::runtime_error(std::string("pbrlab: material upload failed: ") + pbrgpu_last_error(ctx_));
}

const MaterialParameter* Scene::FetchMeshMaterialParameter(const TraceResult& tr) const {
  const std::vector<uint32_t>& ids = instances_[tr.instance_id].material_ids[tr.geom_id];
  if (ids[tr.prim_id] == uint32_t(-1)) return nullptr;
  return &material_params_[ids[tr.prim_id]];
}

float3 Scene::FetchMeshShadingNormal(const TraceResult& tr) const {
  const MeshPtr& m = instances_[tr.instance_id].local_scene->meshes[tr.geom_id];
  if (m.index() == kTriangleMesh) return std::get<kTriangleMesh>(m)->FetchShadingNormal(tr.prim_id, tr.u, tr.v);
  return float3(tr.normal_g);
}

float2 Scene::FetchMeshTexcoord(const TraceResult& tr) const {
  const MeshPtr& m = instances_[tr.instance_id].local_scene->meshes[tr.geom_id];
  if (m.index() == kTriangleMesh) return std::get<kTriangleMesh>(m)->FetchTexcoord(tr.prim_id, tr.u, tr.v);
  return float2(0.f, 0.f);
}

void Scene::FetchSceneAABB(float* bmin, float* bmax) const {
  for (int c = 0; c < 3; ++c) { bmin[c] = bmin_[c]; bmax[c] = bmax_[c]; }
}

TraceResult Scene::TraceFirstHit1(const Ray& ray) const {
  if (!ctx_) throw std::runtime_error("pbrlab: TraceFirstHit1 before CommitScene");
  pbrgpu_ray r = {{ray.ray_org[0], ray.ray_org[1], ray.ray_org[2]}, ray.min_t,
                  {ray.ray_dir[0], ray.ray_dir[1], ray.ray_dir[2]}, ray.max_t};
  pbrgpu_hit h;
  if (pbrgpu_trace(ctx_, &r, 1, &h) != PBRGPU_OK)
    throw std::runtime_error(std::string("pbrlab: trace failed: ") + pbrgpu_last_error(ctx_));
  TraceResult tr;
  if (h.instance_id != PBRGPU_INVALID_ID) {
    for (int c = 0; c < 3; ++c) tr.normal_g[c] = h.normal_g[c];
    tr.t = h.t; tr.u = h.u; tr.v = h.v;
    tr.instance_id = h.instance_id; tr.geom_id = h.geom_id; tr.prim_id = h.prim_id;
  }
  return tr;
}

bool Scene::AnyHit1(const Ray& ray) const {
  if (!ctx_) throw std::runtime_error("pbrlab: AnyHit1 before CommitScene");
  pbrgpu_ray r = {{ray.ray_org[0], ray.ray_org[1], ray.ray_org[2]}, ray.min_t,
                  {ray.ray_dir[0], ray.ray_dir[1], ray.ray_dir[2]}, ray.max_t};
  uint8_t occ = 0;
  if (pbrgpu_occluded(ctx_, &r, 1, &occ) != PBRGPU_OK)
    throw std::runtime_error(std::string("pbrlab: occlusion query failed: ") + pbrgpu_last_error(ctx_));
  return occ != 0;
}

}  // namespace pbrlab
