#include "light-manager.h"

#include <algorithm>
#include <iostream>

namespace pbrlab {

void LightManager::Clear(void) {
  lights_.clear();
  cumulative_probability_.clear();
  area_lights_.clear();
  light_params_.clear();
}

// reference src/light-manager.cc:79-184
void LightManager::RegisterInstanceMesh(const MeshInstance& instance, const uint32_t instance_id) {
  if (area_lights_.size() <= instance_id) area_lights_.resize(instance_id + 1);
  auto& slot = area_lights_[instance_id];
  const auto& meshes = instance.local_scene->meshes;
  if (slot.size() <= meshes.size()) slot.resize(meshes.size());

  for (uint32_t g = 0; g < meshes.size(); ++g) {
    const std::vector<uint32_t>& ids = instance.light_param_ids[g];
    if (ids.empty() || meshes[g].index() != kTriangleMesh) continue;
    const std::shared_ptr<TriangleMesh>& mesh = std::get<kTriangleMesh>(meshes[g]);
    const uint32_t nf = mesh->GetNumFaces();
    if (nf != ids.size()) std::cerr << "warning: invalid light param ids" << std::endl;
    if (std::find_if(ids.begin(), ids.begin() + nf, [](uint32_t v) { return v != uint32_t(-1); }) == ids.begin() + nf)
      continue;

    std::unique_ptr<AreaLight> al(new AreaLight);
    al->mesh_ptr = meshes[g];
    al->light_param_ids = ids;
    al->choose_primitive_probability.resize(nf);
    for (uint32_t f = 0; f < nf; ++f) {
      float intensity = 0.0f;
      if (ids[f] != uint32_t(-1)) {
        const float3& e = AreaParam(ids[f]).emission;
        intensity = std::max({e[0], e[1], e[2]});
      }
      al->choose_primitive_probability[f] = intensity * mesh->FetchFaceArea(f);
    }
    float sum = 0.0f;   // std::accumulate(..., 0.0f): float running sum in index order
    for (float p : al->choose_primitive_probability) sum = sum + p;
    al->intensity_sum = sum;
    for (float& p : al->choose_primitive_probability) p = p / sum;
    al->cumulative_probability = al->choose_primitive_probability;
    for (uint32_t f = 0; f + 1 < nf; ++f) al->cumulative_probability[f + 1] += al->cumulative_probability[f];
    al->prim_area_measure_pdf.assign(nf, 0.f);
    for (uint32_t f = 0; f < nf; ++f) {
      if (ids[f] != uint32_t(-1)) al->prim_area_measure_pdf[f] = 1.0f / mesh->FetchFaceArea(f);
    }
    slot[g] = std::move(al);
  }
}

// reference src/light-manager.cc:29-77: the normalisation runs in double
void LightManager::Commit(void) {
  lights_.clear();
  cumulative_probability_.clear();
  double intensity_sum = 0.0;
  for (uint32_t i = 0; i < area_lights_.size(); ++i) {
    for (uint32_t g = 0; g < area_lights_[i].size(); ++g) {
      AreaLight* al = area_lights_[i][g].get();
      if (!al) continue;
      al->global_id = uint32_t(lights_.size());
      lights_.push_back({kAreaLight, al->intensity_sum, i, g});
      intensity_sum += double(al->intensity_sum);
    }
  }
  for (Light& l : lights_) l.choose_light_probability = float(double(l.choose_light_probability) / intensity_sum);
  cumulative_probability_.resize(lights_.size());
  for (size_t k = 0; k < lights_.size(); ++k) cumulative_probability_[k] = lights_[k].choose_light_probability;
  for (size_t k = 0; k + 1 < cumulative_probability_.size(); ++k) cumulative_probability_[k + 1] += cumulative_probability_[k];
}

bool LightManager::ImplicitAreaLight(const uint32_t instance_id, const uint32_t local_geom_id, const uint32_t prim_id,
                                     float3* emission, float* pdf) const {
  if (instance_id >= area_lights_.size() || local_geom_id >= area_lights_[instance_id].size()) return false;
  const AreaLight* al = area_lights_[instance_id][local_geom_id].get();
  if (!al || al->light_param_ids[prim_id] == uint32_t(-1)) return false;
  *emission = AreaParam(al->light_param_ids[prim_id]).emission;
  *pdf = lights_[al->global_id].choose_light_probability * al->choose_primitive_probability[prim_id] *
         al->prim_area_measure_pdf[prim_id];
  return true;
}

LightManager::SampledLight LightManager::SampleAllLight(const RNG& rng) const {
  SampledLight ret;
  ret.light_type = kLightNone;
  ret.v1 = ret.v2 = ret.emission = float3(0.f);
  ret.pdf = 0.f;
  if (cumulative_probability_.empty()) return ret;
  const float u0 = rng.Draw();
  size_t li = size_t(std::lower_bound(cumulative_probability_.begin(), cumulative_probability_.end(), u0) -
                     cumulative_probability_.begin());
  li = std::min(li, lights_.size() - 1);
  const Light& light = lights_[li];
  const AreaLight* al = area_lights_[light.instance_id][light.local_id].get();
  const float u1 = rng.Draw();
  size_t pi = size_t(std::lower_bound(al->cumulative_probability.begin(), al->cumulative_probability.end(), u1) -
                     al->cumulative_probability.begin());
  pi = std::min(pi, al->cumulative_probability.size() - 1);
  const float u2 = rng.Draw(), u3 = rng.Draw();
  const bool flag = (u2 > u3);
  const float M = flag ? u2 : u3, m = (!flag) ? u2 : u3;
  const std::shared_ptr<TriangleMesh>& mesh = std::get<kTriangleMesh>(al->mesh_ptr);
  ret.light_type = kAreaLight;
  ret.v1 = mesh->FetchLocalPosition(uint32_t(pi), 1.0f - M, M - m);
  ret.v2 = mesh->FetchGeometryNormal(uint32_t(pi));
  ret.emission = AreaParam(al->light_param_ids[pi]).emission;
  ret.pdf = light.choose_light_probability * al->choose_primitive_probability[pi] * al->prim_area_measure_pdf[pi];
  return ret;
}

LightManager::Tables LightManager::ExportTables(void) const {
  Tables t;
  t.light_prim_offset.push_back(0);
  for (size_t k = 0; k < lights_.size(); ++k) {
    const Light& l = lights_[k];
    const AreaLight* al = area_lights_[l.instance_id][l.local_id].get();
    t.light_probability.push_back(l.choose_light_probability);
    t.light_cdf.push_back(cumulative_probability_[k]);
    t.light_instance.push_back(l.instance_id);
    t.light_geom.push_back(l.local_id);
    const size_t nf = al->choose_primitive_probability.size();
    for (size_t f = 0; f < nf; ++f) {
      const bool em = al->light_param_ids[f] != uint32_t(-1);
      t.prim_probability.push_back(al->choose_primitive_probability[f]);
      t.prim_cdf.push_back(al->cumulative_probability[f]);
      t.prim_area_pdf.push_back(al->prim_area_measure_pdf[f]);
      const float3 e = em ? AreaParam(al->light_param_ids[f]).emission : float3(0.f);
      t.prim_emission.push_back(e[0]); t.prim_emission.push_back(e[1]); t.prim_emission.push_back(e[2]);
      t.prim_is_emissive.push_back(em ? 1u : 0u);
    }
    t.light_prim_offset.push_back(uint32_t(t.prim_probability.size()));
  }
  return t;
}

}  // namespace pbrlab
