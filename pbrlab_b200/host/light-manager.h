// Area-light tables of the public API (reference src/light-manager.h:14-193, src/light-manager.cc:29-184):
// per light mesh a primitive CDF proportional to max(emission) * area, a global light CDF proportional to the
// summed intensity.  The device samples from the flattened copy of exactly these numbers (ExportTables ->
// pbrgpu_set_lights); the host methods are kept for API parity and for the CPU tests.
#ifndef PBRLAB_B200_LIGHT_MANAGER_H_
#define PBRLAB_B200_LIGHT_MANAGER_H_
#include <memory>
#include <vector>

#include "light-param.h"
#include "mesh-instance.h"
#include "random/rng.h"
#include "type.h"

namespace pbrlab {

class LightManager {
public:
  struct SampledLight {
    LightType light_type;
    float3 v1, v2, emission;
    float pdf{0.0f};
  };
  // flattened copy for the device: see include/pbrgpu.h pbrgpu_light_tables
  struct Tables {
    std::vector<float> light_probability, light_cdf;
    std::vector<uint32_t> light_prim_offset;           // num_lights + 1
    std::vector<uint32_t> light_instance, light_geom;  // where each light lives
    std::vector<float> prim_probability, prim_cdf, prim_area_pdf, prim_emission;
    std::vector<uint32_t> prim_is_emissive;
  };

  LightManager() {}

  template <class... Args>
  uint32_t AddLightParam(Args&&... args) {
    light_params_.emplace_back(args...);
    return uint32_t(light_params_.size() - 1);
  }
  void Clear(void);
  void Commit(void);
  void RegisterInstanceMesh(const MeshInstance& instance, const uint32_t instance_id);
  bool ImplicitAreaLight(const uint32_t instance_id, const uint32_t local_geom_id, const uint32_t prim_id,
                         float3* emission, float* pdf) const;
  SampledLight SampleAllLight(const RNG& rng) const;
  Tables ExportTables(void) const;

private:
  struct AreaLight {
    MeshPtr mesh_ptr;
    std::vector<uint32_t> light_param_ids;
    std::vector<float> choose_primitive_probability, cumulative_probability, prim_area_measure_pdf;
    float intensity_sum = 0.f;
    uint32_t global_id = uint32_t(-1);
  };
  struct Light {
    LightType light_type;
    float choose_light_probability;
    uint32_t instance_id, local_id;
  };
  const AreaLightParameter& AreaParam(uint32_t id) const { return std::get<kAreaLight>(light_params_[id]); }

  std::vector<Light> lights_;
  std::vector<float> cumulative_probability_;
  std::vector<std::vector<std::unique_ptr<AreaLight>>> area_lights_;
  std::vector<LightParameter> light_params_;
};

}  // namespace pbrlab
#endif  // PBRLAB_B200_LIGHT_MANAGER_H_
