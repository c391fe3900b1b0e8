// Material parameter structs of the public API: same member names, defaults and variant order as the reference
// (reference src/material-param.h:20-108) so callers (OBJ loader, GUI-style live edits) compile unchanged.
// std::variant (C++17) stands in for the reference's vendored mpark::variant.
#ifndef PBRLAB_B200_MATERIAL_PARAM_H_
#define PBRLAB_B200_MATERIAL_PARAM_H_
#include <cstdint>
#include <string>
#include <variant>

#include "type.h"

namespace pbrlab {

enum MaterialParameterType { kCyclesPrincipledBsdfParameter = 0, kHairBsdfParameter };

struct CyclesPrincipledBsdfParameter {
  float3 base_color = float3(0.8f, 0.8f, 0.8f);
  float subsurface = 0.0f;
  float3 subsurface_radius = float3(1.0f, 1.0f, 1.0f);
  float3 subsurface_color = float3(0.7f, 0.1f, 0.1f);
  float metallic = 0.0f;
  float specular = 0.5f;
  float specular_tint = 0.0f;
  float roughness = 0.5f;
  float anisotropic = 0.0f;
  float anisotropic_rotation = 0.0f;
  float sheen = 0.0f;
  float sheen_tint = 0.5f;
  float clearcoat = 0.0f;
  float clearcoat_roughness = 0.03f;
  float ior = 1.45f;
  float transmission = 0.0f;
  float transmission_roughness = 0.0f;
  uint32_t base_color_tex_id = static_cast<uint32_t>(-1);
  uint32_t subsurface_color_tex_id = static_cast<uint32_t>(-1);
  std::string name;
};

struct HairBsdfParameter {
  enum ColoringHair { kRGB = 0, kMelanin };
  ColoringHair coloring_hair = kMelanin;
  float3 base_color = float3(0.18f, 0.06f, 0.02f);
  float melanin = 0.5f;
  float melanin_redness = 0.8f;
  float melanin_randomize = 0.f;
  float roughness = 0.2f;
  float azimuthal_roughness = 0.3f;
  float ior = 1.55f;
  float shift = 2.f;
  float3 specular_tint = float3(1.f, 1.f, 1.f);
  float3 second_specular_tint = float3(1.f, 1.f, 1.f);
  float3 transmission_tint = float3(1.f, 1.f, 1.f);
  std::string name;
};

using MaterialParameter = std::variant<CyclesPrincipledBsdfParameter, HairBsdfParameter>;

inline void SetMaterialName(const std::string& name, MaterialParameter* m) {
  std::visit([&name](auto& p) { p.name = name; }, *m);
}
inline std::string GetMaterialName(const MaterialParameter& m) {
  return std::visit([](const auto& p) { return p.name; }, m);
}

}  // namespace pbrlab
#endif  // PBRLAB_B200_MATERIAL_PARAM_H_
