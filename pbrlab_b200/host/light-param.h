// Light parameter variant of the public API (reference src/light-param.h:20-47).
#ifndef PBRLAB_B200_LIGHT_PARAM_H_
#define PBRLAB_B200_LIGHT_PARAM_H_
#include <string>
#include <variant>

#include "type.h"

namespace pbrlab {
struct AreaLightParameter {
  float3 emission = float3(0.8f);
  std::string name;
};
enum LightType { kAreaLight = 0, kLightNone };
using LightParameter = std::variant<AreaLightParameter>;
inline void SetLightName(const std::string& name, LightParameter* p) { std::get<kAreaLight>(*p).name = name; }
inline std::string GetLightName(const LightParameter& p) { return std::get<kAreaLight>(p).name; }
}  // namespace pbrlab
#endif  // PBRLAB_B200_LIGHT_PARAM_H_
