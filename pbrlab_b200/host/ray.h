// Ray as the reference's public API spells it (reference src/ray.h:9-14).
#ifndef PBRLAB_B200_RAY_H_
#define PBRLAB_B200_RAY_H_
#include "type.h"
namespace pbrlab {
struct Ray {
  float3 ray_dir;
  float3 ray_org;
  float min_t = 0.0f;
  float max_t = kInf;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_RAY_H_
