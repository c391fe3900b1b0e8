// The small plain-data types of the reference's public API, kept field for field so that code written against
// pbrlab compiles against this host layer unchanged: Ray (reference src/ray.h:9-14), RenderConfig
// (src/render-config.h:9-18), Attribute / CurveAttribute (src/mesh/attribute.h:6-15), in one header.
#ifndef PBRLAB_B200_API_TYPES_H_
#define PBRLAB_B200_API_TYPES_H_
#include <cstdint>
#include <string>
#include <vector>

#include "type.h"

namespace pbrlab {

// closest-hit / occlusion query as Scene::TraceFirstHit1 / AnyHit1 take it
struct Ray {
  float3 ray_dir;
  float3 ray_org;
  float min_t = 0.0f;
  float max_t = kInf;
};

// `thread` is as unused here as it is in the reference (src/render.cc:203-204 never consults it): the GPU backend has
// no host worker pool
struct RenderConfig {
  std::vector<std::string> scene_filepaths;
  uint32_t width = 512, height = 512;
  uint32_t max_pass = 32;
  int thread = -1;
};

// shared vertex attribute pools of the meshes of one file
struct Attribute {
  std::vector<float> vertices;   // xyzw per vertex (w = 1)
  std::vector<float> normals;    // xyzw per normal (w = 1), may be empty
  std::vector<float> texcoords;  // uv per texcoord, may be empty
};
struct CurveAttribute {
  std::vector<float> vertices;   // xyz + thickness per control point
};

}  // namespace pbrlab
#endif  // PBRLAB_B200_API_TYPES_H_
