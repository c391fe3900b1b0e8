// Instance and local-scene records (reference src/mesh-instance.h:11-24, src/local-scene.h:8-10).
#ifndef PBRLAB_B200_MESH_INSTANCE_H_
#define PBRLAB_B200_MESH_INSTANCE_H_
#include <memory>
#include <vector>

#include "mesh/mesh.h"

namespace pbrlab {
struct LocalScene {
  std::vector<MeshPtr> meshes;
};
struct MeshInstance {
  std::shared_ptr<LocalScene> local_scene;
  std::vector<std::vector<uint32_t>> material_ids;     // [geom][prim]
  std::vector<std::vector<uint32_t>> light_param_ids;  // [geom][prim]; empty = mesh does not emit
  float transform_lg[4][4];
  float transform_gl[4][4];
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_MESH_INSTANCE_H_
