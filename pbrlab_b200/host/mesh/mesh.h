// MeshPtr variant (reference src/mesh/mesh.h:19-43).
#ifndef PBRLAB_B200_MESH_H_
#define PBRLAB_B200_MESH_H_
#include <memory>
#include <string>
#include <variant>

#include "cubic-bezier-curve-mesh.h"
#include "triangle-mesh.h"

namespace pbrlab {
enum MeshType { kTriangleMesh = 0, kCubicBezierCurveMesh, kMeshNone };
using MeshPtr = std::variant<std::shared_ptr<TriangleMesh>, std::shared_ptr<CubicBezierCurveMesh>>;
inline std::string GetName(const MeshPtr& m) {
  return m.index() == kTriangleMesh ? std::get<kTriangleMesh>(m)->GetName() : std::get<kCubicBezierCurveMesh>(m)->GetName();
}
inline uint32_t GetNumPrimitive(const MeshPtr& m) {
  return m.index() == kTriangleMesh ? std::get<kTriangleMesh>(m)->GetNumFaces()
                                    : std::get<kCubicBezierCurveMesh>(m)->GetNumSegments();
}
}  // namespace pbrlab
#endif  // PBRLAB_B200_MESH_H_
