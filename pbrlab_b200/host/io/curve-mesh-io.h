// .hair -> CubicBezierCurveMesh (reference src/io/curve-mesh-io.h:12-19, src/io/curve-mesh-io.cc:32-138).
#ifndef PBRLAB_B200_CURVE_MESH_IO_H_
#define PBRLAB_B200_CURVE_MESH_IO_H_
#include <cstdint>
#include <string>
#include <vector>

#include "../mesh/cubic-bezier-curve-mesh.h"

namespace pbrlab {
namespace io {
// memory_saving_mode shares the joint control point between consecutive segments of a strand (3 control points per
// segment + 1); the command-line front-end always loads with it off: 4 control points per segment, indices 0,4,8...
bool LoadCurveMeshAsCubicBezierCurve(const std::string& filepath, const bool memory_saving_mode,
                                     std::vector<float>* vertices_thickness, std::vector<uint32_t>* indices);
bool LoadCurveMeshAsCubicBezierCurve(const std::string& filepath, const bool memory_saving_mode,
                                     CubicBezierCurveMesh* curve_mesh);
}  // namespace io
}  // namespace pbrlab
#endif  // PBRLAB_B200_CURVE_MESH_IO_H_
