// Cubic Bezier curve mesh of the public API (reference src/mesh/cubic-bezier-curve-mesh.h:12-38): control points
// (xyz + thickness) and, per segment, the index of its first control point.
#ifndef PBRLAB_B200_CUBIC_BEZIER_CURVE_MESH_H_
#define PBRLAB_B200_CUBIC_BEZIER_CURVE_MESH_H_
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../api-types.h"

namespace pbrlab {
class CubicBezierCurveMesh {
public:
  CubicBezierCurveMesh() {}
  CubicBezierCurveMesh(const std::string& name, const std::shared_ptr<CurveAttribute> attribute,
                       const std::vector<uint32_t>& indices, const std::vector<uint32_t>& material_ids)
      : attribute_(attribute), indices_(indices), material_ids_(material_ids), name_(name) {}
  const std::vector<uint32_t>& GetIndices(void) const { return indices_; }
  uint32_t GetNumSegments(void) const { return uint32_t(indices_.size()); }
  uint32_t GetNumVertices(void) const { return uint32_t(attribute_->vertices.size() / 4); }
  const std::vector<uint32_t>& GetMaterials(void) const { return material_ids_; }
  std::string GetName(void) const { return name_; }
  const std::vector<float>& GetVertices(void) const { return attribute_->vertices; }
  void SetMaterialId(const uint32_t material_id, const uint32_t segment_id) { material_ids_[segment_id] = material_id; }

private:
  std::shared_ptr<CurveAttribute> attribute_;
  std::vector<uint32_t> indices_, material_ids_;
  std::string name_;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_CUBIC_BEZIER_CURVE_MESH_H_
