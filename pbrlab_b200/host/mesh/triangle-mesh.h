// TriangleMesh of the public API (reference src/mesh/triangle-mesh.h:16-62, src/mesh/triangle-mesh.cc:18-184):
// an index list into a shared Attribute plus per-face material ids.  The fetch helpers are the host-side twins of
// what the shading kernels compute per hit (pbrlab_b200/csrc/device/shade.cuh MakeSurface).
// Additions over the reference: const accessors for the normal / texcoord index lists and the attribute pool,
// which the device upload needs (SURVEY §8(b) "what the reference API does not expose").
#ifndef PBRLAB_B200_TRIANGLE_MESH_H_
#define PBRLAB_B200_TRIANGLE_MESH_H_
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../type.h"
#include "../api-types.h"

namespace pbrlab {

class TriangleMesh {
public:
  TriangleMesh() : num_faces_(0) {}
  TriangleMesh(const std::string name, const std::shared_ptr<Attribute>& attribute,
               const std::vector<uint32_t> vertex_ids, const std::vector<uint32_t> normal_ids,
               const std::vector<uint32_t> texcoord_ids, const std::vector<uint32_t> material_ids);

  float3 FetchGeometryNormal(const uint32_t prim_id) const;
  float3 FetchShadingNormal(const uint32_t prim_id, const float u, const float v) const;
  float3 FetchLocalPosition(const uint32_t prim_id, const float u, const float v) const;
  float FetchFaceArea(const uint32_t prim_id) const;
  float2 FetchTexcoord(const uint32_t prim_id, const float u, const float v) const;

  const std::vector<uint32_t>& GetMaterials(void) const { return material_ids_; }
  uint32_t GetNumFaces(void) const { return num_faces_; }
  uint32_t GetNumVertices(void) const { return uint32_t(attribute_->vertices.size() / 4); }
  std::string GetName(void) const { return name_; }
  const std::vector<uint32_t>& GetVertexIds(void) const { return vertex_ids_; }
  const std::vector<float>& GetVertices(void) const { return attribute_->vertices; }
  void SetMaterialId(const uint32_t material_id, const uint32_t prim_id) { material_ids_[prim_id] = material_id; }

  // device-upload accessors (not in the reference)
  const std::vector<uint32_t>& GetNormalIds(void) const { return normal_ids_; }
  const std::vector<uint32_t>& GetTexcoordIds(void) const { return texcoord_ids_; }
  const std::shared_ptr<Attribute>& GetAttribute(void) const { return attribute_; }

private:
  float3 P(uint32_t prim, int k) const { return float3(attribute_->vertices.data() + size_t(vertex_ids_[prim * 3 + k]) * 4); }
  uint32_t num_faces_;
  std::vector<uint32_t> vertex_ids_, normal_ids_, texcoord_ids_, material_ids_;
  std::shared_ptr<Attribute> attribute_;
  std::string name_;
};

}  // namespace pbrlab
#endif  // PBRLAB_B200_TRIANGLE_MESH_H_
