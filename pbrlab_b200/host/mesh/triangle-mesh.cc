#include "triangle-mesh.h"

#include <iostream>

namespace pbrlab {

TriangleMesh::TriangleMesh(const std::string name, const std::shared_ptr<Attribute>& attribute,
                           const std::vector<uint32_t> vertex_ids, const std::vector<uint32_t> normal_ids,
                           const std::vector<uint32_t> texcoord_ids, const std::vector<uint32_t> material_ids)
    : attribute_(attribute), name_(name) {
  if (vertex_ids.size() % 3 != 0) std::cerr << "error! wrong vertex ids size" << std::endl;
  num_faces_ = uint32_t(vertex_ids.size() / 3);
  vertex_ids_ = vertex_ids;
  // index lists of the wrong length are replaced by "none" (reference src/mesh/triangle-mesh.cc:33-53; the
  // material list is then 3*num_faces long there, quirk 18 — any length >= num_faces behaves the same)
  if (normal_ids.size() == size_t(num_faces_) * 3) normal_ids_ = normal_ids;
  else normal_ids_.assign(size_t(num_faces_) * 3, uint32_t(-1));
  if (texcoord_ids.size() == size_t(num_faces_) * 3) texcoord_ids_ = texcoord_ids;
  else texcoord_ids_.assign(size_t(num_faces_) * 3, uint32_t(-1));
  if (material_ids.size() == num_faces_) material_ids_ = material_ids;
  else material_ids_.assign(size_t(num_faces_) * 3, uint32_t(-1));
}

float3 TriangleMesh::FetchGeometryNormal(const uint32_t prim_id) const {
  const float3 p0 = P(prim_id, 0), p1 = P(prim_id, 1), p2 = P(prim_id, 2);
  return vnormalized(vcross(p1 - p0, p2 - p1));   // reference :62-75,181-184
}

float3 TriangleMesh::FetchShadingNormal(const uint32_t prim_id, const float u, const float v) const {
  const uint32_t a = normal_ids_[prim_id * 3], b = normal_ids_[prim_id * 3 + 1], c = normal_ids_[prim_id * 3 + 2];
  if (a == uint32_t(-1) || b == uint32_t(-1) || c == uint32_t(-1)) return FetchGeometryNormal(prim_id);
  const float* n = attribute_->normals.data();
  const float3 n0(n + size_t(a) * 4), n1(n + size_t(b) * 4), n2(n + size_t(c) * 4);
  return vnormalized(((1.0f - u - v) * n0 + u * n1) + v * n2);   // reference :77-101
}

float3 TriangleMesh::FetchLocalPosition(const uint32_t prim_id, const float u, const float v) const {
  return ((1.0f - u - v) * P(prim_id, 0) + u * P(prim_id, 1)) + v * P(prim_id, 2);   // reference :102-112
}

float TriangleMesh::FetchFaceArea(const uint32_t prim_id) const {
  const float3 p0 = P(prim_id, 0), p1 = P(prim_id, 1), p2 = P(prim_id, 2);
  return vlength(vcross(p1 - p0, p2 - p0)) * 0.5f;   // reference :114-124
}

float2 TriangleMesh::FetchTexcoord(const uint32_t prim_id, const float u, const float v) const {
  const uint32_t a = texcoord_ids_[prim_id * 3], b = texcoord_ids_[prim_id * 3 + 1], c = texcoord_ids_[prim_id * 3 + 2];
  if (a == uint32_t(-1) || b == uint32_t(-1) || c == uint32_t(-1)) return float2(u, v);   // reference :126-135
  const float* t = attribute_->texcoords.data();
  return ((1.0f - u - v) * float2(t + size_t(a) * 2) + u * float2(t + size_t(b) * 2)) + v * float2(t + size_t(c) * 2);
}

}  // namespace pbrlab
