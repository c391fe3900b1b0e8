// Host half of pbrgpu_set_* / pbrgpu_commit: keeps the uploaded soups, builds the two 8-wide BVHs and lays out
// every table exactly as the device will see it (std::vector = one cudaMemcpy each).  No CUDA in this file, so the
// same object also feeds the g++-compiled host emulation used by the CPU tests.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pbrgpu.h"
#include "bvh_builder.h"
#include "device/scene_view.cuh"

namespace pbrhost {

struct F4 { float x, y, z, w; };
struct U4 { uint32_t x, y, z, w; };
struct F2 { float x, y; };

struct HostScene {
  // ---- inputs (copied from the caller)
  std::vector<F4> verts, normals;
  std::vector<F2> texcoords;
  std::vector<U4> tri_vidx, tri_nidx, tri_tidx, tri_ids;   // tri_nidx.w = emissive entry, tri_ids = inst,geom,prim,mat
  std::vector<F4> curve_cps;                               // 4 per segment, gathered
  std::vector<U4> curve_ids;
  std::vector<pbrgpu_material> materials;
  std::vector<uint32_t> material_class;                    // per material: pbr::MaterialClass (shading queue it is routed to)
  std::vector<float> tex_pixels;                           // all textures back to back
  std::vector<pbr::TexDesc> tex_desc;
  // lights
  std::vector<float> light_cdf;
  std::vector<pbr::LightRec> lights;
  std::vector<float> lprim_cdf;
  std::vector<F4> lprim_info;
  std::vector<uint32_t> lprim_tri;
  std::vector<F4> emissive;
  std::vector<uint32_t> pending_emissive_;   // prim_is_emissive of the last SetLights()

  // ---- built by Commit()
  pbrbvh::Bvh8 tri_bvh, curve_bvh;
  std::vector<F4> tri_data;      // 3 per triangle, leaf order
  std::vector<F4> curve_data;    // 4 per segment, slot order
  std::vector<uint32_t> curve_prim;   // slot -> curve primitive (segment) index of pbrgpu_set_curves
  std::vector<uint32_t> curve_sub;    // leaf order: (slot << 2) | first quad of the part
  uint32_t curve_part_quads = 4;      // quads (quarter sub-segments) per BVH primitive: 4, 2 or 1
  std::vector<float> curve_cull;   // 8 per segment, slot order (device/traverse.cuh: CurveMayHit)
  float bmin[3], bmax[3];
  // clearance grid for random-walk segments (see device/scene_view.cuh); empty when no material scatters
  std::vector<uint32_t> clear_dist;   // one byte per cell, packed
  float clear_org[3] = {0, 0, 0}, clear_inv_cell = 0.f, clear_quantum = 0.f;
  uint32_t clear_dims[3] = {0, 0, 0};
  bool committed = false;
  uint32_t max_material_id = 0;   // highest material id any committed primitive refers to (live edits check against it)
  bool any_material_id = false;
  double build_seconds = 0.0;
  // Which builder Commit() uses for the two BVHs: PBRGPU_BVH_TRIS / PBRGPU_BVH_CURVES (or PBRGPU_BVH for both) =
  // "sah" (host binned SAH, bvh_builder.cc), "ploc" (the data-parallel builder of bvh_ploc.h: on the device through
  // `device_builder` when the CUDA library installed one, else its host loops), "auto" (default): triangles "ploc"
  // when a device builder is installed, curves "sah".
  using DeviceBuilder = bool (*)(void* user, const pbrbvh::Aabb* boxes, uint32_t n, const pbrbvh::BuildParams& prm,
                                 uint32_t radius, pbrbvh::Bvh8* out, const char** err);
  DeviceBuilder device_builder = nullptr;
  void* device_builder_user = nullptr;
  std::string last_builder;      // what the last Commit() used for the triangle BVH ("sah", "ploc-host", "ploc-device")
  double bvh_seconds = 0.0, clearance_seconds = 0.0;
  bool BuildBvh(const pbrbvh::Aabb* boxes, uint32_t n, const pbrbvh::BuildParams& prm, pbrbvh::Bvh8* out,
                std::string* which, bool curves);

  std::string error;

  bool SetTriangles(const float* xyzw, uint32_t nverts, const uint32_t* vidx, const float* nxyzw, uint32_t nnormals,
                    const uint32_t* nidx, const float* uv, uint32_t nuv, const uint32_t* tidx,
                    const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                    const uint32_t* prim_id, uint64_t ntris);
  bool SetCurves(const float* xyzr, uint32_t nverts, const uint32_t* first_cp, const uint32_t* material_id,
                 const uint32_t* instance_id, const uint32_t* geom_id, const uint32_t* prim_id, uint64_t nsegs);
  bool SetMaterials(const pbrgpu_material* m, uint32_t n);
  bool SetTextures(const pbrgpu_texture* t, uint32_t n);
  bool CheckTextureIds();
  bool SetLights(const pbrgpu_light_tables* t);
  bool Commit(const float* bmin_in, const float* bmax_in);
  void BuildClearance();
  void RefineClearanceNearSurface(float* scratch);

  uint32_t num_tris() const { return uint32_t(tri_vidx.size()); }
  uint32_t num_curves() const { return uint32_t(curve_ids.size()); }

  // view over the host vectors (for the emulation); the CUDA side builds the same struct over device copies
  pbr::SceneView HostView() const;
};

// camera of RenderingTile (reference src/render.cc:132-158): eye.xyz, x_corner, y_corner, z_corner, dx, dy
void MakeCamera(const float* bmin, const float* bmax, uint32_t width, uint32_t height, float* cam8);

}  // namespace pbrhost
