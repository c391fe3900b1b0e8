// pbrlab::Scene with the reference's public surface (reference src/scene.h:14-111) in front of the B200 backend.
// What the reference keeps inside Embree (acceleration structure, scene bounds) lives on the device here: CommitScene()
// flattens every instance into one triangle soup + one curve soup with (instance, geom, prim) side tables, exports
// the LightManager tables and hands all of it to the C ABI (include/pbrgpu.h).  Ray queries and Render() then run on
// the GPU; there is no CPU path — CommitScene() throws if no device context can be created.
#ifndef PBRLAB_B200_SCENE_H_
#define PBRLAB_B200_SCENE_H_
#include <memory>
#include <vector>

#include "light-manager.h"
#include "material-param.h"
#include "mesh-instance.h"
#include "api-types.h"
#include "texture.h"

struct pbrgpu_ctx;

namespace pbrlab {

// the reference's TraceResult (reference src/raytracer/raytracer.h:9-17)
struct TraceResult {
  float normal_g[3] = {1.0f, 0.0f, 0.0f};
  float t = 1.0f;
  float u = 0.0f;
  float v = 0.0f;
  uint32_t instance_id = static_cast<uint32_t>(-1);
  uint32_t geom_id = static_cast<uint32_t>(-1);
  uint32_t prim_id = static_cast<uint32_t>(-1);
};

// Everything pbrgpu_set_* needs, as flat arrays (also what the CPU tests inspect).
struct FlatScene {
  std::vector<float> verts, normals, texcoords;                   // xyzw, xyzw, uv
  std::vector<uint32_t> vidx, nidx, tidx;                         // 3 per triangle (global indices)
  std::vector<uint32_t> tri_material, tri_instance, tri_geom, tri_prim;
  std::vector<float> curve_verts;                                 // xyzr
  std::vector<uint32_t> curve_first, curve_material, curve_instance, curve_geom, curve_prim;
  std::vector<float> materials;                                   // 28 words per material (pbrgpu_material)
  std::vector<float> tex_pixels;                                  // all textures back to back
  std::vector<uint32_t> tex_desc;                                 // per texture: offset (floats), width, height, channels
  LightManager::Tables lights;
  std::vector<uint32_t> light_prim_triangle;                      // per light primitive: flattened triangle index
  float bmin[3], bmax[3];
};

class Scene {
public:
  Scene(void);
  ~Scene(void);
  Scene(const Scene&) = delete;
  Scene& operator=(const Scene&) = delete;

  template <class... Args>
  MeshPtr AddTriangleMesh(Args&&... args) {
    triangle_meshes_.emplace_back(std::make_shared<TriangleMesh>(args...));
    return MeshPtr(triangle_meshes_.back());
  }
  template <class... Args>
  MeshPtr AddCubicBezierCurveMesh(Args&&... args) {
    cubic_bezier_curve_meshes_.emplace_back(std::make_shared<CubicBezierCurveMesh>(args...));
    return MeshPtr(cubic_bezier_curve_meshes_.back());
  }
  template <class... Args>
  uint32_t AddLightParam(Args&&... args) { return light_manager_->AddLightParam(args...); }
  template <class... Args>
  uint32_t AddMaterialParam(Args&&... args) {
    material_params_.emplace_back(args...);
    return uint32_t(material_params_.size() - 1);
  }
  template <class... Args>
  uint32_t AddTexture(Args&&... args) {
    textures_.emplace_back(std::make_shared<Texture>(args...));
    return uint32_t(textures_.size() - 1);
  }

  uint32_t AddMeshToLocalScene(const uint32_t local_scene_id, const MeshPtr& mesh_ptr);
  void AttachLightParamIdsToInstance(const uint32_t instance_id,
                                     const std::vector<std::vector<uint32_t>>& light_param_ids);
  void AttachMaterialParamIdsToInstance(const uint32_t instance_id,
                                        const std::vector<std::vector<uint32_t>>& material_ids);
  void CommitScene(void);
  uint32_t CreateInstance(const uint32_t local_scene_id, const float transform[4][4]);
  uint32_t CreateLocalScene(void);

  const MeshInstance& GetMeshInstance(const uint32_t instance_id) const { return instances_[instance_id]; }
  const LightManager* GetLightManager(void) const { return light_manager_.get(); }
  const Texture* GetTexture(const uint32_t tex_id) const { return textures_[tex_id].get(); }

  const MaterialParameter* FetchMeshMaterialParameter(const TraceResult& trace_result) const;
  std::vector<MaterialParameter>* FetchMeshMaterialParameters(void) { return &material_params_; }
  float3 FetchMeshShadingNormal(const TraceResult& trace_result) const;
  float2 FetchMeshTexcoord(const TraceResult& trace_result) const;
  void FetchSceneAABB(float* bmin, float* bmax) const;

  TraceResult TraceFirstHit1(const Ray& ray) const;
  bool AnyHit1(const Ray& ray) const;

  // ---- additions for the GPU backend (not in the reference)
  // host half of CommitScene(): light tables, bounds, flattening — no device needed
  void CommitHostOnly(void);
  const FlatScene& Flat(void) const { return flat_; }
  // device context the scene was committed to (nullptr before CommitScene); owned by the scene
  pbrgpu_ctx* DeviceContext(void) const { return ctx_; }
  // re-upload the (possibly edited) material table; Render() calls it every time, as the reference's shaders read
  // material_params_ live (reference pc/pbrlab-gui.cc:207-238)
  void SyncMaterialsToDevice(void) const;
  // choose the devices before CommitScene (default: current device only)
  void SetDevices(const std::vector<int>& device_ids) { device_ids_ = device_ids; }

private:
  void PackMaterials(std::vector<float>* out) const;

  std::vector<std::shared_ptr<LocalScene>> local_scenes_;
  std::vector<MeshInstance> instances_;
  std::vector<std::shared_ptr<TriangleMesh>> triangle_meshes_;
  std::vector<std::shared_ptr<CubicBezierCurveMesh>> cubic_bezier_curve_meshes_;
  std::vector<MaterialParameter> material_params_;
  std::vector<std::shared_ptr<Texture>> textures_;
  std::shared_ptr<LightManager> light_manager_;
  float bmax_[3], bmin_[3];
  FlatScene flat_;
  std::vector<int> device_ids_;
  pbrgpu_ctx* ctx_ = nullptr;
};

}  // namespace pbrlab
#endif  // PBRLAB_B200_SCENE_H_
