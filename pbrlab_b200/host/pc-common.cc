#include "pc-common.h"

#include <cstdio>
#include <iostream>
#include <string>

#include "io/curve-mesh-io.h"
#include "io/triangle-mesh-io.h"

namespace {

const float kIdentity[4][4] = {{1.f, 0.f, 0.f, 0.f}, {0.f, 1.f, 0.f, 0.f}, {0.f, 0.f, 1.f, 0.f}, {0.f, 0.f, 0.f, 1.f}};

// reference pc/pc-common.cc:100-190
bool AddObj(const std::string& path, pbrlab::Scene* scene) {
  std::vector<pbrlab::TriangleMesh> meshes;
  std::vector<pbrlab::MaterialParameter> materials;
  std::vector<pbrlab::Texture> textures;
  if (!pbrlab::io::LoadTriangleMeshFromObj(path, &meshes, &materials, &textures)) {
    std::cerr << "Faild loading obj file [" << path << "]" << std::endl;
    return false;
  }
  std::cerr << "Load obj file [" << path << "]" << std::endl;
  // file-local texture index -> scene texture id (reference pc/pc-common.cc:93-98,122-140)
  std::vector<uint32_t> texture_ids;
  for (const auto& t : textures) texture_ids.push_back(scene->AddTexture(t));
  for (auto& m : materials) {
    if (m.index() != pbrlab::kCyclesPrincipledBsdfParameter) continue;
    auto& p = std::get<pbrlab::kCyclesPrincipledBsdfParameter>(m);
    if (p.base_color_tex_id != uint32_t(-1)) p.base_color_tex_id = texture_ids.at(p.base_color_tex_id);
    if (p.subsurface_color_tex_id != uint32_t(-1)) p.subsurface_color_tex_id = texture_ids.at(p.subsurface_color_tex_id);
  }
  std::vector<uint32_t> material_ids;
  for (const auto& m : materials) material_ids.push_back(scene->AddMaterialParam(m));
  std::cerr << "The Number of shapes is " << meshes.size() << " in [" << path << "]" << std::endl;

  for (auto& mesh : meshes) {
    std::cerr << "  add shape [" << mesh.GetName() << "]" << std::endl;
    std::cerr << "    num face : " << mesh.GetNumFaces() << std::endl;
    const uint32_t nf = mesh.GetNumFaces();
    for (uint32_t f = 0; f < nf; ++f) {
      const uint32_t local = mesh.GetMaterials()[f];
      // file-local material index -> scene material id (faces without usemtl keep "no material": absorbed)
      mesh.SetMaterialId(local < material_ids.size() ? material_ids[local] : uint32_t(-1), f);
    }
    const pbrlab::MeshPtr mesh_ptr = scene->AddTriangleMesh(mesh);
    const uint32_t ls = scene->CreateLocalScene();
    scene->AddMeshToLocalScene(ls, mesh_ptr);
    const uint32_t inst = scene->CreateInstance(ls, kIdentity);
    // shapes whose name starts with "light" emit (3,3,3); MTL Ke is ignored (reference pc/pc-common.cc:172-186)
    if (mesh.GetName().substr(0, 5) == "light") {
      pbrlab::AreaLightParameter lp = {};
      lp.emission = pbrlab::float3(3.0f);
      const uint32_t light_id = scene->AddLightParam(lp);
      scene->AttachLightParamIdsToInstance(inst, {std::vector<uint32_t>(nf, light_id)});
    }
  }
  std::cerr << std::endl;
  return true;
}

// reference pc/pc-common.cc:192-237
bool AddHair(const std::string& path, pbrlab::Scene* scene) {
  pbrlab::CubicBezierCurveMesh curves;
  if (!pbrlab::io::LoadCurveMeshAsCubicBezierCurve(path, false, &curves)) return false;
  std::cerr << "Load curve file [" << path << "]" << std::endl;
  std::cerr << "  add shape [" << curves.GetName() << "]" << std::endl;
  std::cerr << "  num segments : " << curves.GetNumSegments() << std::endl;
  pbrlab::MaterialParameter mp = pbrlab::HairBsdfParameter();
  pbrlab::SetMaterialName("hair", &mp);
  const uint32_t material_id = scene->AddMaterialParam(mp);
  for (uint32_t s = 0; s < curves.GetNumSegments(); ++s) curves.SetMaterialId(material_id, s);
  const pbrlab::MeshPtr mesh_ptr = scene->AddCubicBezierCurveMesh(curves);
  const uint32_t ls = scene->CreateLocalScene();
  scene->AddMeshToLocalScene(ls, mesh_ptr);
  scene->CreateInstance(ls, kIdentity);
  return true;
}

std::string Extension(const std::string& p) {
  const size_t dot = p.find_last_of('.');
  const size_t slash = p.find_last_of('/');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
  return p.substr(dot);
}

}  // namespace

bool CreateScene(int argc, char** argv, pbrlab::Scene* scene, bool commit_to_device) {
  if (argc < 2) return false;
  for (int i = 1; i < argc; ++i) {
    const std::string path(argv[i]);
    const std::string ext = Extension(path);
    if (ext == ".obj") {
      if (!AddObj(path, scene)) return false;
    } else if (ext == ".hair") {
      if (!AddHair(path, scene)) return false;
    }
  }
  if (commit_to_device) scene->CommitScene();
  else scene->CommitHostOnly();
  float bmin[3], bmax[3];
  scene->FetchSceneAABB(bmin, bmax);
  printf("bmin: %f %f %f\n  bmax: %f %f %f\n", double(bmin[0]), double(bmin[1]), double(bmin[2]), double(bmax[0]),
         double(bmax[1]), double(bmax[2]));
  return true;
}
