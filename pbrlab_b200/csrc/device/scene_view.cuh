// Device-side view of a committed scene: every pointer is HBM-resident, read-only during rendering and identical
// on every GPU (the scene is replicated).  Built by pbrgpu_commit() (csrc/pbrgpu.cu).
#pragma once
#include "common.cuh"

namespace pbr {

// include/pbrgpu.h: pbrgpu_material (112 bytes)
struct DeviceMaterial {
  uint32_t type;        // 0 = CyclesPrincipledBsdfParameter, 1 = HairBsdfParameter (material-param.h:20-23)
  uint32_t tex_id[2];
  uint32_t reserved;
  float p[24];
};

struct LightRec {        // one entry of LightManager::lights_ (light-manager.h:185-189) flattened
  float choose_probability;
  uint32_t prim_offset;  // first entry in the per-primitive light tables
  uint32_t prim_count;
  uint32_t pad;
};

// Shading queue of a material (scene_host.cc: ClassifyMaterial): the closest-hit kernel routes every hit by this
enum MaterialClass : uint32_t {
  kClassGeneral = 0,   // Principled with any closure set (incl. none, textured, subsurface): ShadeSurfaceKernel<false>
  kClassDiffuse = 1,   // Principled that can only ever enable the Lambert closure: ShadeSurfaceKernel<true>
  kClassHair = 2       // HairBsdfParameter: ShadeHairKernel
};

struct TexDesc {         // one Texture (texture.h:10-44): row-major float pixels, `channels` interleaved
  uint32_t offset;       // first float in SceneView::tex_pixels
  uint32_t width, height, channels;
};

struct SceneView {
  // ---- traversal (see bvh_builder.h for the node format)
  const float4* tri_nodes;     // 5 x float4 per 8-wide compressed node; root = node 0
  const float4* tri_data;      // leaf-ordered, 3 x float4 per triangle: (v0, prim id) (e1=v0-v1, 0) (e2=v2-v0, 0)
  const float4* curve_nodes;
  const float4* curve_data;    // 4 x float4 (xyz, radius) per cubic Bezier segment, in "slot" order (spatially sorted)
  const uint32_t* curve_prim;  // slot -> curve primitive id
  const uint32_t* curve_sub;   // curve BVH leaf order: (slot << 2) | first quad; the BVH primitive is a PART of a segment
  uint32_t curve_part_quads;   // quads (quarter sub-segments) per part: 4 (whole segments), 2 or 1
  uint32_t ribbon_min_lanes;   // traversal engine: lanes holding a curve candidate that trigger a ribbon phase
  uint32_t thin_spread;        // traversal engine: short launches spread their items over all warps (LanesFor)
  uint32_t inside_first;       // traversal engine, curve BVH: children whose box holds the ray origin first (PopChild)
  const float4* curve_cull;    // slot order, 2 per segment: (c0, capsule radius around the line c0c3), (c3 - c0, |c3 - c0|), see CurveMayHit; may be null
  uint32_t num_tris, num_curves;
  uint32_t bias_magic;         // kBiasMagic (traverse.cuh), as a run-time value on purpose
  // ---- per-primitive shading tables (indexed by primitive id = order given to pbrgpu_set_*)
  const uint4* tri_ids;        // instance_id, geom_id, prim_id inside the shape, material_id
  const uint4* tri_nidx;       // 3 normal indices (0xFFFFFFFF = none), index into emissive[] or 0xFFFFFFFF
  const uint4* tri_vidx;       // 3 vertex indices, unused
  const uint4* tri_tidx;       // 3 texcoord indices (0xFFFFFFFF = none), unused
  const float4* verts;         // xyzw
  const float4* normals;       // xyzw
  const float2* texcoords;
  const uint4* curve_ids;      // instance_id, geom_id, segment id inside the shape, material_id
  const DeviceMaterial* materials;
  uint32_t num_materials;
  uint32_t num_hair_materials; // materials of type 1
  const uint32_t* material_class;   // MaterialClass per material
  // ---- textures (Scene::AddTexture; sampled by ParamToBsdf for base_color / subsurface_color)
  const float* tex_pixels;     // all textures back to back
  const TexDesc* tex_desc;
  uint32_t num_textures;
  // ---- lights (light-manager.h:172-193 flattened)
  const float4* emissive;      // per emissive triangle: emission rgb, pdf = P(light) P(prim) / area   (ImplicitAreaLight)
  const float* light_cdf;      // cumulative_probability_ over lights
  const LightRec* lights;
  uint32_t num_lights;
  const float* lprim_cdf;      // per light primitive: AreaLight::cumulative_probability_
  const float4* lprim_info;    // per light primitive: emission rgb, pdf = P(light) P(prim) / area
  const uint32_t* lprim_tri;   // per light primitive: triangle primitive id
  // ---- clearance field for random-walk segments (scene_host.cc: BuildClearance): one byte per cell of an isotropic
  // grid = a lower bound, in units of clear_quantum, of the distance from any point of the cell to any primitive
  const uint8_t* clear_dist;   // null: no subsurface material in the scene
  float clear_org[3];          // grid origin
  float clear_inv_cell;        // 1 / cell edge
  float clear_quantum;         // cell edge / 4
  uint32_t clear_dims[3];      // cells per axis
  uint32_t clear_march_steps;  // SegmentIsClear: sphere-tracing steps along a segment before giving up (>= 1)
};

}  // namespace pbr
