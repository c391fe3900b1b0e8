// Wavefront OBJ + MTL ingest with the reference's entry point and behaviour
// (reference src/io/triangle-mesh-io.h:14, src/io/triangle-mesh-io.cc:143-325, which drives tinyobjloader).
// The parser is our own; what it reproduces from the reference's loader stack:
//   * one TriangleMesh per `o` / `g` block that has faces, all sharing one Attribute (xyzw positions, xyzw normals,
//     uv with v flipped to 1 - v), per-face material ids following `usemtl`;
//   * polygons are triangulated (quads along the shorter diagonal, larger polygons as fans);
//   * MTL: the Principled keys are free-form `key value...` lines; when a key repeats inside one material the FIRST
//     occurrence wins (tinyobj keeps unknown keys in a std::map and uses insert; reference
//     src/io/tiny_obj_loader.h:2413-2428);
//   * map_base_color / map_subsurface_color name texture files (options as tinyobj parses them, `-colorspace`
//     honoured); they are loaded by io/image-io.cc, de-gammaed unless HDR / non-sRGB, appended to `textures`, and the
//     material's tex id is the index in that vector (reference src/io/triangle-mesh-io.cc:80-141).
#ifndef PBRLAB_B200_TRIANGLE_MESH_IO_H_
#define PBRLAB_B200_TRIANGLE_MESH_IO_H_
#include <string>
#include <vector>

#include "../material-param.h"
#include "../mesh/triangle-mesh.h"
#include "../texture.h"

namespace pbrlab {
namespace io {
bool LoadTriangleMeshFromObj(const std::string& filename, std::vector<TriangleMesh>* meshes,
                             std::vector<MaterialParameter>* material_params, std::vector<Texture>* textures);
}  // namespace io
}  // namespace pbrlab
#endif  // PBRLAB_B200_TRIANGLE_MESH_IO_H_
