// Shared vertex attribute pools (reference src/mesh/attribute.h:6-15).
#ifndef PBRLAB_B200_ATTRIBUTE_H_
#define PBRLAB_B200_ATTRIBUTE_H_
#include <vector>
namespace pbrlab {
struct Attribute {
  std::vector<float> vertices;   // xyzw per vertex (w = 1)
  std::vector<float> normals;    // xyzw per normal (w = 1), may be empty
  std::vector<float> texcoords;  // uv per texcoord, may be empty
};
struct CurveAttribute {
  std::vector<float> vertices;   // xyz + thickness per control point
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_ATTRIBUTE_H_
