// Catmull-Rom strand -> chain of cubic Bezier segments (reference src/curve-util.h, src/curve-util.cc:7-199).
#ifndef PBRLAB_B200_CURVE_UTIL_H_
#define PBRLAB_B200_CURVE_UTIL_H_
#include <vector>
namespace pbrlab {
// cvs: xyz per control vertex (>= 3 of them), cv_radiuss: one per control vertex.  Appends 4 control points
// (12 floats) and 4 radii per produced segment; a strand with N vertices gives N-1 segments.
bool ToCubicBezierCurve(const std::vector<float>& cvs, const std::vector<float>& cv_radiuss,
                        std::vector<float>* bezier_vertices, std::vector<float>* bezier_radiuss);
}  // namespace pbrlab
#endif  // PBRLAB_B200_CURVE_UTIL_H_
