// CyHair (.hair) reader with the reference's behaviour (reference src/io/cyhair.h:8-58, src/io/cyhair.cc:20-180):
// 128-byte header, optional per-strand segment counts, points, thickness, transparency, colour arrays.
#ifndef PBRLAB_B200_CYHAIR_H_
#define PBRLAB_B200_CYHAIR_H_
#include <string>
#include <vector>
namespace pbrlab {
struct CyHairHeader {
  char magic[4];
  unsigned int num_strands;
  unsigned int total_points;
  unsigned int flags;
  unsigned int default_segments;
  float default_thickness;
  float default_transparency;
  float default_color[3];
  char infomation[88];
};
// One vector of xyz (and one of thickness) per strand.  is_y_up == false swaps y and z.
bool LoadCyHair(const std::string& filepath, const bool is_y_up, std::vector<std::vector<float>>* vertices,
                std::vector<std::vector<float>>* thicknesses);
}  // namespace pbrlab
#endif  // PBRLAB_B200_CYHAIR_H_
