#include "cyhair.h"

#include <cstdio>
#include <cstring>

namespace pbrlab {

bool LoadCyHair(const std::string& filepath, const bool is_y_up, std::vector<std::vector<float>>* vertices,
                std::vector<std::vector<float>>* thicknesses) {
  static_assert(sizeof(CyHairHeader) == 128, "CyHair header is 128 bytes");
  FILE* fp = fopen(filepath.c_str(), "rb");
  if (!fp) return false;
  CyHairHeader h;
  if (fread(&h, 128, 1, fp) != 1 || memcmp(h.magic, "HAIR", 4) != 0) {
    fclose(fp);
    return false;
  }
  const bool has_segments = h.flags & 0x1, has_points = h.flags & 0x2, has_thickness = h.flags & 0x4;
  const int default_segments = int(h.default_segments);
  if (!has_points || (default_segments < 1 && !has_segments)) {
    fclose(fp);
    return false;
  }
  std::vector<unsigned short> segments;
  std::vector<float> points, thick;
  bool ok = true;
  if (has_segments) {
    segments.resize(h.num_strands);
    ok = ok && fread(segments.data(), sizeof(unsigned short) * h.num_strands, 1, fp) == 1;
  }
  if (ok) {
    points.resize(size_t(3) * h.total_points);
    ok = fread(points.data(), sizeof(float) * 3 * h.total_points, 1, fp) == 1;
  }
  if (ok && has_thickness) {
    thick.resize(h.total_points);
    ok = fread(thick.data(), sizeof(float) * h.total_points, 1, fp) == 1;
  }
  fclose(fp);   // transparency and colour arrays are not used by the renderer
  if (!ok) return false;

  size_t offset = 0;
  for (size_t s = 0; s < h.num_strands; ++s) {
    const size_t nseg = segments.empty() ? size_t(default_segments) : size_t(segments[s]);
    const size_t nv = nseg + 1;
    if (nv < 2) continue;   // as in the reference, the point offset is not advanced for a skipped strand
    if ((offset + nv) * 3 > points.size()) return false;
    vertices->emplace_back();
    thicknesses->emplace_back();
    std::vector<float>& v = vertices->back();
    std::vector<float>& t = thicknesses->back();
    v.reserve(nv * 3);
    t.reserve(nv);
    for (size_t i = 0; i < nv; ++i) {
      const float* p = &points[(offset + i) * 3];
      v.push_back(p[0]);
      v.push_back(is_y_up ? p[1] : p[2]);
      v.push_back(is_y_up ? p[2] : p[1]);
      t.push_back(thick.empty() ? h.default_thickness : thick[offset + i]);
    }
    offset += nv;
  }
  return true;
}

}  // namespace pbrlab
