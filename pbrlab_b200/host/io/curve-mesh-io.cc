#include "curve-mesh-io.h"

#include <cmath>
#include <iostream>
#include <limits>

#include "../curve-util.h"
#include "cyhair.h"

namespace pbrlab {
namespace io {

bool LoadCurveMeshAsCubicBezierCurve(const std::string& filepath, const bool memory_saving_mode,
                                     std::vector<float>* vt, std::vector<uint32_t>* indices) {
  const size_t dot = filepath.find_last_of('.');
  if (dot == std::string::npos || filepath.substr(dot) != ".hair") {
    std::cerr << "unknown data type" << std::endl;
    return false;
  }
  std::vector<std::vector<float>> strands, radii;
  LoadCyHair(filepath, true, &strands, &radii);
  if (radii.size() != strands.size()) return false;

  size_t index_base = 0;
  for (size_t s = 0; s < strands.size(); ++s) {
    std::vector<float> bv, br;
    const bool ok = ToCubicBezierCurve(strands[s], radii[s], &bv, &br);
    const size_t nv = br.size();
    if (!ok || bv.size() != nv * 3 || nv % 4 != 0) return false;   // one bad strand rejects the whole file
    auto push = [&](size_t v) {
      vt->push_back(bv[3 * v]); vt->push_back(bv[3 * v + 1]); vt->push_back(bv[3 * v + 2]);
      vt->push_back(br[v]);
    };
    if (memory_saving_mode) {
      const size_t nseg = nv / 4;
      const float eps = std::numeric_limits<float>::epsilon();
      for (size_t seg = 0; seg < nseg; ++seg) {
        indices->push_back(uint32_t(index_base + seg * 3));
        for (size_t c = 0; c < 3; ++c) push(seg * 4 + c);
        if (seg > 0) {   // consecutive segments must share their joint
          for (int k = 0; k < 3; ++k)
            if (!(std::fabs(bv[3 * seg * 4 + k] - bv[3 * seg * 4 + k - 3]) < eps)) return false;
        }
      }
      push(nv - 1);
    } else {
      for (size_t v = 0; v < nv; ++v) {
        if (v % 4 == 0) indices->push_back(uint32_t(index_base + v));
        push(v);
      }
    }
    index_base = vt->size() / 4;
  }
  return true;
}

bool LoadCurveMeshAsCubicBezierCurve(const std::string& filepath, const bool memory_saving_mode,
                                     CubicBezierCurveMesh* curve_mesh) {
  std::shared_ptr<CurveAttribute> attr(new CurveAttribute());
  std::vector<uint32_t> indices;
  // as in the reference (curve-mesh-io.cc:123-138) a failed read is not reported: the mesh just comes back
  // partially filled or empty
  LoadCurveMeshAsCubicBezierCurve(filepath, memory_saving_mode, &attr->vertices, &indices);
  const std::vector<uint32_t> material_ids(indices.size(), uint32_t(-1));
  *curve_mesh = CubicBezierCurveMesh(filepath, attr, indices, material_ids);
  return curve_mesh != nullptr;
}

}  // namespace io
}  // namespace pbrlab
