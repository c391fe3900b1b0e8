#include "curve-util.h"

#include "type.h"

namespace pbrlab {
namespace {

const float kTau = 0.5f;   // Catmull-Rom tightness

// The three conversions work on anything with float-scalar * and +/-: float3 for positions, float for radii.
template <typename T>
void SegmentStart(const T& p0, const T& p1, const T& p2, T q[4]) {   // reference curve-util.cc:33-55
  const float tau3 = kTau / 3.0f;
  q[0] = p0;
  q[1] = ((kTau + 1.0f) / 3.0f) * p0 + (2.0f / 3.0f) * p1 - tau3 * p2;
  q[2] = tau3 * (p0 - p2) + p1;
  q[3] = p1;
}
template <typename T>
void SegmentMiddle(const T& p0, const T& p1, const T& p2, const T& p3, T q[4]) {   // :58-79
  const float tau3 = kTau / 3.0f;
  q[0] = p1;
  q[1] = tau3 * (p2 - p0) + p1;
  q[2] = tau3 * (p1 - p3) + p2;
  q[3] = p2;
}
template <typename T>
void SegmentEnd(const T& p0, const T& p1, const T& p2, T q[4]) {   // :7-31
  const float tau3 = kTau / 3.0f;
  q[0] = p1;
  q[1] = tau3 * (p2 - p0) + p1;
  q[2] = (-tau3) * p0 + (2.0f / 3.0f) * p1 + ((kTau + 1.0f) / 3.0f) * p2;
  q[3] = p2;
}

inline float3 Cv(const std::vector<float>& cvs, size_t i) { return float3(cvs[3 * i], cvs[3 * i + 1], cvs[3 * i + 2]); }

void Emit(const float3 q[4], const float r[4], std::vector<float>* bv, std::vector<float>* br) {
  for (int i = 0; i < 4; ++i) {
    bv->push_back(q[i][0]); bv->push_back(q[i][1]); bv->push_back(q[i][2]);
    br->push_back(r[i]);
  }
}

}  // namespace

bool ToCubicBezierCurve(const std::vector<float>& cvs, const std::vector<float>& cv_radiuss,
                        std::vector<float>* bezier_vertices, std::vector<float>* bezier_radiuss) {
  if (cvs.empty() || cv_radiuss.empty() || (cvs.size() % 3) != 0) return false;
  if (bezier_vertices->size() % 12 != 0 || bezier_radiuss->size() % 4 != 0 ||
      bezier_vertices->size() != bezier_radiuss->size() * 3)
    return false;
  const size_t n = cvs.size() / 3;
  if (n < 3 || n != cv_radiuss.size()) return false;   // every strand needs >= 3 vertices (:104-106)
  const size_t nseg = n - 1;
  const float* r = cv_radiuss.data();
  float3 q[4];
  float w[4];

  SegmentStart(Cv(cvs, 0), Cv(cvs, 1), Cv(cvs, 2), q);
  SegmentStart(r[0], r[1], r[2], w);
  Emit(q, w, bezier_vertices, bezier_radiuss);

  for (size_t s = 1; s + 1 < nseg; ++s) {
    const size_t k = s - 1;
    SegmentMiddle(Cv(cvs, k), Cv(cvs, k + 1), Cv(cvs, k + 2), Cv(cvs, k + 3), q);
    SegmentMiddle(r[k], r[k + 1], r[k + 2], r[k + 3], w);
    Emit(q, w, bezier_vertices, bezier_radiuss);
  }

  if (nseg > 1) {
    const size_t k = nseg - 2;
    SegmentEnd(Cv(cvs, k), Cv(cvs, k + 1), Cv(cvs, k + 2), q);
    SegmentEnd(r[k], r[k + 1], r[k + 2], w);
    Emit(q, w, bezier_vertices, bezier_radiuss);
  }
  return true;
}

}  // namespace pbrlab
