// float3 / float2 value types of the host-side API.  The reference aliases nanort::real3<float>
// (reference src/type.h:8, src/nanort.h:314-404); this is an independent minimal equivalent with the same
// observable semantics: v[i] access, component-wise * and /, scalar broadcast constructor.
#ifndef PBRLAB_B200_TYPE_H_
#define PBRLAB_B200_TYPE_H_
#include <cmath>
#include <cstdint>
#include <limits>

namespace pbrlab {

struct float3 {
  float v[3];
  float3() {}
  float3(float s) { v[0] = v[1] = v[2] = s; }  // NOLINT: implicit like real3(T x)
  float3(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
  explicit float3(const float* p) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
  float operator[](int i) const { return v[i]; }
  float& operator[](int i) { return v[i]; }
  float3 operator+(const float3& o) const { return float3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  float3 operator-(const float3& o) const { return float3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  float3 operator*(const float3& o) const { return float3(v[0] * o.v[0], v[1] * o.v[1], v[2] * o.v[2]); }
  float3 operator/(const float3& o) const { return float3(v[0] / o.v[0], v[1] / o.v[1], v[2] / o.v[2]); }
  float3 operator*(float f) const { return float3(v[0] * f, v[1] * f, v[2] * f); }
  float3 operator-() const { return float3(-v[0], -v[1], -v[2]); }
};
inline float3 operator*(float f, const float3& a) { return float3(a.v[0] * f, a.v[1] * f, a.v[2] * f); }
inline float vdot(const float3& a, const float3& b) { return a.v[0] * b.v[0] + a.v[1] * b.v[1] + a.v[2] * b.v[2]; }
inline float3 vcross(const float3& a, const float3& b) {
  return float3(a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]);
}
inline float vlength(const float3& a) { return std::sqrt(a.v[0] * a.v[0] + a.v[1] * a.v[1] + a.v[2] * a.v[2]); }
inline float3 vnormalized(const float3& a) {
  const float len = vlength(a);
  if (std::fabs(len) > std::numeric_limits<float>::epsilon()) {
    const float inv = 1.0f / len;
    return float3(a.v[0] * inv, a.v[1] * inv, a.v[2] * inv);
  }
  return a;
}

struct float2 {
  float v[2];
  float2() {}
  float2(float s) { v[0] = v[1] = s; }  // NOLINT
  float2(float x, float y) { v[0] = x; v[1] = y; }
  explicit float2(const float* p) { v[0] = p[0]; v[1] = p[1]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float operator[](int i) const { return v[i]; }
  float& operator[](int i) { return v[i]; }
  float2 operator+(const float2& o) const { return float2(v[0] + o.v[0], v[1] + o.v[1]); }
  float2 operator*(float f) const { return float2(v[0] * f, v[1] * f); }
};
inline float2 operator*(float f, const float2& a) { return float2(a.v[0] * f, a.v[1] * f); }

constexpr float kPi = 3.141592653589793f;
constexpr float kEps = 1e-3f;
constexpr float kInf = 1.844E18f;

}  // namespace pbrlab
#endif  // PBRLAB_B200_TYPE_H_
