// Shared building blocks of the sm_100a path tracer: vector type, constants, bit casts.
//
// Every function in device/*.cuh is a plain per-path (per-ray, per-vertex) routine tagged PBR_HD so that the very
// same source can also be compiled by g++ into the host-emulation test shim (tests/host_emul/) and compared with
// the compiled reference in this GPU-less container.  The shipped library only ever runs them inside CUDA kernels.
//
// Arithmetic contract: the reference is built by g++ for plain x86-64 (-std=c++11, no FMA contraction), so every
// expression here is written in the reference's evaluation order and the library is compiled with -fmad=false;
// fused multiply-adds appear only where the reference itself calls std::fma (fast_math::Madd) or where bit parity
// does not matter (BVH slab tests) and are spelled pbr_fma().
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string.h>

#if defined(__CUDACC__)
#define PBR_HD __host__ __device__ __forceinline__
#define PBR_D __device__ __forceinline__
// large routines with several call sites per kernel: one out-of-line copy keeps the shading kernels inside the
// instruction cache (the fully inlined ShadeSurfaceKernel was 120 KB of SASS and stalled on instruction fetch)
#define PBR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define PBR_HD inline
#define PBR_D inline
#define PBR_HD_NOINLINE inline
#include <algorithm>
#include <cmath>
#endif
#if !defined(__CUDACC__) && !defined(__VECTOR_TYPES_H__)
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
#endif

namespace pbr {

constexpr float kPi    = 3.141592653589793f;  // src/pbrlab_math.h:7
constexpr float kPiInv = 0.318309886183f;     // src/pbrlab_math.h:8
constexpr float kEps   = 1e-3f;               // src/pbrlab_math.h:10
constexpr float kInf   = 1.844E18f;           // src/pbrlab_math.h:11
constexpr float kFltEps = 1.1920928955078125e-07f;  // std::numeric_limits<float>::epsilon()
constexpr uint32_t kInvalid = 0xFFFFFFFFu;

PBR_HD float pbr_fma(float a, float b, float c) { return fmaf(a, b, c); }

PBR_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PBR_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
PBR_HD float fminf_(float a, float b) { return b < a ? b : a; }  // std::min semantics, NaN behaviour included
PBR_HD float fmaxf_(float a, float b) { return a < b ? b : a; }  // std::max semantics, NaN behaviour included
PBR_HD bool finitef_(float x) { return (f2u(x) & 0x7f800000u) != 0x7f800000u; }

// float3 with the semantics of nanort::real3<float> (src/nanort.h:314-404): component-wise * and /, no fusing.
struct vec3 {
  float x, y, z;
  PBR_HD vec3() {}
  PBR_HD explicit vec3(float s) : x(s), y(s), z(s) {}
  PBR_HD vec3(float xx, float yy, float zz) : x(xx), y(yy), z(zz) {}
  PBR_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
PBR_HD vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
PBR_HD vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
PBR_HD vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
PBR_HD vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
PBR_HD vec3 operator*(const vec3& a, float f) { return vec3(a.x * f, a.y * f, a.z * f); }
PBR_HD vec3 operator*(float f, const vec3& a) { return vec3(a.x * f, a.y * f, a.z * f); }
PBR_HD vec3 operator/(const vec3& a, float f) { return vec3(a.x / f, a.y / f, a.z / f); }  // real3 / real3(f)
PBR_HD vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
PBR_HD float vdot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PBR_HD vec3 vcross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PBR_HD float vlength(const vec3& v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
// nanort::vnormalize: vectors no longer than FLT_EPSILON are returned untouched (src/nanort.h:380-390)
PBR_HD vec3 vnormalized(const vec3& v) {
  const float len = vlength(v);
  if (fabsf(len) > kFltEps) {
    const float inv = 1.0f / len;
    return vec3(v.x * inv, v.y * inv, v.z * inv);
  }
  return v;
}
PBR_HD vec3 from4(const float4& v) { return vec3(v.x, v.y, v.z); }

PBR_HD float Sqr(float v) { return v * v; }
PBR_HD float SafeSqrtf(float f) { return sqrtf(fmaxf_(f, 0.0f)); }            // src/pbrlab_math.h:17
PBR_HD float Clampf(float x, float a, float b) { return fmaxf_(a, fminf_(b, x)); }  // src/pbrlab-util.h:9-12
PBR_HD float Saturatef(float x) { return Clampf(x, 0.f, 1.f); }
PBR_HD float Average(const vec3& c) { return (c.x + c.y + c.z) / 3.f; }        // src/pbrlab-util.h:19
PBR_HD float SpectrumNorm(const vec3& c) { return fmaxf_(fmaxf_(c.x, c.y), c.z); }  // std::max({..}) :21-23
PBR_HD float RgbToY(const vec3& c) {                                            // src/pbrlab-util.h:48-51
  return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z;
}
PBR_HD bool IsBlack(const vec3& v) { return (fabsf(v.x) + fabsf(v.y) + fabsf(v.z)) < kFltEps; }  // :53-56
PBR_HD bool IsFinite3(const vec3& v) { return finitef_(v.x) && finitef_(v.y) && finitef_(v.z); }
PBR_HD vec3 SafeDivideSpectrum(const vec3& a, const vec3& b) {                   // src/pbrlab-util.h:25-46
  return vec3(fabsf(b.x) < kFltEps ? 0.f : a.x / b.x, fabsf(b.y) < kFltEps ? 0.f : a.y / b.y,
              fabsf(b.z) < kFltEps ? 0.f : a.z / b.z);
}
PBR_HD vec3 Lerp3v(const vec3& v0, const vec3& v1, float u) { return (1.0f - u) * v0 + u * v1; }  // pbrlab_math.h:29-32
// Lerp3: (1-u-v)*v0 + u*v1 + v*v2, left to right (src/pbrlab_math.h:34-38)
PBR_HD vec3 Lerp3(const vec3& v0, const vec3& v1, const vec3& v2, float u, float v) {
  return ((1.0f - u - v) * v0 + u * v1) + v * v2;
}

}  // namespace pbr
