// Closest-hit / any-hit traversal of one ray through the compressed 8-wide BVH (Ylitie, Karras, Laine 2017 layout,
// 80-byte nodes, see ../bvh_builder.h), with the primitive tests restating what the reference's ray engine does:
//   * triangles: Moeller-Trumbore on precomputed (v0, e1 = v0-v1, e2 = v2-v0), Ng = e2 x e1, inclusive edges,
//     no back-face culling, tnear < t <= tfar  (Embree kernels/geometry/triangle_intersector_moeller.h:69-110,
//     kernels/geometry/triangle.h:40-41);
//   * flat cubic Bezier curves: 4 ray-facing quads per segment in ray space, back-face culled quad test,
//     tnear <= t <= tfar, self-intersection rejection t <= 2 r depth_scale, u = (i+U)/4, v in [-1,1],
//     Ng = dB/du  (kernels/geometry/curve_intersector_ribbon.h:72-177, quad_intersector.h:15-74,
//     curve_intersector_precalculations.h:20-26, kernels/subdiv/bezier_curve.h:12-51).
// The reference reaches these through Scene::TraceFirstHit1 / AnyHit1 (src/scene.cc:261-268,
// src/raytracer/raytracer_impl.cc:268-287).
#pragma once
#include "common.cuh"
#include "scene_view.cuh"

namespace pbr {

struct RayT {
  vec3 o, d;
  float tmin, tmax;
};
struct HitT {
  float t, u, v;
  uint32_t prim;   // triangle: leaf-order index; curve: bit 31 | segment slot; kInvalid = miss
};
constexpr uint32_t kCurveFlag = 0x80000000u;
constexpr int kStackSize = 32;

struct TraverseStats { uint32_t nodes, prims; };

// Embree's Vec3 dot/cross association on the SSE2 path (common/math/vec3.h:205-209, madd = a*b+c unfused)
PBR_HD float edot(const vec3& a, const vec3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
PBR_HD vec3 ecross(const vec3& a, const vec3& b) {
  return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

PBR_HD bool IntersectTriangle(const vec3& O, const vec3& D, float tnear, float tfar, const vec3& v0, const vec3& e1,
                              const vec3& e2, float* t, float* u, float* v) {
  const vec3 Ng = ecross(e2, e1);
  const vec3 C = v0 - O;
  const vec3 R = ecross(C, D);
  const float den = edot(Ng, D);
  const float absDen = fabsf(den);
  const uint32_t sgn = f2u(den) & 0x80000000u;
  const float U = u2f(f2u(edot(R, e2)) ^ sgn);
  const float V = u2f(f2u(edot(R, e1)) ^ sgn);
  if (!((den != 0.0f) & (U >= 0.0f) & (V >= 0.0f) & (U + V <= absDen))) return false;
  const float T = u2f(f2u(edot(Ng, C)) ^ sgn);
  if (!((absDen * tnear < T) & (T <= absDen * tfar))) return false;
  const float rcp = 1.0f / absDen;
  *t = T * rcp;
  *u = U * rcp;
  *v = V * rcp;
  return true;
}

// per-ray precalculation for curves (curve_intersector_precalculations.h:20-26, linearspace3.h:117-124)
struct CurveRaySpace {
  vec3 dx, dy, dz;   // rows of the transposed frame: p_ray = (dot(dx,p), dot(dy,p), dot(dz,p))
  float depth_scale;
};
PBR_HD uint32_t lowest_bit(uint32_t x) {   // x != 0
#if defined(__CUDA_ARCH__)
  return uint32_t(__ffs(int(x)) - 1);
#else
  return uint32_t(__builtin_ctz(x));
#endif
}
PBR_HD vec3 enormalize(const vec3& v) { return v * (1.0f / sqrtf(edot(v, v))); }
PBR_HD CurveRaySpace MakeCurveRaySpace(const vec3& D) {
  CurveRaySpace s;
  s.depth_scale = 1.0f / sqrtf(edot(D, D));
  const vec3 N = s.depth_scale * D;
  const vec3 dx0(0.f, N.z, -N.y);
  const vec3 dx1(-N.z, 0.f, N.x);
  s.dx = enormalize(edot(dx0, dx0) > edot(dx1, dx1) ? dx0 : dx1);
  s.dy = enormalize(ecross(N, s.dx));
  s.dz = N * s.depth_scale;
  return s;
}
// xfmVector(ray_space, p): p.x*col0 + (p.y*col1 + p.z*col2)
PBR_HD vec3 ToRaySpace(const CurveRaySpace& s, const vec3& p) {
  return vec3(p.x * s.dx.x + (p.y * s.dx.y + p.z * s.dx.z), p.x * s.dy.x + (p.y * s.dy.y + p.z * s.dy.z),
              p.x * s.dz.x + (p.y * s.dz.y + p.z * s.dz.z));
}

struct vec4 { float x, y, z, w; };
PBR_HD void BezierBasis(float u, float* b) {             // bezier_curve.h:16-26
  const float t1 = u, t0 = 1.0f - t1;
  b[0] = t0 * t0 * t0;
  b[1] = 3.0f * t1 * (t0 * t0);
  b[2] = 3.0f * (t1 * t1) * t0;
  b[3] = t1 * t1 * t1;
}
PBR_HD void BezierDerivative(float u, float* b) {        // bezier_curve.h:28-38
  const float t1 = u, t0 = 1.0f - t1;
  b[0] = 3.0f * (-(t0 * t0));
  b[1] = 3.0f * (-2.0f * (t0 * t1) + t0 * t0);
  b[2] = 3.0f * (2.0f * (t0 * t1) - t1 * t1);
  b[3] = 3.0f * (t1 * t1);
}
PBR_HD float bz(const float* b, float a0, float a1, float a2, float a3) {   // madd(b0,v0,madd(b1,v1,madd(b2,v2,b3*v3)))
  return b[0] * a0 + (b[1] * a1 + (b[2] * a2 + b[3] * a3));
}

// tangent dB/du at u from the world-space control points (RibbonHit::Ng, curve_intersector_ribbon.h:35)
PBR_HD vec3 CurveTangent(const float4& c0, const float4& c1, const float4& c2, const float4& c3, float u) {
  float b[4];
  BezierDerivative(u, b);
  return vec3(bz(b, c0.x, c1.x, c2.x, c3.x), bz(b, c0.y, c1.y, c2.y, c3.y), bz(b, c0.z, c1.z, c2.z, c3.z));
}

// Cheap conservative rejection ahead of IntersectCurve.  The ribbon of a segment is four ray-facing quads between
// consecutive points B(i/4) of the curve, each quad inside the capsule of radius max(r_i, r_i+1) around its chord
// (the corners are at distance r from the chord's ends and a capsule is convex); every point of the curve, hence of
// those chords, lies within `dev` = max distance of the inner control points from the line c0c3 (convex hull).  So a
// ray that hits the ribbon passes within R = r_max + dev of the LINE through c0 and c3: if the distance between the
// two lines is larger, the full test cannot succeed.  The test reads its own 32-byte record per segment
// (c0, R with its safety factor | c3 - c0, its length), built at commit (scene_host.cc), so the 64 bytes of control
// points are only fetched for the one candidate in six that passes; the second term bounds the rounding error of
// the triple product (2^-22 |w| |D| |e|).
PBR_HD bool CurveMayHit(const vec3& O, const vec3& D, const float4& a, const float4& b) {
  // a = (c0.xyz, R), b = (c3 - c0, |c3 - c0|)
  const vec3 e(b.x, b.y, b.z);
  const vec3 w(a.x - O.x, a.y - O.y, a.z - O.z);
  const vec3 n = ecross(D, e);
  const float q = edot(w, n);
  const float nn = edot(n, n), ww = edot(w, w), dd = edot(D, D);
  return !(fabsf(q) > a.w * sqrtf(nn) + 3e-7f * (sqrtf(ww * dd) * b.w));
}

// One cubic segment against the ray: Embree's ribbon intersector (curve_intersector_ribbon.h:72-177) — the curve is
// cut at u = 0, 1/4, .. 1 into four ray-facing quads; the nearest accepted quad hit wins (lowest sub-segment on ties).
// Two stages so that the lanes of a warp stay together: (1) the five points in ray space and, per sub-segment, the
// two 2-D rejections (a bit mask); (2) the quad test, once per surviving sub-segment.
//   * cylinder_culling_test: distance from the ray (the 2-D origin) to the LINE p0p1 <= max(r0, r1) — Embree's own;
//   * ours: the foot of that perpendicular must lie within max(r0, r1) of the SEGMENT p0p1.  The quad's corners are
//     p0 +- r0 n0, p1 +- r1 n1 with unit n in the xy plane, all inside the (convex) 2-D capsule of radius max(r0, r1)
//     around p0p1, and so is the quad; a ray outside the capsule cannot hit it.  Near-collinear sub-segments all pass
//     the line test, this one keeps the one or two the ray actually crosses.
// [first_quad, first_quad + num_quads): the quads this BVH primitive stands for (0, 4 = the whole segment).
PBR_HD bool IntersectCurve(const vec3& O, const CurveRaySpace& rs, float tnear, float tfar, const float4& c0,
                           const float4& c1, const float4& c2, const float4& c3, uint32_t first_quad,
                           uint32_t num_quads, float* t_out, float* u_out, float* v_out) {
  // control points in ray space (xfm_pr): position relative to the origin, radius carried in w
  const vec3 q0 = ToRaySpace(rs, vec3(c0.x, c0.y, c0.z) - O);
  const vec3 q1 = ToRaySpace(rs, vec3(c1.x, c1.y, c1.z) - O);
  const vec3 q2 = ToRaySpace(rs, vec3(c2.x, c2.y, c2.z) - O);
  const vec3 q3 = ToRaySpace(rs, vec3(c3.x, c3.y, c3.z) - O);
  float m = 0.f;
  m = fmaxf(m, fmaxf(fmaxf(fabsf(q0.x), fabsf(q0.y)), fabsf(q0.z)));
  m = fmaxf(m, fmaxf(fmaxf(fabsf(q1.x), fabsf(q1.y)), fabsf(q1.z)));
  m = fmaxf(m, fmaxf(fmaxf(fabsf(q2.x), fabsf(q2.y)), fabsf(q2.z)));
  m = fmaxf(m, fmaxf(fmaxf(fabsf(q3.x), fabsf(q3.y)), fabsf(q3.z)));
  const float eps = 4.0f * kFltEps * m;

  // ---- stage 1: which sub-segments can be hit (xy and radius of the five points only)
  uint32_t mask = 0u;
  {
    float px0, py0, pr0;
    {
      float b[4];
      BezierBasis(float(first_quad) / 4.0f, b);
      px0 = bz(b, q0.x, q1.x, q2.x, q3.x); py0 = bz(b, q0.y, q1.y, q2.y, q3.y); pr0 = bz(b, c0.w, c1.w, c2.w, c3.w);
    }
    for (uint32_t i = first_quad; i < first_quad + num_quads; ++i) {
      float b[4];
      BezierBasis(float(i + 1) / 4.0f, b);
      const float px1 = bz(b, q0.x, q1.x, q2.x, q3.x), py1 = bz(b, q0.y, q1.y, q2.y, q3.y),
                  pr1 = bz(b, c0.w, c1.w, c2.w, c3.w);
      // cylinder_culling_test(0, p0.xy, p1.xy, max(r0, r1))
      const float ax = px1 - px0, ay = py1 - py0;
      const float bx = px0 - 0.f, by = py0 - 0.f;
      const float num = ax * by - ay * bx;
      const float den2 = ax * ax + ay * ay;
      const float r = fmaxf(pr0, pr1);
      bool keep = num * num <= r * r * den2;
      // capsule ends: s = -(b.a) is the foot's parameter times |a|^2; outside [0, |a|^2] by more than r |a| -> no hit
      const float sfoot = -(bx * ax + by * ay);
      const float lim = r * r * den2 * 1.001f + 1e-12f * (m * m) * den2;
      const float over = sfoot - den2;
      if ((sfoot < 0.f && sfoot * sfoot > lim) || (over > 0.f && over * over > lim)) keep = false;
      if (keep) mask |= 1u << i;
      px0 = px1; py0 = py1; pr0 = pr1;
    }
  }

  // ---- stage 2: the quad of every surviving sub-segment, in ascending order
  bool found = false;
  float best_t = 0.f, best_u = 0.f, best_v = 0.f;
  // (straight-line body, no early exits: the lanes of a warp that are in this loop stay converged)
  while (mask != 0u) {
    const int i = int(lowest_bit(mask));
    mask &= mask - 1u;
    float b0[4], b1[4], d0[4], d1[4];
    BezierBasis(float(i) / 4.0f, b0);
    BezierBasis(float(i + 1) / 4.0f, b1);
    const vec4 p0 = {bz(b0, q0.x, q1.x, q2.x, q3.x), bz(b0, q0.y, q1.y, q2.y, q3.y), bz(b0, q0.z, q1.z, q2.z, q3.z),
                     bz(b0, c0.w, c1.w, c2.w, c3.w)};
    const vec4 p1 = {bz(b1, q0.x, q1.x, q2.x, q3.x), bz(b1, q0.y, q1.y, q2.y, q3.y), bz(b1, q0.z, q1.z, q2.z, q3.z),
                     bz(b1, c0.w, c1.w, c2.w, c3.w)};
    BezierDerivative(float(i) / 4.0f, d0);
    BezierDerivative(float(i + 1) / 4.0f, d1);
    vec3 dp0(bz(d0, q0.x, q1.x, q2.x, q3.x), bz(d0, q0.y, q1.y, q2.y, q3.y), bz(d0, q0.z, q1.z, q2.z, q3.z));
    vec3 dp1(bz(d1, q0.x, q1.x, q2.x, q3.x), bz(d1, q0.y, q1.y, q2.y, q3.y), bz(d1, q0.z, q1.z, q2.z, q3.z));
    const vec3 chord(p1.x - p0.x, p1.y - p0.y, p1.z - p0.z);
    if (fmaxf(fmaxf(fabsf(dp0.x), fabsf(dp0.y)), fabsf(dp0.z)) < eps) dp0 = chord;
    if (fmaxf(fmaxf(fabsf(dp1.x), fabsf(dp1.y)), fabsf(dp1.z)) < eps) dp1 = chord;
    const vec3 nn0 = enormalize(vec3(dp0.y, -dp0.x, 0.0f));
    const vec3 nn1 = enormalize(vec3(dp1.y, -dp1.x, 0.0f));
    const vec3 P0(p0.x, p0.y, p0.z), P1(p1.x, p1.y, p1.z);
    const vec3 lp0 = p0.w * nn0 + P0, lp1 = p1.w * nn1 + P1;       // madd(w, nn, p)
    const vec3 up0 = P0 - p0.w * nn0, up1 = P1 - p1.w * nn1;       // nmadd(w, nn, p)

    // intersect_quad_backface_culling(O = 0, D = (0,0,1), quad = lp0, lp1, up1, up0)
    const vec3 va = lp0, vb = lp1, vc = up1, vd = up0;
    const vec3 edb = vb - vd;
    const float WW = ecross(vd, edb).z;
    const bool first = WW <= 0.0f;
    const vec3 v0 = first ? va : vc;
    const vec3 v1 = first ? vb : vd;
    const vec3 v2 = first ? vd : vb;
    const vec3 e0 = v2 - v0;
    const vec3 e1 = v0 - v1;
    const float U = ecross(v0, e0).z;
    const float V = ecross(v1, e1).z;
    bool ok = fmaxf(U, V) <= 0.0f;
    const vec3 Ng = ecross(e1, e0);
    const float den = Ng.z;
    const float rcpDen = 1.0f / den;
    const float t = rcpDen * edot(v0, Ng);
    ok = ok & (tnear <= t) & (t <= tfar) & (den != 0.0f);
    float u = U * rcpDen, v = V * rcpDen;
    u = first ? u : 1.0f - u;
    v = first ? v : 1.0f - v;
    // self-intersection avoidance (EMBREE_CURVE_SELF_INTERSECTION_AVOIDANCE_FACTOR 2.0)
    const float r = u * (p1.w - p0.w) + p0.w;                        // lerp = madd(t, b-a, a)
    ok = ok & (t > 2.0f * r * rs.depth_scale);
    if (ok && (!found || t < best_t)) {   // select_min over the 4 lanes: lowest t, lowest lane on ties
      found = true;
      best_t = t;
      best_u = (float(i) + u + 0.0f) * (1.0f / 4.0f);
      best_v = 2.0f * v + -1.0f;
    }
  }
  if (found) { *t_out = best_t; *u_out = best_u; *v_out = best_v; }
  return found;
}

PBR_HD uint32_t extract_byte(uint32_t x, uint32_t i) { return (x >> (i * 8)) & 0xffu; }
PBR_HD uint32_t sign_extend_s8x4(uint32_t x) {
  // each byte: 0x80 -> 0xff, else 0x00 (only ever called with bytes in {0x00, 0x80})
  return ((x >> 7) & 0x01010101u) * 0xffu;
}
PBR_HD uint32_t msb(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return 31u - uint32_t(__clz(int(x)));
#else
  return 31u - uint32_t(__builtin_clz(x));
#endif
}
PBR_HD uint32_t popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return uint32_t(__popc(x));
#else
  return uint32_t(__builtin_popcount(x));
#endif
}

// byte j of x as the float 32768 + byte: the byte goes to mantissa bits 8..15 of 2^15 (one PRMT on the device instead
// of an int->float conversion on the quarter-rate XU pipe, which ncu showed as the busiest pipe of the traversal)
// `magic` is 0x47000000 and MUST reach the kernel as a run-time value (SceneView::bias_magic): PRMT takes one
// immediate, and when both the constant and the selector are known at compile time ptxas keeps the constant as the
// immediate and re-materialises the selector in front of every one of the 48 PRMTs of a node test (UMOV + IMAD.U32,
// seen in the SASS: 17 % of the node step).  With the constant in a register the selector is the immediate.
constexpr uint32_t kBiasMagic = 0x47000000u;
PBR_HD float byte_as_biased_float(uint32_t x, uint32_t j, uint32_t magic) {
#if defined(__CUDA_ARCH__)
  // result bytes: b0 = 0x00 (magic.b0), b1 = x.bj, b2 = 0x00 (magic.b2), b3 = 0x47 (magic.b3)
  return __uint_as_float(__byte_perm(x, magic, 0x7604u | (j << 4)));
#else
  return u2f(0x47000000u | (((x >> (8u * j)) & 0xffu) << 8));
#endif
}

// Slab-test the 8 quantised child boxes of one node; returns the hit mask: bits 24..31 = internal children in
// octant-adjusted traversal order, bits 0..23 = primitives of hit leaf children (relative to the node's prim base).
// t = (p + q*2^e - o) / d is evaluated as fma(32768 + q, s, c) with s = 2^e/d and c = (p - o)/d - 32768 s.  c carries a
// rounding error of up to |s|/512 (1/512 of a grid step), so the near side is pushed out by |s|/256 and the far side
// by the same plus 4 ulps: the test stays conservative, which is all a box test has to be.
// WANT_INSIDE: *inside receives, at bits 16..23 (the hit bits of the inner children shifted down by 8), the hit inner
// children whose box holds the ray origin (the ray enters the box before tmin) — see PopChild.
template <bool WANT_INSIDE>
PBR_HD uint32_t NodeIntersectT(const vec3& o_over_d, const vec3& inv_d, uint32_t oct_inv4, bool neg_x, bool neg_y,
                               bool neg_z, float tmin, float tmax, const float4& n0, const float4& n1,
                               const float4& n2, const float4& n3, const float4& n4, uint32_t magic,
                               uint32_t* inside) {
  const uint32_t ew = f2u(n0.w);
  const float sx = u2f(extract_byte(ew, 0) << 23) * inv_d.x;
  const float sy = u2f(extract_byte(ew, 1) << 23) * inv_d.y;
  const float sz = u2f(extract_byte(ew, 2) << 23) * inv_d.z;
  const float ox = pbr_fma(-32768.0f, sx, pbr_fma(n0.x, inv_d.x, -o_over_d.x));   // (p - o)/d - 32768 s
  const float oy = pbr_fma(-32768.0f, sy, pbr_fma(n0.y, inv_d.y, -o_over_d.y));
  const float oz = pbr_fma(-32768.0f, sz, pbr_fma(n0.z, inv_d.z, -o_over_d.z));
  const float mx = fabsf(sx) * (1.0f / 256.0f), my = fabsf(sy) * (1.0f / 256.0f), mz = fabsf(sz) * (1.0f / 256.0f);
  const float olx = ox - mx, oly = oy - my, olz = oz - mz;   // near side
  const float ohx = ox + mx, ohy = oy + my, ohz = oz + mz;   // far side
  const float tmax_s = tmax * 1.0000004f + 1e-30f;
  uint32_t hit_mask = 0, inside_mask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < 2; ++i) {
    const uint32_t meta4 = f2u(i == 0 ? n1.z : n1.w);
    const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = sign_extend_s8x4(is_inner4 << 3);
    const uint32_t bit_index4 = (meta4 ^ (oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
    const uint32_t q_lo_x = f2u(i == 0 ? n2.x : n2.y), q_lo_y = f2u(i == 0 ? n2.z : n2.w);
    const uint32_t q_lo_z = f2u(i == 0 ? n3.x : n3.y), q_hi_x = f2u(i == 0 ? n3.z : n3.w);
    const uint32_t q_hi_y = f2u(i == 0 ? n4.x : n4.y), q_hi_z = f2u(i == 0 ? n4.z : n4.w);
    const uint32_t x_min = neg_x ? q_hi_x : q_lo_x, x_max = neg_x ? q_lo_x : q_hi_x;
    const uint32_t y_min = neg_y ? q_hi_y : q_lo_y, y_max = neg_y ? q_lo_y : q_hi_y;
    const uint32_t z_min = neg_z ? q_hi_z : q_lo_z, z_max = neg_z ? q_lo_z : q_hi_z;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t j = 0; j < 4; ++j) {
      const float tlx = pbr_fma(byte_as_biased_float(x_min, j, magic), sx, olx);
      const float tly = pbr_fma(byte_as_biased_float(y_min, j, magic), sy, oly);
      const float tlz = pbr_fma(byte_as_biased_float(z_min, j, magic), sz, olz);
      const float thx = pbr_fma(byte_as_biased_float(x_max, j, magic), sx, ohx);
      const float thy = pbr_fma(byte_as_biased_float(y_max, j, magic), sy, ohy);
      const float thz = pbr_fma(byte_as_biased_float(z_max, j, magic), sz, ohz);
      if (WANT_INSIDE) {
        const float m3 = fmaxf(fmaxf(tlx, tly), tlz);
        const float tn = fmaxf(m3, tmin);   // (the same value as below: max is exact)
        const float tf = fminf(fminf(thx, thy), fminf(thz, tmax_s));
        if (tn <= tf) {
          const uint32_t c = extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
          hit_mask |= c;
          if (m3 <= tmin) inside_mask |= c;   // (one predicated OR; the leaves' bits, below bit 24, are masked off at the end)
        }
      } else {
        const float tn = fmaxf(fmaxf(tlx, tly), fmaxf(tlz, tmin));
        const float tf = fminf(fminf(thx, thy), fminf(thz, tmax_s));
        if (tn <= tf) {
          hit_mask |= extract_byte(child_bits4, j) << extract_byte(bit_index4, j);
        }
      }
    }
  }
  if (WANT_INSIDE) *inside = (inside_mask & 0xff000000u) >> 8;
  return hit_mask;
}

PBR_HD uint32_t NodeIntersect(const vec3& o_over_d, const vec3& inv_d, uint32_t oct_inv4, bool neg_x, bool neg_y,
                              bool neg_z, float tmin, float tmax, const float4& n0, const float4& n1,
                              const float4& n2, const float4& n3, const float4& n4,
                              uint32_t magic = kBiasMagic) {
  return NodeIntersectT<false>(o_over_d, inv_d, oct_inv4, neg_x, neg_y, neg_z, tmin, tmax, n0, n1, n2, n3, n4, magic, nullptr);
}

// Which hit child of a node group a ray visits next.  group_y = hits << 24 | inside << 16 | imask.  The static order —
// the highest hit bit, i.e. (slot XOR inverse ray octant) — is near-to-far for boxes that do not overlap.  The boxes
// of a curve BVH do (thin diagonal quads): a ray that starts inside the hair volume lies INSIDE several child boxes,
// and the static order often descends into a farther child before the one around the origin, where the nearest hit
// is.  Curve BVHs therefore visit the children whose box holds the ray origin first, the rest in the static order:
// 28.8 -> 23.1 node visits and 9.8 -> 7.0 candidate tests per secondary hair ray on the C3 / C4 hair ball
// (scripts/hair_visit_order.py; sorting all hit children by entry distance would give 22.6 / 6.5 at the price of a
// sorting network per node).  The closest hit does not depend on the order (up to exactly tied t).
template <bool INSIDE_FIRST>
PBR_HD uint32_t PopChild(uint32_t* group_y) {
  uint32_t bit;
  if (INSIDE_FIRST) {
    const uint32_t inside = (*group_y >> 16) & 0xffu;
    const uint32_t p = inside ? msb(inside) : msb(*group_y >> 24);
    bit = 24u + p;
    *group_y &= ~((1u << bit) | (1u << (16u + p)));
  } else {
    bit = msb(*group_y);
    *group_y &= ~(1u << bit);
  }
  return bit;
}

#ifdef PBR_TRI_INSIDE_FIRST   // measurement switch (tests/host_emul): the visit order of PopChild for triangle BVHs too
constexpr bool kTriInsideFirst = true;
#else
constexpr bool kTriInsideFirst = false;
#endif
// Generic traversal of one BVH.  CURVES selects the leaf test, ANY the early-out.
template <bool CURVES, bool ANY, bool STATS>
PBR_HD bool TraverseBvh(const float4* __restrict__ nodes, const float4* __restrict__ prims,
                        const float4* __restrict__ cull, const uint32_t* __restrict__ sub, uint32_t part_quads,
                        const RayT& ray, float* tfar_io, HitT* hit, TraverseStats* st) {
  const vec3 O = ray.o, D = ray.d;
  // slab-test direction: zero components are nudged so 1/d stays finite (sign kept)
  const float tiny = 1e-30f;
  const vec3 ds(fabsf(D.x) < tiny ? copysignf(tiny, D.x) : D.x, fabsf(D.y) < tiny ? copysignf(tiny, D.y) : D.y,
                fabsf(D.z) < tiny ? copysignf(tiny, D.z) : D.z);
  const vec3 inv_d(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
  const vec3 o_over_d(O.x * inv_d.x, O.y * inv_d.y, O.z * inv_d.z);
  const bool neg_x = ds.x < 0.f, neg_y = ds.y < 0.f, neg_z = ds.z < 0.f;
  const uint32_t oct_inv4 = (neg_x ? 0u : 0x04040404u) | (neg_y ? 0u : 0x02020202u) | (neg_z ? 0u : 0x01010101u);
  CurveRaySpace rs;
  if (CURVES) rs = MakeCurveRaySpace(D);

  float tfar = *tfar_io;
  bool found = false;
  uint2 stack[kStackSize];
  int sp = 0;
  uint2 group = make_uint2(0u, 0x80000000u);   // root: node base 0, one "child" at bit 31 -> slot 7^oct... see below
  // The root is addressed as if it were child slot (7 ^ oct_inv) of a virtual parent whose only child it is:
  // relative index = popc(imask below slot) = 0 because the low 8 bits (imask) are 0.
  for (;;) {
    uint2 pgroup;
    if (group.y & 0xff000000u) {
      const uint32_t hits_imask = group.y;
      const uint32_t child_bit = PopChild<CURVES || kTriInsideFirst>(&group.y);
      if (group.y & 0xff000000u) {
        if (sp < kStackSize) stack[sp++] = group;
      }
      const uint32_t slot = (child_bit - 24u) ^ (oct_inv4 & 0xffu);
      const uint32_t rel = popc(hits_imask & ~(0xffffffffu << slot));
      const uint32_t node = group.x + rel;
      const float4 n0 = nodes[node * 5 + 0], n1 = nodes[node * 5 + 1], n2 = nodes[node * 5 + 2];
      const float4 n3 = nodes[node * 5 + 3], n4 = nodes[node * 5 + 4];
      if (STATS) st->nodes++;
      uint32_t inside = 0;
      const uint32_t hitmask = NodeIntersectT<CURVES || kTriInsideFirst>(o_over_d, inv_d, oct_inv4, neg_x, neg_y, neg_z, ray.tmin, tfar, n0,
                                                      n1, n2, n3, n4, kBiasMagic, &inside);
      group.x = f2u(n1.x);
      group.y = (hitmask & 0xff000000u) | inside | extract_byte(f2u(n0.w), 3);
      pgroup.x = f2u(n1.y);
      pgroup.y = hitmask & 0x00ffffffu;
    } else {
      pgroup = group;
      group = make_uint2(0u, 0u);
    }
    while (pgroup.y != 0u) {
      const uint32_t bit = msb(pgroup.y);
      pgroup.y &= ~(1u << bit);
      uint32_t idx = pgroup.x + bit;
      if (STATS) st->prims++;
      float t, u, v;
      bool h;
      if (CURVES) {
        h = false;
        const uint32_t code = sub[idx];   // the BVH primitive is a part of a segment: (slot << 2) | first quad
        idx = code >> 2;
        const bool may = !cull || CurveMayHit(O, D, cull[idx * 2], cull[idx * 2 + 1]);
#ifdef PBR_CURVE_PROBE   // tests/host_emul only
        PBR_CURVE_PROBE(may);
#endif
        if (may) {
          const float4 c0 = prims[idx * 4 + 0], c1 = prims[idx * 4 + 1], c2 = prims[idx * 4 + 2],
                       c3 = prims[idx * 4 + 3];
          h = IntersectCurve(O, rs, ray.tmin, tfar, c0, c1, c2, c3, code & 3u, part_quads, &t, &u, &v);
        }
      } else {
        const float4 a = prims[idx * 3 + 0], b = prims[idx * 3 + 1], c = prims[idx * 3 + 2];
        h = IntersectTriangle(O, D, ray.tmin, tfar, from4(a), from4(b), from4(c), &t, &u, &v);
      }
      if (h) {
        if (ANY) return true;
        found = true;
        tfar = t;
        hit->t = t; hit->u = u; hit->v = v;
        hit->prim = CURVES ? (idx | kCurveFlag) : idx;
      }
    }
    if ((group.y & 0xff000000u) == 0u) {
      if (sp == 0) break;
      group = stack[--sp];
    }
  }
  *tfar_io = tfar;
  return found;
}

// Scene::TraceFirstHit1: triangles first, then curves against the shortened ray.
template <bool STATS>
PBR_HD bool TraceClosest(const SceneView& s, const RayT& ray, HitT* hit, TraverseStats* st) {
  float tfar = ray.tmax;
  hit->prim = kInvalid;
  bool found = false;
  if (s.num_tris) found |= TraverseBvh<false, false, STATS>(s.tri_nodes, s.tri_data, nullptr, nullptr, 0u, ray, &tfar, hit, st);
  if (s.num_curves) found |= TraverseBvh<true, false, STATS>(s.curve_nodes, s.curve_data, s.curve_cull, s.curve_sub, s.curve_part_quads, ray, &tfar, hit, st);
  return found;
}

// Scene::AnyHit1
template <bool STATS>
PBR_HD bool TraceAny(const SceneView& s, const RayT& ray, TraverseStats* st) {
  float tfar = ray.tmax;
  HitT hit;
  if (s.num_tris && TraverseBvh<false, true, STATS>(s.tri_nodes, s.tri_data, nullptr, nullptr, 0u, ray, &tfar, &hit, st)) return true;
  if (s.num_curves && TraverseBvh<true, true, STATS>(s.curve_nodes, s.curve_data, s.curve_cull, s.curve_sub, s.curve_part_quads, ray, &tfar, &hit, st)) return true;
  return false;
}

}  // namespace pbr
