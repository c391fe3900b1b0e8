// TEST INFRASTRUCTURE ONLY — host emulation of the device code.
// The per-ray / per-vertex routines of the CUDA path tracer (pbrlab_b200/csrc/device/*.cuh, kat.cuh) are plain
// functions; this shim compiles them with g++ so their arithmetic can be checked against the compiled reference in
// a container without a GPU.  It mirrors the pbrgpu_* entry points one to one (same argument meaning) but is never
// built into, loaded by, or a fallback for the product: libpbrgpu.so has no CPU path.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

// walk segments seen by PathRadiance: [0] all, [1] the clearance field says "cannot hit", [2] of those, segments that hit
static std::atomic<uint64_t> g_clear_probe[3];
#define PBR_CLEARANCE_PROBE(clear, hit) \
  do { const bool c_ = (clear); ++g_clear_probe[0]; if (c_) { ++g_clear_probe[1]; if (hit) ++g_clear_probe[2]; } } while (0)

// walk segments of PathRadiance that the clearance field does NOT answer, recorded (up to a cap) while enabled:
// 8 floats each (origin, tmin, direction, tmax) — scripts/walk_segment_stats.py
#include <mutex>
static std::mutex g_seg_mutex;
static std::vector<float> g_segments;
static std::atomic<uint64_t> g_seg_cap(0);
#define PBR_WALK_SEGMENT_PROBE(ray, clear)                                                              \
  do {                                                                                                  \
    if (!(clear) && g_seg_cap.load(std::memory_order_relaxed)) {                                        \
      std::lock_guard<std::mutex> lock_(g_seg_mutex);                                                   \
      if (g_segments.size() / 8 < g_seg_cap.load()) {                                                   \
        const float r_[8] = {(ray).o.x, (ray).o.y, (ray).o.z, (ray).tmin, (ray).d.x, (ray).d.y, (ray).d.z, (ray).tmax}; \
        g_segments.insert(g_segments.end(), r_, r_ + 8);                                                \
      }                                                                                                 \
    }                                                                                                   \
  } while (0)

// curve leaf tests: [0] all, [1] those that pass CurveMayHit and run the full ribbon test
static std::atomic<uint64_t> g_curve_probe[2];
#define PBR_CURVE_PROBE(may) do { ++g_curve_probe[0]; if (may) ++g_curve_probe[1]; } while (0)

#include "../../pbrlab_b200/csrc/kat.cuh"
#include "../../pbrlab_b200/csrc/scene_host.h"
#include "../../pbrlab_b200/csrc/job_split.h"

using namespace pbr;  // NOLINT

namespace {
struct Emul {
  pbrhost::HostScene scene;
  SceneView view;
};

template <typename F>
void ParallelFor(uint64_t n, F f) {
  const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<uint64_t> next(0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (;;) {
        const uint64_t b = next.fetch_add(1024);
        if (b >= n) break;
        const uint64_t e = std::min(n, b + 1024);
        for (uint64_t i = b; i < e; ++i) f(i);
      }
    });
  for (auto& t : th) t.join();
}

RayT ToRay(const pbrgpu_ray& r) {
  RayT ray;
  ray.o = vec3(r.org[0], r.org[1], r.org[2]); ray.tmin = r.tmin;
  ray.d = vec3(r.dir[0], r.dir[1], r.dir[2]); ray.tmax = r.tmax;
  return ray;
}
}  // namespace

extern "C" {

void* emul_create() { return new Emul(); }
void emul_destroy(void* h) { delete static_cast<Emul*>(h); }
const char* emul_last_error(void* h) { return static_cast<Emul*>(h)->scene.error.c_str(); }

int emul_set_triangles(void* h, const float* xyzw, uint32_t nverts, const uint32_t* vidx, const float* nxyzw,
                       uint32_t nnormals, const uint32_t* nidx, const float* uv, uint32_t nuv, const uint32_t* tidx,
                       const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                       const uint32_t* prim_id, uint64_t ntris) {
  return static_cast<Emul*>(h)->scene.SetTriangles(xyzw, nverts, vidx, nxyzw, nnormals, nidx, uv, nuv, tidx,
                                                   material_id, instance_id, geom_id, prim_id, ntris) ? 0 : 1;
}
int emul_set_curves(void* h, const float* xyzr, uint32_t nverts, const uint32_t* first_cp, const uint32_t* material_id,
                    const uint32_t* instance_id, const uint32_t* geom_id, const uint32_t* prim_id, uint64_t nsegs) {
  return static_cast<Emul*>(h)->scene.SetCurves(xyzr, nverts, first_cp, material_id, instance_id, geom_id, prim_id,
                                                nsegs) ? 0 : 1;
}
int emul_set_materials(void* h, const pbrgpu_material* m, uint32_t n) {
  return static_cast<Emul*>(h)->scene.SetMaterials(m, n) ? 0 : 1;
}
int emul_set_textures(void* h, const pbrgpu_texture* t, uint32_t n) {
  return static_cast<Emul*>(h)->scene.SetTextures(t, n) ? 0 : 1;
}
// Texture::FetchFloat3 through the device function
int emul_texture_fetch3(void* h, uint32_t tex, const float* uv, uint64_t n, float* out) {
  Emul* e = static_cast<Emul*>(h);
  if (tex >= e->view.num_textures) return 1;
  for (uint64_t k = 0; k < n; ++k) {
    const pbr::vec3 c = pbr::TextureFetch3(e->view, tex, uv[2 * k], uv[2 * k + 1]);
    out[3 * k] = c.x; out[3 * k + 1] = c.y; out[3 * k + 2] = c.z;
  }
  return 0;
}
int emul_material_classes(void* h, uint32_t* out) {
  Emul* e = static_cast<Emul*>(h);
  for (size_t i = 0; i < e->scene.material_class.size(); ++i) out[i] = e->scene.material_class[i];
  return int(e->scene.material_class.size());
}
int emul_set_lights(void* h, const pbrgpu_light_tables* t) { return static_cast<Emul*>(h)->scene.SetLights(t) ? 0 : 1; }
int emul_commit(void* h, const float* bmin, const float* bmax) {
  Emul* e = static_cast<Emul*>(h);
  if (!e->scene.Commit(bmin, bmax)) return 3;
  e->view = e->scene.HostView();
  return 0;
}
int emul_scene_bounds(void* h, float* bmin, float* bmax) {
  Emul* e = static_cast<Emul*>(h);
  for (int k = 0; k < 3; ++k) { bmin[k] = e->scene.bmin[k]; bmax[k] = e->scene.bmax[k]; }
  return 0;
}
// nodes, max depth, sah cost of the two BVHs: out[0..2] triangles, out[3..5] curves
void emul_bvh_info(void* h, double* out6) {
  Emul* e = static_cast<Emul*>(h);
  out6[0] = e->scene.tri_bvh.num_nodes; out6[1] = e->scene.tri_bvh.max_depth; out6[2] = e->scene.tri_bvh.sah_cost;
  out6[3] = e->scene.curve_bvh.num_nodes; out6[4] = e->scene.curve_bvh.max_depth; out6[5] = e->scene.curve_bvh.sah_cost;
}

int emul_trace(void* h, const pbrgpu_ray* rays, uint64_t n, pbrgpu_hit* hits, uint64_t* stats2) {
  Emul* e = static_cast<Emul*>(h);
  const SceneView& s = e->view;
  std::atomic<uint64_t> nodes(0), prims(0);
  ParallelFor(n, [&](uint64_t i) {
    HitT hit;
    TraverseStats st = {0, 0};
    pbrgpu_hit out;
    out.normal_g[0] = 1.f; out.normal_g[1] = 0.f; out.normal_g[2] = 0.f;
    out.t = 1.f; out.u = 0.f; out.v = 0.f;
    out.instance_id = out.geom_id = out.prim_id = kInvalid;
    if (TraceClosest<true>(s, ToRay(rays[i]), &hit, &st)) {
      const vec3 ng = HitGeometricNormal(s, hit);
      out.normal_g[0] = ng.x; out.normal_g[1] = ng.y; out.normal_g[2] = ng.z;
      out.t = hit.t; out.u = hit.u; out.v = hit.v;
      uint4 ids;
      if (hit.prim & kCurveFlag) ids = s.curve_ids[s.curve_prim[hit.prim & ~kCurveFlag]];
      else ids = s.tri_ids[f2u(s.tri_data[hit.prim * 3].w)];
      out.instance_id = ids.x; out.geom_id = ids.y; out.prim_id = ids.z;
    }
    hits[i] = out;
    nodes += st.nodes;
    prims += st.prims;
  });
  if (stats2) { stats2[0] = nodes; stats2[1] = prims; }
  return 0;
}

// MEASUREMENT ONLY: closest hit through the curve BVH with the children visited best-first — always the
// box with the smallest entry distance next, over the whole tree (a priority queue) — instead of the static octant
// order of TraverseBvh.  Same boxes, same leaf tests; stats2 = (inner nodes tested, curve candidates tested): the
// lower bound any ordering of the visits could reach on this tree.
int emul_trace_curves_best_first(void* h, const pbrgpu_ray* rays, uint64_t n, float* t_out, uint64_t* stats2, int mode) {
  Emul* e = static_cast<Emul*>(h);
  const SceneView& s = e->view;
  if (!s.num_curves) return 1;
  std::atomic<uint64_t> nodes(0), prims(0);
  const uint32_t* nw = reinterpret_cast<const uint32_t*>(s.curve_nodes);
  ParallelFor(n, [&](uint64_t i) {
    const RayT ray = ToRay(rays[i]);
    const CurveRaySpace rs = MakeCurveRaySpace(ray.d);
    float tfar = ray.tmax;
    uint64_t nn = 0, np = 0;
    struct Item { float t; uint32_t id; uint32_t leaf_count; uint32_t key; };   // leaf_count 0: inner node id; else first prim + count; key: static visit rank
    std::vector<Item> heap;
    auto push = [&](Item it) { heap.push_back(it); std::push_heap(heap.begin(), heap.end(), [](const Item& a, const Item& b) { return a.t > b.t; }); };
    auto pop = [&]() { std::pop_heap(heap.begin(), heap.end(), [](const Item& a, const Item& b) { return a.t > b.t; }); Item it = heap.back(); heap.pop_back(); return it; };
    // mode 1: depth-first, the hit children of a node in order of entry distance (a LIFO stack: pushed far to near)
    std::vector<Item> batch;
    if (mode >= 1) heap.push_back({ray.tmin, 0u, 0u, 0u}); else push({ray.tmin, 0u, 0u, 0u});
    const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
    while (!heap.empty()) {
      Item it;
      if (mode >= 1) { it = heap.back(); heap.pop_back(); if (it.t > tfar && !(mode >= 3 && it.leaf_count == 0)) continue; }
      else { it = pop(); if (it.t > tfar) break; }
      if (it.leaf_count) {
        for (uint32_t j = 0; j < it.leaf_count; ++j) {
          ++np;
          const uint32_t code = s.curve_sub[it.id + j], idx = code >> 2;
          if (s.curve_cull && !CurveMayHit(ray.o, ray.d, s.curve_cull[idx * 2], s.curve_cull[idx * 2 + 1])) continue;
          float t, u, v;
          if (IntersectCurve(ray.o, rs, ray.tmin, tfar, s.curve_data[idx * 4], s.curve_data[idx * 4 + 1], s.curve_data[idx * 4 + 2],
                             s.curve_data[idx * 4 + 3], code & 3u, s.curve_part_quads, &t, &u, &v))
            tfar = t;
        }
        continue;
      }
      ++nn;
      const uint32_t* w = nw + size_t(20) * it.id;
      float p[3], step[3];
      memcpy(p, w, 12);
      for (int k = 0; k < 3; ++k) { const uint32_t eb = (w[3] >> (8 * k)) & 0xffu; const uint32_t bits = eb << 23; memcpy(&step[k], &bits, 4); }
      const uint32_t imask = w[3] >> 24, child_base = w[4], prim_base = w[5];
      uint32_t inner_rank = 0;
      batch.clear();
      for (int sl = 0; sl < 8; ++sl) {
        const uint32_t meta = (w[6 + sl / 4] >> (8 * (sl % 4))) & 0xffu;
        if (!meta) continue;
        const bool inner = (imask >> sl) & 1u;
        const uint32_t my_rank = inner_rank;
        if (inner) ++inner_rank;
        float tn = ray.tmin, tf = tfar;
        for (int k = 0; k < 3; ++k) {
          const uint32_t qlo = (w[8 + 2 * k + sl / 4] >> (8 * (sl % 4))) & 0xffu, qhi = (w[14 + 2 * k + sl / 4] >> (8 * (sl % 4))) & 0xffu;
          const float lo = p[k] + float(qlo) * step[k], hi = p[k] + float(qhi) * step[k];
          const float inv = 1.0f / (std::fabs(d[k]) < 1e-30f ? std::copysign(1e-30f, d[k]) : d[k]);
          float t0 = (lo - o[k]) * inv, t1 = (hi - o[k]) * inv;
          if (t0 > t1) std::swap(t0, t1);
          tn = std::max(tn, t0 * 0.999999f - 1e-6f); tf = std::min(tf, t1 * 1.000001f + 1e-6f);
        }
        if (tn > tf) continue;
        Item c;
        const uint32_t oct_inv = (d[0] < 0.f ? 0u : 4u) | (d[1] < 0.f ? 0u : 2u) | (d[2] < 0.f ? 0u : 1u);
        const uint32_t key = uint32_t(sl) ^ oct_inv;   // the engine visits the hit children by descending key
        if (inner) c = {tn, child_base + my_rank, 0u, key};
        else {
          const uint32_t unary = meta >> 5, cnt = unary == 1 ? 1u : (unary == 3 ? 2u : 3u);
          c = {tn, prim_base + (meta & 0x1fu), cnt, key};
        }
        if (mode >= 1) batch.push_back(c); else push(c);
      }
      if (mode >= 1) {
        if (mode >= 4) {
          // modes 4 / 5: the static octant order (4), and the same with the children whose box holds the ray origin
          // first (5); popped last-pushed first, so sort ascending by visit priority
          const float t0 = ray.tmin;
          std::sort(batch.begin(), batch.end(), [&](const Item& a, const Item& b) {
            const uint32_t pa = ((mode == 5 && a.t <= t0) ? 8u : 0u) + a.key, pb = ((mode == 5 && b.t <= t0) ? 8u : 0u) + b.key;
            return pa < pb;
          });
        } else
        std::sort(batch.begin(), batch.end(), [](const Item& a, const Item& b) { return a.t > b.t; });   // far first
        if (mode >= 2)   // the node's leaf candidates first (as the traversal engine does), then its inner children near to far
          std::stable_partition(batch.begin(), batch.end(), [](const Item& a) { return a.leaf_count == 0; });
        for (const Item& c : batch) heap.push_back(c);
      }
    }
    if (t_out) t_out[i] = tfar;
    nodes += nn; prims += np;
  });
  stats2[0] = nodes; stats2[1] = prims;
  return 0;
}

int emul_occluded(void* h, const pbrgpu_ray* rays, uint64_t n, uint8_t* occ) {
  Emul* e = static_cast<Emul*>(h);
  ParallelFor(n, [&](uint64_t i) { occ[i] = TraceAny<false>(e->view, ToRay(rays[i]), nullptr) ? 1 : 0; });
  return 0;
}

int emul_radiance(void* h, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* out, uint64_t* counts3) {
  Emul* e = static_cast<Emul*>(h);
  std::atomic<uint64_t> c0(0), c1(0), c2(0);
  ParallelFor(n, [&](uint64_t i) {
    Pcg32 rng;
    pcg32_srandom(&rng, seeds[2 * i], seeds[2 * i + 1]);
    uint64_t rc[3] = {0, 0, 0};
    const vec3 L = PathRadiance(e->view, ToRay(rays[i]), &rng, rc);
    out[3 * i] = L.x; out[3 * i + 1] = L.y; out[3 * i + 2] = L.z;
    c0 += rc[0]; c1 += rc[1]; c2 += rc[2];
  });
  if (counts3) { counts3[0] = c0; counts3[1] = c1; counts3[2] = c2; }
  return 0;
}

void emul_curve_probe(uint64_t* out2) {
  for (int k = 0; k < 2; ++k) out2[k] = g_curve_probe[k].exchange(0);
}

// start recording traced walk segments (cap > 0) / fetch what was recorded (returns the count, copies up to max_n)
void emul_record_segments(uint64_t cap) {
  std::lock_guard<std::mutex> lock(g_seg_mutex);
  g_segments.clear();
  g_seg_cap = cap;
}
uint64_t emul_fetch_segments(float* out8, uint64_t max_n) {
  std::lock_guard<std::mutex> lock(g_seg_mutex);
  const uint64_t n = std::min<uint64_t>(g_segments.size() / 8, max_n);
  if (out8) memcpy(out8, g_segments.data(), sizeof(float) * 8 * n);
  return g_segments.size() / 8;
}

// MEASUREMENT ONLY: for closest-hit queries through the TRIANGLE BVH, how many of the nodes a ray tests lie
// on the chain from the root on which exactly one inner child (and no leaf) is hit — the nodes a traversal that
// started at the end of that chain would not have to test.  out4 = (rays, nodes tested, nodes on the initial chain,
// rays whose chain reaches a node with leaf hits only).  Depth-first in the static slot order, same boxes and
// triangle test as TraverseBvh.
int emul_tri_chain_stats(void* h, const pbrgpu_ray* rays, uint64_t n, uint64_t* out4) {
  Emul* e = static_cast<Emul*>(h);
  const SceneView& s = e->view;
  if (!s.num_tris) return 1;
  const uint32_t* nw = reinterpret_cast<const uint32_t*>(s.tri_nodes);
  std::atomic<uint64_t> tot_nodes(0), tot_chain(0), tot_leafend(0);
  ParallelFor(n, [&](uint64_t i) {
    const RayT ray = ToRay(rays[i]);
    float tfar = ray.tmax;
    const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
    std::vector<uint32_t> stack(1, 0u);
    uint64_t nn = 0, chain = 0;
    bool on_chain = true, leaf_end = false;
    while (!stack.empty()) {
      const uint32_t id = stack.back(); stack.pop_back();
      ++nn;
      const uint32_t* w = nw + size_t(20) * id;
      float p[3], step[3];
      memcpy(p, w, 12);
      for (int k = 0; k < 3; ++k) { const uint32_t bits = ((w[3] >> (8 * k)) & 0xffu) << 23; memcpy(&step[k], &bits, 4); }
      const uint32_t imask = w[3] >> 24, child_base = w[4], prim_base = w[5];
      uint32_t inner_rank = 0, inner_hits = 0, leaf_hits = 0;
      uint32_t hit_inner[8];
      for (int sl = 0; sl < 8; ++sl) {
        const uint32_t meta = (w[6 + sl / 4] >> (8 * (sl % 4))) & 0xffu;
        if (!meta) continue;
        const bool inner = (imask >> sl) & 1u;
        const uint32_t my_rank = inner_rank;
        if (inner) ++inner_rank;
        float tn = ray.tmin, tf = tfar;
        for (int k = 0; k < 3; ++k) {
          const uint32_t qlo = (w[8 + 2 * k + sl / 4] >> (8 * (sl % 4))) & 0xffu, qhi = (w[14 + 2 * k + sl / 4] >> (8 * (sl % 4))) & 0xffu;
          const float lo = p[k] + float(qlo) * step[k], hi = p[k] + float(qhi) * step[k];
          const float inv = 1.0f / (std::fabs(d[k]) < 1e-30f ? std::copysign(1e-30f, d[k]) : d[k]);
          float t0 = (lo - o[k]) * inv, t1 = (hi - o[k]) * inv;
          if (t0 > t1) std::swap(t0, t1);
          tn = std::max(tn, t0 * 0.999999f - 1e-6f); tf = std::min(tf, t1 * 1.000001f + 1e-6f);
        }
        if (tn > tf) continue;
        if (inner) hit_inner[inner_hits++] = child_base + my_rank;
        else {
          ++leaf_hits;
          const uint32_t unary = meta >> 5, cnt = unary == 1 ? 1u : (unary == 3 ? 2u : 3u);
          for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t idx = prim_base + (meta & 0x1fu) + j;
            float t, u, v;
            if (IntersectTriangle(ray.o, ray.d, ray.tmin, tfar, from4(s.tri_data[idx * 3]), from4(s.tri_data[idx * 3 + 1]),
                                  from4(s.tri_data[idx * 3 + 2]), &t, &u, &v))
              tfar = t;
          }
        }
      }
      if (on_chain) {
        if (inner_hits == 1 && leaf_hits == 0) ++chain;                 // this node could have been skipped
        else { on_chain = false; leaf_end = (inner_hits == 0); }
      }
      for (uint32_t k = inner_hits; k > 0; --k) stack.push_back(hit_inner[k - 1]);
    }
    tot_nodes += nn; tot_chain += chain; tot_leafend += leaf_end ? 1 : 0;
  });
  out4[0] = n; out4[1] = tot_nodes; out4[2] = tot_chain; out4[3] = tot_leafend;
  return 0;
}

// the clearance field of the committed scene: geo7 = (org.xyz, inv_cell, quantum, -, -), dims3, and (bytes != null) the cells
uint64_t emul_clearance_field(void* h, float* geo7, uint32_t* dims3, uint8_t* bytes) {
  const pbrhost::HostScene& sc = static_cast<Emul*>(h)->scene;
  for (int k = 0; k < 3; ++k) { geo7[k] = sc.clear_org[k]; dims3[k] = sc.clear_dims[k]; }
  geo7[3] = sc.clear_inv_cell; geo7[4] = sc.clear_quantum;
  const uint64_t n = uint64_t(sc.clear_dims[0]) * sc.clear_dims[1] * sc.clear_dims[2];
  if (bytes && !sc.clear_dist.empty()) memcpy(bytes, sc.clear_dist.data(), n);
  return sc.clear_dist.empty() ? 0 : n;
}

// counters of PBR_CLEARANCE_PROBE since the last call (and reset)
void emul_clearance_probe(uint64_t* out3) {
  for (int k = 0; k < 3; ++k) out3[k] = g_clear_probe[k].exchange(0);
}

// one shading vertex, layout of pbrgpu_shade
int emul_shade(void* h, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* out16) {
  Emul* e = static_cast<Emul*>(h);
  const SceneView& s = e->view;
  ParallelFor(n, [&](uint64_t i) {
    float* o = out16 + 16 * i;
    for (int k = 0; k < 16; ++k) o[k] = 0.f;
    const RayT ray = ToRay(rays[i]);
    HitT hit;
    if (!TraceClosest<false>(s, ray, &hit, nullptr)) return;
    const Surface si = MakeSurface(s, ray, hit);
    Pcg32 rng;
    pcg32_srandom(&rng, seeds[2 * i], seeds[2 * i + 1]);
    VertexResult vr;
    const vec3 wo = -ray.d;
    const int kind = MaterialKind(s, si);
    if (kind == 1) {
      if (PrincipledVertex(s, si, wo, &rng, &vr)) SubsurfaceVertex(s, si, &rng, &vr, nullptr);
    } else if (kind == 2) {
      HairVertex(s, si, wo, &rng, &vr);
    } else {
      AbsorbVertex(wo, si.P, &vr);
    }
    vec3 direct(0.f);
    for (int k = 0; k < 2; ++k)
      if (vr.shadow[k].active && !TraceAny<false>(s, vr.shadow[k].ray, nullptr)) direct = direct + vr.shadow[k].contribute;
    o[0] = 1.f;
    o[1] = vr.wi.x; o[2] = vr.wi.y; o[3] = vr.wi.z;
    o[4] = vr.throughput.x; o[5] = vr.throughput.y; o[6] = vr.throughput.z;
    o[7] = direct.x; o[8] = direct.y; o[9] = direct.z;
    o[10] = vr.pdf;
    o[11] = vr.P.x; o[12] = vr.P.y; o[13] = vr.P.z;
    o[14] = float(si.face);
    o[15] = hit.t;
  });
  return 0;
}

// hit -> SurfaceInfo: out n x 12: P(3) Ns(3) Ng(3) uv(2) face (-1 on miss)
int emul_surface(void* h, const pbrgpu_ray* rays, uint64_t n, float* out12) {
  Emul* e = static_cast<Emul*>(h);
  const SceneView& s = e->view;
  ParallelFor(n, [&](uint64_t i) {
    float* o = out12 + 12 * i;
    for (int k = 0; k < 12; ++k) o[k] = 0.f;
    const RayT ray = ToRay(rays[i]);
    HitT hit;
    if (!TraceClosest<false>(s, ray, &hit, nullptr)) { o[11] = -1.f; return; }
    const Surface si = MakeSurface(s, ray, hit);
    o[0] = si.P.x; o[1] = si.P.y; o[2] = si.P.z;
    o[3] = si.Ns.x; o[4] = si.Ns.y; o[5] = si.Ns.z;
    o[6] = si.Ng.x; o[7] = si.Ng.y; o[8] = si.Ng.z;
    o[9] = si.tex_u; o[10] = si.tex_v;
    o[11] = float(si.face);
  });
  return 0;
}

// light sampling: out n x 10 (pos3, normal3, emission3, pdf)
int emul_sample_light(void* h, const uint64_t* seeds, uint64_t n, float* out10) {
  Emul* e = static_cast<Emul*>(h);
  for (uint64_t i = 0; i < n; ++i) {
    Pcg32 rng;
    pcg32_srandom(&rng, seeds[2 * i], seeds[2 * i + 1]);
    const LightSample ls = SampleAllLight(e->view, &rng);
    float* o = out10 + 10 * i;
    o[0] = ls.pos.x; o[1] = ls.pos.y; o[2] = ls.pos.z;
    o[3] = ls.normal.x; o[4] = ls.normal.y; o[5] = ls.normal.z;
    o[6] = ls.emission.x; o[7] = ls.emission.y; o[8] = ls.emission.z;
    o[9] = ls.pdf;
  }
  return 0;
}

int emul_eval_closure(int op, const float* params, const float* in, uint32_t in_stride, uint64_t n, float* out,
                      uint32_t out_stride) {
  for (uint64_t i = 0; i < n; ++i) KatEval(op, params, in + size_t(i) * in_stride, out + size_t(i) * out_stride, out_stride);
  return 0;
}

// the library's own job-split arithmetic (csrc/job_split.h, used by pbrgpu.cu: RenderImpl): out = {offset, stride, count}
// pbrjob::SampleOfId over every id of a worker's frame: out_pixel / out_sample [npix * (probe + rest)]
void emul_sample_order(const uint32_t* order, uint32_t npix, uint32_t probe, uint32_t rest, uint32_t block,
                       uint32_t* out_pixel, uint32_t* out_sample) {
  const pbrjob::SampleOrder so = {order, npix, probe, rest, block};
  const uint64_t n = uint64_t(npix) * (probe + rest);
  for (uint64_t id = 0; id < n; ++id) pbrjob::SampleOfId(so, id, &out_pixel[id], &out_sample[id]);
}

int emul_job_share(uint32_t job_offset, uint32_t job_stride, uint32_t rank, uint32_t world, uint32_t device,
                   uint32_t num_devices, uint32_t spp, uint32_t* out3) {
  const pbrjob::Share s = pbrjob::ShareOf(job_offset, job_stride, rank, world, device, num_devices);
  out3[0] = s.offset; out3[1] = s.stride; out3[2] = pbrjob::CountOf(s, spp);
  return s.ok ? 1 : 0;
}

// Render() on the host cores with the GPU's path numbering: path (pixel p, sample s) = pcg32_srandom(seed + s, p)
int emul_render(void* h, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed, uint32_t sample_offset,
                uint32_t sample_stride, float* rgba, uint32_t* count, uint64_t* counts3) {
  Emul* e = static_cast<Emul*>(h);
  float cam[8];
  pbrhost::MakeCamera(e->scene.bmin, e->scene.bmax, width, height, cam);
  const uint64_t npix = uint64_t(width) * height;
  memset(rgba, 0, sizeof(float) * 4 * npix);
  memset(count, 0, sizeof(uint32_t) * npix);
  std::atomic<uint64_t> c0(0), c1(0), c2(0);
  ParallelFor(npix, [&](uint64_t pixel) {
    const uint32_t x = uint32_t(pixel % width), y = uint32_t(pixel / width);
    uint64_t rc[3] = {0, 0, 0};
    for (uint32_t sidx = sample_offset; sidx < spp; sidx += sample_stride) {
      Pcg32 rng;
      pcg32_srandom(&rng, seed + sidx, pixel);
      const float jx = Draw(&rng), jy = Draw(&rng);
      const float tx = cam[3] + cam[6] * (float(x) + jx);
      const float ty = cam[4] - cam[7] * (float(y) + jy);
      float dx = tx - cam[0], dy = ty - cam[1], dz = cam[5] - cam[2];
      const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
      RayT ray;
      ray.o = vec3(cam[0], cam[1], cam[2]);
      ray.d = vec3(dx * inv, dy * inv, dz * inv);
      ray.tmin = 0.f;
      ray.tmax = kInf;
      const vec3 L = PathRadiance(e->view, ray, &rng, rc);
      rgba[4 * pixel] += L.x; rgba[4 * pixel + 1] += L.y; rgba[4 * pixel + 2] += L.z; rgba[4 * pixel + 3] += 1.0f;
      count[pixel]++;
    }
    c0 += rc[0]; c1 += rc[1]; c2 += rc[2];
  });
  if (counts3) { counts3[0] = c0; counts3[1] = c1; counts3[2] = c2; }
  return 0;
}

}  // extern "C"
