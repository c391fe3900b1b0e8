// Wavefront path tracer kernels (sm_100a).  One launch of Render() is organised as waves of N paths whose state
// lives in HBM as structure-of-arrays (every field a float4 / 16-byte record so each lane issues 128-bit loads and a
// warp touches contiguous 512-byte spans when the queue is dense):
//
//   ray_o[N]  (org.xyz, tmin)      ray_d[N]  (dir.xyz, tmax)      hit[N]  (t, u, v, leaf-order primitive | curve flag)
//   thr[N]    (throughput.rgb, pdf of the last BSDF sample)       rad[N]  (radiance.rgb, depth)
//   rng[N]    (PCG32 state, inc)
//
// and index queues (u32 path ids) compacted with warp ballots + one atomicAdd per warp:
//
//   q_active[2]  paths that still need a closest-hit query (ping-pong)
//   q_surface    hit a triangle-type material (or none): emission + roulette + Principled vertex
//   q_hair       hit a hair material
//   q_sss        Principled vertex selected the random-walk closure
//   shadow queue (ray + contribution + path id): NEE any-hit queries
//
// One bounce iteration = trace_closest -> shade_surface, shade_hair -> sss_walk -> trace_any.  All kernels are
// persistent: the grid is a fixed multiple of the SM count and warps pull 32-entry batches from the queue with an
// atomic counter, so queue lengths never leave the device inside an iteration.
// Replaces the per-pixel loop of the reference (src/render.cc:24-90,125-190) — see device/shade.cuh for the per-vertex
// functions and their citations.
#pragma once
#include <cuda_runtime.h>

#include "device/shade.cuh"

namespace pbr {

// device-side counters, one cache line apart is not needed: touched once per warp
enum Counter {
  kNumActiveNext = 0, kNumSurface, kNumHair, kNumSss, kNumShadow,
  kFetchTrace, kFetchSurface, kFetchHair, kFetchSss, kFetchShadow,
  kCounterCount
};
enum Stat { kStatClosest = 0, kStatShadow, kStatSss, kStatNodes, kStatPrims, kStatCount };

struct WaveState {
  float4* ray_o;
  float4* ray_d;
  float4* hit;
  float4* thr;
  float4* rad;
  ulonglong2* rng;
  uint32_t* q_active[2];
  uint32_t* q_surface;
  uint32_t* q_hair;
  uint32_t* q_sss;
  float4* sh_o;
  float4* sh_d;
  float4* sh_c;
  uint32_t* counters;            // kCounterCount
  unsigned long long* stats;     // kStatCount
  uint32_t capacity;
};

// ---- warp-aggregated queue append: one atomicAdd per warp
__device__ __forceinline__ uint32_t WarpAppend(uint32_t* counter, bool pred) {
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0u) return 0u;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, uint32_t(__popc(mask)));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + uint32_t(__popc(mask & ((1u << lane) - 1u)));
}

// a warp pulls the next 32 queue slots; returns this lane's slot (>= n when the queue is exhausted)
__device__ __forceinline__ uint32_t WarpFetch(uint32_t* fetch_counter) {
  const int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == 0) base = atomicAdd(fetch_counter, 32u);
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + uint32_t(lane);
}

__device__ __forceinline__ RayT LoadRay(const WaveState& w, uint32_t p) {
  const float4 o = w.ray_o[p], d = w.ray_d[p];
  RayT r;
  r.o = vec3(o.x, o.y, o.z); r.tmin = o.w;
  r.d = vec3(d.x, d.y, d.z); r.tmax = d.w;
  return r;
}
__device__ __forceinline__ void StoreRay(const WaveState& w, uint32_t p, const vec3& o, const vec3& d, float tmin,
                                         float tmax) {
  w.ray_o[p] = make_float4(o.x, o.y, o.z, tmin);
  w.ray_d[p] = make_float4(d.x, d.y, d.z, tmax);
}

__device__ __forceinline__ void PushShadow(const WaveState& w, const ShadowRequest& req, const vec3& throughput,
                                           uint32_t path) {
  const uint32_t slot = WarpAppend(&w.counters[kNumShadow], req.active);
  if (req.active) {
    const vec3 c = throughput * req.contribute;
    w.sh_o[slot] = make_float4(req.ray.o.x, req.ray.o.y, req.ray.o.z, req.ray.tmin);
    w.sh_d[slot] = make_float4(req.ray.d.x, req.ray.d.y, req.ray.d.z, req.ray.tmax);
    w.sh_c[slot] = make_float4(c.x, c.y, c.z, __uint_as_float(path));
  }
}

// ------------------------------------------------------------------------------------------------ camera
// RenderingTile's ray generation (src/render.cc:160-171): target = corner + d*(pixel + xi); the two jitter draws
// are the first two numbers of the path's stream.  cam: eye.xyz, x_corner, y_corner, z_corner, dx, dy.
struct CameraParams {
  float eye[3], x_corner, y_corner, z_corner, dx, dy;
  uint32_t width, height;
};

__global__ void GenCameraRaysKernel(WaveState w, CameraParams cam, uint32_t npix, uint32_t n_paths,
                                    uint64_t seed, uint32_t first_sample, uint32_t sample_stride) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_paths) return;
  const uint32_t s_local = p / npix, pixel = p - s_local * npix;
  const uint32_t x = pixel % cam.width, y = pixel / cam.width;
  Pcg32 rng;
  pcg32_srandom(&rng, seed + uint64_t(first_sample + s_local * sample_stride), uint64_t(pixel));
  const float jx = Draw(&rng), jy = Draw(&rng);
  const float tx = cam.x_corner + cam.dx * (float(x) + jx);
  const float ty = cam.y_corner - cam.dy * (float(y) + jy);
  float dx = tx - cam.eye[0], dy = ty - cam.eye[1], dz = cam.z_corner - cam.eye[2];
  const float inv_norm = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);   // Normalize (render.cc:243-249)
  dx *= inv_norm; dy *= inv_norm; dz *= inv_norm;
  w.ray_o[p] = make_float4(cam.eye[0], cam.eye[1], cam.eye[2], 0.0f);
  w.ray_d[p] = make_float4(dx, dy, dz, kInf);
  w.thr[p] = make_float4(1.f, 1.f, 1.f, 0.f);
  w.rad[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
  w.rng[p] = make_ulonglong2(rng.state, rng.inc);
  w.q_active[0][p] = p;
}

// caller-supplied rays + seeds (pbrgpu_radiance / pbrgpu_shade hooks)
__global__ void InitPathsFromRaysKernel(WaveState w, const float4* rays, const uint64_t* seeds, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  w.ray_o[p] = rays[2 * p];
  w.ray_d[p] = rays[2 * p + 1];
  Pcg32 rng;
  pcg32_srandom(&rng, seeds[2 * p], seeds[2 * p + 1]);
  w.thr[p] = make_float4(1.f, 1.f, 1.f, 0.f);
  w.rad[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
  w.rng[p] = make_ulonglong2(rng.state, rng.inc);
  w.q_active[0][p] = p;
}

// ------------------------------------------------------------------------------------------------ iteration set-up
__global__ void BeginIterationKernel(uint32_t* counters) {
  if (threadIdx.x < kCounterCount) counters[threadIdx.x] = 0u;
}

// ------------------------------------------------------------------------------------------------ closest hit
// Scene::TraceFirstHit1 for every active path; routes the path by the material kind of what it hit.
__global__ void __launch_bounds__(128) TraceClosestKernel(SceneView s, WaveState w, const uint32_t* __restrict__ queue,
                                                          uint32_t n) {
  unsigned long long rays = 0;
  for (;;) {
    const uint32_t slot = WarpFetch(&w.counters[kFetchTrace]);
    if (__all_sync(0xffffffffu, slot >= n)) break;
    const bool valid = slot < n;
    uint32_t p = 0;
    HitT hit;
    hit.prim = kInvalid;
    int kind = -1;   // -1 miss, 0/1 surface queue, 2 hair queue
    if (valid) {
      p = queue[slot];
      const RayT ray = LoadRay(w, p);
      TraceClosest<false>(s, ray, &hit, nullptr);
      ++rays;
      w.hit[p] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
      if (hit.prim != kInvalid) {
        uint32_t mat;
        if (hit.prim & kCurveFlag) mat = s.curve_ids[s.curve_prim[hit.prim & ~kCurveFlag]].w;
        else mat = s.tri_ids[__float_as_uint(s.tri_data[hit.prim * 3].w)].w;
        kind = (mat < s.num_materials && s.materials[mat].type == 1u) ? 2 : 1;
      }
    }
    const uint32_t a = WarpAppend(&w.counters[kNumSurface], kind == 1);
    if (kind == 1) w.q_surface[a] = p;
    const uint32_t b = WarpAppend(&w.counters[kNumHair], kind == 2);
    if (kind == 2) w.q_hair[b] = p;
  }
  if (rays) atomicAdd(&w.stats[kStatClosest], rays);
}

// ------------------------------------------------------------------------------------------------ shading
struct ShadeFlags {
  uint32_t skip_emission_and_roulette;   // pbrgpu_shade hook: call Shader() only
};

__device__ __forceinline__ void CommitVertex(const WaveState& w, uint32_t p, const VertexResult& vr, vec3 throughput,
                                             const vec3& L, uint32_t depth, const Pcg32& rng, uint32_t next_parity) {
  const vec3 new_thr = vr.throughput * throughput;                     // render.cc:80
  w.thr[p] = make_float4(new_thr.x, new_thr.y, new_thr.z, vr.pdf);
  w.rad[p] = make_float4(L.x, L.y, L.z, __uint_as_float(depth + 1u));
  w.rng[p] = make_ulonglong2(rng.state, rng.inc);
  StoreRay(w, p, vr.P, vr.wi, 1e-3f, kInf);                            // render.cc:83-86
}

// emission + MIS, roulette, material dispatch, Principled vertex (everything but the random walk)
__global__ void __launch_bounds__(128) ShadeSurfaceKernel(SceneView s, WaveState w, uint32_t next_parity,
                                                          ShadeFlags flags) {
  const uint32_t n = w.counters[kNumSurface];
  for (;;) {
    const uint32_t slot = WarpFetch(&w.counters[kFetchSurface]);
    if (__all_sync(0xffffffffu, slot >= n)) break;
    const bool valid = slot < n;
    uint32_t p = 0;
    bool to_sss = false, to_next = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    if (valid) {
      p = w.q_surface[slot];
      const RayT ray = LoadRay(w, p);
      const float4 h4 = w.hit[p];
      HitT hit;
      hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
      const float4 t4 = w.thr[p];
      const float4 r4 = w.rad[p];
      throughput = vec3(t4.x, t4.y, t4.z);
      vec3 L(r4.x, r4.y, r4.z);
      const uint32_t depth = __float_as_uint(r4.w);
      const ulonglong2 rs = w.rng[p];
      Pcg32 rng;
      rng.state = rs.x; rng.inc = rs.y;
      const Surface si = MakeSurface(s, ray, hit);
      bool alive = true;
      if (!flags.skip_emission_and_roulette)
        alive = EmissionAndRoulette(s, ray, hit, si, depth, t4.w, &rng, &L, &throughput);
      if (!alive) {
        w.rad[p] = make_float4(L.x, L.y, L.z, __uint_as_float(depth));
      } else {
        const int kind = MaterialKind(s, si);
        VertexResult vr;
        const vec3 wo = -ray.d;
        if (kind == 1) {
          to_sss = PrincipledVertex(s, si, wo, &rng, &vr);
        } else {
          AbsorbVertex(wo, si.P, &vr);   // no material (shader.cc:11-17); hair on this queue cannot happen
          vr.shadow[0].active = false;
        }
        req = vr.shadow[0];
        if (to_sss) {
          // the walk runs in its own kernel: park the path with the post-roulette throughput and the rng positioned
          // right after the closure selector
          w.thr[p] = make_float4(throughput.x, throughput.y, throughput.z, t4.w);
          w.rad[p] = make_float4(L.x, L.y, L.z, __uint_as_float(depth));
          w.rng[p] = make_ulonglong2(rng.state, rng.inc);
        } else {
          CommitVertex(w, p, vr, throughput, L, depth, rng, next_parity);
          to_next = !IsBlack(vr.throughput * throughput);               // render.cc:31
        }
      }
    }
    PushShadow(w, req, throughput, p);
    const uint32_t a = WarpAppend(&w.counters[kNumSss], to_sss);
    if (to_sss) w.q_sss[a] = p;
    const uint32_t b = WarpAppend(&w.counters[kNumActiveNext], to_next);
    if (to_next) w.q_active[next_parity][b] = p;
  }
}

__global__ void __launch_bounds__(128) ShadeHairKernel(SceneView s, WaveState w, uint32_t next_parity,
                                                       ShadeFlags flags) {
  const uint32_t n = w.counters[kNumHair];
  for (;;) {
    const uint32_t slot = WarpFetch(&w.counters[kFetchHair]);
    if (__all_sync(0xffffffffu, slot >= n)) break;
    const bool valid = slot < n;
    uint32_t p = 0;
    bool to_next = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    if (valid) {
      p = w.q_hair[slot];
      const RayT ray = LoadRay(w, p);
      const float4 h4 = w.hit[p];
      HitT hit;
      hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
      const float4 t4 = w.thr[p];
      const float4 r4 = w.rad[p];
      throughput = vec3(t4.x, t4.y, t4.z);
      vec3 L(r4.x, r4.y, r4.z);
      const uint32_t depth = __float_as_uint(r4.w);
      const ulonglong2 rs = w.rng[p];
      Pcg32 rng;
      rng.state = rs.x; rng.inc = rs.y;
      const Surface si = MakeSurface(s, ray, hit);
      bool alive = true;
      if (!flags.skip_emission_and_roulette)
        alive = EmissionAndRoulette(s, ray, hit, si, depth, t4.w, &rng, &L, &throughput);
      if (!alive) {
        w.rad[p] = make_float4(L.x, L.y, L.z, __uint_as_float(depth));
      } else {
        VertexResult vr;
        HairVertex(s, si, -ray.d, &rng, &vr);
        req = vr.shadow[0];
        CommitVertex(w, p, vr, throughput, L, depth, rng, next_parity);
        to_next = !IsBlack(vr.throughput * throughput);
      }
    }
    PushShadow(w, req, throughput, p);
    const uint32_t b = WarpAppend(&w.counters[kNumActiveNext], to_next);
    if (to_next) w.q_active[next_parity][b] = p;
  }
}

// ------------------------------------------------------------------------------------------------ random-walk SSS
// RandomWalkSubsurface (random-walk-sss.h:227-405) for every parked path.  Walks have wildly different lengths
// (1 .. 8192 bounces), so lanes are refilled from the queue as soon as their walk ends instead of waiting for the
// longest walk of the warp: every loop trip runs at most one bounce (= one short closest-hit query) per lane.
__global__ void __launch_bounds__(128) SssWalkKernel(SceneView s, WaveState w, uint32_t next_parity) {
  const uint32_t n = w.counters[kNumSss];
  const int lane = threadIdx.x & 31;
  bool active = false, exhausted = false;
  uint32_t p = 0;
  Pcg32 rng;
  SssWalkState walk;
  Surface entry_si;
  Frame entry_frame;
  vec3 throughput(0.f), L(0.f);
  uint32_t depth = 0;
  unsigned long long rays = 0;

  for (;;) {
    // ---- refill idle lanes
    const unsigned want = __ballot_sync(0xffffffffu, !active && !exhausted);
    if (want) {
      uint32_t base = 0;
      const int leader = __ffs(want) - 1;
      if (lane == leader) base = atomicAdd(&w.counters[kFetchSss], uint32_t(__popc(want)));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (!active && !exhausted) {
        const uint32_t slot = base + uint32_t(__popc(want & ((1u << lane) - 1u)));
        if (slot >= n) {
          exhausted = true;
        } else {
          p = w.q_sss[slot];
          const RayT ray = LoadRay(w, p);
          const float4 h4 = w.hit[p];
          HitT hit;
          hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
          const float4 t4 = w.thr[p];
          const float4 r4 = w.rad[p];
          throughput = vec3(t4.x, t4.y, t4.z);
          L = vec3(r4.x, r4.y, r4.z);
          depth = __float_as_uint(r4.w);
          const ulonglong2 rs = w.rng[p];
          rng.state = rs.x; rng.inc = rs.y;
          entry_si = MakeSurface(s, ray, hit);
          entry_frame = PrincipledFrame(entry_si);
          const PrincipledBsdf bsdf = SurfaceBsdf(s, entry_si);
          active = SssBegin(entry_si, entry_frame, bsdf, &rng, &walk);
          if (!active) {   // walk rejected: the path's throughput becomes 0 and it ends
            VertexResult vr;
            vr.P = entry_si.P;
            FinishPrincipled(entry_frame, vec3(0.f), vec3(0.f), 0.f, &vr);
            CommitVertex(w, p, vr, throughput, L, depth, rng, next_parity);
          }
        }
      }
    }
    if (__all_sync(0xffffffffu, !active && exhausted)) break;

    // ---- one bounce for every walking lane
    bool to_next = false;
    ShadowRequest req;
    req.active = false;
    if (active) {
      HitT hit;
      const SssStep st = SssBounce(s, &rng, &walk, &hit, nullptr);
      ++rays;
      if (st != kSssContinue) {
        VertexResult vr;
        vr.P = entry_si.P;
        vr.shadow[1].active = false;
        if (st == kSssHit) SssFinish(s, entry_si, entry_frame, walk, hit, &rng, &vr);
        else FinishPrincipled(entry_frame, vec3(0.f), vec3(0.f), 0.f, &vr);
        req = vr.shadow[1];
        CommitVertex(w, p, vr, throughput, L, depth, rng, next_parity);
        to_next = !IsBlack(vr.throughput * throughput);
        active = false;
      }
    }
    PushShadow(w, req, throughput, p);
    const uint32_t b = WarpAppend(&w.counters[kNumActiveNext], to_next);
    if (to_next) w.q_active[next_parity][b] = p;
  }
  if (rays) atomicAdd(&w.stats[kStatSss], rays);
}

// ------------------------------------------------------------------------------------------------ shadow rays
// Scene::AnyHit1 for every NEE request; unoccluded contributions are added to their path's radiance.  A path can
// have two requests in flight in one iteration (entry + SSS exit), hence the atomics (never contended).
__global__ void __launch_bounds__(128) TraceAnyKernel(SceneView s, WaveState w) {
  const uint32_t n = w.counters[kNumShadow];
  unsigned long long rays = 0;
  for (;;) {
    const uint32_t slot = WarpFetch(&w.counters[kFetchShadow]);
    if (__all_sync(0xffffffffu, slot >= n)) break;
    if (slot < n) {
      const float4 o = w.sh_o[slot], d = w.sh_d[slot], c = w.sh_c[slot];
      RayT ray;
      ray.o = vec3(o.x, o.y, o.z); ray.tmin = o.w;
      ray.d = vec3(d.x, d.y, d.z); ray.tmax = d.w;
      ++rays;
      if (!TraceAny<false>(s, ray, nullptr)) {
        float* dst = reinterpret_cast<float*>(&w.rad[__float_as_uint(c.w)]);
        atomicAdd(dst + 0, c.x);
        atomicAdd(dst + 1, c.y);
        atomicAdd(dst + 2, c.z);
      }
    }
  }
  if (rays) atomicAdd(&w.stats[kStatShadow], rays);
}

// ------------------------------------------------------------------------------------------------ accumulation
// rgba += (L, 1), count += 1 per sample (render.cc:175-183), samples of a pixel added in sample order.
__global__ void AccumulateKernel(WaveState w, float4* rgba, uint32_t* count, uint32_t npix, uint32_t spp_wave) {
  const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
  if (pixel >= npix) return;
  float4 acc = rgba[pixel];
  for (uint32_t sidx = 0; sidx < spp_wave; ++sidx) {
    const float4 r = w.rad[sidx * npix + pixel];
    acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += 1.0f;
  }
  rgba[pixel] = acc;
  count[pixel] += spp_wave;
}

// ------------------------------------------------------------------------------------------------ test hooks
__global__ void __launch_bounds__(128) TraceBatchKernel(SceneView s, const float4* __restrict__ rays, uint64_t n,
                                                        float4* hits_tuv, uint4* hits_ids, float4* hits_ng,
                                                        uint32_t* fetch, unsigned long long* stats, int collect) {
  TraverseStats st;
  st.nodes = 0; st.prims = 0;
  for (;;) {
    const uint64_t slot = uint64_t(WarpFetch(fetch));
    if (__all_sync(0xffffffffu, slot >= n)) break;
    if (slot < n) {
      const float4 o = rays[2 * slot], d = rays[2 * slot + 1];
      RayT ray;
      ray.o = vec3(o.x, o.y, o.z); ray.tmin = o.w;
      ray.d = vec3(d.x, d.y, d.z); ray.tmax = d.w;
      HitT hit;
      if (collect) TraceClosest<true>(s, ray, &hit, &st);
      else TraceClosest<false>(s, ray, &hit, nullptr);
      if (hits_tuv) {
        uint4 ids = make_uint4(kInvalid, kInvalid, kInvalid, kInvalid);
        vec3 ng(1.f, 0.f, 0.f);
        float t = 1.0f, u = 0.f, v = 0.f;
        if (hit.prim != kInvalid) {
          ng = HitGeometricNormal(s, hit);
          t = hit.t; u = hit.u; v = hit.v;
          if (hit.prim & kCurveFlag) ids = s.curve_ids[s.curve_prim[hit.prim & ~kCurveFlag]];
          else ids = s.tri_ids[__float_as_uint(s.tri_data[hit.prim * 3].w)];
        }
        hits_tuv[slot] = make_float4(t, u, v, 0.f);
        hits_ids[slot] = ids;
        hits_ng[slot] = make_float4(ng.x, ng.y, ng.z, 0.f);
      }
    }
  }
  if (collect) {
    atomicAdd(&stats[kStatNodes], (unsigned long long)st.nodes);
    atomicAdd(&stats[kStatPrims], (unsigned long long)st.prims);
  }
}

__global__ void __launch_bounds__(128) OccludedBatchKernel(SceneView s, const float4* __restrict__ rays, uint64_t n,
                                                           uint8_t* out, uint32_t* fetch) {
  for (;;) {
    const uint64_t slot = uint64_t(WarpFetch(fetch));
    if (__all_sync(0xffffffffu, slot >= n)) break;
    if (slot < n) {
      const float4 o = rays[2 * slot], d = rays[2 * slot + 1];
      RayT ray;
      ray.o = vec3(o.x, o.y, o.z); ray.tmin = o.w;
      ray.d = vec3(d.x, d.y, d.z); ray.tmax = d.w;
      const bool occ = TraceAny<false>(s, ray, nullptr);
      if (out) out[slot] = occ ? 1 : 0;
    }
  }
}

}  // namespace pbr
