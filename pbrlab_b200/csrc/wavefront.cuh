// Wavefront path tracer kernels (sm_100a).
//
// Render() keeps a POOL of N path slots resident in HBM for the whole frame.  A slot's state is structure-of-arrays,
// every field a 16-byte record so each lane issues one 128-bit load per field and a warp reading a dense queue
// touches contiguous 512-byte spans:
//
//   ray_o[N]  (org.xyz, tmin)      ray_d[N]  (dir.xyz, tmax)      hit[N]  (t, u, v, leaf-order primitive | curve flag)
//   thr[N]    (throughput.rgb, pdf of the last BSDF sample)       rad[N]  (radiance.rgb, depth)
//   rng[N]    (PCG32 state, inc)   pixel[N]  (u32)                walk_a..d[N], walk_n[N]  parked random walk
//
// Index queues (u32 slot ids) are compacted with warp ballots + one atomicAdd per warp:
//
//   q_active[2]  slots that need a closest-hit query (ping-pong between iterations)
//   q_surface    hit a triangle-type material (or none): emission + roulette + Principled vertex
//   q_hair       hit a hair material
//   q_sss        the Principled vertex selected the random-walk closure this iteration
//   q_walk[2]    random walks that used up their bounce budget and continue next iteration (ping-pong)
//   q_done[2]    paths that ended this iteration; consumed at the start of the next one (ping-pong)
//   shadow queue (ray, contribution, slot): NEE any-hit queries
//
// One iteration:
//   begin -> regenerate -> trace_closest -> shade_surface, shade_hair -> sss_walk -> trace_any
// `regenerate` retires every slot of q_done (adds its radiance to its pixel: rgba += (L,1), count += 1, the sums
// RenderLayer holds) and immediately starts the next camera sample in the same slot, so the pool stays full until the
// frame runs out of samples: long random walks and deep paths never leave the GPU idle, and the number of iterations
// is (total rays) / N instead of (longest path).  Every kernel is persistent — a grid that is a fixed multiple of
// the SM count, warps pulling 32-entry batches with an atomic counter — and reads its queue length from device
// memory, so nothing but one small counter block crosses PCIe per iteration.
//
// Replaces the per-pixel loops of the reference (src/render.cc:24-90,125-190); the per-vertex functions and their
// citations are in device/shade.cuh.
#pragma once
#include <cuda_runtime.h>

#include "device/shade.cuh"
#include "device/trav_engine.cuh"

namespace pbr {

enum Counter {
  kNumActive0 = 0, kNumActive1,   // length of q_active[parity]
  kNumWalk0, kNumWalk1,           // length of q_walk[parity]
  kNumDone0, kNumDone1,           // length of q_done[parity]
  kNumSurface, kNumHair, kNumSss, kNumShadow,
  kFetchRegen, kFetchTrace, kFetchSurface, kFetchHair, kFetchSss, kFetchShadow,
  kCounterCount
};
// 64-bit counters that live for a whole frame
enum Stat {
  kStatClosest = 0, kStatShadow, kStatSss, kStatNodes, kStatPrims,
  kStatNextSample,      // next camera sample id to hand out
  kStatRetired,         // camera samples accumulated into the frame so far
  kStatCount
};

struct WaveState {
  float4* ray_o;
  float4* ray_d;
  float4* hit;
  float4* thr;
  float4* rad;
  ulonglong2* rng;
  uint32_t* pixel;
  uint32_t* q_active[2];
  uint32_t* q_walk[2];
  uint32_t* q_done[2];
  uint32_t* q_surface;
  uint32_t* q_hair;
  uint32_t* q_sss;
  float4* walk_a;                // parked walk: (sigma_t.xyz, throughput.x)
  float4* walk_b;                //              (sigma_s.xyz, throughput.y)
  float4* walk_c;                //              (ray.org.xyz, throughput.z)
  float4* walk_d;                //              (ray.dir.xyz, tmin)
  uint32_t* walk_n;              //              bounce count
  float4* sh_o;
  float4* sh_d;
  float4* sh_c;
  uint32_t* counters;            // kCounterCount
  unsigned long long* stats;     // kStatCount
  uint32_t capacity;
};

constexpr uint32_t kNoPixel = 0xFFFFFFFFu;

// ---- warp-aggregated queue append: one atomicAdd per warp.  Must be reached by all 32 lanes.
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
__device__ __forceinline__ HitT LoadHit(const WaveState& w, uint32_t p) {
  const float4 h4 = w.hit[p];
  HitT hit;
  hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
  return hit;
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

// where a slot goes after its vertex: next closest-hit query, or retirement
__device__ __forceinline__ void RouteSlot(const WaveState& w, uint32_t next_parity, uint32_t p, bool to_next,
                                          bool to_done) {
  const uint32_t a = WarpAppend(&w.counters[kNumActive0 + next_parity], to_next);
  if (to_next) w.q_active[next_parity][a] = p;
  const uint32_t b = WarpAppend(&w.counters[kNumDone0 + next_parity], to_done);
  if (to_done) w.q_done[next_parity][b] = p;
}

// ------------------------------------------------------------------------------------------------ iteration set-up
// zeroes every per-iteration counter; the three ping-pong lists keep the half that this iteration consumes
__global__ void BeginIterationKernel(uint32_t* counters, uint32_t cur_parity) {
  const uint32_t i = threadIdx.x;
  if (i >= kCounterCount) return;
  const uint32_t keep0 = kNumActive0 + cur_parity, keep1 = kNumWalk0 + cur_parity, keep2 = kNumDone0 + cur_parity;
  if (i == keep0 || i == keep1 || i == keep2) return;
  counters[i] = 0u;
}

// ------------------------------------------------------------------------------------------------ camera + retire
// RenderingTile's ray generation (src/render.cc:160-171): target = corner + d*(pixel + xi); the two jitter draws
// are the first two numbers of the path's stream.
struct CameraParams {
  float eye[3], x_corner, y_corner, z_corner, dx, dy;
  uint32_t width, height;
};
struct FrameParams {
  CameraParams cam;
  uint32_t npix;
  unsigned long long total_samples;   // camera samples this device renders in this call = npix * local_spp
  uint64_t seed;
  uint32_t first_sample;      // global index of local sample 0
  uint32_t sample_stride;     // global sample index = first_sample + local * stride
  float4* rgba;               // frame accumulators (sums)
  uint32_t* count;
};

__device__ __forceinline__ void StartCameraPath(const WaveState& w, const FrameParams& f, uint32_t p,
                                                unsigned long long id) {
  const uint32_t s_local = uint32_t(id / f.npix), pixel = uint32_t(id - (unsigned long long)s_local * f.npix);
  const uint32_t x = pixel % f.cam.width, y = pixel / f.cam.width;
  Pcg32 rng;
  pcg32_srandom(&rng, f.seed + uint64_t(f.first_sample) + uint64_t(s_local) * f.sample_stride, uint64_t(pixel));
  const float jx = Draw(&rng), jy = Draw(&rng);
  const float tx = f.cam.x_corner + f.cam.dx * (float(x) + jx);
  const float ty = f.cam.y_corner - f.cam.dy * (float(y) + jy);
  float dx = tx - f.cam.eye[0], dy = ty - f.cam.eye[1], dz = f.cam.z_corner - f.cam.eye[2];
  const float inv_norm = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);   // Normalize (render.cc:243-249)
  dx *= inv_norm; dy *= inv_norm; dz *= inv_norm;
  w.ray_o[p] = make_float4(f.cam.eye[0], f.cam.eye[1], f.cam.eye[2], 0.0f);
  w.ray_d[p] = make_float4(dx, dy, dz, kInf);
  w.thr[p] = make_float4(1.f, 1.f, 1.f, 0.f);
  w.rad[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
  w.rng[p] = make_ulonglong2(rng.state, rng.inc);
  w.pixel[p] = pixel;
}

// Retire the slots of q_done[cur] (render.cc:175-183: rgba += (L, 1), count += 1) and restart them on the next
// camera samples.  Several samples of one pixel can retire in the same iteration, hence atomics.
__global__ void __launch_bounds__(256) RegenerateKernel(WaveState w, FrameParams f, uint32_t cur_parity) {
  const uint32_t n = w.counters[kNumDone0 + cur_parity];
  unsigned long long retired = 0;
  for (;;) {
    const uint32_t i = WarpFetch(&w.counters[kFetchRegen]);
    if (__all_sync(0xffffffffu, i >= n)) break;
    const bool valid = i < n;
    uint32_t p = 0;
    if (valid) {
      p = w.q_done[cur_parity][i];
      const uint32_t pixel = w.pixel[p];
      if (pixel != kNoPixel) {
        const float4 r = w.rad[p];
        float* dst = reinterpret_cast<float*>(&f.rgba[pixel]);
        atomicAdd(dst + 0, r.x);
        atomicAdd(dst + 1, r.y);
        atomicAdd(dst + 2, r.z);
        atomicAdd(dst + 3, 1.0f);
        atomicAdd(&f.count[pixel], 1u);
        w.pixel[p] = kNoPixel;
        ++retired;
      }
    }
    // hand out new sample ids, one atomic per warp
    const unsigned mask = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (mask) {
      const int leader = __ffs(mask) - 1;
      if (lane == leader) base = atomicAdd(&w.stats[kStatNextSample], (unsigned long long)__popc(mask));
      base = __shfl_sync(0xffffffffu, base, leader);
    }
    const unsigned long long id = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
    const bool start = valid && id < f.total_samples;
    if (start) StartCameraPath(w, f, p, id);
    const uint32_t a = WarpAppend(&w.counters[kNumActive0 + cur_parity], start);
    if (start) w.q_active[cur_parity][a] = p;
  }
  if (retired) atomicAdd(&w.stats[kStatRetired], retired);
}

// all slots idle and queued for (re)generation: the state a frame starts from
__global__ void ResetPoolKernel(WaveState w, uint32_t n_slots) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_slots) {
    w.pixel[p] = kNoPixel;
    w.q_done[0][p] = p;
  }
  if (p < kCounterCount) w.counters[p] = (p == kNumDone0) ? n_slots : 0u;
}

// caller-supplied rays + seeds (pbrgpu_radiance / pbrgpu_shade hooks): slot i = path i = pixel i, no regeneration
__global__ void InitPathsFromRaysKernel(WaveState w, const float4* rays, const uint64_t* seeds, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < kCounterCount) w.counters[p] = (p == kNumActive0) ? n : 0u;
  if (p >= n) return;
  w.ray_o[p] = rays[2 * p];
  w.ray_d[p] = rays[2 * p + 1];
  Pcg32 rng;
  pcg32_srandom(&rng, seeds[2 * p], seeds[2 * p + 1]);
  w.thr[p] = make_float4(1.f, 1.f, 1.f, 0.f);
  w.rad[p] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
  w.rng[p] = make_ulonglong2(rng.state, rng.inc);
  w.pixel[p] = p;
  w.q_active[0][p] = p;
}

// ------------------------------------------------------------------------------------------------ closest hit
// Scene::TraceFirstHit1 for every active slot; routes it by the material kind of what it hit, retires it on a miss.
// Runs in the warp traversal engine (device/trav_engine.cuh): lanes are refilled from q_active while others traverse.
struct ClosestClient {
  const SceneView& s;
  const WaveState& w;
  const uint32_t* __restrict__ queue;
  uint32_t n, next_parity;
  uint32_t p = 0;
  bool has_result = false;
  unsigned long long rays = 0;

  __device__ __forceinline__ ClosestClient(const SceneView& s_, const WaveState& w_, uint32_t cur_parity)
      : s(s_), w(w_), queue(w_.q_active[cur_parity]), n(w_.counters[kNumActive0 + cur_parity]),
        next_parity(cur_parity ^ 1u) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || !exhausted; }
  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    int kind = -1;   // -1 nothing, 0 miss, 1 surface queue, 2 hair queue
    if (!t.active && has_result) {
      has_result = false;
      const HitT hit = t.hit;
      w.hit[p] = make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
      kind = 0;
      if (hit.prim != kInvalid) {
        uint32_t mat;
        if (hit.prim & kCurveFlag) mat = s.curve_ids[s.curve_prim[hit.prim & ~kCurveFlag]].w;
        else mat = s.tri_ids[__float_as_uint(s.tri_data[hit.prim * 3].w)].w;
        kind = (mat < s.num_materials && s.materials[mat].type == 1u) ? 2 : 1;
      }
    }
    const uint32_t done_p = p;
    const uint32_t a = WarpAppend(&w.counters[kNumSurface], kind == 1);
    if (kind == 1) w.q_surface[a] = done_p;
    const uint32_t b = WarpAppend(&w.counters[kNumHair], kind == 2);
    if (kind == 2) w.q_hair[b] = done_p;
    const uint32_t c = WarpAppend(&w.counters[kNumDone0 + next_parity], kind == 0);
    if (kind == 0) w.q_done[next_parity][c] = done_p;
    bool dry = false;
    if (!exhausted) {
      const bool need = !t.active;
      const uint32_t slot = WarpAppend(&w.counters[kFetchTrace], need);
      if (need) {
        if (slot < n) {
          p = queue[slot];
          TravBegin(s, LoadRay(w, p), t);
          has_result = true;
          ++rays;
        } else {
          dry = true;
        }
      }
    }
    return dry;
  }
  __device__ __forceinline__ void End(const Trav&) {
    if (rays) atomicAdd(&w.stats[kStatClosest], rays);
  }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) TraceClosestKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                          uint32_t refill_min_idle) {
  ClosestClient client(s, w, cur_parity);
  TravEngine<false, HAS_CURVES, false>(s, client, refill_min_idle);
}

// ------------------------------------------------------------------------------------------------ shading
struct ShadeFlags {
  uint32_t skip_emission_and_roulette;   // pbrgpu_shade hook: call Shader() only
};

__device__ __forceinline__ void CommitVertex(const WaveState& w, uint32_t p, const VertexResult& vr,
                                             const vec3& throughput, const vec3& L, uint32_t depth, const Pcg32& rng) {
  const vec3 new_thr = vr.throughput * throughput;                     // render.cc:80
  w.thr[p] = make_float4(new_thr.x, new_thr.y, new_thr.z, vr.pdf);
  w.rad[p] = make_float4(L.x, L.y, L.z, __uint_as_float(depth + 1u));
  w.rng[p] = make_ulonglong2(rng.state, rng.inc);
  w.ray_o[p] = make_float4(vr.P.x, vr.P.y, vr.P.z, 1e-3f);             // render.cc:83-86
  w.ray_d[p] = make_float4(vr.wi.x, vr.wi.y, vr.wi.z, kInf);
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
    bool to_sss = false, to_next = false, to_done = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    if (valid) {
      p = w.q_surface[slot];
      const RayT ray = LoadRay(w, p);
      const HitT hit = LoadHit(w, p);
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
        to_done = true;
      } else {
        const int kind = MaterialKind(s, si);
        VertexResult vr;
        const vec3 wo = -ray.d;
        if (kind == 1) to_sss = PrincipledVertex(s, si, wo, &rng, &vr);
        else AbsorbVertex(wo, si.P, &vr);   // no material (shader.cc:11-17)
        req = vr.shadow[0];
        if (to_sss) {
          // the walk runs in its own kernel: park the path with the post-roulette throughput and the rng positioned
          // right after the closure selector
          w.thr[p] = make_float4(throughput.x, throughput.y, throughput.z, t4.w);
          w.rad[p] = make_float4(L.x, L.y, L.z, __uint_as_float(depth));
          w.rng[p] = make_ulonglong2(rng.state, rng.inc);
        } else {
          CommitVertex(w, p, vr, throughput, L, depth, rng);
          to_next = !IsBlack(vr.throughput * throughput);               // render.cc:31
          to_done = !to_next;
        }
      }
    }
    PushShadow(w, req, throughput, p);
    const uint32_t a = WarpAppend(&w.counters[kNumSss], to_sss);
    if (to_sss) w.q_sss[a] = p;
    RouteSlot(w, next_parity, p, to_next, to_done);
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
    bool to_next = false, to_done = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    if (valid) {
      p = w.q_hair[slot];
      const RayT ray = LoadRay(w, p);
      const HitT hit = LoadHit(w, p);
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
        to_done = true;
      } else {
        VertexResult vr;
        HairVertex(s, si, -ray.d, &rng, &vr);
        req = vr.shadow[0];
        CommitVertex(w, p, vr, throughput, L, depth, rng);
        to_next = !IsBlack(vr.throughput * throughput);
        to_done = !to_next;
      }
    }
    PushShadow(w, req, throughput, p);
    RouteSlot(w, next_parity, p, to_next, to_done);
  }
}

// ------------------------------------------------------------------------------------------------ random-walk SSS
// RandomWalkSubsurface (random-walk-sss.h:227-405).  Walk lengths are wildly uneven (1 .. 8192 bounces; Lucy's red
// channel has albedo ~1, so neither absorption nor roulette ends a walk early) and every bounce is a dependent
// short closest-hit query (~10 us of latency), hence:
//   * lanes are refilled from the queue the moment their walk ends — every loop trip runs at most one bounce per
//     lane, no lane waits for the longest walk of its warp;
//   * a walk gets at most `max_bounces` bounces per launch; if it is still inside the medium its 68-byte state is
//     parked in HBM (walk_a..d, walk_n) and the slot goes to q_walk[next]: the next iteration resumes it first.
//     A launch therefore never outlives its queue by more than max_bounces bounces, and because the pool is kept
//     full by regeneration, long walks cost slots, not idle SMs.
__device__ __forceinline__ void ParkWalk(const WaveState& w, uint32_t p, const SssWalkState& k) {
  w.walk_a[p] = make_float4(k.sigma_t.x, k.sigma_t.y, k.sigma_t.z, k.throughput.x);
  w.walk_b[p] = make_float4(k.sigma_s.x, k.sigma_s.y, k.sigma_s.z, k.throughput.y);
  w.walk_c[p] = make_float4(k.ray.o.x, k.ray.o.y, k.ray.o.z, k.throughput.z);
  w.walk_d[p] = make_float4(k.ray.d.x, k.ray.d.y, k.ray.d.z, k.ray.tmin);
  w.walk_n[p] = k.bounce;
}
__device__ __forceinline__ void ResumeWalk(const WaveState& w, uint32_t p, SssWalkState* k) {
  const float4 a = w.walk_a[p], b = w.walk_b[p], c = w.walk_c[p], d = w.walk_d[p];
  k->sigma_t = vec3(a.x, a.y, a.z);
  k->sigma_s = vec3(b.x, b.y, b.z);
  k->throughput = vec3(a.w, b.w, c.w);
  k->ray.o = vec3(c.x, c.y, c.z);
  k->ray.d = vec3(d.x, d.y, d.z);
  k->ray.tmin = d.w;
  k->ray.tmax = kInf;
  k->bounce = w.walk_n[p];
}

struct SssClient {
  const SceneView& s;
  const WaveState& w;
  uint32_t cur_parity, next_parity, n_resume, n, max_bounces;
  uint32_t p = 0, budget = 0;
  bool has_walk = false;
  Pcg32 rng;
  SssWalkState walk;    // walk.ray is rebuilt from the traversal state after every segment
  unsigned long long rays = 0;

  __device__ __forceinline__ SssClient(const SceneView& s_, const WaveState& w_, uint32_t cur, uint32_t max_b)
      : s(s_), w(w_), cur_parity(cur), next_parity(cur ^ 1u), n_resume(w_.counters[kNumWalk0 + cur]),
        n(w_.counters[kNumWalk0 + cur] + w_.counters[kNumSss]), max_bounces(max_b) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_walk || !exhausted; }

  __device__ __forceinline__ void StartSegment(Trav& t) {
    SssPrepareSegment(&rng, &walk);
    TravBegin(s, walk.ray, t);
    if (!t.active) {   // empty scene: the segment ends without a hit
      t.hit.prim = kInvalid;
    }
  }

  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    // ---- (1) walks whose segment query finished: scatter / exit / absorb
    bool to_next = false, to_done = false, to_park = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    uint32_t routed_p = p;
    if (!t.active && has_walk) {
      walk.ray.o = t.O; walk.ray.d = t.D; walk.ray.tmin = t.tmin;   // tmax untouched: the scatter distance
      const bool is_hit = t.hit.prim != kInvalid;
      const SssStep st = SssFinishSegment(is_hit, t.hit.t, &rng, &walk);
      ++rays;
      --budget;
      if (st != kSssContinue) {
        // the entry vertex is only needed now: rebuild it from the parked path's ray + hit
        const RayT ray = LoadRay(w, p);
        const Surface entry_si = MakeSurface(s, ray, LoadHit(w, p));
        const Frame entry_frame = PrincipledFrame(entry_si);
        const float4 t4 = w.thr[p], r4 = w.rad[p];
        throughput = vec3(t4.x, t4.y, t4.z);
        VertexResult vr;
        vr.P = entry_si.P;
        vr.shadow[1].active = false;
        if (st == kSssHit) SssFinish(s, entry_si, entry_frame, walk, t.hit, &rng, &vr);
        else FinishPrincipled(entry_frame, vec3(0.f), vec3(0.f), 0.f, &vr);
        req = vr.shadow[1];
        CommitVertex(w, p, vr, throughput, vec3(r4.x, r4.y, r4.z), __float_as_uint(r4.w), rng);
        to_next = !IsBlack(vr.throughput * throughput);
        to_done = !to_next;
        has_walk = false;
      } else if (budget == 0u) {
        ParkWalk(w, p, walk);
        w.rng[p] = make_ulonglong2(rng.state, rng.inc);
        to_park = true;
        has_walk = false;
      } else {
        StartSegment(t);
      }
    }
    PushShadow(w, req, throughput, routed_p);
    RouteSlot(w, next_parity, routed_p, to_next, to_done);
    const uint32_t c = WarpAppend(&w.counters[kNumWalk0 + next_parity], to_park);
    if (to_park) w.q_walk[next_parity][c] = routed_p;

    // ---- (2) lanes without a walk take the next one: parked walks first, then this iteration's new ones
    bool dry = false, rejected = false;
    if (!exhausted) {
      const bool need = !t.active && !has_walk;
      const uint32_t slot = WarpAppend(&w.counters[kFetchSss], need);
      if (need) {
        if (slot >= n) {
          dry = true;
        } else {
          const bool resume = slot < n_resume;
          p = resume ? w.q_walk[cur_parity][slot] : w.q_sss[slot - n_resume];
          const ulonglong2 rs = w.rng[p];
          rng.state = rs.x; rng.inc = rs.y;
          budget = max_bounces;
          if (resume) {
            ResumeWalk(w, p, &walk);
            has_walk = true;
          } else {
            const RayT ray = LoadRay(w, p);
            const Surface entry_si = MakeSurface(s, ray, LoadHit(w, p));
            const Frame entry_frame = PrincipledFrame(entry_si);
            const PrincipledBsdf bsdf = SurfaceBsdf(s, entry_si);
            has_walk = SssBegin(entry_si, entry_frame, bsdf, &rng, &walk);
            if (!has_walk) {   // walk rejected: the path's throughput becomes 0 and it ends
              const float4 t4 = w.thr[p], r4 = w.rad[p];
              VertexResult vr;
              vr.P = entry_si.P;
              FinishPrincipled(entry_frame, vec3(0.f), vec3(0.f), 0.f, &vr);
              CommitVertex(w, p, vr, vec3(t4.x, t4.y, t4.z), vec3(r4.x, r4.y, r4.z), __float_as_uint(r4.w), rng);
              rejected = true;
            }
          }
          if (has_walk) StartSegment(t);
        }
      }
      RouteSlot(w, next_parity, p, false, rejected);
    }
    return dry;
  }
  __device__ __forceinline__ void End(const Trav&) {
    if (rays) atomicAdd(&w.stats[kStatSss], rays);
  }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) SssWalkKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                     uint32_t max_bounces, uint32_t refill_min_idle) {
  SssClient client(s, w, cur_parity, max_bounces);
  TravEngine<false, HAS_CURVES, false>(s, client, refill_min_idle);
}

// ------------------------------------------------------------------------------------------------ shadow rays
// Scene::AnyHit1 for every NEE request; unoccluded contributions are added to their path's radiance.  A path can
// have two requests in flight in one iteration (entry + SSS exit), hence the atomics (never contended).
struct ShadowClient {
  const SceneView& s;
  const WaveState& w;
  uint32_t n;
  float4 c;
  bool has_result = false;
  unsigned long long rays = 0;

  __device__ __forceinline__ ShadowClient(const SceneView& s_, const WaveState& w_)
      : s(s_), w(w_), n(w_.counters[kNumShadow]) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || !exhausted; }
  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    if (!t.active && has_result) {
      has_result = false;
      if (t.hit.prim == kInvalid) {
        float* dst = reinterpret_cast<float*>(&w.rad[__float_as_uint(c.w)]);
        atomicAdd(dst + 0, c.x);
        atomicAdd(dst + 1, c.y);
        atomicAdd(dst + 2, c.z);
      }
    }
    bool dry = false;
    if (!exhausted) {
      const bool need = !t.active;
      const uint32_t slot = WarpAppend(&w.counters[kFetchShadow], need);
      if (need) {
        if (slot < n) {
          const float4 o = w.sh_o[slot], d = w.sh_d[slot];
          c = w.sh_c[slot];
          RayT ray;
          ray.o = vec3(o.x, o.y, o.z); ray.tmin = o.w;
          ray.d = vec3(d.x, d.y, d.z); ray.tmax = d.w;
          TravBegin(s, ray, t);
          has_result = true;
          ++rays;
        } else {
          dry = true;
        }
      }
    }
    return dry;
  }
  __device__ __forceinline__ void End(const Trav&) {
    if (rays) atomicAdd(&w.stats[kStatShadow], rays);
  }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) TraceAnyKernel(SceneView s, WaveState w, uint32_t refill_min_idle) {
  ShadowClient client(s, w);
  TravEngine<true, HAS_CURVES, false>(s, client, refill_min_idle);
}

// ------------------------------------------------------------------------------------------------ test hooks
// pbrgpu_trace / pbrgpu_occluded: caller-supplied ray batches through the same engine as the render kernels
struct BatchClient {
  const SceneView& s;
  const float4* __restrict__ rays;
  uint64_t n;
  float4* hits_tuv;
  uint4* hits_ids;
  float4* hits_ng;
  uint8_t* occluded;
  uint32_t* fetch;
  unsigned long long* stats;
  uint64_t slot_of_result = 0;
  bool has_result = false;

  __device__ __forceinline__ BatchClient(const SceneView& s_, const float4* r, uint64_t n_, float4* tuv, uint4* ids,
                                         float4* ng, uint8_t* occ, uint32_t* f, unsigned long long* st)
      : s(s_), rays(r), n(n_), hits_tuv(tuv), hits_ids(ids), hits_ng(ng), occluded(occ), fetch(f), stats(st) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || !exhausted; }
  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    if (!t.active && has_result) {
      has_result = false;
      const HitT hit = t.hit;
      const uint64_t slot = slot_of_result;
      if (occluded) occluded[slot] = (hit.prim != kInvalid) ? 1 : 0;
      if (hits_tuv) {
        uint4 ids = make_uint4(kInvalid, kInvalid, kInvalid, kInvalid);
        vec3 ng(1.f, 0.f, 0.f);
        float ht = 1.0f, hu = 0.f, hv = 0.f;
        if (hit.prim != kInvalid) {
          ng = HitGeometricNormal(s, hit);
          ht = hit.t; hu = hit.u; hv = hit.v;
          if (hit.prim & kCurveFlag) ids = s.curve_ids[s.curve_prim[hit.prim & ~kCurveFlag]];
          else ids = s.tri_ids[__float_as_uint(s.tri_data[hit.prim * 3].w)];
        }
        hits_tuv[slot] = make_float4(ht, hu, hv, 0.f);
        hits_ids[slot] = ids;
        hits_ng[slot] = make_float4(ng.x, ng.y, ng.z, 0.f);
      }
    }
    bool dry = false;
    if (!exhausted) {
      const bool need = !t.active;
      const uint64_t slot = uint64_t(WarpAppend(fetch, need));
      if (need) {
        if (slot < n) {
          const float4 o = rays[2 * slot], d = rays[2 * slot + 1];
          RayT ray;
          ray.o = vec3(o.x, o.y, o.z); ray.tmin = o.w;
          ray.d = vec3(d.x, d.y, d.z); ray.tmax = d.w;
          TravBegin(s, ray, t);
          slot_of_result = slot;
          has_result = true;
        } else {
          dry = true;
        }
      }
    }
    return dry;
  }
  __device__ __forceinline__ void End(const Trav& t) {
    if (stats) {
      atomicAdd(&stats[kStatNodes], (unsigned long long)t.n_nodes);
      atomicAdd(&stats[kStatPrims], (unsigned long long)t.n_prims);
    }
  }
};

template <bool HAS_CURVES, bool STATS>
__global__ void __launch_bounds__(128) TraceBatchKernel(SceneView s, const float4* __restrict__ rays, uint64_t n,
                                                        float4* hits_tuv, uint4* hits_ids, float4* hits_ng,
                                                        uint32_t* fetch, unsigned long long* stats,
                                                        uint32_t refill_min_idle) {
  BatchClient client(s, rays, n, hits_tuv, hits_ids, hits_ng, nullptr, fetch, STATS ? stats : nullptr);
  TravEngine<false, HAS_CURVES, STATS>(s, client, refill_min_idle);
}

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) OccludedBatchKernel(SceneView s, const float4* __restrict__ rays, uint64_t n,
                                                           uint8_t* out, uint32_t* fetch, uint32_t refill_min_idle) {
  BatchClient client(s, rays, n, nullptr, nullptr, nullptr, out, fetch, nullptr);
  TravEngine<true, HAS_CURVES, false>(s, client, refill_min_idle);
}

}  // namespace pbr
