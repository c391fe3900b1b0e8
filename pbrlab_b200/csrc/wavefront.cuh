// Wavefront path tracer kernels (sm_100a).
//
// Render() keeps a POOL of N path slots resident in HBM for the whole frame.  Slots are visited through index
// queues, i.e. in random order, so a slot's state is ONE 128-byte line (eight 16-byte records) instead of eight
// separate arrays: a random 16-byte access costs a full 32-byte sector (64 bytes at the DRAM), and measured on B200 the
// one-array-per-field layout moved 1.3 kB of DRAM traffic per shaded vertex for 176 bytes of state
// (profiles/r1a_summary.md).  The line is grouped into 32-byte sectors by WRITER, so that every store replaces whole
// sectors and never forces a read-modify-write:
//
//   sector 0   ray_o (org.xyz, tmin)            ray_d (dir.xyz, tmax)              written by shade / regenerate
//   sector 1   thr   (throughput.rgb, pdf)      rad   (radiance.rgb, depth)        written by shade / regenerate
//   sector 2   hit   (t, u, v, leaf-order prim) pad                                written by trace_closest
//   sector 3   rng   (PCG32 state, inc)         pix   (pixel, -, -, -)             written by shade / regenerate
//
// A second line per slot (`walk`) holds a parked random walk or the exit record of a finished one; only slots
// that are inside a subsurface walk ever touch it.  Slot state is read and written with streaming (evict-first)
// cache hints so that the ~1-2 GB that stream through per iteration do not push the BVH (19 MB on the Cornell scene)
// out of the 126 MB L2.
//
// Index queues (u32 slot ids) are compacted with warp ballots + one atomicAdd per warp:
//
//   q_active[2]  slots that need a closest-hit query (ping-pong between iterations)
//   q_surface    hit a Principled material of the general class (or none): emission + roulette + Principled vertex
//   q_diffuse    hit a Principled material that can only enable the Lambert closure (scene_host.cc: ClassifyMaterial):
//                same vertex function with the other closures compiled out (material-sorted shading)
//   q_hair       hit a hair material
//   q_sss        the Principled vertex selected the random-walk closure this iteration (walk state already parked)
//   q_walk[2]    random walks that used up their bounce budget and continue next iteration (ping-pong)
//   q_exit       walks that left the medium this iteration: exit vertex still to be shaded
//   q_done[2]    paths that ended this iteration; consumed at the start of the next one (ping-pong)
//   shadow queue (ray, contribution, slot) as three dense arrays: NEE any-hit queries, written and read coalesced
//
// One iteration:
//   begin -> regenerate -> trace_closest -> shade_surface, shade_hair -> sss_walk -> sss_exit -> trace_any
// `regenerate` retires every slot of q_done (adds its radiance to its pixel: rgba += (L,1), the sums RenderLayer
// holds) and immediately starts the next camera sample in the same slot, so the pool stays full until the frame runs
// out of samples: long random walks and deep paths never leave the GPU idle, and the number of iterations is
// (total rays) / N instead of (longest path).  Every kernel is persistent — a grid that is a fixed multiple of the SM
// count, warps pulling work with an atomic counter — and reads its queue length from device memory, so nothing but
// one small counter block crosses PCIe per iteration.  The three ray kernels run in the warp traversal engine
// (device/trav_engine.cuh), which refills finished lanes while the rest of the warp keeps traversing.
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
  kNumSurface, kNumDiffuse, kNumHair, kNumSss, kNumExit, kNumShadow,
  kFetchRegen, kFetchTrace, kFetchSurface, kFetchDiffuse, kFetchHair, kFetchSss, kFetchExit, kFetchShadow,
  kCounterCount
};
// 64-bit counters that live for a whole frame
enum Stat {
  kStatClosest = 0, kStatShadow, kStatSss, kStatNodes, kStatPrims,
  kStatNextSample,      // next camera sample id to hand out
  kStatSampleBase,      // first camera sample id of the current iteration (set by BeginIterationKernel)
  kStatRetired,         // camera samples accumulated into the frame so far
  kStatSssSkipped,      // walk segments answered by the clearance grid
  kStatVertices,        // path vertices shaded (surface + diffuse + hair + SSS exit kernels)
  kStatCount
};

// records of a slot line / walk line (units of float4)
enum SlotField { kRayO = 0, kRayD = 1, kThr = 2, kRad = 3, kHit = 4, kHitPad = 5, kRng = 6, kPix = 7, kSlotStride = 8 };
enum WalkField { kWalkA = 0, kWalkB = 1, kWalkC = 2, kWalkD = 3, kWalkN = 4, kWalkStride = 8 };

struct WaveState {
  float4* slot;                  // kSlotStride x float4 per path slot
  float4* walk;                  // kWalkStride x float4 per path slot
  uint32_t* q_active[2];
  uint32_t* q_walk[2];
  uint32_t* q_done[2];
  uint32_t* q_surface;
  uint32_t* q_diffuse;
  uint32_t* q_hair;
  uint32_t* q_sss;
  uint32_t* q_exit;
  float4* sh_o;
  float4* sh_d;
  float4* sh_c;
  uint32_t* counters;            // kCounterCount
  unsigned long long* stats;     // kStatCount
  uint32_t capacity;
};

constexpr uint32_t kNoPixel = 0xFFFFFFFFu;
constexpr int kShadeBlock = 512;   // most threads per block of the shading kernels (block-synchronous batches)

// ---- streaming access to slot state (ld/st.global.cs: evict-first in L2)
__device__ __forceinline__ float4 LdSlot(const WaveState& w, uint32_t p, int field) {
  return __ldcs(&w.slot[size_t(p) * kSlotStride + field]);
}
__device__ __forceinline__ void StSlot(const WaveState& w, uint32_t p, int field, const float4& v) {
  __stcs(&w.slot[size_t(p) * kSlotStride + field], v);
}
__device__ __forceinline__ float4 LdWalk(const WaveState& w, uint32_t p, int field) {
  return __ldcs(&w.walk[size_t(p) * kWalkStride + field]);
}
__device__ __forceinline__ void StWalk(const WaveState& w, uint32_t p, int field, const float4& v) {
  __stcs(&w.walk[size_t(p) * kWalkStride + field], v);
}

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

// Five queue reservations with ONE atomic instruction: lanes 0..4 each reserve the range of one queue (different
// addresses, so the L2 handles them in parallel) instead of five dependent atomic round trips — waiting for atomic
// results was 18 % of the closest-hit kernel's stall samples (profiles/r1e_ncu.md).  Issue early, resolve late:
// whatever is issued in between overlaps the round trip.
struct Append5 {
  unsigned m[5];
  uint32_t base;   // lane k < 5: first index reserved in queue k
};
__device__ __forceinline__ Append5 Append5Issue(uint32_t* c0, uint32_t* c1, uint32_t* c2, uint32_t* c3, uint32_t* c4,
                                                bool p0, bool p1, bool p2, bool p3, bool p4) {
  Append5 a;
  a.m[0] = __ballot_sync(0xffffffffu, p0);
  a.m[1] = __ballot_sync(0xffffffffu, p1);
  a.m[2] = __ballot_sync(0xffffffffu, p2);
  a.m[3] = __ballot_sync(0xffffffffu, p3);
  a.m[4] = c4 ? __ballot_sync(0xffffffffu, p4) : 0u;
  a.base = 0;
  const int lane = threadIdx.x & 31;
  const unsigned mine = lane == 0 ? a.m[0] : (lane == 1 ? a.m[1] : (lane == 2 ? a.m[2] : (lane == 3 ? a.m[3] : a.m[4])));
  const uint32_t cnt = uint32_t(__popc(mine));
  uint32_t* ctr = lane == 0 ? c0 : (lane == 1 ? c1 : (lane == 2 ? c2 : (lane == 3 ? c3 : c4)));
  if (lane < 5 && cnt) a.base = atomicAdd(ctr, cnt);
  return a;
}
__device__ __forceinline__ uint32_t Append5Index(const Append5& a, int k) {
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  return __shfl_sync(0xffffffffu, a.base, k) + uint32_t(__popc(a.m[k] & lt));
}

// one 64-bit atomic per warp for a per-lane tally (kernel epilogues: 32 same-address atomics per warp otherwise)
__device__ __forceinline__ void WarpTally(unsigned long long* counter, uint32_t v) {
  const uint32_t sum = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(counter, (unsigned long long)sum);
}

// A whole block pulls the next blockDim.x queue slots and re-converges: the shading kernels are long straight-line
// code (70-120 KB of SASS, every instruction executed once per path), so warps that drift apart each stream the
// code from L2 on their own — measured 67 % of warp time waiting on instruction fetch.  Warps that start every batch
// together share the fetched lines.
__device__ __forceinline__ uint32_t BlockFetch(uint32_t* fetch_counter) {
  __shared__ uint32_t s_base;
  __syncthreads();
  if (threadIdx.x == 0) s_base = atomicAdd(fetch_counter, blockDim.x);
  __syncthreads();
  return s_base + threadIdx.x;
}

__device__ __forceinline__ RayT LoadRay(const WaveState& w, uint32_t p) {
  const float4 o = LdSlot(w, p, kRayO), d = LdSlot(w, p, kRayD);
  RayT r;
  r.o = vec3(o.x, o.y, o.z); r.tmin = o.w;
  r.d = vec3(d.x, d.y, d.z); r.tmax = d.w;
  return r;
}
__device__ __forceinline__ HitT LoadHit(const WaveState& w, uint32_t p) {
  const float4 h4 = LdSlot(w, p, kHit);
  HitT hit;
  hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
  return hit;
}
__device__ __forceinline__ Pcg32 LoadRng(const WaveState& w, uint32_t p) {
  const float4 r = LdSlot(w, p, kRng);
  Pcg32 rng;
  rng.state = (uint64_t(__float_as_uint(r.y)) << 32) | __float_as_uint(r.x);
  rng.inc = (uint64_t(__float_as_uint(r.w)) << 32) | __float_as_uint(r.z);
  return rng;
}
__device__ __forceinline__ float4 PackRng(const Pcg32& rng) {
  return make_float4(__uint_as_float(uint32_t(rng.state)), __uint_as_float(uint32_t(rng.state >> 32)),
                     __uint_as_float(uint32_t(rng.inc)), __uint_as_float(uint32_t(rng.inc >> 32)));
}

__device__ __forceinline__ void PushShadow(const WaveState& w, const ShadowRequest& req, const vec3& throughput,
                                           uint32_t path) {
  const uint32_t slot = WarpAppend(&w.counters[kNumShadow], req.active);
  if (req.active) {
    const vec3 c = throughput * req.contribute;
    __stcs(&w.sh_o[slot], make_float4(req.ray.o.x, req.ray.o.y, req.ray.o.z, req.ray.tmin));
    __stcs(&w.sh_d[slot], make_float4(req.ray.d.x, req.ray.d.y, req.ray.d.z, req.ray.tmax));
    __stcs(&w.sh_c[slot], make_float4(c.x, c.y, c.z, __uint_as_float(path)));
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
// In frame mode it also hands this iteration's camera sample ids to the slots of q_done[cur]: the closest-hit kernel
// gives work item (n_active + k) the sample id (base + k), so regeneration needs no atomic of its own.
__global__ void BeginIterationKernel(uint32_t* counters, unsigned long long* stats, uint32_t cur_parity,
                                     uint32_t frame_mode) {
  const uint32_t i = threadIdx.x;
  if (i == 0 && frame_mode) {
    const unsigned long long base = stats[kStatNextSample];
    stats[kStatSampleBase] = base;
    stats[kStatNextSample] = base + counters[kNumDone0 + cur_parity];
  }
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
  float4* rgba;               // frame accumulator (sums); alpha counts the samples (render.cc:175-183)
};

__device__ __forceinline__ RayT StartCameraPath(const WaveState& w, const FrameParams& f, uint32_t p,
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
  StSlot(w, p, kRayO, make_float4(f.cam.eye[0], f.cam.eye[1], f.cam.eye[2], 0.0f));
  StSlot(w, p, kRayD, make_float4(dx, dy, dz, kInf));
  StSlot(w, p, kThr, make_float4(1.f, 1.f, 1.f, 0.f));
  StSlot(w, p, kRad, make_float4(0.f, 0.f, 0.f, __uint_as_float(0u)));
  StSlot(w, p, kRng, PackRng(rng));
  StSlot(w, p, kPix, make_float4(__uint_as_float(pixel), 0.f, 0.f, 0.f));
  RayT ray;
  ray.o = vec3(f.cam.eye[0], f.cam.eye[1], f.cam.eye[2]); ray.tmin = 0.0f;
  ray.d = vec3(dx, dy, dz); ray.tmax = kInf;
  return ray;
}

// RenderLayer::count (render-layer.h:11-26) from the alpha sums: both are incremented together per sample
__global__ void FinishFrameKernel(const float4* rgba, uint32_t* count, uint32_t npix) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix) count[i] = uint32_t(rgba[i].w);
}

// all slots idle and queued for (re)generation: the state a frame starts from
__global__ void ResetPoolKernel(WaveState w, uint32_t n_slots) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_slots) {
    StSlot(w, p, kPix, make_float4(__uint_as_float(kNoPixel), 0.f, 0.f, 0.f));
    w.q_done[0][p] = p;
  }
  if (p < kCounterCount) w.counters[p] = (p == kNumDone0) ? n_slots : 0u;
}

// caller-supplied rays + seeds (pbrgpu_radiance / pbrgpu_shade hooks): slot i = path i = pixel i, no regeneration
__global__ void InitPathsFromRaysKernel(WaveState w, const float4* rays, const uint64_t* seeds, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < kCounterCount) w.counters[p] = (p == kNumActive0) ? n : 0u;
  if (p >= n) return;
  StSlot(w, p, kRayO, rays[2 * p]);
  StSlot(w, p, kRayD, rays[2 * p + 1]);
  Pcg32 rng;
  pcg32_srandom(&rng, seeds[2 * p], seeds[2 * p + 1]);
  StSlot(w, p, kThr, make_float4(1.f, 1.f, 1.f, 0.f));
  StSlot(w, p, kRad, make_float4(0.f, 0.f, 0.f, __uint_as_float(0u)));
  StSlot(w, p, kRng, PackRng(rng));
  StSlot(w, p, kPix, make_float4(__uint_as_float(p), 0.f, 0.f, 0.f));
  w.q_active[0][p] = p;
}

// ------------------------------------------------------------------------------------------------ closest hit
// Scene::TraceFirstHit1 for every active slot; routes it by the material kind of what it hit, retires it on a miss.
// Runs in the warp traversal engine (device/trav_engine.cuh): lanes are refilled from q_active while others traverse.
struct ClosestClient {
  const SceneView& s;
  const WaveState& w;
  const FrameParams& frame;      // frame mode: retire + regenerate the slots of q_done[cur] in here
  const bool regen;
  const bool sort_materials;     // route diffuse-only materials to their own shading queue
  const uint32_t* __restrict__ queue;
  const uint32_t* __restrict__ done;
  uint32_t n_active, n, next_parity;
  unsigned long long sample_base;
  uint32_t p = 0;
  bool has_result = false;
  uint32_t rays = 0, retired = 0;

  __device__ __forceinline__ ClosestClient(const SceneView& s_, const WaveState& w_, uint32_t cur_parity,
                                           const FrameParams& frame_, bool regen_, bool sort_)
      : s(s_), w(w_), frame(frame_), regen(regen_), sort_materials(sort_), queue(w_.q_active[cur_parity]), done(w_.q_done[cur_parity]),
        n_active(w_.counters[kNumActive0 + cur_parity]),
        n(w_.counters[kNumActive0 + cur_parity] + (regen_ ? w_.counters[kNumDone0 + cur_parity] : 0u)),
        next_parity(cur_parity ^ 1u),
        sample_base(regen_ ? w_.stats[kStatSampleBase] : 0ull) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || !exhausted; }

  // work item -> slot: first the slots that shading sent on, then (frame mode) last iteration's finished slots
  __device__ __forceinline__ uint32_t SlotOf(uint32_t item) const {
    return item < n_active ? queue[item] : done[item - n_active];
  }

  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    const bool finished = !t.active && has_result;
    const bool need = !exhausted && !t.active;
    // ---- (1) finished rays: miss -> retire, else the shading queue of the material's class
    int kind = -1;   // -1 nothing, 0 miss, 1 general surface queue, 2 hair queue, 3 diffuse-only queue
    const HitT hit = t.hit;
    const uint32_t done_p = p;
    if (finished) {
      has_result = false;
      kind = 0;
      if (hit.prim != kInvalid) {
        kind = 1;
        uint32_t mat = kInvalid;   // triangles carry their material id in the spare lane of e1 (scene_host.cc: Commit)
        if (sort_materials || s.num_hair_materials) {   // scene-uniform
          if (hit.prim & kCurveFlag) mat = s.curve_ids[s.curve_prim[hit.prim & ~kCurveFlag]].w;
          else mat = __float_as_uint(s.tri_data[hit.prim * 3 + 1].w);
        }
        if (mat < s.num_materials) {
          const uint32_t cls = s.material_class[mat];
          kind = (cls == kClassHair) ? 2 : ((cls == kClassDiffuse && sort_materials) ? 3 : 1);
        }
      }
    }
    // ---- (2) the four output queues and the work fetch: one atomic instruction
    const Append5 app = Append5Issue(&w.counters[kNumSurface], &w.counters[kNumHair],
                                     &w.counters[kNumDone0 + next_parity], &w.counters[kFetchTrace],
                                     &w.counters[kNumDiffuse], kind == 1, kind == 2, kind == 0, need, kind == 3);
    const uint32_t i_surf = Append5Index(app, 0), i_hair = Append5Index(app, 1), i_done = Append5Index(app, 2),
                   item = Append5Index(app, 3), i_diff = Append5Index(app, 4);
    // ---- (3) new work: the loads are issued here and consumed after the finished rays have been written out
    const bool take = need && item < n_active;          // a path that continues
    const bool renew = need && item >= n_active && item < n;   // a finished slot: retire it, start the next sample
    float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = ro, rrad = ro;
    uint32_t np = 0, rpix = kNoPixel;
    if (take || renew) np = SlotOf(item);
    if (take) {
      ro = LdSlot(w, np, kRayO);
      rd = LdSlot(w, np, kRayD);
    } else if (renew) {
      rpix = __float_as_uint(LdSlot(w, np, kPix).x);
      rrad = LdSlot(w, np, kRad);
    }
    // ---- (4) write out the finished rays
    if (kind >= 0) {
      StSlot(w, done_p, kHit, make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim)));
      StSlot(w, done_p, kHitPad, make_float4(0.f, 0.f, 0.f, 0.f));   // completes the sector: no read-modify-write
      if (kind == 1) w.q_surface[i_surf] = done_p;
      else if (kind == 2) w.q_hair[i_hair] = done_p;
      else if (kind == 3) w.q_diffuse[i_diff] = done_p;
      else w.q_done[next_parity][i_done] = done_p;
    }
    // ---- (5) start the new rays
    if (take) {
      p = np;
      RayT ray;
      ray.o = vec3(ro.x, ro.y, ro.z); ray.tmin = ro.w;
      ray.d = vec3(rd.x, rd.y, rd.z); ray.tmax = rd.w;
      TravBegin(s, ray, t);
      has_result = true;
      ++rays;
    } else if (renew) {
      // render.cc:175-183 for the path that ended in this slot, then the next camera sample in the same slot: the
      // camera ray never makes a round trip through memory before its first traversal
      if (rpix != kNoPixel) {
        atomicAdd(&frame.rgba[rpix], make_float4(rrad.x, rrad.y, rrad.z, 1.0f));
        ++retired;
      }
      const unsigned long long id = sample_base + (item - n_active);
      if (id < frame.total_samples) {
        p = np;
        const RayT ray = StartCameraPath(w, frame, p, id);
        TravBegin(s, ray, t);
        has_result = true;
        ++rays;
      } else {
        StSlot(w, np, kPix, make_float4(__uint_as_float(kNoPixel), 0.f, 0.f, 0.f));
      }
    }
    return need && item >= n;
  }
  __device__ __forceinline__ void End(const Trav&) {
    WarpTally(&w.stats[kStatClosest], rays);
    if (regen) WarpTally(&w.stats[kStatRetired], retired);
  }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) TraceClosestKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                          uint32_t refill_min_idle, uint32_t prim_min_lanes,
                                                          FrameParams frame, uint32_t regenerate,
                                                          uint32_t sort_materials) {
  ClosestClient client(s, w, cur_parity, frame, regenerate != 0u, sort_materials != 0u);
  TravEngine<false, HAS_CURVES, false>(s, client, refill_min_idle, prim_min_lanes);
}

// ------------------------------------------------------------------------------------------------ shading
struct ShadeFlags {
  uint32_t skip_emission_and_roulette;   // pbrgpu_shade hook: call Shader() only
};

// what a shading kernel holds of its path between load and commit
struct PathRegs {
  RayT ray;
  HitT hit;
  vec3 throughput, L;
  float pdf_prev;
  uint32_t depth, pixel;
  Pcg32 rng;
};

// the seven records of a slot line that a shading kernel reads (kHitPad is never read)
__device__ __forceinline__ PathRegs PathFromRecords(const float4& o, const float4& d, const float4& t4, const float4& r4,
                                                    const float4& h4, const float4& g4, const float4& x4) {
  PathRegs r;
  r.ray.o = vec3(o.x, o.y, o.z); r.ray.tmin = o.w;
  r.ray.d = vec3(d.x, d.y, d.z); r.ray.tmax = d.w;
  r.hit.t = h4.x; r.hit.u = h4.y; r.hit.v = h4.z; r.hit.prim = __float_as_uint(h4.w);
  r.throughput = vec3(t4.x, t4.y, t4.z);
  r.pdf_prev = t4.w;
  r.L = vec3(r4.x, r4.y, r4.z);
  r.depth = __float_as_uint(r4.w);
  r.rng.state = (uint64_t(__float_as_uint(g4.y)) << 32) | __float_as_uint(g4.x);
  r.rng.inc = (uint64_t(__float_as_uint(g4.w)) << 32) | __float_as_uint(g4.z);
  r.pixel = __float_as_uint(x4.x);
  return r;
}
__device__ __forceinline__ PathRegs LoadPath(const WaveState& w, uint32_t p) {
  return PathFromRecords(LdSlot(w, p, kRayO), LdSlot(w, p, kRayD), LdSlot(w, p, kThr), LdSlot(w, p, kRad),
                         LdSlot(w, p, kHit), LdSlot(w, p, kRng), LdSlot(w, p, kPix));
}

// after a vertex: sectors 0, 1 and 3 of the line are rewritten whole (render.cc:79-86)
__device__ __forceinline__ void CommitVertex(const WaveState& w, uint32_t p, const VertexResult& vr,
                                             const vec3& throughput, const vec3& L, uint32_t depth, const Pcg32& rng,
                                             uint32_t pixel) {
  const vec3 new_thr = vr.throughput * throughput;                     // render.cc:80
  StSlot(w, p, kRayO, make_float4(vr.P.x, vr.P.y, vr.P.z, 1e-3f));     // render.cc:83-86
  StSlot(w, p, kRayD, make_float4(vr.wi.x, vr.wi.y, vr.wi.z, kInf));
  StSlot(w, p, kThr, make_float4(new_thr.x, new_thr.y, new_thr.z, vr.pdf));
  StSlot(w, p, kRad, make_float4(L.x, L.y, L.z, __uint_as_float(depth + 1u)));
  StSlot(w, p, kRng, PackRng(rng));
  StSlot(w, p, kPix, make_float4(__uint_as_float(pixel), 0.f, 0.f, 0.f));
}

// the path ends here: only its radiance (and pixel) are read again
__device__ __forceinline__ void CommitEnd(const WaveState& w, uint32_t p, const vec3& L, uint32_t depth) {
  StSlot(w, p, kThr, make_float4(0.f, 0.f, 0.f, 0.f));
  StSlot(w, p, kRad, make_float4(L.x, L.y, L.z, __uint_as_float(depth)));
}

__device__ __forceinline__ void ParkWalk(const WaveState& w, uint32_t p, const SssWalkState& k) {
  StWalk(w, p, kWalkA, make_float4(k.sigma_t.x, k.sigma_t.y, k.sigma_t.z, k.throughput.x));
  StWalk(w, p, kWalkB, make_float4(k.sigma_s.x, k.sigma_s.y, k.sigma_s.z, k.throughput.y));
  StWalk(w, p, kWalkC, make_float4(k.ray.o.x, k.ray.o.y, k.ray.o.z, k.throughput.z));
  StWalk(w, p, kWalkD, make_float4(k.ray.d.x, k.ray.d.y, k.ray.d.z, k.ray.tmin));
  StWalk(w, p, kWalkN, make_float4(__uint_as_float(k.bounce), 0.f, 0.f, 0.f));
}
__device__ __forceinline__ void ResumeWalk(const WaveState& w, uint32_t p, SssWalkState* k) {
  const float4 a = LdWalk(w, p, kWalkA), b = LdWalk(w, p, kWalkB), c = LdWalk(w, p, kWalkC), d = LdWalk(w, p, kWalkD);
  k->sigma_t = vec3(a.x, a.y, a.z);
  k->sigma_s = vec3(b.x, b.y, b.z);
  k->throughput = vec3(a.w, b.w, c.w);
  k->ray.o = vec3(c.x, c.y, c.z);
  k->ray.d = vec3(d.x, d.y, d.z);
  k->ray.tmin = d.w;
  k->ray.tmax = kInf;
  k->bounce = __float_as_uint(LdWalk(w, p, kWalkN).x);
}

// emission + MIS, roulette, material dispatch, Principled vertex.  When the vertex selects the random-walk closure
// the walk is set up here (entry direction + coefficients, random-walk-sss.h:227-279) and parked for sss_walk.
// Launch shapes: the general kernel needs ~110-128 registers, so one 512-thread block per SM; the diffuse-only kernel is
// 7x smaller (2.1k vs 15.7k SASS instructions) and fits three 256-thread blocks per SM.
constexpr int kDiffuseBlock = 256, kDiffuseBlocksPerSm = 3;
// One Principled vertex (render.cc:33-86 + shader.cc:8-34) of path slot p whose state is in r; called by every lane of
// a warp (lanes without an item pass valid = false): the queue appends are warp collectives.
template <bool DIFFUSE_ONLY>
__device__ __forceinline__ void ShadeSurfaceItem(const SceneView& s, const WaveState& w, uint32_t next_parity,
                                                 const ShadeFlags& flags, bool valid, uint32_t p, PathRegs& r) {
  bool to_sss = false, to_next = false, to_done = false;
  ShadowRequest req;
  req.active = false;
  vec3 throughput(0.f);
  if (valid) {
    throughput = r.throughput;
    vec3 L = r.L;
    const Surface si = MakeSurface(s, r.ray, r.hit);
    bool alive = true;
    if (!flags.skip_emission_and_roulette)
      alive = EmissionAndRoulette(s, r.ray, r.hit, si, r.depth, r.pdf_prev, &r.rng, &L, &throughput);
    if (!alive) {
      CommitEnd(w, p, L, r.depth);
      to_done = true;
    } else {
      const int kind = DIFFUSE_ONLY ? 1 : MaterialKind(s, si);
      VertexResult vr;
      const vec3 wo = -r.ray.d;
      Frame fr;
      PrincipledBsdf bsdf;
      bool sss = false;
      if (kind == 1) sss = PrincipledVertexT<DIFFUSE_ONLY>(s, si, wo, &r.rng, &vr, &fr, &bsdf);
      else AbsorbVertex(wo, si.P, &vr);   // no material (shader.cc:11-17)
      req = vr.shadow[0];
      if (!DIFFUSE_ONLY && sss) {
        SssWalkState walk;
        if (SssBegin(si, fr, bsdf, &r.rng, &walk)) {
          // the walk runs in its own kernel: park it, and the path with the post-roulette throughput
          ParkWalk(w, p, walk);
          StSlot(w, p, kThr, make_float4(throughput.x, throughput.y, throughput.z, r.pdf_prev));
          StSlot(w, p, kRad, make_float4(L.x, L.y, L.z, __uint_as_float(r.depth)));
          StSlot(w, p, kRng, PackRng(r.rng));
          StSlot(w, p, kPix, make_float4(__uint_as_float(r.pixel), 0.f, 0.f, 0.f));
          to_sss = true;
        } else {   // walk rejected: throughput 0, the path ends (cycles-principled-shader.cc:217-220)
          CommitEnd(w, p, L, r.depth + 1u);
          to_done = true;
        }
      } else {
        CommitVertex(w, p, vr, throughput, L, r.depth, r.rng, r.pixel);
        to_next = !IsBlack(vr.throughput * throughput);               // render.cc:31
        to_done = !to_next;
      }
    }
  }
  PushShadow(w, req, throughput, p);
  if (!DIFFUSE_ONLY) {
    const uint32_t a = WarpAppend(&w.counters[kNumSss], to_sss);
    if (to_sss) w.q_sss[a] = p;
  }
  RouteSlot(w, next_parity, p, to_next, to_done);
}

template <bool DIFFUSE_ONLY>
__global__ void __launch_bounds__(DIFFUSE_ONLY ? kDiffuseBlock : kShadeBlock, DIFFUSE_ONLY ? kDiffuseBlocksPerSm : 1)
ShadeSurfaceKernel(SceneView s, WaveState w, uint32_t next_parity,
                                                          ShadeFlags flags) {
  const uint32_t n = w.counters[DIFFUSE_ONLY ? kNumDiffuse : kNumSurface];
  const uint32_t* __restrict__ queue = DIFFUSE_ONLY ? w.q_diffuse : w.q_surface;
  uint32_t shaded = 0;
  for (;;) {
    const uint32_t slot = BlockFetch(&w.counters[DIFFUSE_ONLY ? kFetchDiffuse : kFetchSurface]);
    if (slot - threadIdx.x >= n) break;   // block-uniform
    const bool valid = slot < n;
    shaded += valid ? 1u : 0u;
    uint32_t p = 0;
    PathRegs r;
    if (valid) {
      p = queue[slot];
      r = LoadPath(w, p);
    }
    ShadeSurfaceItem<DIFFUSE_ONLY>(s, w, next_parity, flags, valid, p, r);
  }
  WarpTally(&w.stats[kStatVertices], shaded);
}

// The diffuse-only vertex is short (2.1 k SASS instructions) and its kernel was bound by the latency of one dependent
// chain per thread — work fetch (atomic) -> queue entry -> slot line (a random 128-byte line in HBM) -> primitive ->
// normals -> material -> light tables — with 15 % of the issue slots busy and DRAM at 16 % (profiles/r1_final_ncu.md).
// This version takes the first three links off the chain: a block reserves kPipeBatches batches with ONE atomic, reads
// the queue entries (dense, coalesced) one batch ahead, and copies the slot line of batch b+1 into shared memory with
// cp.async (L2 evict-first, no registers held) while batch b is shaded.  A thread reads back only the row it copied
// itself, so the pipeline needs no block barrier.  Rows are swizzled by (record ^ row) so that the 128-bit shared
// loads of a quarter warp fall into different banks.
constexpr int kPipeBatches = 8;
__device__ __forceinline__ void CpAsync16(void* smem_dst, const void* gmem_src, uint64_t policy) {
  const uint32_t dst = uint32_t(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void CpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kDiffuseBlock, kDiffuseBlocksPerSm)
ShadeDiffusePipelinedKernel(SceneView s, WaveState w, uint32_t next_parity, ShadeFlags flags) {
  extern __shared__ float4 stage[];   // [2][blockDim.x][8 records]
  __shared__ uint32_t s_base;
  const uint32_t n = w.counters[kNumDiffuse];
  const uint32_t* __restrict__ queue = w.q_diffuse;
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  float4* row0 = stage + size_t(tid) * 8;
  float4* row1 = stage + size_t(nthr + tid) * 8;
  const uint32_t sw = tid & 7u;
  auto issue = [&](float4* row, uint32_t p) {
    const float4* src = w.slot + size_t(p) * kSlotStride;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k)
      if (k != uint32_t(kHitPad)) CpAsync16(row + (k ^ sw), src + k, policy);
  };
  uint32_t shaded = 0;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_base = atomicAdd(&w.counters[kFetchDiffuse], nthr * uint32_t(kPipeBatches));
    __syncthreads();
    const uint32_t base = s_base;
    if (base >= n) break;   // block-uniform
    uint32_t item = base + tid;
    uint32_t p_cur = item < n ? queue[item] : kInvalid;
    if (p_cur != kInvalid) issue(row0, p_cur);
    CpAsyncCommit();
    uint32_t p_next = (item + nthr < n) ? queue[item + nthr] : kInvalid;
#pragma unroll 1
    for (int b = 0; b < kPipeBatches; ++b) {
      if (base + uint32_t(b) * nthr >= n) break;   // block-uniform: nothing left in this chunk
      float4* cur = (b & 1) ? row1 : row0;
      float4* nxt = (b & 1) ? row0 : row1;
      uint32_t p_next2 = kInvalid;
      if (b + 1 < kPipeBatches) {
        if (p_next != kInvalid) issue(nxt, p_next);
        const uint32_t item2 = base + uint32_t(b + 2) * nthr + tid;
        if (b + 2 < kPipeBatches && item2 < n) p_next2 = queue[item2];
      }
      CpAsyncCommit();
      CpAsyncWait<1>();   // everything but the group just committed has landed: batch b is in `cur`
      const bool valid = p_cur != kInvalid;
      shaded += valid ? 1u : 0u;
      PathRegs r;
      if (valid)
        r = PathFromRecords(cur[kRayO ^ sw], cur[kRayD ^ sw], cur[kThr ^ sw], cur[kRad ^ sw], cur[kHit ^ sw],
                            cur[kRng ^ sw], cur[kPix ^ sw]);
      ShadeSurfaceItem<true>(s, w, next_parity, flags, valid, valid ? p_cur : 0u, r);
      p_cur = p_next;
      p_next = p_next2;
    }
    CpAsyncWait<0>();
  }
  WarpTally(&w.stats[kStatVertices], shaded);
}

__global__ void __launch_bounds__(kShadeBlock) ShadeHairKernel(SceneView s, WaveState w, uint32_t next_parity,
                                                       ShadeFlags flags) {
  const uint32_t n = w.counters[kNumHair];
  uint32_t shaded = 0;
  for (;;) {
    const uint32_t slot = BlockFetch(&w.counters[kFetchHair]);
    if (slot - threadIdx.x >= n) break;   // block-uniform
    const bool valid = slot < n;
    shaded += valid ? 1u : 0u;
    uint32_t p = 0;
    bool to_next = false, to_done = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    if (valid) {
      p = w.q_hair[slot];
      PathRegs r = LoadPath(w, p);
      throughput = r.throughput;
      vec3 L = r.L;
      const Surface si = MakeSurface(s, r.ray, r.hit);
      bool alive = true;
      if (!flags.skip_emission_and_roulette)
        alive = EmissionAndRoulette(s, r.ray, r.hit, si, r.depth, r.pdf_prev, &r.rng, &L, &throughput);
      if (!alive) {
        CommitEnd(w, p, L, r.depth);
        to_done = true;
      } else {
        VertexResult vr;
        HairVertex(s, si, -r.ray.d, &r.rng, &vr);
        req = vr.shadow[0];
        CommitVertex(w, p, vr, throughput, L, r.depth, r.rng, r.pixel);
        to_next = !IsBlack(vr.throughput * throughput);
        to_done = !to_next;
      }
    }
    PushShadow(w, req, throughput, p);
    RouteSlot(w, next_parity, p, to_next, to_done);
  }
  WarpTally(&w.stats[kStatVertices], shaded);
}

// ------------------------------------------------------------------------------------------------ random-walk SSS
// RandomWalkSubsurface (random-walk-sss.h:281-383), the bounce loop only.  Walk lengths are wildly uneven (1 .. 8192
// bounces; Lucy's red channel has albedo ~1, so neither absorption nor roulette ends a walk early) and every bounce
// is a dependent short closest-hit query, hence:
//   * the queries run in the warp traversal engine; a lane whose segment ends waits until `refill_min_idle` lanes are
//     in that state, then they all do the scatter step (transmittance, roulette, new direction and distance)
//     converged and go back to traversing — no lane waits for the longest walk of its warp;
//   * a walk gets at most `max_bounces` bounces per launch; if it is still inside the medium its state is parked in
//     its walk line and the slot goes to q_walk[next]: the next iteration resumes it first.  A launch therefore
//     never outlives its queue by more than max_bounces bounces, and because the pool is kept full by
//     regeneration, long walks cost slots, not idle SMs;
//   * entering the medium is part of shade_surface, leaving it (exit vertex: NEE + diffuse bounce,
//     cycles-principled-shader.cc:187-216) is sss_exit: both are rare per bounce and ran at 2-3 lanes per warp when
//     they were inlined here (profiles/r1b_summary.md).
struct SssClient {
  const SceneView& s;
  const WaveState& w;
  uint32_t cur_parity, next_parity, n_resume, n, max_bounces;
  uint32_t p = 0, budget = 0, pixel = 0;
  bool has_walk = false;
  bool skipped_seg = false;   // the current segment was answered by the clearance grid, not traced
  Pcg32 rng;
  SssWalkState walk;    // walk.ray is rebuilt from the traversal state after every segment
  vec3 cpdf;            // channel probabilities of the segment in flight (SssPrepareSegment -> SssFinishSegment)
  uint32_t rays = 0, skipped = 0;

  __device__ __forceinline__ SssClient(const SceneView& s_, const WaveState& w_, uint32_t cur, uint32_t max_b)
      : s(s_), w(w_), cur_parity(cur), next_parity(cur ^ 1u), n_resume(w_.counters[kNumWalk0 + cur]),
        n(w_.counters[kNumWalk0 + cur] + w_.counters[kNumSss]), max_bounces(max_b) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_walk || !exhausted; }

  __device__ __forceinline__ void StartSegment(Trav& t) {
    SssPrepareSegment(&rng, &walk, &cpdf);
    TravBegin(s, walk.ray, t);
    // Segments that end far from any surface: the clearance grid answers those ("no hit") without a traversal; the
    // lane then waits for the next converged section like any lane whose query has finished.  (Finishing chains of
    // such segments inline was measured and is slower: the converged section serialises on the longest chain.)
    skipped_seg = SegmentIsClear(s, walk.ray.o, walk.ray.d, walk.ray.tmax * 1.001f);
    if (skipped_seg) t.active = false;
  }

  __device__ __forceinline__ uint32_t SlotOf(uint32_t item) const {
    return item < n_resume ? w.q_walk[cur_parity][item] : w.q_sss[item - n_resume];
  }

  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    // ---- (1) walks whose segment query finished: scatter / exit / absorb
    bool to_exit = false, to_done = false, to_park = false, go_on = false;
    const uint32_t routed_p = p;
    const HitT hit = t.hit;
    if (!t.active && has_walk) {
      walk.ray.o = t.O; walk.ray.d = t.D; walk.ray.tmin = t.tmin;   // tmax untouched: the scatter distance
      const bool is_hit = hit.prim != kInvalid;
      const SssStep st = SssFinishSegment(is_hit, hit.t, &rng, &walk, cpdf);
      if (skipped_seg) ++skipped; else ++rays;
      --budget;
      if (st == kSssHit) to_exit = true;
      else if (st == kSssAbsorbed) to_done = true;   // throughput 0: the path ends with the radiance it already holds
      else if (budget == 0u) to_park = true;
      else go_on = true;
      has_walk = go_on;
    }
    // ---- (2) the three output queues and the work fetch (lanes without a walk): one atomic instruction
    const bool need = !exhausted && !t.active && !has_walk;
    const Append5 app = Append5Issue(&w.counters[kNumExit], &w.counters[kNumDone0 + next_parity],
                                     &w.counters[kNumWalk0 + next_parity], &w.counters[kFetchSss], nullptr, to_exit,
                                     to_done, to_park, need, false);
    const uint32_t i_exit = Append5Index(app, 0), i_done = Append5Index(app, 1), i_park = Append5Index(app, 2),
                   item = Append5Index(app, 3);
    // ---- (3) write out the walks that stopped
    if (to_exit) {
      // exit record for sss_exit: the segment ray, its hit and the walk throughput
      StWalk(w, routed_p, kWalkA, make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim)));
      StWalk(w, routed_p, kWalkB, make_float4(walk.throughput.x, walk.throughput.y, walk.throughput.z, 0.f));
      StWalk(w, routed_p, kWalkC, make_float4(walk.ray.o.x, walk.ray.o.y, walk.ray.o.z, walk.ray.tmin));
      StWalk(w, routed_p, kWalkD, make_float4(walk.ray.d.x, walk.ray.d.y, walk.ray.d.z, walk.ray.tmax));
      w.q_exit[i_exit] = routed_p;
    } else if (to_park) {
      ParkWalk(w, routed_p, walk);
      w.q_walk[next_parity][i_park] = routed_p;
    } else if (to_done) {
      w.q_done[next_parity][i_done] = routed_p;
    }
    if (to_exit || to_park) {
      StSlot(w, routed_p, kRng, PackRng(rng));
      StSlot(w, routed_p, kPix, make_float4(__uint_as_float(pixel), 0.f, 0.f, 0.f));
    }
    // ---- (4) new walks: parked ones first, then this iteration's new ones (loaded straight into the walk registers
    // that (3) has just finished with)
    const bool take = need && item < n;
    if (take) {
      p = SlotOf(item);
      rng = LoadRng(w, p);
      pixel = __float_as_uint(LdSlot(w, p, kPix).x);
      budget = max_bounces;
      ResumeWalk(w, p, &walk);
      has_walk = true;
    }
    // ---- (5) next segment: of the walk that goes on, or of the one just taken
    if (take || go_on) StartSegment(t);
    return need && item >= n;
  }
  __device__ __forceinline__ void End(const Trav&) {
    WarpTally(&w.stats[kStatSss], rays);
    WarpTally(&w.stats[kStatSssSkipped], skipped);
  }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128, 5) SssWalkKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                     uint32_t max_bounces, uint32_t refill_min_idle,
                                                     uint32_t prim_min_lanes) {
  SssClient client(s, w, cur_parity, max_bounces);
  TravEngine<false, HAS_CURVES, false>(s, client, refill_min_idle, prim_min_lanes);
}

// The exit vertex of every walk that left the medium this iteration (random-walk-sss.h:385-404 +
// cycles-principled-shader.cc:187-216): same-instance / back-face acceptance, NEE at the exit point, diffuse bounce.
__global__ void __launch_bounds__(kShadeBlock) SssExitKernel(SceneView s, WaveState w, uint32_t next_parity) {
  const uint32_t n = w.counters[kNumExit];
  uint32_t shaded = 0;
  for (;;) {
    const uint32_t slot = BlockFetch(&w.counters[kFetchExit]);
    if (slot - threadIdx.x >= n) break;   // block-uniform
    const bool valid = slot < n;
    shaded += valid ? 1u : 0u;
    uint32_t p = 0;
    bool to_next = false, to_done = false;
    ShadowRequest req;
    req.active = false;
    vec3 throughput(0.f);
    if (valid) {
      p = w.q_exit[slot];
      PathRegs r = LoadPath(w, p);   // ray + hit are still those of the ENTRY vertex
      throughput = r.throughput;
      const Surface entry_si = MakeSurface(s, r.ray, r.hit);
      const Frame entry_frame = PrincipledFrame(entry_si);
      const float4 a = LdWalk(w, p, kWalkA), b = LdWalk(w, p, kWalkB), c = LdWalk(w, p, kWalkC),
                   d = LdWalk(w, p, kWalkD);
      HitT hit;
      hit.t = a.x; hit.u = a.y; hit.v = a.z; hit.prim = __float_as_uint(a.w);
      SssWalkState walk;
      walk.throughput = vec3(b.x, b.y, b.z);
      walk.ray.o = vec3(c.x, c.y, c.z); walk.ray.tmin = c.w;
      walk.ray.d = vec3(d.x, d.y, d.z); walk.ray.tmax = d.w;
      VertexResult vr;
      vr.P = entry_si.P;
      vr.shadow[1].active = false;
      SssFinish(s, entry_si, entry_frame, walk, hit, &r.rng, &vr);
      req = vr.shadow[1];
      CommitVertex(w, p, vr, throughput, r.L, r.depth, r.rng, r.pixel);
      to_next = !IsBlack(vr.throughput * throughput);
      to_done = !to_next;
    }
    PushShadow(w, req, throughput, p);
    RouteSlot(w, next_parity, p, to_next, to_done);
  }
  WarpTally(&w.stats[kStatVertices], shaded);
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
  uint32_t rays = 0;

  __device__ __forceinline__ ShadowClient(const SceneView& s_, const WaveState& w_)
      : s(s_), w(w_), n(w_.counters[kNumShadow]) {}
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || !exhausted; }
  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    const unsigned lane = threadIdx.x & 31u;
    const bool need = !exhausted && !t.active;
    const unsigned m_need = __ballot_sync(0xffffffffu, need);
    uint32_t fetch_base = 0;
    if (lane == 0u && m_need) fetch_base = atomicAdd(&w.counters[kFetchShadow], uint32_t(__popc(m_need)));
    if (!t.active && has_result) {   // overlaps the fetch round trip
      has_result = false;
      if (t.hit.prim == kInvalid) {
        float* dst = reinterpret_cast<float*>(&w.slot[size_t(__float_as_uint(c.w)) * kSlotStride + kRad]);
        atomicAdd(dst + 0, c.x);
        atomicAdd(dst + 1, c.y);
        atomicAdd(dst + 2, c.z);
      }
    }
    fetch_base = __shfl_sync(0xffffffffu, fetch_base, 0);
    const uint32_t slot = fetch_base + uint32_t(__popc(m_need & ((1u << lane) - 1u)));
    if (need && slot < n) {
      const float4 o = __ldcs(&w.sh_o[slot]), d = __ldcs(&w.sh_d[slot]);
      c = __ldcs(&w.sh_c[slot]);
      RayT ray;
      ray.o = vec3(o.x, o.y, o.z); ray.tmin = o.w;
      ray.d = vec3(d.x, d.y, d.z); ray.tmax = d.w;
      TravBegin(s, ray, t);
      has_result = true;
      ++rays;
    }
    return need && slot >= n;
  }
  __device__ __forceinline__ void End(const Trav&) { WarpTally(&w.stats[kStatShadow], rays); }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) TraceAnyKernel(SceneView s, WaveState w, uint32_t refill_min_idle,
                                                      uint32_t prim_min_lanes) {
  ShadowClient client(s, w);
  TravEngine<true, HAS_CURVES, false>(s, client, refill_min_idle, prim_min_lanes);
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
                                                        uint32_t refill_min_idle, uint32_t prim_min_lanes) {
  BatchClient client(s, rays, n, hits_tuv, hits_ids, hits_ng, nullptr, fetch, STATS ? stats : nullptr);
  TravEngine<false, HAS_CURVES, STATS>(s, client, refill_min_idle, prim_min_lanes);
}

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) OccludedBatchKernel(SceneView s, const float4* __restrict__ rays, uint64_t n,
                                                           uint8_t* out, uint32_t* fetch, uint32_t refill_min_idle,
                                                           uint32_t prim_min_lanes) {
  BatchClient client(s, rays, n, nullptr, nullptr, nullptr, out, fetch, nullptr);
  TravEngine<true, HAS_CURVES, false>(s, client, refill_min_idle, prim_min_lanes);
}

}  // namespace pbr
