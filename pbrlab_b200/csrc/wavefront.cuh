// Wavefront path tracer kernels (sm_100a): STREAMING path state.
//
// Render() keeps up to N paths in flight for the whole frame.  A path has no home: its state travels with it
// through dense structure-of-arrays queues that every kernel reads and writes in (nearly) sequential order, as
// north_star describes ("SoA ray/hit/path-state queues read with coalesced 128-bit loads and compacted with warp
// ballot/prefix-sum").  Round 1 kept one 128-byte line per path SLOT and visited the slots through index queues, i.e.
// in random order; measured on the B200 that design ran every kernel at the RANDOM-access throughput of HBM3e
// (~1-1.5 TB/s: 7.2 GB per closest-hit launch, 4.4 GB per diffuse-shading launch at 13-15 % issue utilisation —
// profiles/r1_final_ncu.md; 80-byte random gathers over 4 GB reach 0.94 TB/s, a seventh of the streaming figure),
// and hiding the LATENCY of those lines with a cp.async pipeline changed nothing (profiles/r2a_*).  Now:
//
//   S[2]   path state, ping-pong by iteration parity, one array per field (ray_o, ray_d, thr, rad, hit, rng, pix):
//          iteration `it` traces and shades the paths S[cur][0 .. n_active + n_new); shading APPENDS the paths that go
//          on to S[next] (warp ballot + one atomic per warp), so S[next] is dense and in shading order
//   W[2]   random walks (hot: the walk state + rng; cold: the entry vertex and the path state it will resume with),
//          ping-pong: the walk kernel steps W[cur][0 .. n_walk) and appends the walks that are still inside the
//          medium to W[next]; shading appends NEW walks to W[next] as well (they start one iteration later)
//   E      exit records of the walks that left the medium this iteration (the cold part is read from W[cur])
//   D[2]   (radiance, pixel) of the paths that ended this iteration; retired into the frame at the start of the next
//   q_surface / q_diffuse / q_hair   positions in S[cur] of the paths whose closest hit has a material of that class:
//          the only index queues left; closest-hit rays finish roughly in fetch order, so these positions are nearly
//          ascending and the shading kernels' gathers walk S[cur] front to back
//   shadow queue (ray, contribution, target) as three dense arrays; the target says where an unoccluded contribution
//          is added: rad of S[next][j], D[next][j] or W[next][j] — all written by this iteration and stable until
//          the next one consumes them
//
// One iteration:   begin -> retire D[cur] -> trace_closest (+ camera rays for the free capacity, misses retired in
//                  place) -> shade_surface, shade_diffuse, shade_hair -> sss_walk -> sss_exit -> trace_any
//                  (sss_walk + sss_exit on their own stream beside trace_closest + shading: they only share
//                  atomically appended output streams; joined before trace_any)
// The end of a frame: camera samples start longest paths first (FrameParams::order), launches with
// fewer items than lanes spread them over all warps (trav_engine.cuh: LanesFor), and the last few thousand paths
// and walks are run to their end by FinishPathsKernel instead of one iteration per vertex.
// Every kernel is persistent — a grid that is a fixed multiple of the SM count, warps pulling work with an atomic
// counter — and reads its queue length from device memory, so nothing but one small counter block crosses PCIe per
// iteration.  The three ray kernels run in the warp traversal engine (device/trav_engine.cuh), which refills finished
// lanes while the rest of the warp keeps traversing.  The pool stays full until the frame runs out of samples: the
// number of iterations is (total rays) / N instead of (longest path).
//
// Replaces the per-pixel loops of the reference (src/render.cc:24-90,125-190); the per-vertex functions and their
// citations are in device/shade.cuh.
#pragma once
#include <cuda_runtime.h>

#include "device/shade.cuh"
#include "device/trav_engine.cuh"
#include "job_split.h"
#include "kat.cuh"

namespace pbr {

enum Counter {
  kNumActive0 = 0, kNumActive1,   // paths in S[parity]
  kNumWalk0, kNumWalk1,           // walks in W[parity]
  kNumDone0, kNumDone1,           // records in D[parity]
  kNumNew,                        // camera samples started by this iteration's closest-hit kernel
  kNumSurface, kNumDiffuse, kNumHair, kNumExit, kNumShadow,
  kFetchRetire, kFetchTrace, kFetchSurface, kFetchDiffuse, kFetchHair, kFetchWalk, kFetchExit, kFetchShadow,
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

// fields of the three record streams (one float4 array of `capacity` entries per field)
enum StateField { kRayO = 0, kRayD, kThr, kRad, kHit, kRng, kPix, kStateFields };
enum WalkField {
  kWalkA = 0, kWalkB, kWalkC, kWalkD, kWalkN, kWalkRng,                 // hot: stepped by the walk kernel
  kWalkRayO, kWalkRayD, kWalkHit, kWalkThr, kWalkRad,                   // cold: entry vertex + the path it resumes
  kWalkFields
};
enum ExitField { kExHit = 0, kExThr, kExO, kExD, kExRng, kExitFields };

// where an unoccluded NEE contribution is added (ShadowRequest target): tag << 30 | position
constexpr uint32_t kTargetState = 0u, kTargetDone = 1u, kTargetWalk = 2u, kTargetNone = 3u;

struct WaveState {
  float4* state[2];              // kStateFields x capacity
  float4* walk[2];               // kWalkFields x capacity
  float4* exit_rec;              // kExitFields x capacity
  float4* done[2];               // capacity: (radiance rgb, pixel)
  uint32_t* q_surface;
  uint32_t* q_diffuse;
  uint32_t* q_hair;
  float4* sh_o;
  float4* sh_d;
  float4* sh_c;
  uint32_t* counters;            // kCounterCount
  unsigned long long* stats;     // kStatCount
  uint32_t capacity;
  uint32_t* heavy;               // per pixel: random walks started by its first samples (sample ordering, FrameParams::order); null = off
};

constexpr uint32_t kNoPixel = 0xFFFFFFFFu;
constexpr int kShadeBlock = 512;   // most threads per block of the shading kernels (block-synchronous batches)

// ---- streaming access to path records (ld/st.global.cs: evict-first in L2, the BVH stays resident)
// (the ping-pong halves are picked with a select, not an index: indexing a kernel-parameter array with a run-time value
// makes the compiler copy the struct to local memory)
__device__ __forceinline__ float4* StateBuf(const WaveState& w, uint32_t parity) { return parity ? w.state[1] : w.state[0]; }
__device__ __forceinline__ float4* WalkBuf(const WaveState& w, uint32_t parity) { return parity ? w.walk[1] : w.walk[0]; }
__device__ __forceinline__ float4* DoneBuf(const WaveState& w, uint32_t parity) { return parity ? w.done[1] : w.done[0]; }
__device__ __forceinline__ float4 LdS(const WaveState& w, uint32_t parity, uint32_t i, int field) {
  return __ldcs(&StateBuf(w, parity)[size_t(field) * w.capacity + i]);
}
// Gathers by queue position (shading kernels) and the scattered 16-byte hit stores: the positions a kernel touches at
// any time lie in a window of a few hundred thousand neighbours, but neighbours in the same 32-byte sector are
// touched by different warps at different times — with the evict-first hint the sector went back to HBM in between
// (L2 hit 21 %, 2x the DRAM bytes).  Default policy: the window (tens of MB over all fields) lives in L2.
__device__ __forceinline__ float4 LdSGather(const WaveState& w, uint32_t parity, uint32_t i, int field) {
  return __ldg(&StateBuf(w, parity)[size_t(field) * w.capacity + i]);
}
__device__ __forceinline__ void StSScatter(const WaveState& w, uint32_t parity, uint32_t i, int field, const float4& v) {
  StateBuf(w, parity)[size_t(field) * w.capacity + i] = v;
}
__device__ __forceinline__ void StS(const WaveState& w, uint32_t parity, uint32_t i, int field, const float4& v) {
  __stcs(&StateBuf(w, parity)[size_t(field) * w.capacity + i], v);
}
__device__ __forceinline__ float4 LdW(const WaveState& w, uint32_t parity, uint32_t i, int field) {
  return __ldcs(&WalkBuf(w, parity)[size_t(field) * w.capacity + i]);
}
__device__ __forceinline__ void StW(const WaveState& w, uint32_t parity, uint32_t i, int field, const float4& v) {
  __stcs(&WalkBuf(w, parity)[size_t(field) * w.capacity + i], v);
}
__device__ __forceinline__ float4 LdE(const WaveState& w, uint32_t i, int field) {
  return __ldcs(&w.exit_rec[size_t(field) * w.capacity + i]);
}
__device__ __forceinline__ void StE(const WaveState& w, uint32_t i, int field, const float4& v) {
  __stcs(&w.exit_rec[size_t(field) * w.capacity + i], v);
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

// Five queue reservations with ONE atomic instruction: lanes 0..4 each reserve the range of one queue (different
// addresses, so the L2 handles them in parallel) instead of five dependent atomic round trips — waiting for atomic
// results was 18 % of the closest-hit kernel's stall samples (profiles/r1e_ncu.md).  Issue early, resolve late:
// whatever is issued in between overlaps the round trip.  A null counter is a queue that is not used.
struct Append5 {
  unsigned m[5];
  uint32_t base;   // lane k < 5: first index reserved in queue k
};
__device__ __forceinline__ Append5 Append5Issue(uint32_t* c0, uint32_t* c1, uint32_t* c2, uint32_t* c3, uint32_t* c4,
                                                bool p0, bool p1, bool p2, bool p3, bool p4) {
  Append5 a;
  a.m[0] = c0 ? __ballot_sync(0xffffffffu, p0) : 0u;
  a.m[1] = c1 ? __ballot_sync(0xffffffffu, p1) : 0u;
  a.m[2] = c2 ? __ballot_sync(0xffffffffu, p2) : 0u;
  a.m[3] = c3 ? __ballot_sync(0xffffffffu, p3) : 0u;
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

// A whole block pulls the next blockDim.x queue entries and re-converges: the shading kernels are long straight-line
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

// The shading kernels append to up to four output streams per batch and then fetch the next batch.  As one atomic per
// warp and stream these were 4-5 dependent round trips per batch to counters that every warp of the GPU hammers:
// 67 % of the diffuse kernel's stall samples were waits for atomic results (profiles/r2c_ncu.md).  Here a BLOCK
// reserves all of it with ONE atomic instruction per batch: the warps leave their ballot counts in shared memory,
// lanes 0..4 of warp 0 add the column totals to the four counters and the fetch counter (five addresses, served in
// parallel), and every thread computes its positions from the bases and its warp's offsets.
struct BlockSlots {
  uint32_t idx[4];       // position of this thread's record in stream q (valid if its predicate q was set)
  uint32_t next_fetch;   // first queue entry of the block's next batch
};
__device__ __forceinline__ BlockSlots BlockReserve(uint32_t* c0, uint32_t* c1, uint32_t* c2, uint32_t* c3,
                                                   uint32_t* fetch_counter, bool p0, bool p1, bool p2, bool p3) {
  __shared__ uint32_t s_cnt[kShadeBlock / 32][4];
  __shared__ uint32_t s_base[5];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1),
                 m2 = __ballot_sync(0xffffffffu, p2), m3 = __ballot_sync(0xffffffffu, p3);
  if (lane == 0u) {
    s_cnt[warp][0] = uint32_t(__popc(m0)); s_cnt[warp][1] = uint32_t(__popc(m1));
    s_cnt[warp][2] = uint32_t(__popc(m2)); s_cnt[warp][3] = uint32_t(__popc(m3));
  }
  __syncthreads();
  if (warp == 0u) {
    if (lane < 4u) {
      uint32_t acc = 0;
      for (unsigned k = 0; k < nwarps; ++k) {
        const uint32_t c = s_cnt[k][lane];
        s_cnt[k][lane] = acc;   // exclusive offset of warp k in stream `lane`
        acc += c;
      }
      uint32_t* ctr = lane == 0u ? c0 : (lane == 1u ? c1 : (lane == 2u ? c2 : c3));
      s_base[lane] = (acc && ctr) ? atomicAdd(ctr, acc) : 0u;
    } else if (lane == 4u) {
      s_base[4] = atomicAdd(fetch_counter, blockDim.x);
    }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  BlockSlots r;
  r.idx[0] = s_base[0] + s_cnt[warp][0] + uint32_t(__popc(m0 & lt));
  r.idx[1] = s_base[1] + s_cnt[warp][1] + uint32_t(__popc(m1 & lt));
  r.idx[2] = s_base[2] + s_cnt[warp][2] + uint32_t(__popc(m2 & lt));
  r.idx[3] = s_base[3] + s_cnt[warp][3] + uint32_t(__popc(m3 & lt));
  r.next_fetch = s_base[4];
  __syncwarp();   // lane 0 overwrites this warp's row in the next batch
  return r;
}

__device__ __forceinline__ RayT RayFrom(const float4& o, const float4& d) {
  RayT r;
  r.o = vec3(o.x, o.y, o.z); r.tmin = o.w;
  r.d = vec3(d.x, d.y, d.z); r.tmax = d.w;
  return r;
}
__device__ __forceinline__ HitT HitFrom(const float4& h4) {
  HitT hit;
  hit.t = h4.x; hit.u = h4.y; hit.v = h4.z; hit.prim = __float_as_uint(h4.w);
  return hit;
}
__device__ __forceinline__ float4 PackHit(const HitT& hit) {
  return make_float4(hit.t, hit.u, hit.v, __uint_as_float(hit.prim));
}
__device__ __forceinline__ Pcg32 RngFrom(const float4& r) {
  Pcg32 rng;
  rng.state = (uint64_t(__float_as_uint(r.y)) << 32) | __float_as_uint(r.x);
  rng.inc = (uint64_t(__float_as_uint(r.w)) << 32) | __float_as_uint(r.z);
  return rng;
}
__device__ __forceinline__ float4 PackRng(const Pcg32& rng) {
  return make_float4(__uint_as_float(uint32_t(rng.state)), __uint_as_float(uint32_t(rng.state >> 32)),
                     __uint_as_float(uint32_t(rng.inc)), __uint_as_float(uint32_t(rng.inc >> 32)));
}
__device__ __forceinline__ RayT LoadRay(const WaveState& w, uint32_t parity, uint32_t i) {
  return RayFrom(LdS(w, parity, i, kRayO), LdS(w, parity, i, kRayD));
}
__device__ __forceinline__ HitT LoadHit(const WaveState& w, uint32_t parity, uint32_t i) {
  return HitFrom(LdS(w, parity, i, kHit));
}

__device__ __forceinline__ uint32_t MakeTarget(uint32_t tag, uint32_t index) { return (tag << 30) | index; }

// NEE request of a vertex: the any-hit kernel adds `throughput * contribute` to the target if the ray is unoccluded
__device__ __forceinline__ void PushShadow(const WaveState& w, const ShadowRequest& req, const vec3& throughput,
                                           uint32_t target) {
  const bool push = req.active && (target >> 30) != kTargetNone;
  const uint32_t slot = WarpAppend(&w.counters[kNumShadow], push);
  if (push) {
    const vec3 c = throughput * req.contribute;
    __stcs(&w.sh_o[slot], make_float4(req.ray.o.x, req.ray.o.y, req.ray.o.z, req.ray.tmin));
    __stcs(&w.sh_d[slot], make_float4(req.ray.d.x, req.ray.d.y, req.ray.d.z, req.ray.tmax));
    __stcs(&w.sh_c[slot], make_float4(c.x, c.y, c.z, __uint_as_float(target)));
  }
}

// the same with a position reserved by BlockReserve
__device__ __forceinline__ void StoreShadow(const WaveState& w, uint32_t slot, const ShadowRequest& req,
                                            const vec3& throughput, uint32_t target) {
  const vec3 c = throughput * req.contribute;
  __stcs(&w.sh_o[slot], make_float4(req.ray.o.x, req.ray.o.y, req.ray.o.z, req.ray.tmin));
  __stcs(&w.sh_d[slot], make_float4(req.ray.d.x, req.ray.d.y, req.ray.d.z, req.ray.tmax));
  __stcs(&w.sh_c[slot], make_float4(c.x, c.y, c.z, __uint_as_float(target)));
}

// where a path goes after its vertex: S[next] (it continues) or D[next] (it ended); returns the NEE target
struct Routed {
  uint32_t index;    // position in S[next] / D[next]
  uint32_t target;
};
__device__ __forceinline__ Routed RoutePath(const WaveState& w, uint32_t next_parity, bool to_next, bool to_done) {
  const uint32_t a = WarpAppend(&w.counters[kNumActive0 + next_parity], to_next);
  const uint32_t b = WarpAppend(&w.counters[kNumDone0 + next_parity], to_done);
  Routed r;
  r.index = to_next ? a : b;
  r.target = to_next ? MakeTarget(kTargetState, a) : (to_done ? MakeTarget(kTargetDone, b) : MakeTarget(kTargetNone, 0u));
  return r;
}

// ------------------------------------------------------------------------------------------------ iteration set-up
// zeroes every per-iteration counter; the three ping-pong lists keep the half that this iteration consumes.  In frame
// mode it also decides how many camera samples this iteration starts: whatever capacity the paths and walks in flight
// leave free, as long as the frame has samples left.  The closest-hit kernel gives work item (n_active + k) the sample
// id (base + k), so generation needs no atomic of its own.
__global__ void BeginIterationKernel(uint32_t* counters, unsigned long long* stats, uint32_t cur_parity,
                                     uint32_t frame_mode, uint32_t capacity, unsigned long long total_samples) {
  // total_samples: the frame's samples that may have been handed out by the end of this iteration (the host holds the
  // frame at the end of its probing passes until the pixel order is built, FrameParams::order)
  const uint32_t i = threadIdx.x;
  __shared__ uint32_t n_new;
  if (i == 0) {
    n_new = 0u;
    if (frame_mode) {
      const unsigned long long used = (unsigned long long)counters[kNumActive0 + cur_parity] + counters[kNumWalk0 + cur_parity];
      const unsigned long long space = capacity > used ? capacity - used : 0ull;
      const unsigned long long base = stats[kStatNextSample];
      const unsigned long long left = total_samples > base ? total_samples - base : 0ull;
      const unsigned long long take = space < left ? space : left;
      n_new = uint32_t(take);
      stats[kStatSampleBase] = base;
      stats[kStatNextSample] = base + take;
    }
  }
  __syncthreads();
  if (i >= kCounterCount) return;
  const uint32_t keep0 = kNumActive0 + cur_parity, keep1 = kNumWalk0 + cur_parity, keep2 = kNumDone0 + cur_parity;
  if (i == keep0 || i == keep1 || i == keep2) return;
  counters[i] = (i == kNumNew) ? n_new : 0u;
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
  // Longest paths first.  The reference hands out (tile, sample) jobs in raster order (render.cc:211-222); which
  // sample runs when does not change the image (every (pixel, sample) has its own random stream).  Here the first
  // `probe_passes` samples of every pixel run in that order and count, per pixel, the random walks they start; the
  // remaining `rest_passes` samples run pixel block by pixel block (`order_block` pixels, sample-major inside a
  // block) through `order` = the pixels sorted by that count, descending: the paths that live for hundreds of
  // iterations start early and the frame drains with short ones.  order == null or rest_passes == 0: raster order.
  const uint32_t* order;
  uint32_t probe_passes, rest_passes, order_block;
};

__device__ __forceinline__ RayT StartCameraPath(const WaveState& w, const FrameParams& f, uint32_t parity, uint32_t i,
                                                unsigned long long id, uint32_t* pixel_out) {
  uint32_t s_local, pixel;
  const pbrjob::SampleOrder so = {f.order, f.npix, f.probe_passes, f.rest_passes, f.order_block};
  pbrjob::SampleOfId(so, id, &pixel, &s_local);   // job_split.h
  const uint32_t x = pixel % f.cam.width, y = pixel / f.cam.width;
  Pcg32 rng;
  pcg32_srandom(&rng, f.seed + uint64_t(f.first_sample) + uint64_t(s_local) * f.sample_stride, uint64_t(pixel));
  const float jx = Draw(&rng), jy = Draw(&rng);
  const float tx = f.cam.x_corner + f.cam.dx * (float(x) + jx);
  const float ty = f.cam.y_corner - f.cam.dy * (float(y) + jy);
  float dx = tx - f.cam.eye[0], dy = ty - f.cam.eye[1], dz = f.cam.z_corner - f.cam.eye[2];
  const float inv_norm = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);   // Normalize (render.cc:243-249)
  dx *= inv_norm; dy *= inv_norm; dz *= inv_norm;
  StS(w, parity, i, kRayO, make_float4(f.cam.eye[0], f.cam.eye[1], f.cam.eye[2], 0.0f));
  StS(w, parity, i, kRayD, make_float4(dx, dy, dz, kInf));
  StS(w, parity, i, kThr, make_float4(1.f, 1.f, 1.f, 0.f));
  StS(w, parity, i, kRad, make_float4(0.f, 0.f, 0.f, __uint_as_float(0u)));
  StS(w, parity, i, kRng, PackRng(rng));
  StS(w, parity, i, kPix, make_float4(__uint_as_float(pixel), 0.f, 0.f, 0.f));
  *pixel_out = pixel;
  RayT ray;
  ray.o = vec3(f.cam.eye[0], f.cam.eye[1], f.cam.eye[2]); ray.tmin = 0.0f;
  ray.d = vec3(dx, dy, dz); ray.tmax = kInf;
  return ray;
}

// render.cc:175-183 for every path that ended in the previous iteration: rgba[pixel] += (L, 1).  D[cur] is read front
// to back; the frame accumulator (16 B per pixel) is L2-resident.
__global__ void __launch_bounds__(256) RetireKernel(WaveState w, uint32_t cur_parity, float4* rgba) {
  const uint32_t n = w.counters[kNumDone0 + cur_parity];
  const float4* __restrict__ done = DoneBuf(w, cur_parity);
  uint32_t retired = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 r = __ldcs(&done[i]);
    const uint32_t pixel = __float_as_uint(r.w);
    if (pixel != kNoPixel) {
      atomicAdd(&rgba[pixel], make_float4(r.x, r.y, r.z, 1.0f));
      ++retired;
    }
  }
  WarpTally(&w.stats[kStatRetired], retired);
}

// RenderLayer::count (render-layer.h:11-26) from the alpha sums: both are incremented together per sample
__global__ void FinishFrameKernel(const float4* rgba, uint32_t* count, uint32_t npix) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix) count[i] = uint32_t(rgba[i].w);
}

// nothing in flight: the state a frame starts from
__global__ void ResetPoolKernel(WaveState w) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < kCounterCount) w.counters[p] = 0u;
}

// caller-supplied rays + seeds (pbrgpu_radiance / pbrgpu_shade hooks): path i = pixel i, no camera samples
__global__ void InitPathsFromRaysKernel(WaveState w, const float4* rays, const uint64_t* seeds, uint32_t n) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < kCounterCount) w.counters[p] = (p == kNumActive0) ? n : 0u;
  if (p >= n) return;
  StS(w, 0u, p, kRayO, rays[2 * p]);
  StS(w, 0u, p, kRayD, rays[2 * p + 1]);
  Pcg32 rng;
  pcg32_srandom(&rng, seeds[2 * p], seeds[2 * p + 1]);
  StS(w, 0u, p, kThr, make_float4(1.f, 1.f, 1.f, 0.f));
  StS(w, 0u, p, kRad, make_float4(0.f, 0.f, 0.f, __uint_as_float(0u)));
  StS(w, 0u, p, kRng, PackRng(rng));
  StS(w, 0u, p, kPix, make_float4(__uint_as_float(p), 0.f, 0.f, 0.f));
}

// ------------------------------------------------------------------------------------------------ closest hit
// Scene::TraceFirstHit1 for every path of S[cur]; a hit is routed by the material class of what it hit, a miss is
// retired on the spot (render.cc:28-30: the path ends with the radiance it has).  Items past n_active are this
// iteration's new camera samples: generated here, so a camera ray never makes a round trip through memory before its
// first traversal.  Runs in the warp traversal engine: lanes are refilled while others traverse.
struct ClosestClient {
  static constexpr bool kOrderByOrigin = true;
  const SceneView& s;
  const WaveState& w;
  const FrameParams& frame;      // frame mode: how camera samples are numbered
  float4* rgba;                  // accumulator of retired paths (frame.rgba, or the hook's per-path buffer)
  const bool sort_materials;     // route diffuse-only materials to their own shading queue
  uint32_t cur, n_active, n, lanes;
  unsigned long long sample_base;
  uint32_t item = 0, pixel = kNoPixel;
  bool has_result = false, fresh = false;   // fresh: a camera ray (radiance 0, pixel known)
  uint32_t rays = 0, retired = 0;

  __device__ __forceinline__ ClosestClient(const SceneView& s_, const WaveState& w_, uint32_t cur_parity,
                                           const FrameParams& frame_, float4* rgba_, bool sort_)
      : s(s_), w(w_), frame(frame_), rgba(rgba_), sort_materials(sort_), cur(cur_parity),
        n_active(w_.counters[kNumActive0 + cur_parity]),
        n(w_.counters[kNumActive0 + cur_parity] + w_.counters[kNumNew]),
        lanes(LanesFor(s_, w_.counters[kNumActive0 + cur_parity] + w_.counters[kNumNew])),
        sample_base(w_.stats[kStatSampleBase]) {}
  __device__ __forceinline__ bool Fetches() const { return (threadIdx.x & 31u) < lanes; }
  __device__ __forceinline__ uint32_t RefillThreshold(uint32_t dflt) const { return ThresholdFor(lanes, dflt); }
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || (!exhausted && Fetches()); }

  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    const bool finished = !t.active && has_result;
    const bool need = !exhausted && !t.active && Fetches();
    // ---- (1) finished rays: miss -> retire, else the shading queue of the material's class
    int kind = -1;   // -1 nothing, 0 miss, 1 general surface queue, 2 hair queue, 3 diffuse-only queue
    const HitT hit = t.hit;
    const uint32_t done_item = item;
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
    // ---- (2) the three output queues and the work fetch: one atomic instruction
    const Append5 app = Append5Issue(&w.counters[kNumSurface], &w.counters[kNumHair], nullptr, &w.counters[kFetchTrace],
                                     &w.counters[kNumDiffuse], kind == 1, kind == 2, false, need, kind == 3);
    const uint32_t i_surf = Append5Index(app, 0), i_hair = Append5Index(app, 1), next_item = Append5Index(app, 3),
                   i_diff = Append5Index(app, 4);
    // ---- (3) new work: the loads are issued here and consumed after the finished rays have been written out
    const bool take = need && next_item < n_active;                      // a path that continues
    const bool start = need && next_item >= n_active && next_item < n;   // a new camera sample
    float4 ro = make_float4(0.f, 0.f, 0.f, 0.f), rd = ro;
    if (take) {
      ro = LdS(w, cur, next_item, kRayO);
      rd = LdS(w, cur, next_item, kRayD);
    }
    // ---- (4) write out the finished rays
    if (kind > 0) {
      StSScatter(w, cur, done_item, kHit, PackHit(hit));
      if (kind == 1) w.q_surface[i_surf] = done_item;
      else if (kind == 2) w.q_hair[i_hair] = done_item;
      else w.q_diffuse[i_diff] = done_item;
    } else if (kind == 0) {
      // render.cc:175-183 for the path that just left the scene
      float4 rrad = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t rpix = pixel;
      if (!fresh) {
        rrad = LdS(w, cur, done_item, kRad);
        rpix = __float_as_uint(LdS(w, cur, done_item, kPix).x);
      }
      if (rpix != kNoPixel) {
        atomicAdd(&rgba[rpix], make_float4(rrad.x, rrad.y, rrad.z, 1.0f));
        ++retired;
      }
    }
    // ---- (5) start the new rays
    if (take) {
      item = next_item;
      fresh = false;
      TravBegin(s, RayFrom(ro, rd), t);
      has_result = true;
      ++rays;
    } else if (start) {
      item = next_item;
      fresh = true;
      const RayT ray = StartCameraPath(w, frame, cur, item, sample_base + (next_item - n_active), &pixel);
      TravBegin(s, ray, t);
      has_result = true;
      ++rays;
    }
    return need && next_item >= n;
  }
  __device__ __forceinline__ void End(const Trav&) {
    WarpTally(&w.stats[kStatClosest], rays);
    WarpTally(&w.stats[kStatRetired], retired);
  }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) TraceClosestKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                          uint32_t refill_min_idle, uint32_t prim_min_lanes,
                                                          FrameParams frame, float4* rgba, uint32_t sort_materials) {
  ClosestClient client(s, w, cur_parity, frame, rgba, sort_materials != 0u);
  TravEngine<false, HAS_CURVES, false>(s, client, refill_min_idle, prim_min_lanes);
}

// ------------------------------------------------------------------------------------------------ shading
struct ShadeFlags {
  uint32_t skip_emission_and_roulette;   // pbrgpu_shade hook: call Shader() only, and keep every path in S[next]
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

__device__ __forceinline__ PathRegs LoadPath(const WaveState& w, uint32_t parity, uint32_t i) {
  const float4 o = LdSGather(w, parity, i, kRayO), d = LdSGather(w, parity, i, kRayD), t4 = LdSGather(w, parity, i, kThr),
               r4 = LdSGather(w, parity, i, kRad), h4 = LdSGather(w, parity, i, kHit), g4 = LdSGather(w, parity, i, kRng),
               x4 = LdSGather(w, parity, i, kPix);
  PathRegs r;
  r.ray = RayFrom(o, d);
  r.hit = HitFrom(h4);
  r.throughput = vec3(t4.x, t4.y, t4.z);
  r.pdf_prev = t4.w;
  r.L = vec3(r4.x, r4.y, r4.z);
  r.depth = __float_as_uint(r4.w);
  r.rng = RngFrom(g4);
  r.pixel = __float_as_uint(x4.x);
  return r;
}

// after a vertex: the path's next state, appended to S[next] (render.cc:79-86)
__device__ __forceinline__ void CommitVertex(const WaveState& w, uint32_t next_parity, uint32_t j, const VertexResult& vr,
                                             const vec3& throughput, const vec3& L, uint32_t depth, const Pcg32& rng,
                                             uint32_t pixel) {
  const vec3 new_thr = vr.throughput * throughput;                                  // render.cc:80
  StS(w, next_parity, j, kRayO, make_float4(vr.P.x, vr.P.y, vr.P.z, 1e-3f));        // render.cc:83-86
  StS(w, next_parity, j, kRayD, make_float4(vr.wi.x, vr.wi.y, vr.wi.z, kInf));
  StS(w, next_parity, j, kThr, make_float4(new_thr.x, new_thr.y, new_thr.z, vr.pdf));
  StS(w, next_parity, j, kRad, make_float4(L.x, L.y, L.z, __uint_as_float(depth + 1u)));
  StS(w, next_parity, j, kRng, PackRng(rng));
  StS(w, next_parity, j, kPix, make_float4(__uint_as_float(pixel), 0.f, 0.f, 0.f));
}

// the path ends here: its radiance waits in D[next] for this iteration's shadow rays, then for the next retire
__device__ __forceinline__ void CommitDone(const WaveState& w, uint32_t next_parity, uint32_t j, const vec3& L,
                                           uint32_t pixel) {
  __stcs(&DoneBuf(w, next_parity)[j], make_float4(L.x, L.y, L.z, __uint_as_float(pixel)));
}

// a walk that starts (or goes on) next iteration: hot part
__device__ __forceinline__ void StoreWalkHot(const WaveState& w, uint32_t parity, uint32_t k, const SssWalkState& st,
                                             const Pcg32& rng, uint32_t pixel) {
  StW(w, parity, k, kWalkA, make_float4(st.sigma_t.x, st.sigma_t.y, st.sigma_t.z, st.throughput.x));
  StW(w, parity, k, kWalkB, make_float4(st.sigma_s.x, st.sigma_s.y, st.sigma_s.z, st.throughput.y));
  StW(w, parity, k, kWalkC, make_float4(st.ray.o.x, st.ray.o.y, st.ray.o.z, st.throughput.z));
  StW(w, parity, k, kWalkD, make_float4(st.ray.d.x, st.ray.d.y, st.ray.d.z, st.ray.tmin));
  StW(w, parity, k, kWalkN, make_float4(__uint_as_float(st.bounce), __uint_as_float(pixel), 0.f, 0.f));
  StW(w, parity, k, kWalkRng, PackRng(rng));
}

// What a vertex decided; the kernel then reserves the output positions for the whole block and commits.
struct VertexOutcome {
  bool to_next = false, to_done = false, to_sss = false;
  ShadowRequest req;
  vec3 throughput, L;
  VertexResult vr;
  SssWalkState walk;
};

// One Principled vertex (render.cc:33-86 + shader.cc:8-34) of a path of S[cur] whose state is in r.
template <bool DIFFUSE_ONLY>
__device__ __forceinline__ void ShadeSurfaceVertex(const SceneView& s, const ShadeFlags& flags, PathRegs& r,
                                                   VertexOutcome& o) {
  o.throughput = r.throughput;
  o.L = r.L;
  const Surface si = MakeSurface(s, r.ray, r.hit);
  bool alive = true;
  if (!flags.skip_emission_and_roulette)
    alive = EmissionAndRoulette(s, r.ray, r.hit, si, r.depth, r.pdf_prev, &r.rng, &o.L, &o.throughput);
  if (!alive) {
    o.to_done = true;
    return;
  }
  const int kind = DIFFUSE_ONLY ? 1 : MaterialKind(s, si);
  const vec3 wo = -r.ray.d;
  Frame fr;
  PrincipledBsdf bsdf;
  bool sss = false;
  if (kind == 1) sss = PrincipledVertexT<DIFFUSE_ONLY>(s, si, wo, &r.rng, &o.vr, &fr, &bsdf);
  else AbsorbVertex(wo, si.P, &o.vr);   // no material (shader.cc:11-17)
  o.req = o.vr.shadow[0];
  if (!DIFFUSE_ONLY && sss) {
    // the walk runs in its own kernel; rejected: throughput 0, the path ends (cycles-principled-shader.cc:217-220)
    if (SssBegin(si, fr, bsdf, &r.rng, &o.walk)) o.to_sss = true;
    else o.to_done = true;
  } else {
    o.to_next = flags.skip_emission_and_roulette ? true : !IsBlack(o.vr.throughput * o.throughput);   // render.cc:31
    o.to_done = !o.to_next;
  }
}

// writes the outcome of a vertex to the positions BlockReserve handed out: S[next] / D[next] / W[next] + shadow queue
__device__ __forceinline__ void CommitOutcome(const WaveState& w, uint32_t next_parity, const BlockSlots& slots,
                                              const VertexOutcome& o, const PathRegs& r, bool shadow) {
  uint32_t target = MakeTarget(kTargetNone, 0u);
  if (o.to_next) {
    CommitVertex(w, next_parity, slots.idx[0], o.vr, o.throughput, o.L, r.depth, r.rng, r.pixel);
    target = MakeTarget(kTargetState, slots.idx[0]);
  } else if (o.to_done) {
    CommitDone(w, next_parity, slots.idx[1], o.L, r.pixel);
    target = MakeTarget(kTargetDone, slots.idx[1]);
  } else if (o.to_sss) {
    // the walk, the entry vertex (ray + hit: the exit vertex rebuilds the entry surface from them) and the path it
    // resumes with (post-roulette throughput)
    const uint32_t k = slots.idx[2];
    StoreWalkHot(w, next_parity, k, o.walk, r.rng, r.pixel);
    StW(w, next_parity, k, kWalkRayO, make_float4(r.ray.o.x, r.ray.o.y, r.ray.o.z, r.ray.tmin));
    StW(w, next_parity, k, kWalkRayD, make_float4(r.ray.d.x, r.ray.d.y, r.ray.d.z, r.ray.tmax));
    StW(w, next_parity, k, kWalkHit, PackHit(r.hit));
    StW(w, next_parity, k, kWalkThr, make_float4(o.throughput.x, o.throughput.y, o.throughput.z, r.pdf_prev));
    StW(w, next_parity, k, kWalkRad, make_float4(o.L.x, o.L.y, o.L.z, __uint_as_float(r.depth)));
    target = MakeTarget(kTargetWalk, k);
    if (w.heavy) atomicAdd(&w.heavy[r.pixel], 1u);   // probing passes of a frame: this pixel's paths are long ones
  }
  if (shadow) StoreShadow(w, slots.idx[3], o.req, o.throughput, target);
}

// Launch shapes: the general kernel needs ~110-128 registers, so one 512-thread block per SM; the diffuse-only kernel is
// 7x smaller (2.1k vs 15.7k SASS instructions): 128-thread blocks, 6 per SM (sweep: profiles/r1z_diffuse_shape_sweep.log)
constexpr int kDiffuseBlock = 256, kDiffuseBlocksPerSm = 3;
template <bool DIFFUSE_ONLY>
__global__ void __launch_bounds__(DIFFUSE_ONLY ? kDiffuseBlock : kShadeBlock, DIFFUSE_ONLY ? kDiffuseBlocksPerSm : 1)
ShadeSurfaceKernel(SceneView s, WaveState w, uint32_t cur_parity, ShadeFlags flags) {
  const uint32_t n = w.counters[DIFFUSE_ONLY ? kNumDiffuse : kNumSurface];
  const uint32_t* __restrict__ queue = DIFFUSE_ONLY ? w.q_diffuse : w.q_surface;
  uint32_t* fetch = &w.counters[DIFFUSE_ONLY ? kFetchDiffuse : kFetchSurface];
  const uint32_t next_parity = cur_parity ^ 1u;
  uint32_t shaded = 0;
  uint32_t base = BlockFetch(fetch) - threadIdx.x;
  while (base < n) {   // block-uniform
    const uint32_t slot = base + threadIdx.x;
    const bool valid = slot < n;
    shaded += valid ? 1u : 0u;
    PathRegs r;
    VertexOutcome o;
    o.req.active = false;
    if (valid) {
      r = LoadPath(w, cur_parity, queue[slot]);
      ShadeSurfaceVertex<DIFFUSE_ONLY>(s, flags, r, o);
    }
    const bool shadow = valid && o.req.active && (o.to_next || o.to_done || o.to_sss);
    const BlockSlots slots = BlockReserve(&w.counters[kNumActive0 + next_parity], &w.counters[kNumDone0 + next_parity],
                                          DIFFUSE_ONLY ? nullptr : &w.counters[kNumWalk0 + next_parity],
                                          &w.counters[kNumShadow], fetch, o.to_next, o.to_done, o.to_sss, shadow);
    if (valid) CommitOutcome(w, next_parity, slots, o, r, shadow);
    base = slots.next_fetch;
  }
  WarpTally(&w.stats[kStatVertices], shaded);
}

__global__ void __launch_bounds__(kShadeBlock) ShadeHairKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                               ShadeFlags flags) {
  const uint32_t n = w.counters[kNumHair];
  const uint32_t next_parity = cur_parity ^ 1u;
  uint32_t shaded = 0;
  uint32_t base = BlockFetch(&w.counters[kFetchHair]) - threadIdx.x;
  while (base < n) {   // block-uniform
    const uint32_t slot = base + threadIdx.x;
    const bool valid = slot < n;
    shaded += valid ? 1u : 0u;
    PathRegs r;
    VertexOutcome o;
    o.req.active = false;
    if (valid) {
      r = LoadPath(w, cur_parity, w.q_hair[slot]);
      o.throughput = r.throughput;
      o.L = r.L;
      const Surface si = MakeSurface(s, r.ray, r.hit);
      bool alive = true;
      if (!flags.skip_emission_and_roulette)
        alive = EmissionAndRoulette(s, r.ray, r.hit, si, r.depth, r.pdf_prev, &r.rng, &o.L, &o.throughput);
      if (!alive) {
        o.to_done = true;
      } else {
        HairVertex(s, si, -r.ray.d, &r.rng, &o.vr);
        o.req = o.vr.shadow[0];
        o.to_next = flags.skip_emission_and_roulette ? true : !IsBlack(o.vr.throughput * o.throughput);
        o.to_done = !o.to_next;
      }
    }
    const bool shadow = valid && o.req.active;
    const BlockSlots slots = BlockReserve(&w.counters[kNumActive0 + next_parity], &w.counters[kNumDone0 + next_parity],
                                          nullptr, &w.counters[kNumShadow], &w.counters[kFetchHair], o.to_next,
                                          o.to_done, false, shadow);
    if (valid) CommitOutcome(w, next_parity, slots, o, r, shadow);
    base = slots.next_fetch;
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
//   * a walk gets at most `max_bounces` bounces per launch; if it is still inside the medium its record is appended
//     to W[next] and the next iteration goes on with it.  A launch therefore never outlives its queue by more than
//     max_bounces bounces (and the walks are re-balanced over the warps every time: larger budgets were measured and
//     are slower, profiles/r2a_tune_diffuse_pipe_walk_budget_pool.log), and because the pool is kept full by new camera samples, long walks cost
//     capacity, not idle SMs;
//   * entering the medium is part of shade_surface, leaving it (exit vertex: NEE + diffuse bounce,
//     cycles-principled-shader.cc:187-216) is sss_exit: both are rare per bounce and ran at 2-3 lanes per warp when
//     they were inlined here.
struct SssClient {
  static constexpr bool kOrderByOrigin = false;
  const SceneView& s;
  const WaveState& w;
  uint32_t cur, next, n, max_bounces, lanes;
  uint32_t item = 0, budget = 0, pixel = 0;
  bool has_walk = false;
  bool skipped_seg = false;   // the current segment was answered by the clearance grid, not traced
  Pcg32 rng;
  SssWalkState walk;    // walk.ray is rebuilt from the traversal state after every segment
  vec3 cpdf;            // channel probabilities of the segment in flight (SssPrepareSegment -> SssFinishSegment)
  uint32_t rays = 0, skipped = 0;

  __device__ __forceinline__ SssClient(const SceneView& s_, const WaveState& w_, uint32_t cur_parity, uint32_t max_b)
      : s(s_), w(w_), cur(cur_parity), next(cur_parity ^ 1u), n(w_.counters[kNumWalk0 + cur_parity]), max_bounces(max_b),
        lanes(LanesFor(s_, w_.counters[kNumWalk0 + cur_parity])) {}
  __device__ __forceinline__ bool Fetches() const { return (threadIdx.x & 31u) < lanes; }
  __device__ __forceinline__ uint32_t RefillThreshold(uint32_t dflt) const { return ThresholdFor(lanes, dflt); }
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_walk || (!exhausted && Fetches()); }

  __device__ __forceinline__ void StartSegment(Trav& t) {
    SssPrepareSegment(&rng, &walk, &cpdf);
    TravBegin(s, walk.ray, t);
    // Segments that end far from any surface: the clearance grid answers those ("no hit") without a traversal; the
    // lane then waits for the next converged section like any lane whose query has finished.  (Finishing chains of
    // such segments inline was measured and is slower: the converged section serialises on the longest chain.)
    skipped_seg = SegmentIsClear(s, walk.ray.o, walk.ray.d, walk.ray.tmax * 1.001f);
    if (skipped_seg) t.active = false;
  }

  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    // ---- (1) walks whose segment query finished: scatter / exit / absorb
    bool to_exit = false, to_done = false, to_park = false, go_on = false;
    const uint32_t src = item;
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
    // ---- (2) the three output streams and the work fetch (lanes without a walk): one atomic instruction
    const bool need = !exhausted && !t.active && !has_walk && Fetches();
    const Append5 app = Append5Issue(&w.counters[kNumExit], &w.counters[kNumDone0 + next],
                                     &w.counters[kNumWalk0 + next], &w.counters[kFetchWalk], nullptr, to_exit,
                                     to_done, to_park, need, false);
    const uint32_t i_exit = Append5Index(app, 0), i_done = Append5Index(app, 1), i_park = Append5Index(app, 2),
                   next_item = Append5Index(app, 3);
    // ---- (3) write out the walks that stopped
    if (to_exit) {
      // exit record for sss_exit: the segment ray, its hit, the walk throughput; the cold part stays in W[cur][src]
      StE(w, i_exit, kExHit, PackHit(hit));
      StE(w, i_exit, kExThr, make_float4(walk.throughput.x, walk.throughput.y, walk.throughput.z, __uint_as_float(src)));
      StE(w, i_exit, kExO, make_float4(walk.ray.o.x, walk.ray.o.y, walk.ray.o.z, walk.ray.tmin));
      StE(w, i_exit, kExD, make_float4(walk.ray.d.x, walk.ray.d.y, walk.ray.d.z, walk.ray.tmax));
      StE(w, i_exit, kExRng, PackRng(rng));
    } else if (to_park) {
      StoreWalkHot(w, next, i_park, walk, rng, pixel);
#pragma unroll
      for (int f = kWalkRayO; f < kWalkFields; ++f) StW(w, next, i_park, f, LdW(w, cur, src, f));
    } else if (to_done) {
      const float4 rad = LdW(w, cur, src, kWalkRad);
      CommitDone(w, next, i_done, vec3(rad.x, rad.y, rad.z), pixel);
    }
    // ---- (4) new walks: W[cur] front to back, loaded straight into the walk registers that (3) has finished with
    const bool take = need && next_item < n;
    if (take) {
      item = next_item;
      const float4 a = LdW(w, cur, item, kWalkA), b = LdW(w, cur, item, kWalkB), c = LdW(w, cur, item, kWalkC),
                   d = LdW(w, cur, item, kWalkD), nn = LdW(w, cur, item, kWalkN);
      rng = RngFrom(LdW(w, cur, item, kWalkRng));
      walk.sigma_t = vec3(a.x, a.y, a.z);
      walk.sigma_s = vec3(b.x, b.y, b.z);
      walk.throughput = vec3(a.w, b.w, c.w);
      walk.ray.o = vec3(c.x, c.y, c.z);
      walk.ray.d = vec3(d.x, d.y, d.z);
      walk.ray.tmin = d.w;
      walk.ray.tmax = kInf;
      walk.bounce = __float_as_uint(nn.x);
      pixel = __float_as_uint(nn.y);
      budget = max_bounces;
      has_walk = true;
    }
    // ---- (5) next segment: of the walk that goes on, or of the one just taken
    if (take || go_on) StartSegment(t);
    return need && next_item >= n;
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
__global__ void __launch_bounds__(kShadeBlock) SssExitKernel(SceneView s, WaveState w, uint32_t cur_parity,
                                                             ShadeFlags flags) {
  const uint32_t n = w.counters[kNumExit];
  const uint32_t next_parity = cur_parity ^ 1u;
  uint32_t shaded = 0;
  uint32_t base = BlockFetch(&w.counters[kFetchExit]) - threadIdx.x;
  while (base < n) {   // block-uniform
    const uint32_t e = base + threadIdx.x;
    const bool valid = e < n;
    shaded += valid ? 1u : 0u;
    PathRegs r;          // the path as it resumes: depth, pixel, rng (ray / hit unused)
    VertexOutcome o;
    o.req.active = false;
    if (valid) {
      const float4 xh = LdE(w, e, kExHit), xt = LdE(w, e, kExThr), xo = LdE(w, e, kExO), xd = LdE(w, e, kExD);
      r.rng = RngFrom(LdE(w, e, kExRng));
      const uint32_t src = __float_as_uint(xt.w);
      // the entry vertex and the path state, left in W[cur] by the iteration that started (or last parked) the walk
      const RayT entry_ray = RayFrom(LdW(w, cur_parity, src, kWalkRayO), LdW(w, cur_parity, src, kWalkRayD));
      const HitT entry_hit = HitFrom(LdW(w, cur_parity, src, kWalkHit));
      const float4 t4 = LdW(w, cur_parity, src, kWalkThr), r4 = LdW(w, cur_parity, src, kWalkRad);
      r.pixel = __float_as_uint(LdW(w, cur_parity, src, kWalkN).y);
      r.depth = __float_as_uint(r4.w);
      o.throughput = vec3(t4.x, t4.y, t4.z);
      o.L = vec3(r4.x, r4.y, r4.z);
      const Surface entry_si = MakeSurface(s, entry_ray, entry_hit);
      const Frame entry_frame = PrincipledFrame(entry_si);
      SssWalkState walk;
      walk.throughput = vec3(xt.x, xt.y, xt.z);
      walk.ray = RayFrom(xo, xd);
      o.vr.P = entry_si.P;
      o.vr.shadow[1].active = false;
      SssFinish(s, entry_si, entry_frame, walk, HitFrom(xh), &r.rng, &o.vr);
      o.req = o.vr.shadow[1];
      o.to_next = flags.skip_emission_and_roulette ? true : !IsBlack(o.vr.throughput * o.throughput);
      o.to_done = !o.to_next;
    }
    const bool shadow = valid && o.req.active;
    const BlockSlots slots = BlockReserve(&w.counters[kNumActive0 + next_parity], &w.counters[kNumDone0 + next_parity],
                                          nullptr, &w.counters[kNumShadow], &w.counters[kFetchExit], o.to_next,
                                          o.to_done, false, shadow);
    if (valid) CommitOutcome(w, next_parity, slots, o, r, shadow);
    base = slots.next_fetch;
  }
  WarpTally(&w.stats[kStatVertices], shaded);
}

// ------------------------------------------------------------------------------------------------ the last paths
// The end of a frame is a handful of paths, each a chain of dependent steps; an iteration of the wavefront moves every
// one of them by ONE vertex (or one slice of a walk) and costs eight kernel launches and a host round trip — 80 to
// 350 us for 16 us of work on a path's critical chain.  Once fewer than a few thousand paths and walks are in flight
// the host launches this kernel instead of an iteration: every remaining path of S[cur] and every remaining walk of
// W[cur] is taken up where it stands by ONE thread and run to its end with the megakernel form of the path loop
// (kat.cuh: PathRadianceFrom — the same per-vertex functions, so the same radiance; the walk segments are answered
// by the clearance field or traced exactly as in SssWalkKernel), items spread one per warp first (a warp runs the
// union of its lanes' instruction streams).  The result goes straight into the frame (render.cc:175-183).
__global__ void __launch_bounds__(128, 4) FinishPathsKernel(SceneView s, WaveState w, uint32_t cur, float4* rgba) {
  const uint32_t n_act = w.counters[kNumActive0 + cur], n_walk = w.counters[kNumWalk0 + cur];
  const uint32_t warps = gridDim.x * (blockDim.x >> 5), warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  uint64_t counts[3] = {0ull, 0ull, 0ull};
  uint32_t skipped = 0, retired = 0;
  for (uint32_t item = (threadIdx.x & 31u) * warps + warp; item < n_act + n_walk; item += 32u * warps) {
    RayT ray;
    vec3 L, thr;
    float pdf_prev;
    uint32_t depth, pixel;
    Pcg32 rng;
    bool alive = true;
    if (item < n_act) {
      ray = LoadRay(w, cur, item);
      const float4 t4 = LdS(w, cur, item, kThr), r4 = LdS(w, cur, item, kRad);
      thr = vec3(t4.x, t4.y, t4.z); pdf_prev = t4.w;
      L = vec3(r4.x, r4.y, r4.z); depth = __float_as_uint(r4.w);
      rng = RngFrom(LdS(w, cur, item, kRng));
      pixel = __float_as_uint(LdS(w, cur, item, kPix).x);
    } else {
      // a walk: the rest of its bounces (random-walk-sss.h:281-383), then the exit vertex as in SssExitKernel
      const uint32_t k = item - n_act;
      const float4 a = LdW(w, cur, k, kWalkA), b = LdW(w, cur, k, kWalkB), c = LdW(w, cur, k, kWalkC),
                   d = LdW(w, cur, k, kWalkD), nn = LdW(w, cur, k, kWalkN);
      rng = RngFrom(LdW(w, cur, k, kWalkRng));
      SssWalkState walk;
      walk.sigma_t = vec3(a.x, a.y, a.z);
      walk.sigma_s = vec3(b.x, b.y, b.z);
      walk.throughput = vec3(a.w, b.w, c.w);
      walk.ray.o = vec3(c.x, c.y, c.z);
      walk.ray.d = vec3(d.x, d.y, d.z);
      walk.ray.tmin = d.w;
      walk.ray.tmax = kInf;
      walk.bounce = __float_as_uint(nn.x);
      pixel = __float_as_uint(nn.y);
      const float4 t4 = LdW(w, cur, k, kWalkThr), r4 = LdW(w, cur, k, kWalkRad);
      thr = vec3(t4.x, t4.y, t4.z); pdf_prev = t4.w;
      L = vec3(r4.x, r4.y, r4.z); depth = __float_as_uint(r4.w);
      HitT hit;
      SssStep st;
      do {
        vec3 cpdf;
        SssPrepareSegment(&rng, &walk, &cpdf);
        bool is_hit = false;
        if (SegmentIsClear(s, walk.ray.o, walk.ray.d, walk.ray.tmax * 1.001f)) {
          ++skipped;
        } else {
          is_hit = TraceClosest<false>(s, walk.ray, &hit, nullptr);
          ++counts[2];
        }
        st = SssFinishSegment(is_hit, hit.t, &rng, &walk, cpdf);
      } while (st == kSssContinue);
      if (st == kSssAbsorbed) {
        alive = false;   // throughput 0: the path ends with the radiance it holds
      } else {
        const RayT entry_ray = RayFrom(LdW(w, cur, k, kWalkRayO), LdW(w, cur, k, kWalkRayD));
        const HitT entry_hit = HitFrom(LdW(w, cur, k, kWalkHit));
        const Surface entry_si = MakeSurface(s, entry_ray, entry_hit);
        const Frame entry_frame = PrincipledFrame(entry_si);
        VertexResult vr;
        vr.P = entry_si.P;
        vr.shadow[1].active = false;
        SssFinish(s, entry_si, entry_frame, walk, hit, &rng, &vr);
        if (vr.shadow[1].active) {
          ++counts[1];
          if (!TraceAny<false>(s, vr.shadow[1].ray, nullptr)) L = L + thr * vr.shadow[1].contribute;
        }
        thr = vr.throughput * thr;
        pdf_prev = vr.pdf;
        ray.o = vr.P; ray.d = vr.wi; ray.tmin = 1e-3f; ray.tmax = kInf;
        depth += 1u;
      }
    }
    if (alive) L = PathRadianceFrom(s, ray, &rng, L, thr, pdf_prev, depth, counts);
    if (pixel != kNoPixel) {
      atomicAdd(&rgba[pixel], make_float4(L.x, L.y, L.z, 1.0f));
      ++retired;
    }
  }
  WarpTally(&w.stats[kStatClosest], uint32_t(counts[0]));
  WarpTally(&w.stats[kStatShadow], uint32_t(counts[1]));
  WarpTally(&w.stats[kStatSss], uint32_t(counts[2]));
  WarpTally(&w.stats[kStatSssSkipped], skipped);
  WarpTally(&w.stats[kStatRetired], retired);
}

// ------------------------------------------------------------------------------------------------ shadow rays
// Scene::AnyHit1 for every NEE request of this iteration; an unoccluded contribution is added to the radiance of its
// path wherever that path is now (S[next], D[next] or W[next]; the positions are nearly ascending).
struct ShadowClient {
  static constexpr bool kOrderByOrigin = false;
  const SceneView& s;
  const WaveState& w;
  uint32_t n, next, lanes;
  float4 c;
  bool has_result = false;
  uint32_t rays = 0;

  __device__ __forceinline__ ShadowClient(const SceneView& s_, const WaveState& w_, uint32_t next_parity)
      : s(s_), w(w_), n(w_.counters[kNumShadow]), next(next_parity), lanes(LanesFor(s_, w_.counters[kNumShadow])) {}
  __device__ __forceinline__ bool Fetches() const { return (threadIdx.x & 31u) < lanes; }
  __device__ __forceinline__ uint32_t RefillThreshold(uint32_t dflt) const { return ThresholdFor(lanes, dflt); }
  __device__ __forceinline__ bool Wants(const Trav&, bool exhausted) const { return has_result || (!exhausted && Fetches()); }
  __device__ __forceinline__ bool Refill(Trav& t, bool exhausted) {
    const unsigned lane = threadIdx.x & 31u;
    const bool need = !exhausted && !t.active && Fetches();
    const unsigned m_need = __ballot_sync(0xffffffffu, need);
    uint32_t fetch_base = 0;
    if (lane == 0u && m_need) fetch_base = atomicAdd(&w.counters[kFetchShadow], uint32_t(__popc(m_need)));
    if (!t.active && has_result) {   // overlaps the fetch round trip
      has_result = false;
      if (t.hit.prim == kInvalid) {
        const uint32_t code = __float_as_uint(c.w), tag = code >> 30, idx = code & 0x3fffffffu;
        float4* rec = tag == kTargetState ? &StateBuf(w, next)[size_t(kRad) * w.capacity + idx]
                                          : (tag == kTargetDone ? &DoneBuf(w, next)[idx]
                                                                : &WalkBuf(w, next)[size_t(kWalkRad) * w.capacity + idx]);
        float* dst = reinterpret_cast<float*>(rec);
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
      TravBegin(s, RayFrom(o, d), t);
      has_result = true;
      ++rays;
    }
    return need && slot >= n;
  }
  __device__ __forceinline__ void End(const Trav&) { WarpTally(&w.stats[kStatShadow], rays); }
};

template <bool HAS_CURVES>
__global__ void __launch_bounds__(128) TraceAnyKernel(SceneView s, WaveState w, uint32_t next_parity,
                                                      uint32_t refill_min_idle, uint32_t prim_min_lanes) {
  ShadowClient client(s, w, next_parity);
  TravEngine<true, HAS_CURVES, false>(s, client, refill_min_idle, prim_min_lanes);
}

// ------------------------------------------------------------------------------------------------ test hooks
// pbrgpu_shade: after the single iteration every path sits somewhere in S[p] / D[p] (p = 0, 1); scatter the records
// back to path order.  out16 layout: see pbrgpu.h (hit flag and face / t columns are filled by the caller).
__global__ void GatherVertexKernel(WaveState w, uint32_t parity, uint32_t n_paths, float* out16) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_state = w.counters[kNumActive0 + parity], n_done = w.counters[kNumDone0 + parity];
  if (i < n_state) {
    const uint32_t pixel = __float_as_uint(LdS(w, parity, i, kPix).x);
    if (pixel < n_paths) {
      const float4 o = LdS(w, parity, i, kRayO), d = LdS(w, parity, i, kRayD), t = LdS(w, parity, i, kThr),
                   r = LdS(w, parity, i, kRad);
      float* q = out16 + size_t(16) * pixel;
      q[1] = d.x; q[2] = d.y; q[3] = d.z;
      q[4] = t.x; q[5] = t.y; q[6] = t.z;
      q[7] = r.x; q[8] = r.y; q[9] = r.z;
      q[10] = t.w;
      q[11] = o.x; q[12] = o.y; q[13] = o.z;
    }
  }
  if (i < n_done) {
    const float4 r = __ldcs(&DoneBuf(w, parity)[i]);
    const uint32_t pixel = __float_as_uint(r.w);
    if (pixel < n_paths) {
      float* q = out16 + size_t(16) * pixel;
      q[7] = r.x; q[8] = r.y; q[9] = r.z;
    }
  }
}

// pbrgpu_trace / pbrgpu_occluded: caller-supplied ray batches through the same engine as the render kernels
struct BatchClient {
  static constexpr bool kOrderByOrigin = true;
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
  __device__ __forceinline__ uint32_t RefillThreshold(uint32_t dflt) const { return dflt; }
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
          TravBegin(s, RayFrom(o, d), t);
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
