// Warp-level traversal engine (device only): the same closest-hit / any-hit walk of the compressed 8-wide BVH as
// TraverseBvh (traverse.cuh) — same node test, same primitive tests, same visiting order, hence the same hits — but
// written as a RESUMABLE per-lane state machine so that a persistent warp can replace a finished ray by a new one
// while its neighbours are still traversing.
//
// Why: measured on B200 (profiles/r1a_ncu.md) the one-ray-per-lane-until-all-32-finish form executed its node step
// with 7.0 and its triangle test with 3.0 of 32 lanes active: path-tracing rays have wildly different traversal
// lengths (2 nodes for a wall, 40 for Lucy) and the warp waits for its slowest lane.  Here a lane that finishes goes
// idle only until `refill_min_idle` lanes are idle; then the warp runs one converged section that consumes the
// finished results and fetches new rays (one atomicAdd per warp), and traversal resumes with a full warp.
//
// Replaces Scene::TraceFirstHit1 / AnyHit1 -> rtcIntersect1 / rtcOccluded1 (reference src/scene.cc:261-268,
// src/raytracer/raytracer_impl.cc:268-287); hit semantics are Embree's, see traverse.cuh.
#pragma once
#include "traverse.cuh"

#if defined(__CUDACC__)
namespace pbr {

struct Trav {
  vec3 O, D, inv_d, ood;   // origin, direction, 1/d (zero components nudged), o/d
  float tmin, tfar;
  uint32_t oct_inv4;       // octant-inversion mask replicated in 4 bytes (bit 2 = +x, 1 = +y, 0 = +z)
  uint2 group;             // current node group: (child base, hit mask << 24 | imask)
  uint2 pgroup;            // pending primitives of the last node: (leaf-order base, bit mask)
  int sp;                  // stack entries in use
  bool active;             // still traversing
  bool curves;             // walking the curve BVH (after the triangle BVH)
  bool checked;            // curves: the pending primitives in pgroup have been through CurveMayHit (survivors only)
  uint32_t held;           // curve part ((slot << 2) | first quad) that passed CurveMayHit and waits for the ribbon test; kInvalid: none
  HitT hit;                // closest hit so far (prim == kInvalid: none)
  uint32_t n_nodes, n_prims;   // STATS only
};

__device__ __forceinline__ void TravBegin(const SceneView& s, const RayT& ray, Trav& t) {
  t.O = ray.o; t.D = ray.d;
  t.tmin = ray.tmin; t.tfar = ray.tmax;
  const float tiny = 1e-30f;
  const vec3 ds(fabsf(ray.d.x) < tiny ? copysignf(tiny, ray.d.x) : ray.d.x,
                fabsf(ray.d.y) < tiny ? copysignf(tiny, ray.d.y) : ray.d.y,
                fabsf(ray.d.z) < tiny ? copysignf(tiny, ray.d.z) : ray.d.z);
  t.inv_d = vec3(1.0f / ds.x, 1.0f / ds.y, 1.0f / ds.z);
  t.ood = vec3(ray.o.x * t.inv_d.x, ray.o.y * t.inv_d.y, ray.o.z * t.inv_d.z);
  t.oct_inv4 = (ds.x < 0.f ? 0u : 0x04040404u) | (ds.y < 0.f ? 0u : 0x02020202u) | (ds.z < 0.f ? 0u : 0x01010101u);
  t.hit.prim = kInvalid;
  t.hit.t = 0.f; t.hit.u = 0.f; t.hit.v = 0.f;
  t.sp = 0;
  t.curves = (s.num_tris == 0u);
  t.group = make_uint2(0u, 0x80000000u);   // the root as the only child of a virtual parent (see TraverseBvh)
  t.pgroup = make_uint2(0u, 0u);
  t.held = kInvalid;
  t.checked = false;
  t.active = (s.num_tris | s.num_curves) != 0u;
}

// TraverseBvh's loop body split in two so that a warp can batch the primitive tests of its lanes:
//   TravNodeStep   (precondition: no pending primitives) pops one child of the current group, tests its 8 boxes and
//                  leaves the hit leaves' primitives in t.pgroup;
//   TravPrimStep   tests ONE pending primitive.
// A lane always finishes the primitives of a node before its next node step, exactly like TraverseBvh, so tfar
// shrinks in the same order and the reported hit is the same.
// BY_ORIGIN: curve-BVH nodes hand the children whose box holds the ray origin to the ray first (PopChild); the engine
// asks for it where the order pays, the closest-hit rays of a path (not for shadow rays: any hit ends them; not for
// walk segments: they live inside triangle meshes and would only pay for the extra mask, measured +14 % walk time on C4)
template <bool HAS_CURVES, bool STATS, bool PREFETCH, bool BY_ORIGIN = false>
__device__ __forceinline__ void TravNodeStep(const SceneView& s, Trav& t, uint2* __restrict__ stack) {
  const bool in_curves = HAS_CURVES && t.curves;
  const float4* __restrict__ nodes = in_curves ? s.curve_nodes : s.tri_nodes;
  const uint32_t hits_imask = t.group.y;
  // (curve BVH: the children whose box holds the ray origin first, traverse.cuh: PopChild)
  const uint32_t child_bit = in_curves ? PopChild<true>(&t.group.y) : PopChild<false>(&t.group.y);
  if (t.group.y & 0xff000000u) {
    if (t.sp < kStackSize) stack[t.sp++] = t.group;
  }
  const uint32_t slot = (child_bit - 24u) ^ (t.oct_inv4 & 0xffu);
  const uint32_t rel = popc(hits_imask & ~(0xffffffffu << slot));
  const uint32_t node = t.group.x + rel;
  const float4 n0 = nodes[node * 5 + 0], n1 = nodes[node * 5 + 1], n2 = nodes[node * 5 + 2];
  const float4 n3 = nodes[node * 5 + 3], n4 = nodes[node * 5 + 4];
  if (STATS) t.n_nodes++;
  const bool neg_x = !(t.oct_inv4 & 0x04u), neg_y = !(t.oct_inv4 & 0x02u), neg_z = !(t.oct_inv4 & 0x01u);
  uint32_t inside = 0;
  const uint32_t hitmask = (HAS_CURVES && BY_ORIGIN) ? NodeIntersectT<true>(t.ood, t.inv_d, t.oct_inv4, neg_x, neg_y, neg_z, t.tmin, t.tfar, n0,
                                                             n1, n2, n3, n4, s.bias_magic, &inside)
                                      : NodeIntersectT<false>(t.ood, t.inv_d, t.oct_inv4, neg_x, neg_y, neg_z, t.tmin, t.tfar,
                                                              n0, n1, n2, n3, n4, s.bias_magic, nullptr);
  t.group.x = f2u(n1.x);
  t.group.y = (hitmask & 0xff000000u) | ((in_curves && BY_ORIGIN && s.inside_first) ? inside : 0u) | extract_byte(f2u(n0.w), 3);
  t.pgroup.x = f2u(n1.y);
  t.pgroup.y = hitmask & 0x00ffffffu;
  t.checked = false;
  if ((t.group.y & 0xff000000u) == 0u) {
    if (t.sp > 0) t.group = stack[--t.sp];
    else t.group.y = 0u;
  }
  // The node this lane will test next is known now; its primitives (if any) are tested first.  Pull the node's 80
  // bytes towards L1 in the meantime (two lines: nodes are 16-byte aligned): the long-scoreboard wait on the node
  // fetch was the top stall of the ray kernels.
  if (PREFETCH && (t.group.y & 0xff000000u)) {
    const uint32_t nbit = msb(t.group.y);
    const uint32_t nslot = (nbit - 24u) ^ (t.oct_inv4 & 0xffu);
    const uint32_t nnode = t.group.x + popc(t.group.y & ~(0xffffffffu << nslot));
    const float4* addr = nodes + size_t(nnode) * 5;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(addr));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(addr + 4));
  }
}

// the ray is through with the current BVH: go on to the curves, or finish
template <bool HAS_CURVES>
__device__ __forceinline__ void TravAdvance(const SceneView& s, Trav& t) {
  if ((t.group.y & 0xff000000u) == 0u && t.pgroup.y == 0u) {
    if (HAS_CURVES && !t.curves && s.num_curves) {   // triangles done: the shortened ray goes through the curves
      t.curves = true;
      t.group = make_uint2(0u, 0x80000000u);
    } else {
      t.active = false;
    }
  }
}

// One pending TRIANGLE of the lane.
template <bool ANY, bool STATS>
__device__ __forceinline__ void TravTriStep(const SceneView& s, Trav& t) {
  const uint32_t bit = msb(t.pgroup.y);
  t.pgroup.y &= ~(1u << bit);
  const uint32_t idx = t.pgroup.x + bit;
  if (STATS) t.n_prims++;
  const float4* __restrict__ prims = s.tri_data;
  const float4 a = prims[idx * 3 + 0], b = prims[idx * 3 + 1], c = prims[idx * 3 + 2];
  float ht, hu, hv;
  if (IntersectTriangle(t.O, t.D, t.tmin, t.tfar, from4(a), from4(b), from4(c), &ht, &hu, &hv)) {
    t.hit.t = ht; t.hit.u = hu; t.hit.v = hv;
    t.hit.prim = idx;
    t.tfar = ht;
    if (ANY) t.active = false;
  }
}

// Pending CURVE candidates, warp-cooperatively.  Leaf boxes of thin diagonal segments are much fatter than the
// ribbon: two of three candidates fail the cheap line-distance test (CurveMayHit).  Candidates are spread unevenly
// (0 .. 10 per lane and node step), so a per-lane loop ran with 3-4 of 32 lanes and as many rounds as the busiest
// lane had candidates (profiles/r1o_hair_c3_ncu.md: a quarter of the kernel's instructions, a third of its stall samples).  Here the
// candidates of ALL lanes are numbered by a prefix sum and tested 32 at a time, one per lane: the tester fetches the
// owner's ray by shuffle and reports a pass by setting the candidate's bit in the owner's word of `pass_row` (shared
// memory, one word per lane).  Afterwards a lane's pgroup holds the survivors only (t.checked).
// Called by all 32 lanes, converged.  `fresh`: this lane has unchecked candidates.
template <bool STATS>
__device__ __forceinline__ void TravCurveCullWarp(const SceneView& s, Trav& t, bool fresh, uint32_t* pass_row) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t cnt = fresh ? popc(t.pgroup.y) : 0u;
  uint32_t incl = cnt;
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  const uint32_t excl = incl - cnt;
  pass_row[lane] = 0u;
  __syncwarp();
  const float4* __restrict__ cull = s.curve_cull;
  for (uint32_t base = 0; base < total; base += 32u) {
    const uint32_t j = base + lane;
    // owner of candidate j: the first lane whose inclusive count exceeds j
    uint32_t lo = 0;
#pragma unroll
    for (uint32_t step = 16; step >= 1; step >>= 1) {
      const uint32_t v = __shfl_sync(0xffffffffu, incl, lo + step - 1u);
      if (v <= j) lo += step;
    }
    const uint32_t owner = lo & 31u;
    const uint32_t k = j - __shfl_sync(0xffffffffu, excl, owner);
    const uint32_t oy = __shfl_sync(0xffffffffu, t.pgroup.y, owner);
    const uint32_t ox = __shfl_sync(0xffffffffu, t.pgroup.x, owner);
    const vec3 O(__shfl_sync(0xffffffffu, t.O.x, owner), __shfl_sync(0xffffffffu, t.O.y, owner),
                 __shfl_sync(0xffffffffu, t.O.z, owner));
    const vec3 D(__shfl_sync(0xffffffffu, t.D.x, owner), __shfl_sync(0xffffffffu, t.D.y, owner),
                 __shfl_sync(0xffffffffu, t.D.z, owner));
    if (j < total) {
      const uint32_t bit = __fns(oy, 0u, int(k) + 1);            // the (k+1)-th pending primitive of the owner
      const uint32_t slot = s.curve_sub[ox + bit] >> 2;
      if (!cull || CurveMayHit(O, D, cull[slot * 2], cull[slot * 2 + 1])) atomicOr(&pass_row[owner], 1u << bit);
    }
  }
  __syncwarp();
  if (fresh) {
    if (STATS) t.n_prims += cnt;
    t.pgroup.y &= pass_row[lane];
    t.checked = true;
  }
  __syncwarp();
}

// next survivor of the lane's (checked) pending candidates -> held for the ribbon test
__device__ __forceinline__ void TravCurveTake(const SceneView& s, Trav& t) {
  const uint32_t bit = msb(t.pgroup.y);
  t.pgroup.y &= ~(1u << bit);
  t.held = s.curve_sub[t.pgroup.x + bit];   // (slot << 2) | first quad of the part
}

template <bool ANY>
__device__ __forceinline__ void TravRibbonStep(const SceneView& s, Trav& t) {
  const uint32_t code = t.held, idx = code >> 2;
  t.held = kInvalid;
  const float4* __restrict__ prims = s.curve_data;
  const float4 c0 = prims[idx * 4 + 0], c1 = prims[idx * 4 + 1], c2 = prims[idx * 4 + 2], c3 = prims[idx * 4 + 3];
  const CurveRaySpace rs = MakeCurveRaySpace(t.D);
  float ht, hu, hv;
  if (IntersectCurve(t.O, rs, t.tmin, t.tfar, c0, c1, c2, c3, code & 3u, s.curve_part_quads, &ht, &hu, &hv)) {
    t.hit.t = ht; t.hit.u = hu; t.hit.v = hv;
    t.hit.prim = idx | kCurveFlag;
    t.tfar = ht;
    if (ANY) t.active = false;
  }
}

// One trip of the engine between two converged sections: a node phase (lanes without pending primitives), the curve
// candidate rejection, a primitive phase and the ribbon phase (see TravEngine).  Called by all 32 lanes.
template <bool ANY, bool HAS_CURVES, bool STATS, bool BY_ORIGIN = false>
__device__ __forceinline__ void TravTrip(const SceneView& s, Trav& t, uint2* __restrict__ stack, uint32_t* pass_row,
                                         uint32_t prim_min_lanes) {
  const bool node_work = t.active && t.pgroup.y == 0u && (!HAS_CURVES || t.held == kInvalid);
  if (node_work) {
    TravNodeStep<HAS_CURVES, STATS, false, BY_ORIGIN && !ANY>(s, t, stack);   // prefetching the next node was measured: 1.6x slower
    TravAdvance<HAS_CURVES>(s, t);
  }
  if (HAS_CURVES) {
    // curve candidates: cheap rejection now, across the warp; a survivor (if any) is held for the ribbon phase
    const bool fresh = t.active && t.curves && !t.checked && t.held == kInvalid && t.pgroup.y != 0u;
    if (__ballot_sync(0xffffffffu, fresh) != 0u) TravCurveCullWarp<STATS>(s, t, fresh, pass_row);
    if (t.active && t.curves && t.checked && t.held == kInvalid) {
      if (t.pgroup.y != 0u) TravCurveTake(s, t);
      else TravAdvance<HAS_CURVES>(s, t);
    }
  }
  const bool tri_work = t.active && t.pgroup.y != 0u && !(HAS_CURVES && t.curves);
  const unsigned prim = __ballot_sync(0xffffffffu, tri_work);
  const unsigned node = __ballot_sync(0xffffffffu, t.active && t.pgroup.y == 0u && (!HAS_CURVES || t.held == kInvalid));
  if (prim != 0u && (uint32_t(__popc(prim)) >= prim_min_lanes || node == 0u)) {
    if (tri_work) {
      TravTriStep<ANY, STATS>(s, t);
      if (t.active) TravAdvance<HAS_CURVES>(s, t);
    }
  }
  if (HAS_CURVES) {
    // ribbon phase: when enough lanes hold a candidate, or when nobody can make progress without it
    const bool holding = t.active && t.held != kInvalid;
    const unsigned hold = __ballot_sync(0xffffffffu, holding);
    const unsigned free_lanes = __ballot_sync(0xffffffffu, t.active && t.held == kInvalid);
    if (hold != 0u && (uint32_t(__popc(hold)) >= s.ribbon_min_lanes || free_lanes == 0u)) {
      if (holding) {
        TravRibbonStep<ANY>(s, t);
        if (t.active) TravAdvance<HAS_CURVES>(s, t);
      }
    }
  }
}

// Lanes of a warp that fetch work when a launch has `n` items for `gridDim.x * blockDim.x / 32` warps: all 32 as long
// as there is more work than lanes; with less (the last iterations of a frame: a few thousand paths and walks, each a
// long dependent chain) the items are spread over ALL warps, n / warps per warp.  A warp that carries one walk runs
// only that walk's instructions; a warp that carries 32 runs the scatter step and seven traversal trips for every
// bounce of any of them, and the frame ends when the slowest chain does.
__device__ __forceinline__ uint32_t LanesFor(const SceneView& s, uint32_t n) {
  if (!s.thin_spread) return 32u;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  const uint32_t per_warp = (n + warps - 1u) / warps;
  return per_warp < 1u ? 1u : (per_warp > 32u ? 32u : per_warp);
}
__device__ __forceinline__ uint32_t ThresholdFor(uint32_t lanes, uint32_t dflt) {
  const uint32_t t = (lanes * 3u + 3u) / 4u;   // three quarters of the lanes in use
  return t < dflt ? t : dflt;
}

// The engine loop.  `Client` supplies the converged section:
//   bool Wants(const Trav& t, bool exhausted)  — per lane, for idle lanes: could a Refill give this lane work
//        (a finished walk segment to continue, or a work source that is not dry yet)?
//   bool Refill(Trav& t, bool exhausted)  — called by ALL 32 lanes, converged, so warp collectives are allowed.
//        A lane with !t.active (1) consumes the result of the ray it just finished, if any, and (2) tries to obtain
//        its next ray: on success it calls TravBegin (t.active == true).  Returns true when this lane found the
//        work source empty.
//   void End(const Trav& t)  — after the loop.
//   static constexpr bool kOrderByOrigin  — see TravNodeStep.
//   uint32_t RefillThreshold(uint32_t dflt)  — idle lanes that trigger a Refill (dflt, or fewer when the client lets
//        only some lanes of a warp fetch work, see LanesFor).
// A Refill happens when at least `refill_min_idle` idle lanes want one, or when nobody is traversing.  The loop ends
// when no lane is active after a Refill and the source is dry.
// Between refills every trip runs a node phase (lanes without pending primitives) and, when at least
// `prim_min_lanes` lanes hold pending primitives or no lane has node work, a primitive phase in which each of them
// tests one: leaf hits are rare per lane and step (measured: 3-5 of 32 lanes), so testing them the moment they
// appear would run the intersection code nearly serially.
template <bool ANY, bool HAS_CURVES, bool STATS, class Client>
__device__ __forceinline__ void TravEngine(const SceneView& s, Client& client, uint32_t refill_min_idle,
                                           uint32_t prim_min_lanes) {
  uint2 stack[kStackSize];
  __shared__ uint32_t pass_words[HAS_CURVES ? 128 : 1];   // TravCurveCullWarp: one word per thread of the (128-thread) block
  uint32_t* pass_row = pass_words + (HAS_CURVES ? (threadIdx.x & ~31u) : 0u);
  Trav t;
  t.active = false;
  t.checked = false;
  t.pgroup = make_uint2(0u, 0u);
  t.held = kInvalid;
  t.n_nodes = 0; t.n_prims = 0;
  bool exhausted = false;   // warp-uniform: the work source ran dry
  refill_min_idle = client.RefillThreshold(refill_min_idle);
  for (;;) {
    const unsigned act = __ballot_sync(0xffffffffu, t.active);
    const unsigned want = __ballot_sync(0xffffffffu, !t.active && client.Wants(t, exhausted));
    if (act == 0u || uint32_t(__popc(want)) >= refill_min_idle) {
      const bool dry = client.Refill(t, exhausted);
      exhausted = __any_sync(0xffffffffu, dry) || exhausted;
      if (__ballot_sync(0xffffffffu, t.active) == 0u) {
        // nobody traverses: done once the source is dry and no idle lane still holds work for another Refill
        // (a walk whose segment was answered without a query is idle but not finished)
        if (exhausted && __ballot_sync(0xffffffffu, client.Wants(t, exhausted)) == 0u) break;
        continue;
      }
    }
    TravTrip<ANY, HAS_CURVES, STATS, Client::kOrderByOrigin>(s, t, stack, pass_row, prim_min_lanes);
  }
  client.End(t);
}

}  // namespace pbr
#endif
