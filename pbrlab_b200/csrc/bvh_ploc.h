// Parallel builder of the 8-wide compressed BVH (the node format of bvh_builder.h), written as per-element BODIES that
// compile both for the GPU (bvh_device.cuh wraps each body in a kernel: one thread per element) and for the host
// (BuildBvh8PlocHost below runs the same bodies in loops: the CPU tests check the trees it produces against Embree
// through the emulated traversal, so the algorithm and the node encoding are validated without a GPU).
//
// Why: Embree's rtcCommitScene (reference src/raytracer/raytracer_impl.cc:81,147,192) is replaced on the host by a
// binned-SAH builder (bvh_builder.cc) that needs 18 s for the 20 M triangles of configuration C5 on 16 threads, in front
// of a frame that renders in seconds.  This builder is made of data-parallel passes only:
//   1. scene bounds, 63-bit Morton codes of the primitive centroids, radix sort;
//   2. PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): every cluster looks `radius` neighbours to
//      each side in Morton order for the partner with the smallest merged surface area, mutual pairs merge, the cluster
//      array is compacted; repeated until one cluster is left.  Bottom-up agglomeration guided by surface area:
//      trees close to SAH quality, unlike the plain Morton splits of an LBVH;
//   3. level-synchronous collapse into 8-wide nodes: a wide node opens the inner child of largest area until it has 8
//      children, then uses spare slots to split multi-primitive leaves; slots by octant, boxes quantised to the
//      node's 8-bit grid; children and leaf primitives of a node are contiguous (breadth-first layout).
// Every pass is deterministic (scans instead of atomics for allocation), so all ranks and devices build the same tree.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "bvh_builder.h"

#if defined(__CUDACC__)
#define PLOC_HD __host__ __device__ __forceinline__
#else
#define PLOC_HD inline
#endif

namespace pbrploc {

struct alignas(16) F4 { float x, y, z, w; };
constexpr uint32_t kLeafTag = 0xFFFFFFFFu;   // nlo.w of a binary leaf; its nhi.w holds the primitive index
constexpr int kMaxLevels = 31;               // wide levels the traversal stack can hold (bvh_builder.cc: kErrDepth)

PLOC_HD uint32_t FBits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PLOC_HD float BFloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
PLOC_HD float Min2(float a, float b) { return a < b ? a : b; }
PLOC_HD float Max2(float a, float b) { return a > b ? a : b; }
PLOC_HD float HalfArea(const F4& lo, const F4& hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
PLOC_HD float UnionHalfArea(const F4& alo, const F4& ahi, const F4& blo, const F4& bhi) {
  const float dx = Max2(ahi.x, bhi.x) - Min2(alo.x, blo.x), dy = Max2(ahi.y, bhi.y) - Min2(alo.y, blo.y),
              dz = Max2(ahi.z, bhi.z) - Min2(alo.z, blo.z);
  return dx * dy + dy * dz + dz * dx;
}

// ---- pass 1: Morton codes
PLOC_HD uint64_t Spread21(uint64_t v) {   // bit i -> bit 3 i
  v &= 0x1fffffull;
  v = (v | (v << 32)) & 0x1f00000000ffffull;
  v = (v | (v << 16)) & 0x1f0000ff0000ffull;
  v = (v | (v << 8)) & 0x100f00f00f00f00full;
  v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
// blo: scene lower corner; scale: 2^21 / extent per axis (0 for a flat axis)
PLOC_HD void MortonBody(const pbrbvh::Aabb* prim, const float* blo, const float* scale, uint32_t i, uint64_t* key,
                        uint32_t* idx) {
  const pbrbvh::Aabb& b = prim[i];
  uint64_t q[3];
  for (int k = 0; k < 3; ++k) {
    const float c = 0.5f * (b.lo[k] + b.hi[k]);
    float f = (c - blo[k]) * scale[k];
    f = f < 0.f ? 0.f : (f > 2097151.f ? 2097151.f : f);
    q[k] = uint64_t(uint32_t(f));
  }
  key[i] = (Spread21(q[0]) << 2) | (Spread21(q[1]) << 1) | Spread21(q[2]);
  idx[i] = i;
}

// ---- pass 2: PLOC.  Binary nodes: leaves 0 .. n-1 in Morton order, inner nodes n .. 2n-2 in creation order (the
// root is the last).  nlo = (box.lo, left child | kLeafTag), nhi = (box.hi, right child | primitive index).
PLOC_HD void InitClusterBody(const pbrbvh::Aabb* prim, const uint32_t* sorted_idx, uint32_t i, F4* nlo, F4* nhi,
                             uint32_t* ncount, uint32_t* cnode, F4* clo, F4* chi) {
  const uint32_t p = sorted_idx[i];
  const pbrbvh::Aabb& b = prim[p];
  F4 lo = {b.lo[0], b.lo[1], b.lo[2], BFloat(kLeafTag)};
  F4 hi = {b.hi[0], b.hi[1], b.hi[2], BFloat(p)};
  nlo[i] = lo; nhi[i] = hi;
  ncount[i] = 1u;
  cnode[i] = i;
  clo[i] = lo; chi[i] = hi;
}

// nearest neighbour of cluster i within `radius` positions: smallest merged half-area, lowest index on ties (with
// that rule the pair of globally smallest distance whose lower member has the lowest index is always mutual: every
// round merges at least one pair)
PLOC_HD void NearestBody(const F4* clo, const F4* chi, uint32_t m, uint32_t radius, uint32_t i, uint32_t* nn) {
  const F4 alo = clo[i], ahi = chi[i];
  const uint32_t j0 = i > radius ? i - radius : 0u;
  const uint32_t j1 = (i + radius < m - 1u) ? i + radius : m - 1u;
  float best = FLT_MAX;
  uint32_t bj = i;
  for (uint32_t j = j0; j <= j1; ++j) {
    if (j == i) continue;
    const float a = UnionHalfArea(alo, ahi, clo[j], chi[j]);
    if (a < best) { best = a; bj = j; }
  }
  nn[i] = bj;
}

// create[i] = 1: cluster i is the lower member of a mutual pair (allocates the new node); keep[i] = 0: the upper member
PLOC_HD void PairFlagsBody(const uint32_t* nn, uint32_t i, uint32_t* create_i, uint32_t* keep_i) {
  const uint32_t j = nn[i];
  const bool mutual = (j != i) && (nn[j] == i);
  *create_i = (mutual && i < j) ? 1u : 0u;
  *keep_i = (mutual && i > j) ? 0u : 1u;
}

// create_i / keep_i: the flags of element i; scan_create_i / scan_keep_i: their exclusive prefix sums
PLOC_HD void MergeBody(const uint32_t* nn, uint32_t create_i, uint32_t keep_i, uint32_t scan_create_i,
                       uint32_t scan_keep_i, uint32_t first_new_node, const uint32_t* cnode, const F4* clo,
                       const F4* chi, uint32_t i, F4* nlo, F4* nhi, uint32_t* ncount, uint32_t* cnode_out, F4* clo_out,
                       F4* chi_out) {
  if (!keep_i) return;
  const uint32_t dst = scan_keep_i;
  if (create_i) {
    const uint32_t j = nn[i];
    const uint32_t id = first_new_node + scan_create_i;
    const F4 a0 = clo[i], a1 = chi[i], b0 = clo[j], b1 = chi[j];
    F4 lo = {Min2(a0.x, b0.x), Min2(a0.y, b0.y), Min2(a0.z, b0.z), BFloat(cnode[i])};
    F4 hi = {Max2(a1.x, b1.x), Max2(a1.y, b1.y), Max2(a1.z, b1.z), BFloat(cnode[j])};
    nlo[id] = lo; nhi[id] = hi;
    ncount[id] = ncount[cnode[i]] + ncount[cnode[j]];
    cnode_out[dst] = id;
    clo_out[dst] = lo; chi_out[dst] = hi;
  } else {
    cnode_out[dst] = cnode[i];
    clo_out[dst] = clo[i]; chi_out[dst] = chi[i];
  }
}

// ---- pass 3: collapse into 8-wide nodes
struct WideChildren {
  uint32_t node[8];
  int n;
};
PLOC_HD bool IsWideLeaf(const uint32_t* ncount, uint32_t node, uint32_t max_leaf) { return ncount[node] <= max_leaf; }

PLOC_HD WideChildren GatherChildren(const F4* nlo, const F4* nhi, const uint32_t* ncount, uint32_t node,
                                    uint32_t max_leaf) {
  WideChildren c;
  c.n = 0;
  if (IsWideLeaf(ncount, node, max_leaf)) {   // a lone leaf root
    c.node[c.n++] = node;
    return c;
  }
  c.node[c.n++] = FBits(nlo[node].w);
  c.node[c.n++] = FBits(nhi[node].w);
  // (1) open the inner child of largest surface area (bvh_builder.cc does the same on the SAH tree)
  while (c.n < 8) {
    int pick = -1;
    float pick_area = -1.f;
    for (int i = 0; i < c.n; ++i) {
      if (!IsWideLeaf(ncount, c.node[i], max_leaf)) {
        const float a = HalfArea(nlo[c.node[i]], nhi[c.node[i]]);
        if (a > pick_area) { pick_area = a; pick = i; }
      }
    }
    if (pick < 0) break;
    const uint32_t open = c.node[pick];
    c.node[pick] = FBits(nlo[open].w);
    c.node[c.n++] = FBits(nhi[open].w);
  }
  // (2) spare slots: split the multi-primitive leaf whose box is largest — its primitives get boxes of their own in
  // the same node test, at no extra node
  while (c.n < 8) {
    int pick = -1;
    float pick_area = -1.f;
    for (int i = 0; i < c.n; ++i) {
      const uint32_t nd = c.node[i];
      if (IsWideLeaf(ncount, nd, max_leaf) && ncount[nd] > 1u) {
        const float a = HalfArea(nlo[nd], nhi[nd]) * float(ncount[nd]);
        if (a > pick_area) { pick_area = a; pick = i; }
      }
    }
    if (pick < 0) break;
    const uint32_t open = c.node[pick];
    c.node[pick] = FBits(nlo[open].w);
    c.node[c.n++] = FBits(nhi[open].w);
  }
  return c;
}

PLOC_HD void WideCountBody(const F4* nlo, const F4* nhi, const uint32_t* ncount, const uint32_t* wide_item,
                           uint32_t level_begin, uint32_t t, uint32_t max_leaf, uint32_t* n_inner_t, uint32_t* n_prims_t) {
  const WideChildren c = GatherChildren(nlo, nhi, ncount, wide_item[level_begin + t], max_leaf);
  uint32_t ni = 0, np = 0;
  for (int i = 0; i < c.n; ++i) {
    if (IsWideLeaf(ncount, c.node[i], max_leaf)) np += ncount[c.node[i]];
    else ++ni;
  }
  *n_inner_t = ni;
  *n_prims_t = np;
}

// primitives of a leaf subtree (<= 3 of them), left to right
PLOC_HD uint32_t LeafPrims(const F4* nlo, const F4* nhi, uint32_t node, uint32_t out[4]) {
  uint32_t stack[4];
  int sp = 0;
  uint32_t n = 0;
  stack[sp++] = node;
  while (sp > 0 && n < 4u) {
    const uint32_t nd = stack[--sp];
    if (FBits(nlo[nd].w) == kLeafTag) {
      out[n++] = FBits(nhi[nd].w);
    } else {
      if (sp < 3) { stack[sp++] = FBits(nhi[nd].w); stack[sp++] = FBits(nlo[nd].w); }
    }
  }
  return n;
}

// writes wide node (level_begin + t): 20 words, see bvh_builder.h for the layout
PLOC_HD void WideEmitBody(const F4* nlo, const F4* nhi, const uint32_t* ncount, uint32_t* wide_item,
                          uint32_t level_begin, uint32_t level_end, uint32_t t, uint32_t max_leaf,
                          uint32_t scan_inner_t, uint32_t scan_prims_t, uint32_t prim_total,
                          uint32_t* nodes_out, uint32_t* prim_order) {
  const uint32_t self = wide_item[level_begin + t];
  const WideChildren c = GatherChildren(nlo, nhi, ncount, self, max_leaf);
  const F4 blo = nlo[self], bhi = nhi[self];
  // slot assignment: greedy on dot(child centre - node centre, octant direction of the slot)
  int slot_of[8];
  {
    float cost[8][8];
    const float ncx = 0.5f * (blo.x + bhi.x), ncy = 0.5f * (blo.y + bhi.y), ncz = 0.5f * (blo.z + bhi.z);
    for (int k = 0; k < c.n; ++k) {
      const F4 lo = nlo[c.node[k]], hi = nhi[c.node[k]];
      const float dx = 0.5f * (lo.x + hi.x) - ncx, dy = 0.5f * (lo.y + hi.y) - ncy, dz = 0.5f * (lo.z + hi.z) - ncz;
      for (int s = 0; s < 8; ++s) cost[k][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
    }
    uint32_t cdone = 0, sdone = 0;
    for (int k = 0; k < c.n; ++k) {
      int bc = -1, bs = -1;
      float bv = -FLT_MAX;
      for (int cc = 0; cc < c.n; ++cc) {
        if (cdone & (1u << cc)) continue;
        for (int s = 0; s < 8; ++s) {
          if (sdone & (1u << s)) continue;
          if (bc < 0 || cost[cc][s] > bv) { bv = cost[cc][s]; bc = cc; bs = s; }
        }
      }
      cdone |= 1u << bc;
      sdone |= 1u << bs;
      slot_of[bc] = bs;
    }
  }
  int child_in_slot[8];
  for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
  for (int k = 0; k < c.n; ++k) child_in_slot[slot_of[k]] = k;

  // quantisation grid: step 2^e with extent / 2^e <= 255
  uint32_t ebyte[3];
  double step[3];
  const double nlo3[3] = {double(blo.x), double(blo.y), double(blo.z)};
  const double ext3[3] = {double(bhi.x) - double(blo.x), double(bhi.y) - double(blo.y), double(bhi.z) - double(blo.z)};
  for (int k = 0; k < 3; ++k) {
    int e = -126;
    if (ext3[k] > 0.0) {
      e = int(ceil(log2(ext3[k] / 255.0)));
      while (ext3[k] / ldexp(1.0, e) > 255.0) ++e;
      if (e < -126) e = -126;
    }
    ebyte[k] = uint32_t(e + 127);
    step[k] = ldexp(1.0, e);
  }
  uint32_t meta[8], q[6][8];
  for (int s = 0; s < 8; ++s) { meta[s] = 0; for (int k = 0; k < 6; ++k) q[k][s] = 0; }
  uint32_t imask = 0;
  const uint32_t child_base = level_end + scan_inner_t;
  const uint32_t prim_base = prim_total + scan_prims_t;
  uint32_t n_inner = 0, n_prims = 0;
  for (int s = 0; s < 8; ++s) {
    const int k = child_in_slot[s];
    if (k < 0) continue;
    const uint32_t cn = c.node[k];
    const F4 lo = nlo[cn], hi = nhi[cn];
    const double clo3[3] = {double(lo.x), double(lo.y), double(lo.z)}, chi3[3] = {double(hi.x), double(hi.y), double(hi.z)};
    for (int a = 0; a < 3; ++a) {
      double l = floor((clo3[a] - nlo3[a]) / step[a]);
      double h = ceil((chi3[a] - nlo3[a]) / step[a]);
      l = l < 0.0 ? 0.0 : (l > 255.0 ? 255.0 : l);
      h = h < 0.0 ? 0.0 : (h > 255.0 ? 255.0 : h);
      q[a][s] = uint32_t(l);
      q[3 + a][s] = uint32_t(h);
    }
    if (!IsWideLeaf(ncount, cn, max_leaf)) {
      imask |= 1u << s;
      meta[s] = (1u << 5) | (24u + uint32_t(s));
      wide_item[child_base + n_inner] = cn;
      ++n_inner;
    } else {
      uint32_t prims[4];
      const uint32_t cnt = LeafPrims(nlo, nhi, cn, prims);
      const uint32_t unary = (cnt == 1u) ? 1u : (cnt == 2u ? 3u : 7u);
      meta[s] = (unary << 5) | n_prims;
      for (uint32_t i = 0; i < cnt; ++i) prim_order[prim_base + n_prims + i] = prims[i];
      n_prims += cnt;
    }
  }
  uint32_t* w = nodes_out + size_t(20) * (level_begin + t);
  w[0] = FBits(blo.x);
  w[1] = FBits(blo.y);
  w[2] = FBits(blo.z);
  w[3] = ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24);
  w[4] = child_base;
  w[5] = prim_base;
  w[6] = meta[0] | (meta[1] << 8) | (meta[2] << 16) | (meta[3] << 24);
  w[7] = meta[4] | (meta[5] << 8) | (meta[6] << 16) | (meta[7] << 24);
  for (int k = 0; k < 6; ++k) {
    w[8 + 2 * k] = q[k][0] | (q[k][1] << 8) | (q[k][2] << 16) | (q[k][3] << 24);
    w[9 + 2 * k] = q[k][4] | (q[k][5] << 8) | (q[k][6] << 16) | (q[k][7] << 24);
  }
}

// Morton grid of a scene box
inline void MortonGrid(const pbrbvh::Aabb& scene, float* blo, float* scale) {
  for (int k = 0; k < 3; ++k) {
    blo[k] = scene.lo[k];
    const float ext = scene.hi[k] - scene.lo[k];
    scale[k] = ext > 0.f ? 2097152.0f / ext : 0.f;
  }
}

// ---- the same passes on the host, one element after the other (tests, and the fallback when no device builder runs)
inline bool BuildBvh8PlocHost(const pbrbvh::Aabb* prim, uint32_t n, const pbrbvh::BuildParams& params, uint32_t radius,
                              pbrbvh::Bvh8* out, const char** err) {
  static const char* kErrEmpty = "BuildBvh8Ploc: no primitives";
  static const char* kErrDepth = "BuildBvh8Ploc: tree deeper than the traversal stack (31 wide levels)";
  static const char* kErrBounds = "BuildBvh8Ploc: non-finite primitive bounds";
  if (n == 0) { if (err) *err = kErrEmpty; return false; }
  pbrbvh::Aabb scene;
  for (int k = 0; k < 3; ++k) { scene.lo[k] = FLT_MAX; scene.hi[k] = -FLT_MAX; }
  for (uint32_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      if (!std::isfinite(prim[i].lo[k]) || !std::isfinite(prim[i].hi[k])) { if (err) *err = kErrBounds; return false; }
      scene.lo[k] = std::min(scene.lo[k], prim[i].lo[k]);
      scene.hi[k] = std::max(scene.hi[k], prim[i].hi[k]);
    }
  float blo[3], scale[3];
  MortonGrid(scene, blo, scale);
  std::vector<uint64_t> key(n);
  std::vector<uint32_t> idx(n);
  for (uint32_t i = 0; i < n; ++i) MortonBody(prim, blo, scale, i, key.data(), idx.data());
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });   // = LSD radix sort
  const size_t n2 = size_t(2) * n;
  std::vector<F4> nlo(n2), nhi(n2);
  std::vector<uint32_t> ncount(n2, 0u);
  std::vector<uint32_t> cnode[2] = {std::vector<uint32_t>(n), std::vector<uint32_t>(n)};
  std::vector<F4> clo[2] = {std::vector<F4>(n), std::vector<F4>(n)}, chi[2] = {std::vector<F4>(n), std::vector<F4>(n)};
  std::vector<uint32_t> nn(n), create(n), keep(n), sc(n), sk(n);
  for (uint32_t i = 0; i < n; ++i)
    InitClusterBody(prim, idx.data(), i, nlo.data(), nhi.data(), ncount.data(), cnode[0].data(), clo[0].data(), chi[0].data());
  uint32_t m = n, next_node = n;
  int cur = 0;
  while (m > 1) {
    for (uint32_t i = 0; i < m; ++i) NearestBody(clo[cur].data(), chi[cur].data(), m, radius, i, nn.data());
    for (uint32_t i = 0; i < m; ++i) PairFlagsBody(nn.data(), i, &create[i], &keep[i]);
    uint32_t a = 0, b = 0;
    for (uint32_t i = 0; i < m; ++i) { sc[i] = a; a += create[i]; sk[i] = b; b += keep[i]; }
    for (uint32_t i = 0; i < m; ++i)
      MergeBody(nn.data(), create[i], keep[i], sc[i], sk[i], next_node, cnode[cur].data(), clo[cur].data(),
                chi[cur].data(), i, nlo.data(), nhi.data(), ncount.data(), cnode[cur ^ 1].data(), clo[cur ^ 1].data(),
                chi[cur ^ 1].data());
    next_node += a;
    m = b;
    cur ^= 1;
  }
  const uint32_t root = cnode[cur][0];
  const uint32_t max_leaf = uint32_t(params.max_leaf_prims);
  std::vector<uint32_t> wide_item(std::max<uint32_t>(n, 1u));
  out->nodes.assign(size_t(20) * std::max<uint32_t>(n, 1u), 0u);
  out->prim_order.assign(n, 0u);
  wide_item[0] = root;
  uint32_t lb = 0, le = 1, prim_total = 0, depth = 0;
  std::vector<uint32_t> ci, cp, si, sp;
  while (le > lb) {
    if (++depth > uint32_t(kMaxLevels)) { if (err) *err = kErrDepth; return false; }
    const uint32_t cnt = le - lb;
    ci.resize(cnt); cp.resize(cnt); si.resize(cnt); sp.resize(cnt);
    for (uint32_t t = 0; t < cnt; ++t)
      WideCountBody(nlo.data(), nhi.data(), ncount.data(), wide_item.data(), lb, t, max_leaf, &ci[t], &cp[t]);
    uint32_t a = 0, b = 0;
    for (uint32_t t = 0; t < cnt; ++t) { si[t] = a; a += ci[t]; sp[t] = b; b += cp[t]; }
    for (uint32_t t = 0; t < cnt; ++t)
      WideEmitBody(nlo.data(), nhi.data(), ncount.data(), wide_item.data(), lb, le, t, max_leaf, si[t], sp[t],
                   prim_total, out->nodes.data(), out->prim_order.data());
    prim_total += b;
    lb = le;
    le += a;
  }
  out->num_nodes = lb;
  out->nodes.resize(size_t(20) * lb);
  out->max_depth = depth;
  out->bounds = scene;
  out->sah_cost = 0.0;
  return prim_total == n;
}

}  // namespace pbrploc
