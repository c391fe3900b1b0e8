// See bvh_builder.h.  Plain C++ (no CUDA): runs once per pbrgpu_commit() on the host cores.
#include "bvh_builder.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>

namespace pbrbvh {
namespace {

inline void Grow(Aabb* a, const Aabb& b) {
  for (int k = 0; k < 3; ++k) {
    a->lo[k] = std::min(a->lo[k], b.lo[k]);
    a->hi[k] = std::max(a->hi[k], b.hi[k]);
  }
}
inline Aabb Empty() {
  Aabb a;
  for (int k = 0; k < 3; ++k) { a.lo[k] = FLT_MAX; a.hi[k] = -FLT_MAX; }
  return a;
}
inline float HalfArea(const Aabb& a) {
  const float dx = a.hi[0] - a.lo[0], dy = a.hi[1] - a.lo[1], dz = a.hi[2] - a.lo[2];
  if (dx < 0.f || dy < 0.f || dz < 0.f) return 0.f;
  return dx * dy + dy * dz + dz * dx;
}

struct Node2 {
  Aabb box;
  uint32_t left = 0, right = 0;   // children (inner)
  uint32_t first = 0, count = 0;  // primitive range in idx[] (leaf when count > 0)
};

constexpr int kBins = 16;

struct Builder {
  const Aabb* pb;
  uint32_t n;
  BuildParams prm;
  std::vector<uint32_t> idx;
  std::vector<float> cx, cy, cz;   // primitive centroids
  std::vector<Node2> nodes;
  std::atomic<uint32_t> next_node{0};
  std::atomic<int> spare_threads{0};

  uint32_t Alloc() { return next_node.fetch_add(1); }
  const float* C(int axis) const { return axis == 0 ? cx.data() : (axis == 1 ? cy.data() : cz.data()); }

  void BuildRange(uint32_t node_id, uint32_t begin, uint32_t end) {
    Node2& node = nodes[node_id];
    const uint32_t cnt = end - begin;
    Aabb box = Empty(), cbox = Empty();
    for (uint32_t i = begin; i < end; ++i) {
      const uint32_t p = idx[i];
      Grow(&box, pb[p]);
      const float c[3] = {cx[p], cy[p], cz[p]};
      for (int k = 0; k < 3; ++k) {
        cbox.lo[k] = std::min(cbox.lo[k], c[k]);
        cbox.hi[k] = std::max(cbox.hi[k], c[k]);
      }
    }
    node.box = box;
    if (cnt == 1) {
      node.first = begin;
      node.count = 1;
      return;
    }

    // binned SAH over the three axes
    float best_cost = FLT_MAX;
    int best_axis = -1, best_bin = -1;
    for (int axis = 0; axis < 3; ++axis) {
      const float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
      if (!(ext > 0.f)) continue;
      const float scale = float(kBins) / ext;
      Aabb bb[kBins];
      uint32_t bc[kBins];
      for (int b = 0; b < kBins; ++b) { bb[b] = Empty(); bc[b] = 0; }
      const float* c = C(axis);
      for (uint32_t i = begin; i < end; ++i) {
        const uint32_t p = idx[i];
        int b = int((c[p] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        Grow(&bb[b], pb[p]);
        bc[b]++;
      }
      float right_area[kBins];
      uint32_t right_cnt[kBins];
      Aabb acc = Empty();
      uint32_t c_acc = 0;
      for (int b = kBins - 1; b > 0; --b) {
        Grow(&acc, bb[b]);
        c_acc += bc[b];
        right_area[b] = HalfArea(acc);
        right_cnt[b] = c_acc;
      }
      acc = Empty();
      c_acc = 0;
      for (int b = 0; b < kBins - 1; ++b) {
        Grow(&acc, bb[b]);
        c_acc += bc[b];
        if (c_acc == 0 || right_cnt[b + 1] == 0) continue;
        const float cost = HalfArea(acc) * float(c_acc) + right_area[b + 1] * float(right_cnt[b + 1]);
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = axis;
          best_bin = b;
        }
      }
    }

    const float area = HalfArea(box);
    if (int(cnt) <= prm.max_leaf_prims) {
      const float leaf_cost = prm.prim_cost * float(cnt) * area;
      const float split_cost =
          (best_axis >= 0) ? prm.traversal_cost * area + prm.prim_cost * best_cost : FLT_MAX;
      if (leaf_cost <= split_cost) {
        node.first = begin;
        node.count = cnt;
        return;
      }
    }

    uint32_t mid;
    if (best_axis >= 0) {
      const float lo = cbox.lo[best_axis], ext = cbox.hi[best_axis] - cbox.lo[best_axis];
      const float scale = float(kBins) / ext;
      const float* c = C(best_axis);
      const int bin = best_bin;
      uint32_t* m = std::partition(idx.data() + begin, idx.data() + end, [&](uint32_t p) {
        int b = int((c[p] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        return b <= bin;
      });
      mid = uint32_t(m - idx.data());
    } else {
      mid = begin;   // all centroids coincide
    }
    if (mid == begin || mid == end) mid = begin + cnt / 2;   // degenerate: split the index range in half

    const uint32_t l = Alloc(), r = Alloc();
    node.left = l;
    node.right = r;
    node.count = 0;
    bool forked = false;
    std::thread th;
    if (cnt > 65536) {
      if (spare_threads.fetch_sub(1) > 0) {
        forked = true;
        th = std::thread([this, l, begin, mid]() {
          BuildRange(l, begin, mid);
          spare_threads.fetch_add(1);
        });
      } else {
        spare_threads.fetch_add(1);
      }
    }
    if (!forked) BuildRange(l, begin, mid);
    BuildRange(r, mid, end);
    if (forked) th.join();
  }
};

struct Child {
  uint32_t node2;   // index into nodes2
  int slot;
};

inline uint32_t FloatBits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

}  // namespace

bool BuildBvh8(const Aabb* prim_bounds, uint32_t n, const BuildParams& params, Bvh8* out, const char** err) {
  static const char* kErrEmpty = "BuildBvh8: no primitives";
  static const char* kErrLeaf = "BuildBvh8: max_leaf_prims must be in 1..3";
  static const char* kErrDepth = "BuildBvh8: tree deeper than the traversal stack (32 wide levels)";
  static const char* kErrBounds = "BuildBvh8: non-finite primitive bounds";
  if (n == 0) { if (err) *err = kErrEmpty; return false; }
  if (params.max_leaf_prims < 1 || params.max_leaf_prims > 3) { if (err) *err = kErrLeaf; return false; }

  Builder b;
  b.pb = prim_bounds;
  b.n = n;
  b.prm = params;
  b.idx.resize(n);
  b.cx.resize(n); b.cy.resize(n); b.cz.resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    b.idx[i] = i;
    const Aabb& a = prim_bounds[i];
    for (int k = 0; k < 3; ++k) {
      if (!std::isfinite(a.lo[k]) || !std::isfinite(a.hi[k])) { if (err) *err = kErrBounds; return false; }
    }
    b.cx[i] = 0.5f * (a.lo[0] + a.hi[0]);
    b.cy[i] = 0.5f * (a.lo[1] + a.hi[1]);
    b.cz[i] = 0.5f * (a.lo[2] + a.hi[2]);
  }
  b.nodes.resize(size_t(2) * n);
  int threads = params.threads > 0 ? params.threads : int(std::max(1u, std::thread::hardware_concurrency()));
  b.spare_threads = threads - 1;
  const uint32_t root2 = b.Alloc();
  b.BuildRange(root2, 0, n);
  const std::vector<Node2>& n2 = b.nodes;

  // ---- collapse to 8-wide and emit breadth-first
  out->nodes.clear();
  out->prim_order.clear();
  out->prim_order.reserve(n);
  out->bounds = n2[root2].box;
  out->max_depth = 0;
  out->sah_cost = 0.0;

  struct Item { uint32_t node2; uint32_t out_index; uint32_t depth; };
  std::vector<Item> queue;
  queue.push_back({root2, 0u, 1u});
  out->nodes.resize(20);
  size_t head = 0;
  const double root_area = std::max(1e-30, double(HalfArea(n2[root2].box)));

  while (head < queue.size()) {
    const Item it = queue[head++];
    out->max_depth = std::max(out->max_depth, it.depth);
    const Node2& nd = n2[it.node2];
    out->sah_cost += double(params.traversal_cost) * double(HalfArea(nd.box)) / root_area;

    // gather up to 8 children by repeatedly opening the inner child with the largest surface area
    uint32_t ch[8];
    int nch = 0;
    if (nd.count > 0) {
      ch[nch++] = it.node2;   // a lone leaf root
    } else {
      ch[nch++] = nd.left;
      ch[nch++] = nd.right;
      while (nch < 8) {
        int pick = -1;
        float pick_area = -1.f;
        for (int i = 0; i < nch; ++i) {
          if (n2[ch[i]].count == 0) {
            const float a = HalfArea(n2[ch[i]].box);
            if (a > pick_area) { pick_area = a; pick = i; }
          }
        }
        if (pick < 0) break;
        const uint32_t open = ch[pick];
        ch[pick] = n2[open].left;
        ch[nch++] = n2[open].right;
      }
    }

    // slot assignment: greedy on dot(child centroid - node centroid, octant direction of the slot)
    int slot_of[8];
    {
      float cost[8][8];
      const float ncx = 0.5f * (nd.box.lo[0] + nd.box.hi[0]), ncy = 0.5f * (nd.box.lo[1] + nd.box.hi[1]),
                  ncz = 0.5f * (nd.box.lo[2] + nd.box.hi[2]);
      for (int c = 0; c < nch; ++c) {
        const Aabb& cb = n2[ch[c]].box;
        const float dx = 0.5f * (cb.lo[0] + cb.hi[0]) - ncx, dy = 0.5f * (cb.lo[1] + cb.hi[1]) - ncy,
                    dz = 0.5f * (cb.lo[2] + cb.hi[2]) - ncz;
        for (int s = 0; s < 8; ++s) {
          cost[c][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
        }
      }
      bool cdone[8] = {false, false, false, false, false, false, false, false};
      bool sdone[8] = {false, false, false, false, false, false, false, false};
      for (int k = 0; k < nch; ++k) {
        int bc = -1, bs = -1;
        float bv = -FLT_MAX;
        for (int c = 0; c < nch; ++c) {
          if (cdone[c]) continue;
          for (int s = 0; s < 8; ++s) {
            if (sdone[s]) continue;
            if (cost[c][s] > bv) { bv = cost[c][s]; bc = c; bs = s; }
          }
        }
        cdone[bc] = true;
        sdone[bs] = true;
        slot_of[bc] = bs;
      }
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
    for (int c = 0; c < nch; ++c) child_in_slot[slot_of[c]] = c;

    // quantisation grid
    uint32_t ebyte[3];
    double step[3];
    for (int k = 0; k < 3; ++k) {
      const double ext = double(nd.box.hi[k]) - double(nd.box.lo[k]);
      int e = -126;
      if (ext > 0.0) {
        e = int(std::ceil(std::log2(ext / 255.0)));
        while (ext / std::ldexp(1.0, e) > 255.0) ++e;
        e = std::max(e, -126);
      }
      ebyte[k] = uint32_t(e + 127);
      step[k] = std::ldexp(1.0, e);
    }

    uint8_t meta[8], q[6][8];
    memset(meta, 0, sizeof(meta));
    memset(q, 0, sizeof(q));
    uint32_t imask = 0;
    const uint32_t child_base = uint32_t(out->nodes.size() / 20);
    const uint32_t prim_base = uint32_t(out->prim_order.size());
    uint32_t n_inner = 0, n_prims = 0;
    for (int s = 0; s < 8; ++s) {
      const int c = child_in_slot[s];
      if (c < 0) continue;
      const Node2& cn = n2[ch[c]];
      for (int k = 0; k < 3; ++k) {
        double lo = std::floor((double(cn.box.lo[k]) - double(nd.box.lo[k])) / step[k]);
        double hi = std::ceil((double(cn.box.hi[k]) - double(nd.box.lo[k])) / step[k]);
        lo = std::min(std::max(lo, 0.0), 255.0);
        hi = std::min(std::max(hi, 0.0), 255.0);
        q[k][s] = uint8_t(lo);
        q[3 + k][s] = uint8_t(hi);
      }
      if (cn.count == 0) {
        imask |= 1u << s;
        meta[s] = uint8_t((1u << 5) | (24u + uint32_t(s)));
        queue.push_back({ch[c], child_base + n_inner, it.depth + 1});
        ++n_inner;
      } else {
        const uint32_t unary = (cn.count == 1) ? 1u : (cn.count == 2 ? 3u : 7u);
        meta[s] = uint8_t((unary << 5) | n_prims);
        for (uint32_t i = 0; i < cn.count; ++i) out->prim_order.push_back(b.idx[cn.first + i]);
        n_prims += cn.count;
        out->sah_cost += double(params.prim_cost) * double(cn.count) * double(HalfArea(cn.box)) / root_area;
      }
    }
    out->nodes.resize(out->nodes.size() + size_t(20) * n_inner);

    uint32_t* w = out->nodes.data() + size_t(20) * it.out_index;
    w[0] = FloatBits(nd.box.lo[0]);
    w[1] = FloatBits(nd.box.lo[1]);
    w[2] = FloatBits(nd.box.lo[2]);
    w[3] = ebyte[0] | (ebyte[1] << 8) | (ebyte[2] << 16) | (imask << 24);
    w[4] = child_base;
    w[5] = prim_base;
    auto pack4 = [](const uint8_t* p) {
      return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
    };
    w[6] = pack4(meta);
    w[7] = pack4(meta + 4);
    for (int k = 0; k < 6; ++k) {
      w[8 + 2 * k] = pack4(q[k]);
      w[9 + 2 * k] = pack4(q[k] + 4);
    }
  }
  out->num_nodes = uint32_t(out->nodes.size() / 20);
  if (out->max_depth > 31) { if (err) *err = kErrDepth; return false; }
  return true;
}

}  // namespace pbrbvh
