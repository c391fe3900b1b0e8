// See bvh_builder.h.  Plain C++ (no CUDA): runs once per pbrgpu_commit() on the host cores.
#include "bvh_builder.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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

  std::vector<uint32_t> tmp;       // scratch of the parallel partition (same size as idx)
  bool serial_top = false;         // PBRGPU_BVH_SERIAL_TOP: A/B switch for the chunked passes

  uint32_t Alloc() { return next_node.fetch_add(1); }

  // Near the root there are fewer ranges than host threads: the passes over a large range are cut into chunks and
  // run on the threads nobody uses yet.  Every reduction below is a min / max / integer sum and both partition paths
  // are stable, so neither the tree nor the primitive order depends on the chunking or on thread timing.
  static constexpr uint32_t kParallelMin = 1u << 18;
  int GrabThreads(int want) {
    int got = 0;
    while (got < want) {
      if (spare_threads.fetch_sub(1) > 0) ++got;
      else { spare_threads.fetch_add(1); break; }
    }
    return got;
  }
  template <class F>
  void RunChunks(uint32_t begin, uint32_t end, int nchunks, F fn) {   // fn(chunk, chunk_begin, chunk_end)
    const uint64_t cnt = end - begin;
    std::vector<std::thread> th;
    for (int c = 1; c < nchunks; ++c)
      th.emplace_back([=]() { fn(c, begin + uint32_t(cnt * c / nchunks), begin + uint32_t(cnt * (c + 1) / nchunks)); });
    fn(0, begin, begin + uint32_t(cnt / nchunks));
    for (auto& t : th) t.join();
  }
  const float* C(int axis) const { return axis == 0 ? cx.data() : (axis == 1 ? cy.data() : cz.data()); }

  void BuildRange(uint32_t node_id, uint32_t begin, uint32_t end) {
    Node2& node = nodes[node_id];
    const uint32_t cnt = end - begin;
    Aabb box = Empty(), cbox = Empty();
    const int helpers = (cnt >= kParallelMin && !serial_top) ? GrabThreads(15) : 0;
    const int nchunks = helpers + 1;
    auto bounds_of = [&](uint32_t b0, uint32_t e0, Aabb* bx, Aabb* cbx) {
      for (uint32_t i = b0; i < e0; ++i) {
        const uint32_t p = idx[i];
        Grow(bx, pb[p]);
        const float c[3] = {cx[p], cy[p], cz[p]};
        for (int k = 0; k < 3; ++k) {
          cbx->lo[k] = std::min(cbx->lo[k], c[k]);
          cbx->hi[k] = std::max(cbx->hi[k], c[k]);
        }
      }
    };
    if (helpers == 0) {
      bounds_of(begin, end, &box, &cbox);
    } else {
      std::vector<Aabb> pbx(nchunks, Empty()), pcb(nchunks, Empty());
      RunChunks(begin, end, nchunks, [&](int c, uint32_t b0, uint32_t e0) { bounds_of(b0, e0, &pbx[c], &pcb[c]); });
      for (int c = 0; c < nchunks; ++c) { Grow(&box, pbx[c]); Grow(&cbox, pcb[c]); }
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
    struct Bins { Aabb bb[3][kBins]; uint32_t bc[3][kBins]; };
    std::vector<Bins> pbins;
    if (helpers > 0) {   // all three axes in one pass over each chunk
      pbins.resize(nchunks);
      RunChunks(begin, end, nchunks, [&](int ch, uint32_t b0, uint32_t e0) {
        Bins& B = pbins[ch];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < kBins; ++b) { B.bb[a][b] = Empty(); B.bc[a][b] = 0; }
        for (int a = 0; a < 3; ++a) {
          const float lo = cbox.lo[a], ext = cbox.hi[a] - cbox.lo[a];
          if (!(ext > 0.f)) continue;
          const float scale = float(kBins) / ext;
          const float* c = C(a);
          for (uint32_t i = b0; i < e0; ++i) {
            const uint32_t p = idx[i];
            int b = int((c[p] - lo) * scale);
            b = std::min(std::max(b, 0), kBins - 1);
            Grow(&B.bb[a][b], pb[p]);
            B.bc[a][b]++;
          }
        }
      });
    }
    for (int axis = 0; axis < 3; ++axis) {
      const float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
      if (!(ext > 0.f)) continue;
      const float scale = float(kBins) / ext;
      Aabb bb[kBins];
      uint32_t bc[kBins];
      for (int b = 0; b < kBins; ++b) { bb[b] = Empty(); bc[b] = 0; }
      const float* c = C(axis);
      if (helpers > 0) {
        for (int ch = 0; ch < nchunks; ++ch)
          for (int b = 0; b < kBins; ++b) { Grow(&bb[b], pbins[ch].bb[axis][b]); bc[b] += pbins[ch].bc[axis][b]; }
      } else {
        for (uint32_t i = begin; i < end; ++i) {
          const uint32_t p = idx[i];
          int b = int((c[p] - lo) * scale);
          b = std::min(std::max(b, 0), kBins - 1);
          Grow(&bb[b], pb[p]);
          bc[b]++;
        }
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
    auto release_helpers = [&]() { if (helpers > 0) spare_threads.fetch_add(helpers); };
    if (int(cnt) <= prm.max_leaf_prims) {
      const float leaf_cost = prm.prim_cost * float(cnt) * area;
      const float split_cost =
          (best_axis >= 0) ? prm.traversal_cost * area + prm.prim_cost * best_cost : FLT_MAX;
      if (leaf_cost <= split_cost) {
        node.first = begin;
        node.count = cnt;
        release_helpers();
        return;
      }
    }

    uint32_t mid;
    if (best_axis >= 0) {
      const float lo = cbox.lo[best_axis], ext = cbox.hi[best_axis] - cbox.lo[best_axis];
      const float scale = float(kBins) / ext;
      const float* c = C(best_axis);
      const int bin = best_bin;
      auto goes_left = [&](uint32_t p) {
        int b = int((c[p] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        return b <= bin;
      };
      if (helpers > 0) {   // stable partition through the scratch array: count per chunk, then scatter
        std::vector<uint32_t> nleft(nchunks, 0u), cb(nchunks + 1, 0u);
        RunChunks(begin, end, nchunks, [&](int ch, uint32_t b0, uint32_t e0) {
          uint32_t k = 0;
          for (uint32_t i = b0; i < e0; ++i) k += goes_left(idx[i]) ? 1u : 0u;
          nleft[ch] = k; cb[ch] = b0; if (ch == nchunks - 1) cb[nchunks] = e0;
        });
        uint32_t total_left = 0;
        for (int ch = 0; ch < nchunks; ++ch) total_left += nleft[ch];
        std::vector<uint32_t> lo_off(nchunks), hi_off(nchunks);
        uint32_t lacc = begin, racc = begin + total_left;
        for (int ch = 0; ch < nchunks; ++ch) {
          lo_off[ch] = lacc; hi_off[ch] = racc;
          lacc += nleft[ch]; racc += (cb[ch + 1] - cb[ch]) - nleft[ch];
        }
        RunChunks(begin, end, nchunks, [&](int ch, uint32_t b0, uint32_t e0) {
          uint32_t l = lo_off[ch], r = hi_off[ch];
          for (uint32_t i = b0; i < e0; ++i) { const uint32_t p = idx[i]; if (goes_left(p)) tmp[l++] = p; else tmp[r++] = p; }
        });
        RunChunks(begin, end, nchunks, [&](int, uint32_t b0, uint32_t e0) {
          memcpy(idx.data() + b0, tmp.data() + b0, sizeof(uint32_t) * size_t(e0 - b0));
        });
        mid = begin + total_left;
      } else {
        // stable like the chunked scatter above: whether helper threads were free depends on timing, the primitive
        // order inside a range (hence leaf order, tie-breaking of coincident hits, the "split in half" fallback) must not
        uint32_t* m = std::stable_partition(idx.data() + begin, idx.data() + end, goes_left);
        mid = uint32_t(m - idx.data());
      }
    } else {
      mid = begin;   // all centroids coincide
    }
    if (mid == begin || mid == end) mid = begin + cnt / 2;   // degenerate: split the index range in half
    release_helpers();

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
  b.serial_top = getenv("PBRGPU_BVH_SERIAL_TOP") != nullptr;
  if (n >= Builder::kParallelMin) b.tmp.resize(n);
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
  const auto tb = std::chrono::steady_clock::now();
  b.BuildRange(root2, 0, n);
  if (getenv("PBRGPU_VERBOSE_COMMIT"))
    fprintf(stderr, "  BuildBvh8: binary SAH build %.3f s (%d threads)\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tb).count(), threads);
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
