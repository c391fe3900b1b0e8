// Host-side builder of the acceleration structure the traversal kernels walk: a binned-SAH binary BVH collapsed
// into 8-wide nodes with quantised child boxes (the 80-byte "compressed wide BVH" node of Ylitie, Karras & Laine,
// HPG 2017).  It replaces what the reference delegates to Embree's rtcCommitScene
// (src/raytracer/raytracer_impl.cc:81,147,192) — the *layout* is ours and B200-oriented (one node = five 128-bit
// loads, children of a node contiguous, primitives of a node contiguous), only the hit semantics follow Embree.
//
// Node layout (20 x 32-bit words):
//   w0..w2  float  p.xyz      origin of the node's quantisation grid (= node box lower corner)
//   w3      u8x4   ex,ey,ez   biased exponents: grid step = 2^(e-127) per axis;  imask: bit s set <=> slot s is an
//                              internal child
//   w4      u32    child_base index of the first internal child (internal children are contiguous, in slot order)
//   w5      u32    prim_base  leaf-order index of the node's first primitive (<= 24 primitives per node)
//   w6..w7  u8x8   meta[s]    0 = empty; internal: 0b001_11000 + s; leaf: (unary prim count << 5) | first prim offset
//   w8..w19 u8x8   qlo_x, qlo_y, qlo_z, qhi_x, qhi_y, qhi_z   child boxes on the grid (lo floored, hi ceiled)
// Child slots are chosen so that slot s holds the child lying towards octant s (bit 2 = +x, bit 1 = +y, bit 0 = +z)
// of the node centre; a ray then visits hit children in (slot XOR inverse ray octant) order without sorting.
#pragma once
#include <cstdint>
#include <vector>

namespace pbrbvh {

struct Aabb {
  float lo[3], hi[3];
};

struct BuildParams {
  float traversal_cost = 1.0f;   // SAH cost of one inner-node step relative to one primitive test
  float prim_cost = 0.6f;
  int max_leaf_prims = 3;        // <= 3: a leaf's primitive count is stored in unary in 3 bits
  int threads = 0;               // 0 = hardware concurrency
};

struct Bvh8 {
  std::vector<uint32_t> nodes;        // 20 words per node, root first
  std::vector<uint32_t> prim_order;   // leaf order -> index of the input primitive
  uint32_t num_nodes = 0;
  uint32_t max_depth = 0;             // in wide nodes; the traversal stack must hold this many entries
  Aabb bounds;
  double sah_cost = 0.0;
};

// Builds over `n` primitive boxes.  Returns false (with *err set) if the tree cannot be encoded.
bool BuildBvh8(const Aabb* prim_bounds, uint32_t n, const BuildParams& params, Bvh8* out, const char** err);

}  // namespace pbrbvh
