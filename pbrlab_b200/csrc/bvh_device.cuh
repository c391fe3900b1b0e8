// The acceleration-structure build on the GPU: the data-parallel passes of bvh_ploc.h, one thread per element
// (replaces Embree's rtcCommitScene, reference src/raytracer/raytracer_impl.cc:81,147,192).
// Morton sort and the prefix sums use CUB (plumbing of a once-per-scene pass; the traversal kernels stay hand-written).
// The result — 80-byte nodes + the leaf order of the primitives — is copied back to the host scene, which lays out
// the primitive records exactly as it does after the host builder, so everything downstream is unchanged.
#pragma once
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include "bvh_ploc.h"

namespace pbrdev {

using pbrploc::F4;

__device__ __forceinline__ int FloatOrdered(float f) {   // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float OrderedFloat(int i) {
  const int j = i >= 0 ? i : i ^ 0x7fffffff;
#if defined(__CUDA_ARCH__)
  return __int_as_float(j);
#else
  float f; memcpy(&f, &j, 4); return f;
#endif
}

// scene box (6 ordered ints: min xyz, max xyz) + a flag for non-finite input
__global__ void BoundsKernel(const pbrbvh::Aabb* prim, uint32_t n, int* box6, uint32_t* bad) {
  __shared__ int s[6];
  if (threadIdx.x < 3) s[threadIdx.x] = 0x7fffffff;
  else if (threadIdx.x < 6) s[threadIdx.x] = int(0x80000000);
  __syncthreads();
  int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {int(0x80000000), int(0x80000000), int(0x80000000)};
  bool any_bad = false;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const pbrbvh::Aabb b = prim[i];
    for (int k = 0; k < 3; ++k) {
      any_bad |= !isfinite(b.lo[k]) || !isfinite(b.hi[k]);
      lo[k] = min(lo[k], FloatOrdered(b.lo[k]));
      hi[k] = max(hi[k], FloatOrdered(b.hi[k]));
    }
  }
  for (int k = 0; k < 3; ++k) {
    lo[k] = __reduce_min_sync(0xffffffffu, lo[k]);
    hi[k] = __reduce_max_sync(0xffffffffu, hi[k]);
  }
  if ((threadIdx.x & 31) == 0) {
    for (int k = 0; k < 3; ++k) { atomicMin(&s[k], lo[k]); atomicMax(&s[3 + k], hi[k]); }
  }
  if (any_bad) atomicOr(bad, 1u);
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(&box6[threadIdx.x], s[threadIdx.x]);
  else if (threadIdx.x < 6) atomicMax(&box6[threadIdx.x], s[threadIdx.x]);
}

struct Grid3 { float blo[3], scale[3]; };

__global__ void MortonKernel(const pbrbvh::Aabb* prim, Grid3 g, uint32_t n, uint64_t* key, uint32_t* idx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pbrploc::MortonBody(prim, g.blo, g.scale, i, key, idx);
}
__global__ void InitClusterKernel(const pbrbvh::Aabb* prim, const uint32_t* sorted_idx, uint32_t n, F4* nlo, F4* nhi,
                                  uint32_t* ncount, uint32_t* cnode, F4* clo, F4* chi) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pbrploc::InitClusterBody(prim, sorted_idx, i, nlo, nhi, ncount, cnode, clo, chi);
}
__global__ void NearestKernel(const F4* clo, const F4* chi, uint32_t m, uint32_t radius, uint32_t* nn) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) pbrploc::NearestBody(clo, chi, m, radius, i, nn);
}
// both flags in one 64-bit word (create << 32 | keep): one prefix sum serves both
__global__ void PairFlagsKernel(const uint32_t* nn, uint32_t m, uint64_t* flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  uint32_t create, keep;
  pbrploc::PairFlagsBody(nn, i, &create, &keep);
  flags[i] = (uint64_t(create) << 32) | keep;
}
__global__ void MergeKernel(const uint32_t* nn, const uint64_t* flags, const uint64_t* scan, uint32_t first_new_node,
                            const uint32_t* cnode, const F4* clo, const F4* chi, uint32_t m, F4* nlo, F4* nhi,
                            uint32_t* ncount, uint32_t* cnode_out, F4* clo_out, F4* chi_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t create = uint32_t(flags[i] >> 32), keep = uint32_t(flags[i]);
  const uint32_t sc = uint32_t(scan[i] >> 32), sk = uint32_t(scan[i]);
  pbrploc::MergeBody(nn, create, keep, sc, sk, first_new_node, cnode, clo, chi, i, nlo, nhi, ncount, cnode_out, clo_out,
                     chi_out);
}
// totals of a 64-bit packed flag array after its exclusive scan: out[0] = high sum, out[1] = low sum
__global__ void TotalsKernel(const uint64_t* flags, const uint64_t* scan, uint32_t m, uint32_t* out2) {
  const uint64_t t = scan[m - 1] + flags[m - 1];
  out2[0] = uint32_t(t >> 32);
  out2[1] = uint32_t(t);
}
__global__ void WideCountKernel(const F4* nlo, const F4* nhi, const uint32_t* ncount, const uint32_t* wide_item,
                                uint32_t level_begin, uint32_t cnt, uint32_t max_leaf, uint64_t* counts) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  uint32_t ni, np;
  pbrploc::WideCountBody(nlo, nhi, ncount, wide_item, level_begin, t, max_leaf, &ni, &np);
  counts[t] = (uint64_t(ni) << 32) | np;
}
__global__ void WideEmitKernel(const F4* nlo, const F4* nhi, const uint32_t* ncount, uint32_t* wide_item,
                               uint32_t level_begin, uint32_t level_end, uint32_t cnt, uint32_t max_leaf,
                               const uint64_t* scan, uint32_t prim_total, uint32_t* nodes_out, uint32_t* prim_order) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const uint32_t si = uint32_t(scan[t] >> 32), sp = uint32_t(scan[t]);
  pbrploc::WideEmitBody(nlo, nhi, ncount, wide_item, level_begin, level_end, t, max_leaf, si, sp, prim_total, nodes_out,
                        prim_order);
}

struct Scratch {   // cudaMalloc'ed pieces, freed on every exit path
  std::vector<void*> ptrs;
  template <class T>
  cudaError_t Get(T** p, size_t count) {
    void* q = nullptr;
    const cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess) { ptrs.push_back(q); *p = static_cast<T*>(q); }
    return e;
  }
  ~Scratch() { for (void* p : ptrs) cudaFree(p); }
};

// Builds on the current device, on `st`.  Returns false with *err set on failure (the caller falls back to the host
// builder).  `seconds3` (optional): upload, build, download.
inline bool BuildBvh8OnDevice(cudaStream_t st, const pbrbvh::Aabb* h_boxes, uint32_t n, const pbrbvh::BuildParams& prm,
                              uint32_t radius, pbrbvh::Bvh8* out, const char** err, double* seconds3) {
  static const char* kErrCuda = "BuildBvh8OnDevice: CUDA error (out of memory?)";
  static const char* kErrBounds = "BuildBvh8OnDevice: non-finite primitive bounds";
  static const char* kErrDepth = "BuildBvh8OnDevice: tree deeper than the traversal stack (31 wide levels)";
  static const char* kErrNodes = "BuildBvh8OnDevice: more wide nodes than reserved";
  static const char* kErrStuck = "BuildBvh8OnDevice: clustering made no progress";
#define DEV_TRY(expr) do { if ((expr) != cudaSuccess) { cudaGetLastError(); if (err) *err = kErrCuda; return false; } } while (0)
  if (n == 0) { if (err) *err = "BuildBvh8OnDevice: no primitives"; return false; }
  const auto t0 = std::chrono::steady_clock::now();
  Scratch mem;
  const uint32_t B = 256;
  auto grid = [&](uint32_t m) { return (m + B - 1) / B; };
  pbrbvh::Aabb* prim = nullptr;
  DEV_TRY(mem.Get(&prim, n));
  DEV_TRY(cudaMemcpyAsync(prim, h_boxes, sizeof(pbrbvh::Aabb) * n, cudaMemcpyHostToDevice, st));
  int* box6 = nullptr; uint32_t* bad = nullptr; uint32_t* totals = nullptr;
  DEV_TRY(mem.Get(&box6, 6)); DEV_TRY(mem.Get(&bad, 1)); DEV_TRY(mem.Get(&totals, 2));
  const int init6[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, int(0x80000000), int(0x80000000), int(0x80000000)};
  DEV_TRY(cudaMemcpyAsync(box6, init6, sizeof(init6), cudaMemcpyHostToDevice, st));
  DEV_TRY(cudaMemsetAsync(bad, 0, 4, st));
  BoundsKernel<<<std::min<uint32_t>(grid(n), 1184u), B, 0, st>>>(prim, n, box6, bad);
  int h6[6]; uint32_t h_bad = 0;
  DEV_TRY(cudaMemcpyAsync(h6, box6, sizeof(h6), cudaMemcpyDeviceToHost, st));
  DEV_TRY(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st));
  DEV_TRY(cudaStreamSynchronize(st));
  const auto t1 = std::chrono::steady_clock::now();
  if (h_bad) { if (err) *err = kErrBounds; return false; }
  pbrbvh::Aabb scene;
  for (int k = 0; k < 3; ++k) { scene.lo[k] = OrderedFloat(h6[k]); scene.hi[k] = OrderedFloat(h6[3 + k]); }
  Grid3 g;
  pbrploc::MortonGrid(scene, g.blo, g.scale);

  // ---- Morton codes, sort
  uint64_t *key = nullptr, *key2 = nullptr;
  uint32_t *idx = nullptr, *idx2 = nullptr;
  DEV_TRY(mem.Get(&key, n)); DEV_TRY(mem.Get(&key2, n)); DEV_TRY(mem.Get(&idx, n)); DEV_TRY(mem.Get(&idx2, n));
  MortonKernel<<<grid(n), B, 0, st>>>(prim, g, n, key, idx);
  {
    size_t tmp_bytes = 0;
    cub::DoubleBuffer<uint64_t> dk(key, key2);
    cub::DoubleBuffer<uint32_t> dv(idx, idx2);
    DEV_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, int(n), 0, 63, st));
    char* tmp = nullptr;
    DEV_TRY(mem.Get(&tmp, tmp_bytes));
    DEV_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, int(n), 0, 63, st));
    idx = dv.Current();
  }
  // ---- PLOC
  const size_t n2 = size_t(2) * n;
  F4 *nlo = nullptr, *nhi = nullptr, *clo[2] = {nullptr, nullptr}, *chi[2] = {nullptr, nullptr};
  uint32_t *ncount = nullptr, *cnode[2] = {nullptr, nullptr}, *nn = nullptr;
  uint64_t *flags = nullptr, *scan = nullptr;
  DEV_TRY(mem.Get(&nlo, n2)); DEV_TRY(mem.Get(&nhi, n2)); DEV_TRY(mem.Get(&ncount, n2));
  for (int k = 0; k < 2; ++k) { DEV_TRY(mem.Get(&cnode[k], n)); DEV_TRY(mem.Get(&clo[k], n)); DEV_TRY(mem.Get(&chi[k], n)); }
  DEV_TRY(mem.Get(&nn, n)); DEV_TRY(mem.Get(&flags, n)); DEV_TRY(mem.Get(&scan, n));
  size_t scan_bytes = 0;
  DEV_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags, scan, int(n), st));
  char* scan_tmp = nullptr;
  DEV_TRY(mem.Get(&scan_tmp, scan_bytes));
  InitClusterKernel<<<grid(n), B, 0, st>>>(prim, idx, n, nlo, nhi, ncount, cnode[0], clo[0], chi[0]);
  uint32_t m = n, next_node = n;
  int cur = 0;
  uint32_t h_tot[2];
  while (m > 1) {
    NearestKernel<<<grid(m), B, 0, st>>>(clo[cur], chi[cur], m, radius, nn);
    PairFlagsKernel<<<grid(m), B, 0, st>>>(nn, m, flags);
    DEV_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flags, scan, int(m), st));
    MergeKernel<<<grid(m), B, 0, st>>>(nn, flags, scan, next_node, cnode[cur], clo[cur], chi[cur], m, nlo, nhi, ncount,
                                        cnode[cur ^ 1], clo[cur ^ 1], chi[cur ^ 1]);
    TotalsKernel<<<1, 1, 0, st>>>(flags, scan, m, totals);
    DEV_TRY(cudaMemcpyAsync(h_tot, totals, 8, cudaMemcpyDeviceToHost, st));
    DEV_TRY(cudaStreamSynchronize(st));
    if (h_tot[0] == 0u || h_tot[1] >= m) { if (err) *err = kErrStuck; return false; }
    next_node += h_tot[0];
    m = h_tot[1];
    cur ^= 1;
  }
  uint32_t root = 0;
  DEV_TRY(cudaMemcpyAsync(&root, cnode[cur], 4, cudaMemcpyDeviceToHost, st));
  // ---- collapse to 8-wide, level by level.  A wide node has at least two children (every one of them a subtree
  // with a binary root of its own), so there are at most n/2 + 1 wide nodes... bounded by n; reserve generously
  const uint32_t max_leaf = uint32_t(prm.max_leaf_prims);
  const uint32_t node_cap = std::max<uint32_t>(n / 2u + 1024u, 1024u);
  uint32_t *wide_item = nullptr, *nodes = nullptr, *prim_order = nullptr;
  DEV_TRY(mem.Get(&wide_item, node_cap)); DEV_TRY(mem.Get(&nodes, size_t(20) * node_cap)); DEV_TRY(mem.Get(&prim_order, n));
  DEV_TRY(cudaMemcpyAsync(wide_item, &root, 4, cudaMemcpyHostToDevice, st));
  uint32_t lb = 0, le = 1, prim_total = 0, depth = 0;
  while (le > lb) {
    if (++depth > uint32_t(pbrploc::kMaxLevels)) { if (err) *err = kErrDepth; return false; }
    const uint32_t cnt = le - lb;
    WideCountKernel<<<grid(cnt), B, 0, st>>>(nlo, nhi, ncount, wide_item, lb, cnt, max_leaf, flags);
    DEV_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, flags, scan, int(cnt), st));
    TotalsKernel<<<1, 1, 0, st>>>(flags, scan, cnt, totals);
    DEV_TRY(cudaMemcpyAsync(h_tot, totals, 8, cudaMemcpyDeviceToHost, st));
    DEV_TRY(cudaStreamSynchronize(st));
    if (uint64_t(le) + h_tot[0] > node_cap) { if (err) *err = kErrNodes; return false; }
    WideEmitKernel<<<grid(cnt), B, 0, st>>>(nlo, nhi, ncount, wide_item, lb, le, cnt, max_leaf, scan, prim_total, nodes,
                                             prim_order);
    prim_total += h_tot[1];
    lb = le;
    le += h_tot[0];
  }
  DEV_TRY(cudaStreamSynchronize(st));
  DEV_TRY(cudaGetLastError());
  const auto t2 = std::chrono::steady_clock::now();
  if (prim_total != n) { if (err) *err = "BuildBvh8OnDevice: primitive count mismatch"; return false; }
  out->num_nodes = lb;
  out->nodes.resize(size_t(20) * lb);
  out->prim_order.resize(n);
  DEV_TRY(cudaMemcpyAsync(out->nodes.data(), nodes, sizeof(uint32_t) * 20 * lb, cudaMemcpyDeviceToHost, st));
  DEV_TRY(cudaMemcpyAsync(out->prim_order.data(), prim_order, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
  DEV_TRY(cudaStreamSynchronize(st));
  out->max_depth = depth;
  out->bounds = scene;
  out->sah_cost = 0.0;
  if (seconds3) {
    const auto t3 = std::chrono::steady_clock::now();
    seconds3[0] = std::chrono::duration<double>(t1 - t0).count();
    seconds3[1] = std::chrono::duration<double>(t2 - t1).count();
    seconds3[2] = std::chrono::duration<double>(t3 - t2).count();
  }
#undef DEV_TRY
  return true;
}

}  // namespace pbrdev
