// Implementation of the C ABI in include/pbrgpu.h on CUDA (sm_100a): device memory management, scene upload,
// the host loop that drives the wavefront kernels (wavefront.cuh), and the parity test hooks.
//
// There is no CPU path in this file: every entry point that computes anything launches kernels, and
// pbrgpu_create() fails when no CUDA device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pbrgpu.h"
#include "bvh_device.cuh"
#include "job_split.h"
#include "kat.cuh"
#include "nccl_shim.h"
#include "scene_host.h"
#include "wavefront.cuh"

struct pbrgpu_ctx;
namespace {

std::string g_create_error;
// the per-device worker threads of a multi-device context may fail at the same time: ctx->error is written under a lock
void SetError(pbrgpu_ctx* ctx, const std::string& msg);

#define CUDA_TRY(ctx, expr)                                                                      \
  do {                                                                                           \
    cudaError_t err__ = (expr);                                                                  \
    if (err__ != cudaSuccess) {                                                                  \
      SetError((ctx), std::string(#expr) + ": " + cudaGetErrorString(err__));                    \
      return PBRGPU_ERR_CUDA;                                                                    \
    }                                                                                            \
  } while (0)
#define NCCL_TRY(ctx, expr)                                                                      \
  do {                                                                                           \
    ncclResult_t err__ = (expr);                                                                 \
    if (err__ != ncclSuccess) {                                                                  \
      SetError((ctx), std::string(#expr) + ": " + pbrnccl::Get().GetErrorString(err__));         \
      return PBRGPU_ERR_CUDA;                                                                    \
    }                                                                                            \
  } while (0)

template <typename T>
struct DevBuf {
  T* ptr = nullptr;
  size_t count = 0;
  cudaError_t Alloc(size_t n) {
    if (n <= count && ptr) return cudaSuccess;
    Free();
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ptr), n * sizeof(T));
    if (e == cudaSuccess) count = n;
    return e;
  }
  cudaError_t Upload(const void* src, size_t n, cudaStream_t st) {
    cudaError_t e = Alloc(n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(ptr, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
  }
  void Free() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
};

// everything one GPU holds
struct Device {
  int id = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t walk_stream = nullptr;              // the random-walk kernels of an iteration run beside closest hit + shading
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  cudaEvent_t kev[12] = {};   // per-kernel-family timing of one iteration (profiling mode)
  // scene
  DevBuf<float4> geom;   // [tri_nodes | tri_data | curve_nodes | curve_data]: one range for the L2 persistence window
  DevBuf<float4> verts, normals, emissive, lprim_info;
  DevBuf<float2> texcoords;
  DevBuf<uint32_t> curve_prim, curve_sub, lprim_tri, clear_dist, material_class;
  DevBuf<float> curve_cull;
  DevBuf<float> tex_pixels;
  DevBuf<pbr::TexDesc> tex_desc;
  DevBuf<uint4> tri_ids, tri_nidx, tri_vidx, tri_tidx, curve_ids;
  DevBuf<pbr::DeviceMaterial> materials;
  DevBuf<float> light_cdf, lprim_cdf;
  DevBuf<pbr::LightRec> lights;
  pbr::SceneView view;
  // wave
  DevBuf<float4> state0, state1, walk0, walk1, exit_rec, done0, done1, sh_o, sh_d, sh_c;
  DevBuf<uint32_t> q_surface, q_diffuse, q_hair, counters;
  DevBuf<unsigned long long> stats;
  DevBuf<uint32_t> heavy, order, order_keys, order_scratch;   // longest-paths-first sample order (FrameParams::order)
  DevBuf<uint8_t> order_tmp;
  pbr::WaveState wave;
  uint32_t wave_capacity = 0;
  // frame accumulators (device side of RenderLayer)
  DevBuf<float4> rgba;
  DevBuf<float4> peer_tmp;    // scratch of the NCCL-less frame-end sum
  DevBuf<uint32_t> count;
  DevBuf<uint8_t> srgb8;      // output stage (pbrgpu_resolve_srgb8)
  uint32_t frame_width = 0, frame_height = 0;   // size of the last rendered frame
  // pinned host mirror of the counters
  // pinned staging of the frame read-back (pbrgpu_render with host buffers): the device writes it by DMA in chunks,
  // host threads copy each chunk on to the caller's pageable buffers while the next one is in flight
  uint8_t* h_frame = nullptr;
  size_t h_frame_bytes = 0;
  cudaEvent_t ev_chunk[8] = {};
  uint32_t* h_counters = nullptr;
  unsigned long long* h_stats = nullptr;

  void Release() {
    geom.Free(); clear_dist.Free(); verts.Free(); normals.Free(); material_class.Free(); tex_pixels.Free();
    tex_desc.Free(); q_diffuse.Free(); srgb8.Free();
    emissive.Free(); lprim_info.Free(); texcoords.Free(); curve_prim.Free(); curve_sub.Free(); curve_cull.Free(); lprim_tri.Free(); tri_ids.Free();
    tri_nidx.Free(); tri_vidx.Free(); tri_tidx.Free(); curve_ids.Free(); materials.Free(); light_cdf.Free();
    lprim_cdf.Free(); lights.Free();
    state0.Free(); state1.Free(); walk0.Free(); walk1.Free(); exit_rec.Free(); done0.Free(); done1.Free();
    sh_o.Free(); sh_d.Free(); sh_c.Free();
    q_surface.Free(); q_hair.Free(); counters.Free(); stats.Free();
    heavy.Free(); order.Free(); order_keys.Free(); order_scratch.Free(); order_tmp.Free();
    rgba.Free(); count.Free(); peer_tmp.Free();
    if (h_counters) cudaFreeHost(h_counters);
    if (h_stats) cudaFreeHost(h_stats);
    h_counters = nullptr; h_stats = nullptr;
    if (h_frame) cudaFreeHost(h_frame);
    h_frame = nullptr; h_frame_bytes = 0;
    for (auto& e : ev_chunk) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (auto& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (auto& e : kev) { if (e) cudaEventDestroy(e); e = nullptr; }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    ev_fork = ev_join = nullptr;
    if (walk_stream) cudaStreamDestroy(walk_stream);
    walk_stream = nullptr;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
};

}  // namespace

struct pbrgpu_ctx {
  std::vector<Device> devices;
  pbrhost::HostScene host;
  std::string error;
  std::mutex error_mutex;
  // frame-end reduce: one communicator per device of a multi-device context (ncclCommInitAll, created
  // by the first multi-device render), and/or one communicator of a multi-process job (pbrgpu_nccl_init)
  std::vector<ncclComm_t> dev_comms;
  ncclComm_t job_comm = nullptr;
  int job_rank = 0, job_world = 1;
  pbrgpu_stats stats;
  uint32_t wave_spp = 0;
  bool committed = false;
  pbrgpu_commit_info commit_info;
  bool profile = false;   // time every kernel family with CUDA events (pbrgpu_set_profiling)
  bool trace_iterations = false;   // PBRGPU_TRACE_ITERATIONS: one line per wavefront iteration on stderr
  // launch tuning (defaults measured on B200; PBRGPU_* environment variables override for sweeps)
  uint32_t tune_refill = 16;       // idle lanes that trigger a refill in the traversal engine
  uint32_t tune_refill_any = 16;      // any-hit kernel in triangle scenes (24 was best on the round-1 trees; with the PLOC trees 16: 43.5 -> 39.9 ms per 128-spp frame, profiles/r2h_tune_engine_knobs.log)
  uint32_t tune_refill_curves = 12;   // same in scenes with curves (lanes wait longer there: held ribbon candidates)
  uint32_t tune_refill_sss = 24;   // same for the random-walk kernel (its converged section is the bounce itself)
  uint32_t tune_prim_lanes = 1, tune_prim_lanes_sss = 1;   // lanes with pending primitives that trigger a primitive phase
  uint32_t tune_ribbon_lanes = 8;  // lanes holding a curve candidate that trigger the (batched) ribbon test
  int tune_l2_persist = 0;         // persisting-L2 window over the traversal data
  int tune_walk_bounces = 16;   // bounces a walk gets per launch before it is parked (pool busy)
  // path slots in flight = clamp(samples of the frame / tune_pool_div, tune_pool_min_mi, tune_pool_mi): a frame
  // that is only a few pool-fills long (strong scaling: 1/8 of the samples per GPU) spends a smaller share of its
  // time ramping up and draining with a smaller pool
  int tune_pool_div = 2, tune_pool_min_mi = 4;   // (8 until the end of a frame got cheap, DESIGN §2.4: now a frame of two pool fills beats one of eight, profiles/r4u_tune_pool_short_frames.log)
  int tune_drain_paths = 1 << 16, tune_drain_bounces = 48;     // fewer paths in flight than this: longer walk slices (sweeps: profiles/r2j_tune_drain_stages.log; after the thin spreading 48 instead of 256, profiles/r3e_tune_*.log)
  int tune_drain2_paths = 0, tune_drain2_bounces = 64;          // an intermediate stage (off by default)
  int tune_clear_march = 4;        // sphere-tracing steps of the clearance test along a walk segment
  int tune_sss_skip = 1;           // clearance grid: random-walk segments that provably hit nothing are not traced
  int tune_shade_threads = 512;    // block size of the shading kernels (<= pbr::kShadeBlock)
  int tune_pool_mi = 32;           // paths kept in flight, in Mi (x ~800 B of queue storage each): 8 -> 32 Mi is +5 % on C2 (fewer, longer launches)
  int tune_trace_blocks = 8, tune_shade_blocks = 1, tune_walk_blocks = 6;   // resident 128-thread blocks per SM
  int tune_diffuse_threads = 128, tune_diffuse_blocks = 6;   // launch shape of the diffuse-only shading kernel (sweep: 128 x 6 beats 256 x 3 by 2 %)
  int tune_sort_materials = 1;     // diffuse-only Principled materials get their own shading queue and kernel
  // longest paths first (FrameParams::order): probing passes per pixel before the order is built, pixels per block
  int tune_order = 1, tune_order_probe = 4, tune_order_block = 1 << 16;
  int tune_finish_blocks = 4;      // its blocks per SM (128 registers per thread: four fit)
  int tune_finish_paths = 8192;    // fewer paths + walks in flight than this at the end of a frame: FinishPathsKernel runs them to their end (0: off)
  int tune_inside_first = 1;       // curve BVH: a ray visits the children whose box holds its origin first (traverse.cuh: PopChild)
  int tune_thin_spread = 1;        // launches with fewer items than lanes give every warp n / warps of them (trav_engine.cuh: LanesFor)
  // Walk kernels on their own stream, beside closest hit + shading of the same iteration (they only share atomically
  // appended output streams).  Measured (profiles/r2r_tune_overlap.log): with the full launch shapes the two kernels
  // do not share an SM — the walk blocks fill the register file — but the closest-hit blocks move in as the walk
  // kernel's last long items drain: +3.5 %.  Shapes that do fit together (3 + 3..6 blocks) are slower than running one
  // after the other: both kernels need all the warps they can get.  Profiling mode runs the families in sequence.
  int tune_overlap = 1, tune_trace_blocks_overlap = 8, tune_walk_blocks_overlap = 6;
};

namespace {

void SetError(pbrgpu_ctx* ctx, const std::string& msg) {
  std::lock_guard<std::mutex> lock(ctx->error_mutex);
  ctx->error = msg;
}

using pbr::SceneView;
using pbr::WaveState;

constexpr int kBlock = 128;
// Bounce budget of one SssWalkKernel launch (see wavefront.cuh).  One bounce is a ~4k-instruction dependent chain
// (tens of microseconds for a single warp), so a launch lasts at least budget x that: while the pool is busy a small
// budget keeps the launch throughput-bound (more, shorter walk slices in flight); once only stragglers are left a
// large budget saves host round-trips.
int PersistentGrid(const Device& d, int blocks_per_sm) { return d.sm_count * blocks_per_sm; }

uint32_t CountHairMaterials(const pbrhost::HostScene& h) {
  uint32_t n = 0;
  for (const auto& m : h.materials) n += (m.type == 1u) ? 1u : 0u;
  return n;
}

// a Principled material whose subsurface weight passes kClosureWeightCutOff (cycles-principled-shader.cc:217): its
// paths run random walks (looked up per Render(): live material edits can switch it on)
bool HasScatteringMaterial(const pbrhost::HostScene& h) {
  for (const auto& m : h.materials)
    if (m.type == 0u && m.p[3] > 1e-3f) return true;
  return false;
}

int UploadScene(pbrgpu_ctx* ctx, Device& d) {
  const pbrhost::HostScene& h = ctx->host;
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  cudaStream_t st = d.stream;
  // traversal data in one allocation: nodes first (hottest), then the leaf-ordered primitives
  const size_t n_tn = h.tri_bvh.nodes.size() / 4, n_td = h.tri_data.size(), n_cn = h.curve_bvh.nodes.size() / 4,
               n_cd = h.curve_data.size();
  CUDA_TRY(ctx, d.geom.Alloc(std::max<size_t>(n_tn + n_td + n_cn + n_cd, 1)));
  float4* g_tn = d.geom.ptr;
  float4* g_cn = g_tn + n_tn;
  float4* g_td = g_cn + n_cn;
  float4* g_cd = g_td + n_td;
  if (n_tn) CUDA_TRY(ctx, cudaMemcpyAsync(g_tn, h.tri_bvh.nodes.data(), n_tn * sizeof(float4), cudaMemcpyHostToDevice, st));
  if (n_cn) CUDA_TRY(ctx, cudaMemcpyAsync(g_cn, h.curve_bvh.nodes.data(), n_cn * sizeof(float4), cudaMemcpyHostToDevice, st));
  if (n_td) CUDA_TRY(ctx, cudaMemcpyAsync(g_td, h.tri_data.data(), n_td * sizeof(float4), cudaMemcpyHostToDevice, st));
  if (n_cd) CUDA_TRY(ctx, cudaMemcpyAsync(g_cd, h.curve_data.data(), n_cd * sizeof(float4), cudaMemcpyHostToDevice, st));
  {
    // Keep the BVH in L2 while path state streams through: persisting window over as much of the traversal data as
    // the device allows (nodes first), everything else in this stream is treated as streaming.
    cudaDeviceProp prop;
    CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, d.id));
    const size_t bytes = (n_tn + n_td + n_cn + n_cd) * sizeof(float4);
    if (ctx->tune_l2_persist && prop.persistingL2CacheMaxSize > 0 && bytes > 0) {
      const size_t carve = std::min<size_t>(size_t(prop.persistingL2CacheMaxSize), bytes);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof(attr));
      const size_t window = std::min<size_t>(bytes, size_t(prop.accessPolicyMaxWindowSize));
      attr.accessPolicyWindow.base_ptr = d.geom.ptr;
      attr.accessPolicyWindow.num_bytes = window;
      attr.accessPolicyWindow.hitRatio = float(std::min(1.0, double(carve) / double(window)));
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();   // best effort
    }
  }
  CUDA_TRY(ctx, d.curve_prim.Upload(h.curve_prim.data(), h.curve_prim.size(), st));
  CUDA_TRY(ctx, d.curve_sub.Upload(h.curve_sub.data(), h.curve_sub.size(), st));
  CUDA_TRY(ctx, d.curve_cull.Upload(h.curve_cull.data(), h.curve_cull.size(), st));
  CUDA_TRY(ctx, d.tri_ids.Upload(h.tri_ids.data(), h.tri_ids.size(), st));
  CUDA_TRY(ctx, d.tri_nidx.Upload(h.tri_nidx.data(), h.tri_nidx.size(), st));
  CUDA_TRY(ctx, d.tri_vidx.Upload(h.tri_vidx.data(), h.tri_vidx.size(), st));
  CUDA_TRY(ctx, d.tri_tidx.Upload(h.tri_tidx.data(), h.tri_tidx.size(), st));
  CUDA_TRY(ctx, d.verts.Upload(h.verts.data(), h.verts.size(), st));
  CUDA_TRY(ctx, d.normals.Upload(h.normals.data(), h.normals.size(), st));
  CUDA_TRY(ctx, d.texcoords.Upload(h.texcoords.data(), h.texcoords.size(), st));
  CUDA_TRY(ctx, d.curve_ids.Upload(h.curve_ids.data(), h.curve_ids.size(), st));
  CUDA_TRY(ctx, d.materials.Upload(h.materials.data(), h.materials.size(), st));
  CUDA_TRY(ctx, d.emissive.Upload(h.emissive.data(), h.emissive.size(), st));
  CUDA_TRY(ctx, d.light_cdf.Upload(h.light_cdf.data(), h.light_cdf.size(), st));
  CUDA_TRY(ctx, d.lights.Upload(h.lights.data(), h.lights.size(), st));
  CUDA_TRY(ctx, d.lprim_cdf.Upload(h.lprim_cdf.data(), h.lprim_cdf.size(), st));
  CUDA_TRY(ctx, d.lprim_info.Upload(h.lprim_info.data(), h.lprim_info.size(), st));
  CUDA_TRY(ctx, d.lprim_tri.Upload(h.lprim_tri.data(), h.lprim_tri.size(), st));
  CUDA_TRY(ctx, d.clear_dist.Upload(h.clear_dist.data(), h.clear_dist.size(), st));
  CUDA_TRY(ctx, d.material_class.Upload(h.material_class.data(), h.material_class.size(), st));
  CUDA_TRY(ctx, d.tex_pixels.Upload(h.tex_pixels.data(), h.tex_pixels.size(), st));
  CUDA_TRY(ctx, d.tex_desc.Upload(h.tex_desc.data(), h.tex_desc.size(), st));
  CUDA_TRY(ctx, cudaStreamSynchronize(st));
  SceneView& v = d.view;
  memset(&v, 0, sizeof(v));
  v.tri_nodes = g_tn; v.tri_data = g_td;
  v.curve_nodes = g_cn; v.curve_data = g_cd; v.curve_prim = d.curve_prim.ptr;
  v.curve_sub = d.curve_sub.ptr; v.curve_part_quads = h.curve_part_quads;
  v.ribbon_min_lanes = ctx->tune_ribbon_lanes;
  v.thin_spread = ctx->tune_thin_spread ? 1u : 0u;
  v.inside_first = ctx->tune_inside_first ? 1u : 0u;
  v.curve_cull = h.curve_cull.empty() ? nullptr : reinterpret_cast<const float4*>(d.curve_cull.ptr);
  v.num_tris = h.num_tris(); v.num_curves = h.num_curves();
  v.bias_magic = pbr::kBiasMagic;
  v.tri_ids = d.tri_ids.ptr; v.tri_nidx = d.tri_nidx.ptr; v.tri_vidx = d.tri_vidx.ptr; v.tri_tidx = d.tri_tidx.ptr;
  v.verts = d.verts.ptr; v.normals = d.normals.ptr; v.texcoords = d.texcoords.ptr; v.curve_ids = d.curve_ids.ptr;
  v.materials = d.materials.ptr; v.num_materials = uint32_t(h.materials.size());
  v.num_hair_materials = CountHairMaterials(h);
  v.material_class = d.material_class.ptr;
  v.tex_pixels = d.tex_pixels.ptr; v.tex_desc = d.tex_desc.ptr; v.num_textures = uint32_t(h.tex_desc.size());
  v.emissive = d.emissive.ptr; v.light_cdf = d.light_cdf.ptr; v.lights = d.lights.ptr;
  v.num_lights = uint32_t(h.lights.size());
  v.lprim_cdf = d.lprim_cdf.ptr; v.lprim_info = d.lprim_info.ptr; v.lprim_tri = d.lprim_tri.ptr;
  v.clear_dist = (ctx->tune_sss_skip && !h.clear_dist.empty()) ? reinterpret_cast<const uint8_t*>(d.clear_dist.ptr) : nullptr;
  for (int k = 0; k < 3; ++k) { v.clear_org[k] = h.clear_org[k]; v.clear_dims[k] = h.clear_dims[k]; }
  v.clear_inv_cell = h.clear_inv_cell;
  v.clear_quantum = h.clear_quantum;
  v.clear_march_steps = uint32_t(ctx->tune_clear_march);
  return PBRGPU_OK;
}

// bytes of wave storage per path in flight: S[2] + W[2] + E + D[2] + three index queues + the shadow queue
constexpr size_t kWaveBytesPerPath =
    sizeof(float4) * (2 * pbr::kStateFields + 2 * pbr::kWalkFields + pbr::kExitFields + 2 + 3 * 2) + 3 * sizeof(uint32_t);

int EnsureWave(pbrgpu_ctx* ctx, Device& d, uint32_t capacity) {
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  if (capacity > d.wave_capacity) {
    // grow: release the old arrays first (the largest pool is tens of GB)
    d.state0.Free(); d.state1.Free(); d.walk0.Free(); d.walk1.Free(); d.exit_rec.Free(); d.done0.Free(); d.done1.Free();
    d.q_surface.Free(); d.q_hair.Free(); d.q_diffuse.Free(); d.sh_o.Free(); d.sh_d.Free(); d.sh_c.Free();
    d.wave_capacity = 0;
    const size_t cap = capacity;
    CUDA_TRY(ctx, d.state0.Alloc(cap * pbr::kStateFields)); CUDA_TRY(ctx, d.state1.Alloc(cap * pbr::kStateFields));
    CUDA_TRY(ctx, d.walk0.Alloc(cap * pbr::kWalkFields)); CUDA_TRY(ctx, d.walk1.Alloc(cap * pbr::kWalkFields));
    CUDA_TRY(ctx, d.exit_rec.Alloc(cap * pbr::kExitFields));
    CUDA_TRY(ctx, d.done0.Alloc(cap)); CUDA_TRY(ctx, d.done1.Alloc(cap));
    CUDA_TRY(ctx, d.q_surface.Alloc(cap)); CUDA_TRY(ctx, d.q_hair.Alloc(cap)); CUDA_TRY(ctx, d.q_diffuse.Alloc(cap));
    CUDA_TRY(ctx, d.sh_o.Alloc(2 * cap)); CUDA_TRY(ctx, d.sh_d.Alloc(2 * cap)); CUDA_TRY(ctx, d.sh_c.Alloc(2 * cap));
    d.wave_capacity = capacity;
  }
  CUDA_TRY(ctx, d.counters.Alloc(pbr::kCounterCount));
  CUDA_TRY(ctx, d.stats.Alloc(pbr::kStatCount));
  if (!d.h_counters) CUDA_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&d.h_counters), sizeof(uint32_t) * pbr::kCounterCount));
  if (!d.h_stats) CUDA_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&d.h_stats), sizeof(unsigned long long) * pbr::kStatCount));
  WaveState& w = d.wave;
  w.state[0] = d.state0.ptr; w.state[1] = d.state1.ptr;
  w.walk[0] = d.walk0.ptr; w.walk[1] = d.walk1.ptr;
  w.exit_rec = d.exit_rec.ptr;
  w.done[0] = d.done0.ptr; w.done[1] = d.done1.ptr;
  w.q_surface = d.q_surface.ptr; w.q_diffuse = d.q_diffuse.ptr; w.q_hair = d.q_hair.ptr;
  w.sh_o = d.sh_o.ptr; w.sh_d = d.sh_d.ptr; w.sh_c = d.sh_c.ptr;
  w.counters = d.counters.ptr; w.stats = d.stats.ptr; w.capacity = d.wave_capacity;
  w.heavy = nullptr;
  return PBRGPU_OK;
}

// FrameParams::order: pixels sorted by the number of random walks their probing samples started (descending, stable:
// raster order among equals), as one key kernel + one 8-bit radix sort on the device's stream
__global__ void PixelOrderKeysKernel(const uint32_t* __restrict__ heavy, uint32_t npix, uint32_t* keys, uint32_t* ids) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  keys[i] = min(heavy[i], 255u);
  ids[i] = i;
}

int BuildPixelOrder(pbrgpu_ctx* ctx, Device& d, uint32_t npix) {
  PixelOrderKeysKernel<<<(npix + 255) / 256, 256, 0, d.stream>>>(d.heavy.ptr, npix, d.order_keys.ptr, d.order_scratch.ptr);
  size_t tmp_bytes = d.order_tmp.count;
  CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairsDescending(d.order_tmp.ptr, tmp_bytes, d.order_keys.ptr, d.order_keys.ptr + npix,
                                                          d.order_scratch.ptr, d.order.ptr, int(npix), 0, 8, d.stream));
  return PBRGPU_OK;
}

int EnsurePixelOrder(pbrgpu_ctx* ctx, Device& d, uint32_t npix) {
  CUDA_TRY(ctx, d.heavy.Alloc(npix));
  CUDA_TRY(ctx, d.order.Alloc(npix));
  CUDA_TRY(ctx, d.order_keys.Alloc(size_t(npix) * 2));
  CUDA_TRY(ctx, d.order_scratch.Alloc(npix));
  size_t tmp_bytes = 0;
  CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, d.order_keys.ptr, d.order_keys.ptr + npix,
                                                          d.order_scratch.ptr, d.order.ptr, int(npix), 0, 8, d.stream));
  CUDA_TRY(ctx, d.order_tmp.Alloc(tmp_bytes));
  CUDA_TRY(ctx, cudaMemsetAsync(d.heavy.ptr, 0, sizeof(uint32_t) * npix, d.stream));
  return PBRGPU_OK;
}

struct LoopTimers {
  double closest_ms = 0, any_ms = 0, shade_ms = 0, sss_ms = 0, regen_ms = 0, device_ms = 0;
  uint64_t launches = 0, closest_launches = 0;
};

// Runs wavefront iterations until nothing is in flight and (in frame mode) the frame has no samples left.
// Entry state: counters set up by ResetPoolKernel (frame) or InitPathsFromRaysKernel (hooks).  `rgba` receives the
// retired paths (frame accumulator, or the hook's one-entry-per-path buffer).
// max_iterations == 1 is the single-vertex mode of pbrgpu_shade: one full iteration, then the walks it started run to
// their end (walk + exit + shadow kernels only); the results are left in S / D of both parities.
int RunPool(pbrgpu_ctx* ctx, Device& d, const pbr::FrameParams* frame, float4* rgba, uint32_t max_in_flight,
            uint32_t max_iterations, pbr::ShadeFlags flags, LoopTimers* tm, const volatile int* cancel, size_t* finish_pass, size_t pass_offset,
            size_t pass_stride, size_t pass_cap) {
  cudaStream_t st = d.stream;
  const SceneView& s = d.view;
  WaveState w = d.wave;
  // frames with a pixel order: the probing passes run first (and count walk starts per pixel), the frame is held at
  // their end until the order is built (one key kernel + one radix sort on this stream, no extra synchronisation)
  const unsigned long long probe_samples =
      (frame && frame->order && frame->rest_passes) ? (unsigned long long)frame->probe_passes * frame->npix : 0ull;
  bool order_ready = probe_samples == 0ull;
  if (!order_ready) w.heavy = d.heavy.ptr;
  uint32_t parity = 0;
  const bool single = (max_iterations == 1u);
  const bool overlap = ctx->tune_overlap && !ctx->profile && !single && frame != nullptr;
  cudaStream_t wst = overlap ? d.walk_stream : st;
  const int grid_trace = PersistentGrid(d, ctx->tune_trace_blocks), grid_shade = PersistentGrid(d, ctx->tune_shade_blocks);
  const int grid_closest = overlap ? PersistentGrid(d, ctx->tune_trace_blocks_overlap) : grid_trace;
  const int grid_walk = PersistentGrid(d, overlap ? ctx->tune_walk_blocks_overlap : ctx->tune_walk_blocks);
  const int grid_diffuse = PersistentGrid(d, ctx->tune_diffuse_blocks);
  const bool curves = s.num_curves != 0u;
  const uint32_t refill = curves ? ctx->tune_refill_curves : ctx->tune_refill;
  const uint32_t refill_any = curves ? refill : ctx->tune_refill_any;
  pbr::FrameParams no_frame;
  memset(&no_frame, 0, sizeof(no_frame));
  const uint32_t regen = frame ? 1u : 0u;
  const uint32_t sort = ctx->tune_sort_materials ? 1u : 0u;
  const unsigned long long total = frame ? frame->total_samples : 0ull;
  // hits are routed by material CLASS (a hair material on a triangle goes to q_hair as well, as the reference's
  // Shader() dispatches on the material type, shader.cc:8-34), so the kernel runs whenever that queue can fill
  const bool hair = s.num_curves != 0u || s.num_hair_materials != 0u;
  auto launch_walk = [&](uint32_t cur, uint32_t budget, cudaStream_t ws) {
    if (curves) pbr::SssWalkKernel<true><<<grid_walk, kBlock, 0, ws>>>(s, w, cur, budget, ctx->tune_refill_sss, ctx->tune_prim_lanes_sss);
    else pbr::SssWalkKernel<false><<<grid_walk, kBlock, 0, ws>>>(s, w, cur, budget, ctx->tune_refill_sss, ctx->tune_prim_lanes_sss);
    pbr::SssExitKernel<<<grid_shade, ctx->tune_shade_threads, 0, ws>>>(s, w, cur, flags);
  };
  auto launch_any = [&](uint32_t next) {
    if (curves) pbr::TraceAnyKernel<true><<<grid_trace, kBlock, 0, st>>>(s, w, next, refill_any, ctx->tune_prim_lanes);
    else pbr::TraceAnyKernel<false><<<grid_trace, kBlock, 0, st>>>(s, w, next, refill_any, ctx->tune_prim_lanes);
  };
  // state after the set-up kernel (host knows it): hooks start with n paths, frames with nothing but samples to start
  bool have_active = (frame == nullptr), have_walk = false, samples_left = (frame != nullptr) && total > 0;
  uint64_t in_flight = ~0ull;   // paths + walks that the coming iteration works on (unknown before the first)
  const double trace_t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  for (uint32_t it = 0; (have_active || have_walk || samples_left) && it < max_iterations; ++it) {
    if (cancel && *cancel) break;
    const uint32_t next = parity ^ 1u;
    // few paths left: longer slices (a launch lasts budget x the per-bounce latency of a lone lane, ~3-5 us)
    uint32_t walk_budget = uint32_t(ctx->tune_walk_bounces);
    if (in_flight < uint64_t(ctx->tune_drain_paths)) walk_budget = uint32_t(ctx->tune_drain_bounces);
    else if (in_flight < uint64_t(ctx->tune_drain2_paths)) walk_budget = uint32_t(ctx->tune_drain2_bounces);
    const bool prof = ctx->profile;
    auto mark = [&](int k) { if (prof) cudaEventRecord(d.kev[k], st); };
    mark(0);
    pbr::BeginIterationKernel<<<1, 32, 0, st>>>(w.counters, w.stats, parity, regen, std::min(max_in_flight, w.capacity),
                                                order_ready ? total : probe_samples);
    if (frame && !single && !samples_left && in_flight < uint64_t(ctx->tune_finish_paths)) {
      // the last paths of the frame: one thread each to the end instead of an iteration per vertex
      pbr::RetireKernel<<<grid_shade, 256, 0, st>>>(w, parity, rgba);
      mark(1);
      // (every block resident at once: a chain that waits for a second wave of blocks is a longer frame)
      pbr::FinishPathsKernel<<<PersistentGrid(d, ctx->tune_finish_blocks), kBlock, 0, st>>>(s, w, parity, rgba);
      mark(2); mark(3); mark(4); mark(5);
      tm->launches += 3;
      CUDA_TRY(ctx, cudaMemcpyAsync(d.h_stats, w.stats, sizeof(unsigned long long) * pbr::kStatCount,
                                    cudaMemcpyDeviceToHost, st));
      CUDA_TRY(ctx, cudaStreamSynchronize(st));
      if (prof) {
        float ms[2] = {0, 0};
        cudaEventElapsedTime(&ms[0], d.kev[0], d.kev[1]);
        cudaEventElapsedTime(&ms[1], d.kev[1], d.kev[2]);
        tm->regen_ms += ms[0]; tm->closest_ms += ms[1];
      }
      if (ctx->trace_iterations) {
        const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        fprintf(stderr, "iter %u t %.3f ms: the last %llu paths and walks run to their end in one launch\n", it,
                (now - trace_t0) * 1e3, (unsigned long long)in_flight);
      }
      have_active = have_walk = false;
      if (finish_pass) {
        const size_t passes = size_t(d.h_stats[pbr::kStatRetired] / frame->npix);
        const size_t global = std::min(pass_cap, pass_offset + passes * pass_stride);
        if (global > *finish_pass) *finish_pass = global;
      }
      parity = next;   // D[next] is empty (BeginIterationKernel): nothing left for the retire after the loop
      break;
    }
    // the walks in flight (W[cur]) do not depend on anything this iteration's closest-hit and shading kernels do
    const bool fork = overlap && have_walk;
    if (fork) {
      CUDA_TRY(ctx, cudaEventRecord(d.ev_fork, st));
      CUDA_TRY(ctx, cudaStreamWaitEvent(wst, d.ev_fork, 0));
      launch_walk(parity, walk_budget, wst);
      CUDA_TRY(ctx, cudaEventRecord(d.ev_join, wst));
    }
    pbr::RetireKernel<<<grid_shade, 256, 0, st>>>(w, parity, rgba);
    mark(1);
    if (curves) pbr::TraceClosestKernel<true><<<grid_closest, kBlock, 0, st>>>(s, w, parity, refill, ctx->tune_prim_lanes, frame ? *frame : no_frame, rgba, sort);
    else pbr::TraceClosestKernel<false><<<grid_closest, kBlock, 0, st>>>(s, w, parity, refill, ctx->tune_prim_lanes, frame ? *frame : no_frame, rgba, sort);
    mark(2);
    pbr::ShadeSurfaceKernel<false><<<grid_shade, ctx->tune_shade_threads, 0, st>>>(s, w, parity, flags);
    if (sort) pbr::ShadeSurfaceKernel<true><<<grid_diffuse, ctx->tune_diffuse_threads, 0, st>>>(s, w, parity, flags);
    if (hair) pbr::ShadeHairKernel<<<grid_shade, ctx->tune_shade_threads, 0, st>>>(s, w, parity, flags);
    mark(3);
    if (fork) CUDA_TRY(ctx, cudaStreamWaitEvent(st, d.ev_join, 0));
    else if (have_walk) launch_walk(parity, walk_budget, st);   // (W[cur] is empty otherwise: the host has its length)
    mark(4);
    launch_any(next);
    mark(5);
    tm->launches += (hair ? 6 : 5) + (sort ? 1 : 0) + (have_walk ? 2 : 0);
    tm->closest_launches += 1;
    CUDA_TRY(ctx, cudaMemcpyAsync(d.h_counters, w.counters, sizeof(uint32_t) * pbr::kCounterCount,
                                  cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(d.h_stats, w.stats, sizeof(unsigned long long) * pbr::kStatCount,
                                  cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (prof) {
      float ms[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&ms[k], d.kev[k], d.kev[k + 1]);
      tm->regen_ms += ms[0]; tm->closest_ms += ms[1]; tm->shade_ms += ms[2]; tm->sss_ms += ms[3]; tm->any_ms += ms[4];
    }
    if (ctx->trace_iterations) {
      const double now = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
      fprintf(stderr, "iter %u t %.3f ms active %u walk %u done %u new %u surface %u diffuse %u exit %u shadow %u budget %u\n", it,
              (now - trace_t0) * 1e3, d.h_counters[pbr::kNumActive0 + next], d.h_counters[pbr::kNumWalk0 + next],
              d.h_counters[pbr::kNumDone0 + next], d.h_counters[pbr::kNumNew], d.h_counters[pbr::kNumSurface],
              d.h_counters[pbr::kNumDiffuse], d.h_counters[pbr::kNumExit], d.h_counters[pbr::kNumShadow], walk_budget);
    }
    have_active = d.h_counters[pbr::kNumActive0 + next] > 0;
    have_walk = d.h_counters[pbr::kNumWalk0 + next] > 0;
    samples_left = frame && d.h_stats[pbr::kStatNextSample] < total;
    if (!order_ready && d.h_stats[pbr::kStatNextSample] >= probe_samples) {
      // every probing sample has been through its first vertex: sort the pixels by the walks they started
      int rc = BuildPixelOrder(ctx, d, frame->npix);
      if (rc != PBRGPU_OK) return rc;
      tm->launches += 2;
      order_ready = true;
      w.heavy = nullptr;
    }
    in_flight = uint64_t(d.h_counters[pbr::kNumActive0 + next]) + d.h_counters[pbr::kNumWalk0 + next];
    if (samples_left) in_flight = ~0ull;
    if (frame && finish_pass) {
      const size_t passes = size_t(d.h_stats[pbr::kStatRetired] / frame->npix);
      const size_t global = std::min(pass_cap, pass_offset + passes * pass_stride);
      if (global > *finish_pass) *finish_pass = global;
    }
    parity = next;
  }
  if (single) {
    // the walks the single vertex started: to their end, then their exit vertices and shadow rays.  The set-up kernel
    // runs in any case: it clears the counters of the other parity, whose records are the paths BEFORE the vertex.
    const uint32_t next = parity ^ 1u;
    pbr::BeginIterationKernel<<<1, 32, 0, st>>>(w.counters, w.stats, parity, 0u, w.capacity, 0ull);
    tm->launches += 1;
    if (have_walk) {
      launch_walk(parity, 0x7fffffffu, st);
      launch_any(next);
      tm->launches += 3;
    }
  } else if (!(cancel && *cancel)) {
    // the paths that ended in the last iteration (a cancelled frame drops what is in flight instead)
    pbr::RetireKernel<<<grid_shade, 256, 0, st>>>(w, parity, rgba);
    tm->launches += 1;
    if (frame && finish_pass) {
      CUDA_TRY(ctx, cudaMemcpyAsync(d.h_stats, w.stats, sizeof(unsigned long long) * pbr::kStatCount,
                                    cudaMemcpyDeviceToHost, st));
      CUDA_TRY(ctx, cudaStreamSynchronize(st));
      const size_t passes = size_t(d.h_stats[pbr::kStatRetired] / frame->npix);
      const size_t global = std::min(pass_cap, pass_offset + passes * pass_stride);
      if (global > *finish_pass) *finish_pass = global;
    }
  }
  CUDA_TRY(ctx, cudaGetLastError());
  return PBRGPU_OK;
}

int FetchStats(pbrgpu_ctx* ctx, Device& d) {
  CUDA_TRY(ctx, cudaMemcpyAsync(d.h_stats, d.stats.ptr, sizeof(unsigned long long) * pbr::kStatCount,
                                cudaMemcpyDeviceToHost, d.stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  return PBRGPU_OK;
}

// number of path slots kept in flight: enough to fill the GPU many times over (each iteration costs one host
// round-trip), small enough that the SoA state (~300 B/slot) streams through HBM quickly
uint32_t ChoosePoolSize(const pbrgpu_ctx* ctx, uint64_t npix, uint64_t total_samples) {
  uint64_t n = uint64_t(ctx->tune_pool_mi) << 20;
  if (ctx->wave_spp) {
    n = uint64_t(ctx->wave_spp) * npix;
  } else {
    const uint64_t lo = uint64_t(ctx->tune_pool_min_mi) << 20;
    n = std::min(n, std::max(lo, total_samples / uint64_t(ctx->tune_pool_div)));
  }
  n = std::min<uint64_t>(n, total_samples);
  n = std::min<uint64_t>(n, 64ull << 20);
  return uint32_t(std::max<uint64_t>(n, 1));
}

// the body of Render() on one device; result left in d.rgba / d.count
int RenderOnDevice(pbrgpu_ctx* ctx, Device& d, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed,
                   uint32_t sample_offset, uint32_t sample_stride, const volatile int* cancel, size_t* finish_pass,
                   LoopTimers* tm) {
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  const uint64_t npix64 = uint64_t(width) * height;
  if (npix64 == 0 || npix64 > 0x7fffffffull) { SetError(ctx, "pbrgpu_render: bad image size"); return PBRGPU_ERR_INVALID; }
  const uint32_t npix = uint32_t(npix64);
  const uint32_t local_spp = pbrjob::CountOf(pbrjob::Share{sample_offset, sample_stride, true}, spp);
  CUDA_TRY(ctx, d.rgba.Alloc(npix));
  CUDA_TRY(ctx, d.count.Alloc(npix));
  d.frame_width = width; d.frame_height = height;
  CUDA_TRY(ctx, cudaMemsetAsync(d.rgba.ptr, 0, sizeof(float4) * npix, d.stream));
  CUDA_TRY(ctx, cudaMemsetAsync(d.count.ptr, 0, sizeof(uint32_t) * npix, d.stream));
  if (local_spp == 0) {   // more devices than samples: this one contributes zeros (and no stale statistics)
    if (d.h_stats) memset(d.h_stats, 0, sizeof(unsigned long long) * pbr::kStatCount);
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    return PBRGPU_OK;
  }
  const uint64_t total = npix64 * local_spp;
  uint32_t n_slots = ChoosePoolSize(ctx, npix, total);
  if (n_slots > d.wave_capacity) {
    // 32 Mi paths in flight are 27 GB of queue storage (kWaveBytesPerPath): never more than half of what the device
    // has free (plus what the current pool would give back), whatever else lives on it
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
    const uint64_t avail = uint64_t(free_b) + uint64_t(d.wave_capacity) * kWaveBytesPerPath;
    const uint64_t fit = std::max<uint64_t>(avail / 2 / kWaveBytesPerPath, 1u << 16);
    n_slots = uint32_t(std::min<uint64_t>(n_slots, std::max<uint64_t>(fit, d.wave_capacity)));
  }
  int rc = EnsureWave(ctx, d, n_slots);
  if (rc != PBRGPU_OK) return rc;
  CUDA_TRY(ctx, cudaMemsetAsync(d.stats.ptr, 0, sizeof(unsigned long long) * pbr::kStatCount, d.stream));

  pbr::FrameParams frame;
  float c8[8];
  pbrhost::MakeCamera(ctx->host.bmin, ctx->host.bmax, width, height, c8);
  frame.cam.eye[0] = c8[0]; frame.cam.eye[1] = c8[1]; frame.cam.eye[2] = c8[2];
  frame.cam.x_corner = c8[3]; frame.cam.y_corner = c8[4]; frame.cam.z_corner = c8[5];
  frame.cam.dx = c8[6]; frame.cam.dy = c8[7];
  frame.cam.width = width; frame.cam.height = height;
  frame.npix = npix;
  frame.total_samples = total;
  frame.seed = seed;
  frame.first_sample = sample_offset;
  frame.sample_stride = sample_stride;
  frame.rgba = d.rgba.ptr;
  frame.order = nullptr;
  frame.probe_passes = frame.rest_passes = 0;
  frame.order_block = uint32_t(std::min(std::max(ctx->tune_order_block, 1024), 1 << 16));
  // Longest paths first: only scenes whose paths differ by orders of magnitude (random walks) and frames long enough
  // to have samples left after the probing passes; the probe is at most one pool fill
  if (ctx->tune_order && HasScatteringMaterial(ctx->host) && local_spp < 65536u) {
    const uint32_t probe = std::min<uint32_t>(uint32_t(ctx->tune_order_probe), std::max<uint32_t>(n_slots / npix, 1u));
    if (local_spp > probe) {
      rc = EnsurePixelOrder(ctx, d, npix);
      if (rc != PBRGPU_OK) return rc;
      frame.order = d.order.ptr;
      frame.probe_passes = probe;
      frame.rest_passes = local_spp - probe;
    }
  }

  CUDA_TRY(ctx, cudaEventRecord(d.ev[0], d.stream));
  pbr::ResetPoolKernel<<<1, 64, 0, d.stream>>>(d.wave);
  tm->launches++;
  pbr::ShadeFlags flags;
  flags.skip_emission_and_roulette = 0;
  // n_slots paths in flight for THIS frame (the allocation, whose size is the stride of the field arrays, may be
  // larger from an earlier one)
  rc = RunPool(ctx, d, &frame, d.rgba.ptr, n_slots, 0xffffffffu, flags, tm, cancel, finish_pass, sample_offset,
               sample_stride, spp);
  if (rc != PBRGPU_OK) return rc;
  CUDA_TRY(ctx, cudaEventRecord(d.ev[1], d.stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
  tm->device_ms = ms;
  return FetchStats(ctx, d);
}

// fallback of the in-process frame-end sum when NCCL cannot be loaded: dst += src (src = a peer copy)
__global__ void AddBuffersKernel(float4* dst, const float4* src, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 a = dst[i];
  const float4 b = src[i];
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  dst[i] = a;
}

// Output stage (pc/pbrlab-cli.cc:47-57): mean = sum / count, LinerTosRGB on r,g,b (image-utils.cc:26-38: 12.92 c
// below 0.0031308, else pow(1.055 c, 1/2.4) - 0.055), alpha untouched, then WritePNG's quantisation
// (unsigned char)clamp(v * 256, 0, 255) (io/image-io.cc:200-206).  One thread per pixel, 16 B in, 4 B out.
__device__ __forceinline__ float LinearToSrgbDev(float c) {
  if (c <= 0.0031308f) return 12.92f * c;
  return powf((1.0f + 0.055f) * c, float(1.0 / 2.4)) - 0.055f;
}
__device__ __forceinline__ unsigned char Quantise8(float v) {
  const float x = fmaxf(0.0f, fminf(255.0f, v * 256.0f));   // Clamp(x, 0, 255) = std::max(a, std::min(b, x)): a NaN becomes 255 in both (std::min keeps b, fminf drops the NaN)
  return static_cast<unsigned char>(x);
}
__global__ void ResolveSrgb8Kernel(const float4* __restrict__ rgba, const uint32_t* __restrict__ count, uint32_t npix,
                                   uchar4* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const float4 a = rgba[i];
  const float n = float(count[i]);
  uchar4 o;
  o.x = Quantise8(LinearToSrgbDev(a.x / n));
  o.y = Quantise8(LinearToSrgbDev(a.y / n));
  o.z = Quantise8(LinearToSrgbDev(a.z / n));
  o.w = Quantise8(a.w / n);
  out[i] = o;
}

__global__ void KatKernel(int op, const float* params, const float* in, uint32_t in_stride, uint64_t n, float* out,
                          uint32_t out_stride) {
  const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  pbr::KatEval(op, params, in + i * in_stride, out + i * out_stride, out_stride);
}

// pbrgpu_shade reports the face direction and t of the first hit: path p = item p of S[0] after a closest-hit pass
// over hit records preset to "miss" (the kernel writes a record only for a hit)
__global__ void SurfaceFaceKernel(pbr::SceneView s, pbr::WaveState w, uint32_t n, float* face_t) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  face_t[2 * p] = -1.f;
  face_t[2 * p + 1] = 0.f;
  const pbr::HitT hit = pbr::LoadHit(w, 0u, p);
  if (hit.prim == pbr::kInvalid) return;
  const pbr::RayT ray = pbr::LoadRay(w, 0u, p);
  const pbr::Surface si = pbr::MakeSurface(s, ray, hit);
  face_t[2 * p] = float(si.face);
  face_t[2 * p + 1] = hit.t;
}

__global__ void MegaRadianceKernel(pbr::SceneView s, const float4* rays, const uint64_t* seeds, uint32_t n, float* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pbr::RayT ray;
  const float4 o = rays[2 * i], d = rays[2 * i + 1];
  ray.o = pbr::vec3(o.x, o.y, o.z); ray.tmin = o.w;
  ray.d = pbr::vec3(d.x, d.y, d.z); ray.tmax = d.w;
  pbr::Pcg32 rng;
  pbr::pcg32_srandom(&rng, seeds[2 * i], seeds[2 * i + 1]);
  const pbr::vec3 L = pbr::PathRadiance(s, ray, &rng, nullptr);
  out[3 * i] = L.x; out[3 * i + 1] = L.y; out[3 * i + 2] = L.z;
}

bool CheckCommitted(pbrgpu_ctx* ctx, const char* who) {
  if (!ctx) return false;
  if (!ctx->committed) {
    ctx->error = std::string(who) + ": scene not committed";
    return false;
  }
  return true;
}

}  // namespace

// ---- gather microbenchmark (measurement hook, see pbrgpu.h)
namespace {
template <int CHAINS>
__global__ void __launch_bounds__(128) GatherKernel(const float4* __restrict__ recs, uint32_t num_recs, uint32_t iters,
                                                     float* sink) {
  uint32_t idx[CHAINS];
  float acc = 0.f;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) idx[c] = (tid * 2654435761u + uint32_t(c) * 40503u) % num_recs;
  for (uint32_t i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      const float4* r = recs + size_t(idx[c]) * 5;
      const float4 a = r[0], b = r[1], d = r[2], e = r[3], f = r[4];
      acc += a.x + b.y + d.z + e.w + f.x;
      // the next position depends on the loaded data (like a child index read from a node)
      idx[c] = (idx[c] * 1664525u + 1013904223u + (__float_as_uint(a.w) & 1u)) % num_recs;
    }
  }
  if (acc == 1.2345e30f) *sink = acc;
}
}  // namespace

extern "C" {

int pbrgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

pbrgpu_ctx* pbrgpu_create(const int* device_ids, int n_devices) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("pbrgpu_create: no CUDA device (") + cudaGetErrorString(e) +
                     "); this backend has no CPU fallback";
    return nullptr;
  }
  std::vector<int> ids;
  if (device_ids && n_devices > 0) ids.assign(device_ids, device_ids + n_devices);
  else {
    int cur = 0;
    cudaGetDevice(&cur);
    ids.push_back(cur);
  }
  pbrgpu_ctx* ctx = new pbrgpu_ctx();
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  auto env_int = [](const char* name, int def) { const char* v = getenv(name); return v && *v ? atoi(v) : def; };
  ctx->tune_refill = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_REFILL", int(ctx->tune_refill)))));
  ctx->tune_refill_any = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_REFILL_ANY", int(ctx->tune_refill_any)))));
  ctx->tune_refill_curves = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_REFILL", int(ctx->tune_refill_curves)))));
  ctx->tune_refill_sss = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_REFILL_SSS", int(ctx->tune_refill_sss)))));
  ctx->tune_prim_lanes = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_PRIM_LANES", int(ctx->tune_prim_lanes)))));
  ctx->tune_prim_lanes_sss = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_PRIM_LANES_SSS", int(ctx->tune_prim_lanes_sss)))));
  ctx->tune_ribbon_lanes = uint32_t(std::min(32, std::max(1, env_int("PBRGPU_RIBBON_LANES", int(ctx->tune_ribbon_lanes)))));
  ctx->tune_trace_blocks = std::max(1, env_int("PBRGPU_TRACE_BLOCKS", ctx->tune_trace_blocks));
  ctx->tune_shade_blocks = std::max(1, env_int("PBRGPU_SHADE_BLOCKS", ctx->tune_shade_blocks));
  ctx->tune_walk_blocks = std::max(1, env_int("PBRGPU_WALK_BLOCKS", ctx->tune_walk_blocks));
  ctx->tune_pool_mi = std::min(64, std::max(1, env_int("PBRGPU_POOL_MI", ctx->tune_pool_mi)));
  ctx->tune_l2_persist = env_int("PBRGPU_L2_PERSIST", ctx->tune_l2_persist);
  ctx->tune_sss_skip = env_int("PBRGPU_SSS_SKIP", ctx->tune_sss_skip);
  ctx->tune_walk_bounces = std::min(8192, std::max(1, env_int("PBRGPU_WALK_BOUNCES", ctx->tune_walk_bounces)));
  ctx->tune_pool_div = std::max(1, env_int("PBRGPU_POOL_DIV", ctx->tune_pool_div));
  ctx->tune_pool_min_mi = std::min(64, std::max(1, env_int("PBRGPU_POOL_MIN_MI", ctx->tune_pool_min_mi)));
  ctx->tune_clear_march = std::min(64, std::max(1, env_int("PBRGPU_CLEAR_MARCH", ctx->tune_clear_march)));
  ctx->tune_drain_paths = std::max(0, env_int("PBRGPU_DRAIN_PATHS", ctx->tune_drain_paths));
  ctx->tune_drain_bounces = std::min(1 << 20, std::max(1, env_int("PBRGPU_DRAIN_BOUNCES", ctx->tune_drain_bounces)));
  ctx->tune_drain2_paths = std::max(0, env_int("PBRGPU_DRAIN2_PATHS", ctx->tune_drain2_paths));
  ctx->tune_drain2_bounces = std::min(1 << 20, std::max(1, env_int("PBRGPU_DRAIN2_BOUNCES", ctx->tune_drain2_bounces)));
  ctx->tune_sort_materials = env_int("PBRGPU_SORT_MATERIALS", ctx->tune_sort_materials);
  ctx->tune_overlap = env_int("PBRGPU_OVERLAP", ctx->tune_overlap);
  ctx->tune_trace_blocks_overlap = std::max(1, env_int("PBRGPU_TRACE_BLOCKS_OVERLAP", ctx->tune_trace_blocks_overlap));
  ctx->tune_walk_blocks_overlap = std::max(1, env_int("PBRGPU_WALK_BLOCKS_OVERLAP", ctx->tune_walk_blocks_overlap));
  ctx->tune_order = env_int("PBRGPU_ORDER", ctx->tune_order);
  ctx->tune_thin_spread = env_int("PBRGPU_THIN", ctx->tune_thin_spread);
  ctx->tune_inside_first = env_int("PBRGPU_INSIDE_FIRST", ctx->tune_inside_first);
  ctx->tune_finish_paths = std::max(0, env_int("PBRGPU_FINISH_PATHS", ctx->tune_finish_paths));
  ctx->tune_finish_blocks = std::max(1, env_int("PBRGPU_FINISH_BLOCKS", ctx->tune_finish_blocks));
  ctx->tune_order_probe = std::max(1, env_int("PBRGPU_ORDER_PROBE", ctx->tune_order_probe));
  ctx->tune_order_block = std::max(1, env_int("PBRGPU_ORDER_BLOCK", ctx->tune_order_block));
  ctx->trace_iterations = env_int("PBRGPU_TRACE_ITERATIONS", 0) != 0;
  ctx->tune_diffuse_blocks = std::max(1, env_int("PBRGPU_DIFFUSE_BLOCKS", ctx->tune_diffuse_blocks));
  ctx->tune_diffuse_threads = std::min(pbr::kDiffuseBlock, std::max(32, env_int("PBRGPU_DIFFUSE_THREADS", ctx->tune_diffuse_threads) & ~31));
  ctx->tune_shade_threads = std::min(pbr::kShadeBlock, std::max(32, env_int("PBRGPU_SHADE_THREADS", ctx->tune_shade_threads) & ~31));
  for (int id : ids) {
    if (id < 0 || id >= ndev) {
      g_create_error = "pbrgpu_create: device id out of range";
      delete ctx;
      return nullptr;
    }
    Device d;
    d.id = id;
    cudaDeviceProp prop;
    if (cudaSetDevice(id) != cudaSuccess || cudaGetDeviceProperties(&prop, id) != cudaSuccess ||
        cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&d.walk_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d.ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&d.ev[0]) != cudaSuccess || cudaEventCreate(&d.ev[1]) != cudaSuccess) {
      g_create_error = std::string("pbrgpu_create: cannot initialise device: ") + cudaGetErrorString(cudaGetLastError());
      delete ctx;
      return nullptr;
    }
    d.sm_count = prop.multiProcessorCount;
    for (auto& e : d.kev) cudaEventCreate(&e);
    ctx->devices.push_back(d);
  }
  // peer access for the frame-end sum on multi-device contexts
  for (size_t a = 1; a < ctx->devices.size(); ++a) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, ctx->devices[0].id, ctx->devices[a].id);
    if (can) {
      cudaSetDevice(ctx->devices[0].id);
      cudaDeviceEnablePeerAccess(ctx->devices[a].id, 0);
      cudaGetLastError();
    }
  }
  cudaSetDevice(ctx->devices[0].id);
  return ctx;
}

void pbrgpu_destroy(pbrgpu_ctx* ctx) {
  if (!ctx) return;
  if (ctx->job_comm) pbrnccl::Get().CommDestroy(ctx->job_comm);
  for (ncclComm_t c : ctx->dev_comms) if (c) pbrnccl::Get().CommDestroy(c);
  for (Device& d : ctx->devices) {
    cudaSetDevice(d.id);
    d.Release();
  }
  delete ctx;
}

const char* pbrgpu_last_error(const pbrgpu_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int pbrgpu_set_triangles(pbrgpu_ctx* ctx, const float* xyzw, uint32_t nverts, const uint32_t* vidx, const float* nxyzw,
                         uint32_t nnormals, const uint32_t* nidx, const float* uv, uint32_t nuv, const uint32_t* tidx,
                         const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                         const uint32_t* prim_id, uint64_t ntris) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->committed = false;
  if (!ctx->host.SetTriangles(xyzw, nverts, vidx, nxyzw, nnormals, nidx, uv, nuv, tidx, material_id, instance_id,
                              geom_id, prim_id, ntris)) {
    ctx->error = ctx->host.error;
    return PBRGPU_ERR_INVALID;
  }
  return PBRGPU_OK;
}

int pbrgpu_set_curves(pbrgpu_ctx* ctx, const float* xyzr, uint32_t nverts, const uint32_t* first_cp,
                      const uint32_t* material_id, const uint32_t* instance_id, const uint32_t* geom_id,
                      const uint32_t* prim_id, uint64_t nsegs) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->committed = false;
  if (!ctx->host.SetCurves(xyzr, nverts, first_cp, material_id, instance_id, geom_id, prim_id, nsegs)) {
    ctx->error = ctx->host.error;
    return PBRGPU_ERR_INVALID;
  }
  return PBRGPU_OK;
}

int pbrgpu_set_materials(pbrgpu_ctx* ctx, const pbrgpu_material* materials, uint32_t n) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  if (ctx->committed) {   // live edit: validate against the committed geometry BEFORE the host table is replaced
    // (Render() re-uploads the table every frame: the highest id in use was recorded by the commit, no scan here)
    if (ctx->host.any_material_id && ctx->host.max_material_id >= n) {
      ctx->error = "pbrgpu_set_materials: table shrank below ids in use";
      return PBRGPU_ERR_INVALID;
    }
  }
  if (!ctx->host.SetMaterials(materials, n)) {   // validates everything else before it mutates (scene_host.cc)
    ctx->error = ctx->host.error;
    return PBRGPU_ERR_INVALID;
  }
  if (ctx->committed) {   // refresh the device tables in place
    for (Device& d : ctx->devices) {
      CUDA_TRY(ctx, cudaSetDevice(d.id));
      CUDA_TRY(ctx, d.materials.Upload(ctx->host.materials.data(), ctx->host.materials.size(), d.stream));
      CUDA_TRY(ctx, d.material_class.Upload(ctx->host.material_class.data(), ctx->host.material_class.size(), d.stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
      d.view.materials = d.materials.ptr;
      d.view.material_class = d.material_class.ptr;
      d.view.num_materials = n;
      d.view.num_hair_materials = CountHairMaterials(ctx->host);
    }
  }
  return PBRGPU_OK;
}

int pbrgpu_set_textures(pbrgpu_ctx* ctx, const pbrgpu_texture* textures, uint32_t n) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->committed = false;
  if (!ctx->host.SetTextures(textures, n)) {
    ctx->error = ctx->host.error;
    return PBRGPU_ERR_INVALID;
  }
  return PBRGPU_OK;
}

int pbrgpu_set_lights(pbrgpu_ctx* ctx, const pbrgpu_light_tables* tables) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->committed = false;
  if (!ctx->host.SetLights(tables)) {
    ctx->error = ctx->host.error;
    return PBRGPU_ERR_INVALID;
  }
  return PBRGPU_OK;
}

// HostScene::DeviceBuilder: the acceleration-structure build on the context's first device (bvh_device.cuh)
static bool DeviceBvhBuilder(void* user, const pbrbvh::Aabb* boxes, uint32_t n, const pbrbvh::BuildParams& prm,
                             uint32_t radius, pbrbvh::Bvh8* out, const char** err) {
  pbrgpu_ctx* ctx = static_cast<pbrgpu_ctx*>(user);
  Device& d = ctx->devices[0];
  if (cudaSetDevice(d.id) != cudaSuccess) { if (err) *err = "cudaSetDevice failed"; return false; }
  double sec[3] = {0, 0, 0};
  const bool ok = pbrdev::BuildBvh8OnDevice(d.stream, boxes, n, prm, radius, out, err, sec);
  if (ok && getenv("PBRGPU_VERBOSE_COMMIT"))
    fprintf(stderr, "  device BVH build (%u prims): upload %.3f s, build %.3f s, download %.3f s, %u nodes, depth %u\n", n,
            sec[0], sec[1], sec[2], out->num_nodes, out->max_depth);
  return ok;
}

int pbrgpu_commit(pbrgpu_ctx* ctx, const float* bmin, const float* bmax) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->host.device_builder = &DeviceBvhBuilder;
  ctx->host.device_builder_user = ctx;
  const auto t0 = std::chrono::steady_clock::now();
  if (!ctx->host.Commit(bmin, bmax)) {
    ctx->error = ctx->host.error;
    return PBRGPU_ERR_BUILD;
  }
  const auto t1 = std::chrono::steady_clock::now();
  for (Device& d : ctx->devices) {
    const int rc = UploadScene(ctx, d);
    if (rc != PBRGPU_OK) return rc;
  }
  const auto t2 = std::chrono::steady_clock::now();
  pbrgpu_commit_info& ci = ctx->commit_info;
  memset(&ci, 0, sizeof(ci));
  ci.commit_s = std::chrono::duration<double>(t2 - t0).count();
  ci.upload_s = std::chrono::duration<double>(t2 - t1).count();
  ci.bvh_s = ctx->host.bvh_seconds;
  ci.clearance_s = ctx->host.clearance_seconds;
  ci.tri_builder = ctx->host.last_builder == "ploc-device" ? 2u : (ctx->host.last_builder == "ploc-host" ? 1u : 0u);
  ci.tri_nodes = ctx->host.tri_bvh.num_nodes;
  ci.curve_nodes = ctx->host.curve_bvh.num_nodes;
  ci.tri_depth = ctx->host.tri_bvh.max_depth;
  if (getenv("PBRGPU_VERBOSE_COMMIT"))
    fprintf(stderr, "pbrgpu_commit: %.3f s (BVH %.3f, clearance %.3f, upload %.3f), triangle BVH by %s\n", ci.commit_s,
            ci.bvh_s, ci.clearance_s, ci.upload_s, ctx->host.last_builder.c_str());
  ctx->committed = true;
  return PBRGPU_OK;
}

int pbrgpu_get_commit_info(const pbrgpu_ctx* ctx, pbrgpu_commit_info* out) {
  if (!ctx || !out || !ctx->committed) return PBRGPU_ERR_INVALID;
  *out = ctx->commit_info;
  return PBRGPU_OK;
}

int pbrgpu_scene_bounds(const pbrgpu_ctx* ctx, float* bmin, float* bmax) {
  if (!ctx || !ctx->committed) return PBRGPU_ERR_INVALID;
  for (int k = 0; k < 3; ++k) { bmin[k] = ctx->host.bmin[k]; bmax[k] = ctx->host.bmax[k]; }
  return PBRGPU_OK;
}

// ---- multi-process jobs: one process per GPU (torchrun), the frame-end reduce runs inside the library
int pbrgpu_nccl_unique_id(uint8_t* id128) {
  const pbrnccl::Api& nccl = pbrnccl::Get();
  if (!id128) return PBRGPU_ERR_INVALID;
  if (!nccl.ok) { g_create_error = nccl.error; return PBRGPU_ERR_CUDA; }
  ncclUniqueId id;
  const ncclResult_t r = nccl.GetUniqueId(&id);
  if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + nccl.GetErrorString(r); return PBRGPU_ERR_CUDA; }
  static_assert(sizeof(id) == PBRGPU_NCCL_ID_BYTES, "ncclUniqueId size");
  memcpy(id128, &id, sizeof(id));
  return PBRGPU_OK;
}

int pbrgpu_nccl_init(pbrgpu_ctx* ctx, const uint8_t* id128, int rank, int world) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) {
    if (ctx) ctx->error = "pbrgpu_nccl_init: bad arguments";
    return PBRGPU_ERR_INVALID;
  }
  if (ctx->devices.size() != 1) { ctx->error = "pbrgpu_nccl_init: a multi-process job uses one device per process"; return PBRGPU_ERR_INVALID; }
  const pbrnccl::Api& nccl = pbrnccl::Get();
  if (!nccl.ok) { ctx->error = nccl.error; return PBRGPU_ERR_CUDA; }
  if (ctx->job_comm) { nccl.CommDestroy(ctx->job_comm); ctx->job_comm = nullptr; ctx->job_rank = 0; ctx->job_world = 1; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->devices[0].id));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_TRY(ctx, nccl.CommInitRank(&ctx->job_comm, world, id, rank));
  ctx->job_rank = rank;
  ctx->job_world = world;
  return PBRGPU_OK;
}

int pbrgpu_job_rank(const pbrgpu_ctx* ctx, int* rank, int* world) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  if (rank) *rank = ctx->job_rank;
  if (world) *world = ctx->job_world;
  return PBRGPU_OK;
}

int pbrgpu_set_wave_spp(pbrgpu_ctx* ctx, uint32_t spp_per_wave) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->wave_spp = spp_per_wave;
  return PBRGPU_OK;
}

int pbrgpu_set_profiling(pbrgpu_ctx* ctx, int enabled) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  ctx->profile = enabled != 0;
  return PBRGPU_OK;
}

int pbrgpu_get_stats(const pbrgpu_ctx* ctx, pbrgpu_stats* out) {
  if (!ctx || !out) return PBRGPU_ERR_INVALID;
  *out = ctx->stats;
  return PBRGPU_OK;
}

// Frame-end sum of the per-device accumulators onto the context's first device.  Only the float4
// sums travel: alpha counts the samples (render.cc:175-183 increments both together), so RenderLayer::count is
// derived after the reduce and ONE ncclReduce of 16 B per pixel does the whole job.
static int ReduceDevices(pbrgpu_ctx* ctx, uint32_t npix) {
  const uint32_t ndev = uint32_t(ctx->devices.size());
  if (ndev < 2) return PBRGPU_OK;
  Device& d0 = ctx->devices[0];
  const pbrnccl::Api& nccl = pbrnccl::Get();
  if (nccl.ok && getenv("PBRGPU_NO_NCCL") == nullptr) {
    if (ctx->dev_comms.empty()) {
      std::vector<int> ids;
      for (const Device& d : ctx->devices) ids.push_back(d.id);
      ctx->dev_comms.assign(ndev, nullptr);
      NCCL_TRY(ctx, nccl.CommInitAll(ctx->dev_comms.data(), int(ndev), ids.data()));
    }
    NCCL_TRY(ctx, nccl.GroupStart());
    for (uint32_t k = 0; k < ndev; ++k) {
      Device& d = ctx->devices[k];
      nccl.Reduce(d.rgba.ptr, d.rgba.ptr, size_t(npix) * 4, ncclFloat, ncclSum, 0, ctx->dev_comms[k], d.stream);
    }
    NCCL_TRY(ctx, nccl.GroupEnd());
    for (Device& d : ctx->devices) {
      CUDA_TRY(ctx, cudaSetDevice(d.id));
      CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    }
    CUDA_TRY(ctx, cudaSetDevice(d0.id));
    return PBRGPU_OK;
  }
  // no NCCL on this host: peer copies into a scratch buffer (kept with the context) + add
  CUDA_TRY(ctx, cudaSetDevice(d0.id));
  CUDA_TRY(ctx, d0.peer_tmp.Alloc(npix));
  for (uint32_t k = 1; k < ndev; ++k) {
    Device& dk = ctx->devices[k];
    CUDA_TRY(ctx, cudaMemcpyPeerAsync(d0.peer_tmp.ptr, d0.id, dk.rgba.ptr, dk.id, sizeof(float4) * npix, d0.stream));
    AddBuffersKernel<<<(npix + 255) / 256, 256, 0, d0.stream>>>(d0.rgba.ptr, d0.peer_tmp.ptr, npix);
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(d0.stream));
  return PBRGPU_OK;
}

// The frame (20 B per pixel: float4 sums + u32 count) from the device into the caller's HOST buffers.  A copy to
// pageable memory is staged by the driver through one thread (~10 GB/s: 4-5 ms for a 1080p frame, a fifth of the PCIe
// rate); here the DMA goes to a pinned buffer of the context in eight chunks and one host thread per chunk copies it on
// as soon as its event has fired.  Small frames take the plain path.
static int ReadBackFrame(pbrgpu_ctx* ctx, Device& d, uint32_t npix, float* rgba_out, uint32_t* count_out) {
  const size_t rgba_bytes = sizeof(float4) * size_t(npix), count_bytes = sizeof(uint32_t) * size_t(npix);
  constexpr int kChunks = 8;   // 6 of the sums, 2 of the counts (4 : 1 in bytes)
  static const bool plain = getenv("PBRGPU_PLAIN_READBACK") != nullptr;   // (A/B switch)
  if (plain || rgba_bytes + count_bytes < (8u << 20)) {
    CUDA_TRY(ctx, cudaMemcpyAsync(rgba_out, d.rgba.ptr, rgba_bytes, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(count_out, d.count.ptr, count_bytes, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    return PBRGPU_OK;
  }
  if (d.h_frame_bytes < rgba_bytes + count_bytes) {
    if (d.h_frame) cudaFreeHost(d.h_frame);
    d.h_frame = nullptr; d.h_frame_bytes = 0;
    CUDA_TRY(ctx, cudaMallocHost(reinterpret_cast<void**>(&d.h_frame), rgba_bytes + count_bytes));
    d.h_frame_bytes = rgba_bytes + count_bytes;
  }
  for (auto& e : d.ev_chunk)
    if (!e) CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  struct Chunk { const uint8_t* src; uint8_t* stage; uint8_t* dst; size_t bytes; };
  Chunk chunks[kChunks];
  for (int c = 0; c < kChunks; ++c) {
    const bool sums = c < 6;
    const size_t total = sums ? rgba_bytes : count_bytes, parts = sums ? 6 : 2, k = sums ? size_t(c) : size_t(c - 6);
    const size_t lo = (total * k / parts) & ~size_t(63), hi = (k + 1 == parts) ? total : ((total * (k + 1) / parts) & ~size_t(63));
    chunks[c].src = reinterpret_cast<const uint8_t*>(sums ? static_cast<const void*>(d.rgba.ptr) : static_cast<const void*>(d.count.ptr)) + lo;
    chunks[c].stage = d.h_frame + (sums ? 0 : rgba_bytes) + lo;
    chunks[c].dst = reinterpret_cast<uint8_t*>(sums ? static_cast<void*>(rgba_out) : static_cast<void*>(count_out)) + lo;
    chunks[c].bytes = hi - lo;
    CUDA_TRY(ctx, cudaMemcpyAsync(chunks[c].stage, chunks[c].src, chunks[c].bytes, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaEventRecord(d.ev_chunk[c], d.stream));
  }
  std::vector<std::thread> th;
  std::vector<cudaError_t> err(kChunks, cudaSuccess);
  for (int c = 0; c < kChunks; ++c)
    th.emplace_back([&, c]() {
      err[c] = cudaEventSynchronize(d.ev_chunk[c]);
      if (err[c] == cudaSuccess) memcpy(chunks[c].dst, chunks[c].stage, chunks[c].bytes);
    });
  for (auto& t : th) t.join();
  for (int c = 0; c < kChunks; ++c) CUDA_TRY(ctx, err[c]);
  return PBRGPU_OK;
}

static int RenderImpl(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed,
                      uint32_t sample_offset, uint32_t sample_stride, const volatile int* cancel, float* rgba_out,
                      uint32_t* count_out, size_t* finish_pass, bool out_on_device) {
  if (!CheckCommitted(ctx, "pbrgpu_render")) return PBRGPU_ERR_INVALID;
  if (sample_stride == 0 || !rgba_out || !count_out) { ctx->error = "pbrgpu_render: bad arguments"; return PBRGPU_ERR_INVALID; }
  const auto t0 = std::chrono::steady_clock::now();
  const uint32_t ndev = uint32_t(ctx->devices.size());
  const uint64_t npix64 = uint64_t(width) * height;
  if (npix64 == 0 || npix64 > 0x7fffffffull) { ctx->error = "pbrgpu_render: bad image size"; return PBRGPU_ERR_INVALID; }
  const uint32_t npix = uint32_t(npix64);
  if (finish_pass) *finish_pass = 0;
  // The job's samples sample_offset + j*sample_stride are dealt out round-robin: first over the ranks of a
  // multi-process job (pbrgpu_nccl_init), then over the devices of this context — the (tile, sample) job split of the
  // reference (render.cc:211-222) with the sample as the unit.  Every path keeps its own PCG32 stream
  // (seed + sample, pixel), so the image does not depend on the split (up to the order of the float sums).
  const uint32_t world = uint32_t(ctx->job_world), rank = uint32_t(ctx->job_rank);
  std::vector<pbrjob::Share> shares(ndev);
  for (uint32_t k = 0; k < ndev; ++k) {
    shares[k] = pbrjob::ShareOf(sample_offset, sample_stride, rank, world, k, ndev);
    if (!shares[k].ok) { ctx->error = "pbrgpu_render: sample stride overflow"; return PBRGPU_ERR_INVALID; }
  }
  std::vector<LoopTimers> tms(ndev);
  std::vector<int> rcs(ndev, PBRGPU_OK);
  std::vector<size_t> progress(ndev, 0);
  if (ndev == 1) {
    rcs[0] = RenderOnDevice(ctx, ctx->devices[0], width, height, spp, seed, shares[0].offset, shares[0].stride, cancel,
                            finish_pass, &tms[0]);
  } else {
    std::vector<std::thread> th;   // one host thread per device
    for (uint32_t k = 0; k < ndev; ++k) {
      th.emplace_back([&, k]() {
        rcs[k] = RenderOnDevice(ctx, ctx->devices[k], width, height, spp, seed, shares[k].offset, shares[k].stride,
                                cancel, &progress[k], &tms[k]);
      });
    }
    for (auto& t : th) t.join();
    if (finish_pass) {
      size_t m = spp;
      for (uint32_t k = 0; k < ndev; ++k) m = std::min(m, progress[k]);
      *finish_pass = m;   // every device finished at least this many of the interleaved passes
    }
  }
  for (uint32_t k = 0; k < ndev; ++k) if (rcs[k] != PBRGPU_OK) return rcs[k];

  // ---- frame end: sums onto device 0 of this context, then onto rank 0 of the job; count from alpha
  Device& d0 = ctx->devices[0];
  int rc = ReduceDevices(ctx, npix);
  if (rc != PBRGPU_OK) return rc;
  CUDA_TRY(ctx, cudaSetDevice(d0.id));
  if (ctx->job_comm) {
    const pbrnccl::Api& nccl = pbrnccl::Get();
    NCCL_TRY(ctx, nccl.Reduce(d0.rgba.ptr, d0.rgba.ptr, size_t(npix) * 4, ncclFloat, ncclSum, 0, ctx->job_comm, d0.stream));
  }
  pbr::FinishFrameKernel<<<(npix + 255) / 256, 256, 0, d0.stream>>>(d0.rgba.ptr, d0.count.ptr, npix);
  tms[0].launches++;
  if (out_on_device) {
    CUDA_TRY(ctx, cudaMemcpyAsync(rgba_out, d0.rgba.ptr, sizeof(float4) * npix, cudaMemcpyDeviceToDevice, d0.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(count_out, d0.count.ptr, sizeof(uint32_t) * npix, cudaMemcpyDeviceToDevice, d0.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d0.stream));
  } else {
    rc = ReadBackFrame(ctx, d0, npix, rgba_out, count_out);
    if (rc != PBRGPU_OK) return rc;
  }
  CUDA_TRY(ctx, cudaGetLastError());

  pbrgpu_stats& s = ctx->stats;
  memset(&s, 0, sizeof(s));
  for (uint32_t k = 0; k < ndev; ++k) {
    const Device& d = ctx->devices[k];
    if (d.h_stats) {
      s.closest_rays += d.h_stats[pbr::kStatClosest];
      s.shadow_rays += d.h_stats[pbr::kStatShadow];
      s.sss_rays += d.h_stats[pbr::kStatSss];
      s.sss_skipped += d.h_stats[pbr::kStatSssSkipped];
      s.shade_vertices += d.h_stats[pbr::kStatVertices];
      s.paths += d.h_stats[pbr::kStatRetired];   // camera samples this process accumulated (< its share after a cancel)
    }
    s.kernel_launches += tms[k].launches;
    s.trace_closest_launches += tms[k].closest_launches;
    s.iterations = std::max<uint64_t>(s.iterations, tms[k].closest_launches);
    s.trace_closest_ms += tms[k].closest_ms; s.trace_any_ms += tms[k].any_ms; s.shade_ms += tms[k].shade_ms;
    s.sss_ms += tms[k].sss_ms; s.regen_ms += tms[k].regen_ms;
    s.device_ms = std::max(s.device_ms, tms[k].device_ms);
  }
  s.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return PBRGPU_OK;
}

int pbrgpu_resolve_srgb8_device(pbrgpu_ctx* ctx, const float* d_rgba, const uint32_t* d_count, uint32_t width,
                                uint32_t height, uint8_t* d_rgba8_out) {
  if (!ctx || ctx->devices.empty() || !d_rgba || !d_count || !d_rgba8_out) return PBRGPU_ERR_INVALID;
  const uint64_t npix = uint64_t(width) * height;
  if (npix == 0 || npix > 0x7fffffffull) { ctx->error = "pbrgpu_resolve_srgb8: bad image size"; return PBRGPU_ERR_INVALID; }
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  ResolveSrgb8Kernel<<<unsigned((npix + 255) / 256), 256, 0, d.stream>>>(
      reinterpret_cast<const float4*>(d_rgba), d_count, uint32_t(npix), reinterpret_cast<uchar4*>(d_rgba8_out));
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  return PBRGPU_OK;
}

int pbrgpu_resolve_srgb8(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint8_t* rgba8_out) {
  if (!ctx || ctx->devices.empty() || !rgba8_out) return PBRGPU_ERR_INVALID;
  Device& d = ctx->devices[0];
  if (!d.rgba.ptr || d.frame_width != width || d.frame_height != height) {
    ctx->error = "pbrgpu_resolve_srgb8: no rendered frame of this size on the device";
    return PBRGPU_ERR_INVALID;
  }
  const size_t npix = size_t(width) * height;
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  CUDA_TRY(ctx, d.srgb8.Alloc(npix * 4));
  const int rc = pbrgpu_resolve_srgb8_device(ctx, reinterpret_cast<const float*>(d.rgba.ptr), d.count.ptr, width,
                                             height, d.srgb8.ptr);
  if (rc != PBRGPU_OK) return rc;
  CUDA_TRY(ctx, cudaMemcpy(rgba8_out, d.srgb8.ptr, npix * 4, cudaMemcpyDeviceToHost));
  return PBRGPU_OK;
}

int pbrgpu_render(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed,
                  uint32_t sample_offset, uint32_t sample_stride, const volatile int* cancel, float* rgba_out,
                  uint32_t* count_out, size_t* finish_pass) {
  return RenderImpl(ctx, width, height, spp, seed, sample_offset, sample_stride, cancel, rgba_out, count_out,
                    finish_pass, false);
}

int pbrgpu_render_device(pbrgpu_ctx* ctx, uint32_t width, uint32_t height, uint32_t spp, uint64_t seed,
                         uint32_t sample_offset, uint32_t sample_stride, const volatile int* cancel, float* d_rgba,
                         uint32_t* d_count, size_t* finish_pass) {
  return RenderImpl(ctx, width, height, spp, seed, sample_offset, sample_stride, cancel, d_rgba, d_count, finish_pass,
                    true);
}

// ------------------------------------------------------------------------------------------------ test hooks
int pbrgpu_trace_device(pbrgpu_ctx* ctx, const pbrgpu_ray* d_rays, uint64_t n, pbrgpu_hit* d_hits, int collect_stats) {
  if (!CheckCommitted(ctx, "pbrgpu_trace_device")) return PBRGPU_ERR_INVALID;
  if (n > 0xfffffff0ull) { ctx->error = "pbrgpu_trace_device: at most 2^32-16 rays per call"; return PBRGPU_ERR_INVALID; }
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  int rc = EnsureWave(ctx, d, 1);
  if (rc != PBRGPU_OK) return rc;
  DevBuf<float4> tuv, ng;
  DevBuf<uint4> ids;
  if (d_hits) {
    CUDA_TRY(ctx, tuv.Alloc(n)); CUDA_TRY(ctx, ng.Alloc(n)); CUDA_TRY(ctx, ids.Alloc(n));
  }
  CUDA_TRY(ctx, cudaMemsetAsync(d.counters.ptr, 0, sizeof(uint32_t) * pbr::kCounterCount, d.stream));
  CUDA_TRY(ctx, cudaMemsetAsync(d.stats.ptr, 0, sizeof(unsigned long long) * pbr::kStatCount, d.stream));
  CUDA_TRY(ctx, cudaEventRecord(d.ev[0], d.stream));
  {
    const int grid = PersistentGrid(d, ctx->tune_trace_blocks);
    const float4* r4 = reinterpret_cast<const float4*>(d_rays);
    uint32_t* fetch = d.counters.ptr + pbr::kFetchTrace;
    const bool curves = d.view.num_curves != 0u;
#define PBR_LAUNCH_BATCH(C, S)                                                                                    \
    pbr::TraceBatchKernel<C, S><<<grid, kBlock, 0, d.stream>>>(d.view, r4, n, tuv.ptr, ids.ptr, ng.ptr, fetch,   \
                                                                d.stats.ptr, ctx->tune_refill, ctx->tune_prim_lanes)
    if (curves && collect_stats) PBR_LAUNCH_BATCH(true, true);
    else if (curves) PBR_LAUNCH_BATCH(true, false);
    else if (collect_stats) PBR_LAUNCH_BATCH(false, true);
    else PBR_LAUNCH_BATCH(false, false);
#undef PBR_LAUNCH_BATCH
  }
  CUDA_TRY(ctx, cudaEventRecord(d.ev[1], d.stream));
  if (d_hits) {
    // interleave the three SoA outputs into the caller's array-of-structs (pbrgpu_hit = 9 words)
    CUDA_TRY(ctx, cudaMemcpy2DAsync(reinterpret_cast<char*>(d_hits) + 0, sizeof(pbrgpu_hit), ng.ptr, sizeof(float4),
                                    12, n, cudaMemcpyDeviceToDevice, d.stream));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(reinterpret_cast<char*>(d_hits) + 12, sizeof(pbrgpu_hit), tuv.ptr, sizeof(float4),
                                    12, n, cudaMemcpyDeviceToDevice, d.stream));
    CUDA_TRY(ctx, cudaMemcpy2DAsync(reinterpret_cast<char*>(d_hits) + 24, sizeof(pbrgpu_hit), ids.ptr, sizeof(uint4),
                                    12, n, cudaMemcpyDeviceToDevice, d.stream));
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  CUDA_TRY(ctx, cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
  rc = FetchStats(ctx, d);
  if (rc != PBRGPU_OK) return rc;
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->stats.closest_rays = n;
  ctx->stats.trace_closest_ms = ms;
  ctx->stats.nodes_visited = d.h_stats[pbr::kStatNodes];
  ctx->stats.prims_tested = d.h_stats[pbr::kStatPrims];
  ctx->stats.kernel_launches = 1;
  tuv.Free(); ng.Free(); ids.Free();
  return PBRGPU_OK;
}

int pbrgpu_occluded_device(pbrgpu_ctx* ctx, const pbrgpu_ray* d_rays, uint64_t n, uint8_t* d_occluded) {
  if (!CheckCommitted(ctx, "pbrgpu_occluded_device")) return PBRGPU_ERR_INVALID;
  if (n > 0xfffffff0ull) { ctx->error = "pbrgpu_occluded_device: at most 2^32-16 rays per call"; return PBRGPU_ERR_INVALID; }
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  int rc = EnsureWave(ctx, d, 1);
  if (rc != PBRGPU_OK) return rc;
  CUDA_TRY(ctx, cudaMemsetAsync(d.counters.ptr, 0, sizeof(uint32_t) * pbr::kCounterCount, d.stream));
  CUDA_TRY(ctx, cudaEventRecord(d.ev[0], d.stream));
  if (d.view.num_curves)
    pbr::OccludedBatchKernel<true><<<PersistentGrid(d, ctx->tune_trace_blocks), kBlock, 0, d.stream>>>(
        d.view, reinterpret_cast<const float4*>(d_rays), n, d_occluded, d.counters.ptr + pbr::kFetchShadow,
        ctx->tune_refill, ctx->tune_prim_lanes);
  else
    pbr::OccludedBatchKernel<false><<<PersistentGrid(d, ctx->tune_trace_blocks), kBlock, 0, d.stream>>>(
        d.view, reinterpret_cast<const float4*>(d_rays), n, d_occluded, d.counters.ptr + pbr::kFetchShadow,
        ctx->tune_refill, ctx->tune_prim_lanes);
  CUDA_TRY(ctx, cudaEventRecord(d.ev[1], d.stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  CUDA_TRY(ctx, cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]);
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->stats.shadow_rays = n;
  ctx->stats.trace_any_ms = ms;
  ctx->stats.kernel_launches = 1;
  return PBRGPU_OK;
}

// ---- gather microbenchmark (measurement hook, see pbrgpu.h; kernel: GatherKernel above)

int pbrgpu_measure_gather(pbrgpu_ctx* ctx, uint64_t working_set_bytes, uint32_t records_per_thread, uint32_t chains,
                          double* gbytes_per_s) {
  if (!ctx || !gbytes_per_s || ctx->devices.empty()) return PBRGPU_ERR_INVALID;
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  const uint32_t num_recs = uint32_t(std::min<uint64_t>(std::max<uint64_t>(working_set_bytes / 80, 1024), 0x7fffffffull / 5));
  DevBuf<float4> recs;
  DevBuf<float> sink;
  CUDA_TRY(ctx, recs.Alloc(size_t(num_recs) * 5));
  CUDA_TRY(ctx, sink.Alloc(1));
  CUDA_TRY(ctx, cudaMemsetAsync(recs.ptr, 0, size_t(num_recs) * 80, d.stream));
  const int grid = PersistentGrid(d, ctx->tune_trace_blocks);
  const uint32_t per_chain = std::max(1u, records_per_thread / std::max(1u, chains));
  float best_ms = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CUDA_TRY(ctx, cudaEventRecord(d.ev[0], d.stream));
    if (chains >= 8) GatherKernel<8><<<grid, kBlock, 0, d.stream>>>(recs.ptr, num_recs, per_chain, sink.ptr);
    else if (chains >= 4) GatherKernel<4><<<grid, kBlock, 0, d.stream>>>(recs.ptr, num_recs, per_chain, sink.ptr);
    else if (chains >= 2) GatherKernel<2><<<grid, kBlock, 0, d.stream>>>(recs.ptr, num_recs, per_chain, sink.ptr);
    else GatherKernel<1><<<grid, kBlock, 0, d.stream>>>(recs.ptr, num_recs, per_chain, sink.ptr);
    CUDA_TRY(ctx, cudaEventRecord(d.ev[1], d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, d.ev[0], d.ev[1]));
    if (rep > 0) best_ms = std::min(best_ms, ms);   // the first launch warms the working set into L2
  }
  const uint32_t c = chains >= 8 ? 8 : chains >= 4 ? 4 : chains >= 2 ? 2 : 1;
  const double bytes = double(grid) * kBlock * double(per_chain) * c * 80.0;
  *gbytes_per_s = bytes / (double(best_ms) * 1e-3) / 1e9;
  recs.Free(); sink.Free();
  return PBRGPU_OK;
}

int pbrgpu_trace(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, uint64_t n, pbrgpu_hit* hits) {
  if (!CheckCommitted(ctx, "pbrgpu_trace")) return PBRGPU_ERR_INVALID;
  if (n == 0) return PBRGPU_OK;
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  DevBuf<pbrgpu_ray> dr;
  DevBuf<pbrgpu_hit> dh;
  CUDA_TRY(ctx, dr.Upload(rays, n, d.stream));
  CUDA_TRY(ctx, dh.Alloc(n));
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  const int rc = pbrgpu_trace_device(ctx, dr.ptr, n, dh.ptr, 1);
  if (rc == PBRGPU_OK) {
    CUDA_TRY(ctx, cudaMemcpy(hits, dh.ptr, sizeof(pbrgpu_hit) * n, cudaMemcpyDeviceToHost));
  }
  dr.Free(); dh.Free();
  return rc;
}

int pbrgpu_occluded(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, uint64_t n, uint8_t* occluded) {
  if (!CheckCommitted(ctx, "pbrgpu_occluded")) return PBRGPU_ERR_INVALID;
  if (n == 0) return PBRGPU_OK;
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  DevBuf<pbrgpu_ray> dr;
  DevBuf<uint8_t> dout;
  CUDA_TRY(ctx, dr.Upload(rays, n, d.stream));
  CUDA_TRY(ctx, dout.Alloc(n));
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  const int rc = pbrgpu_occluded_device(ctx, dr.ptr, n, dout.ptr);
  if (rc == PBRGPU_OK) CUDA_TRY(ctx, cudaMemcpy(occluded, dout.ptr, n, cudaMemcpyDeviceToHost));
  dr.Free(); dout.Free();
  return rc;
}

// mode: 0 = wavefront to termination (pbrgpu_radiance), 1 = single vertex (pbrgpu_shade), 2 = megakernel
static int PathHook(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* out, int mode) {
  if (!CheckCommitted(ctx, "pbrgpu_radiance")) return PBRGPU_ERR_INVALID;
  if (n == 0) return PBRGPU_OK;
  if (n > 0x7fffffffull) { ctx->error = "path hook: too many paths"; return PBRGPU_ERR_INVALID; }
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  int rc = EnsureWave(ctx, d, uint32_t(n));
  if (rc != PBRGPU_OK) return rc;
  DevBuf<pbrgpu_ray> dr;
  DevBuf<uint64_t> ds;
  CUDA_TRY(ctx, dr.Upload(rays, n, d.stream));
  CUDA_TRY(ctx, ds.Upload(seeds, 2 * n, d.stream));
  CUDA_TRY(ctx, cudaMemsetAsync(d.stats.ptr, 0, sizeof(unsigned long long) * pbr::kStatCount, d.stream));
  const uint32_t n32 = uint32_t(n);
  LoopTimers tm;
  if (mode == 2) {
    DevBuf<float> dout;
    CUDA_TRY(ctx, dout.Alloc(3 * n));
    MegaRadianceKernel<<<(n32 + 63) / 64, 64, 0, d.stream>>>(d.view, reinterpret_cast<const float4*>(dr.ptr), ds.ptr, n32, dout.ptr);
    CUDA_TRY(ctx, cudaMemcpyAsync(out, dout.ptr, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    CUDA_TRY(ctx, cudaGetLastError());
    dout.Free(); dr.Free(); ds.Free();
    return PBRGPU_OK;
  }
  pbr::InitPathsFromRaysKernel<<<(std::max(n32, 64u) + 255) / 256, 256, 0, d.stream>>>(d.wave, reinterpret_cast<const float4*>(dr.ptr), ds.ptr, n32);
  pbr::ShadeFlags flags;
  flags.skip_emission_and_roulette = (mode == 1) ? 1u : 0u;
  // retired paths land in a one-entry-per-path accumulator (path i = "pixel" i)
  DevBuf<float4> acc;
  CUDA_TRY(ctx, acc.Alloc(n));
  CUDA_TRY(ctx, cudaMemsetAsync(acc.ptr, 0, sizeof(float4) * n, d.stream));
  std::vector<float> face_t;
  if (mode == 1) {
    // the hook reports face direction and t of the first hit: run the closest-hit stage alone first
    DevBuf<float> dface;
    CUDA_TRY(ctx, dface.Alloc(2 * n));
    CUDA_TRY(ctx, cudaMemsetAsync(d.wave.state[0] + size_t(pbr::kHit) * d.wave.capacity, 0xff, sizeof(float4) * n, d.stream));
    pbr::BeginIterationKernel<<<1, 32, 0, d.stream>>>(d.wave.counters, d.wave.stats, 0u, 0u, d.wave.capacity, 0ull);
    if (d.view.num_curves)
      pbr::TraceClosestKernel<true><<<PersistentGrid(d, ctx->tune_trace_blocks), kBlock, 0, d.stream>>>(d.view, d.wave, 0u, ctx->tune_refill, ctx->tune_prim_lanes, pbr::FrameParams(), acc.ptr, 0u);
    else
      pbr::TraceClosestKernel<false><<<PersistentGrid(d, ctx->tune_trace_blocks), kBlock, 0, d.stream>>>(d.view, d.wave, 0u, ctx->tune_refill, ctx->tune_prim_lanes, pbr::FrameParams(), acc.ptr, 0u);
    SurfaceFaceKernel<<<(n32 + 255) / 256, 256, 0, d.stream>>>(d.view, d.wave, n32, dface.ptr);
    // restore the entry state for the real iteration below
    pbr::InitPathsFromRaysKernel<<<(std::max(n32, 64u) + 255) / 256, 256, 0, d.stream>>>(d.wave, reinterpret_cast<const float4*>(dr.ptr), ds.ptr, n32);
    CUDA_TRY(ctx, cudaMemsetAsync(acc.ptr, 0, sizeof(float4) * n, d.stream));
    face_t.resize(2 * n);
    CUDA_TRY(ctx, cudaMemcpyAsync(face_t.data(), dface.ptr, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    dface.Free();
  }
  rc = RunPool(ctx, d, nullptr, acc.ptr, uint32_t(n), mode == 1 ? 1u : 0xffffffffu, flags, &tm, nullptr, nullptr, 0, 1, 0);
  if (rc != PBRGPU_OK) { dr.Free(); ds.Free(); acc.Free(); return rc; }
  if (mode == 0) {
    std::vector<float4> rad(n);
    CUDA_TRY(ctx, cudaMemcpyAsync(rad.data(), acc.ptr, sizeof(float4) * n, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    for (uint64_t i = 0; i < n; ++i) { out[3 * i] = rad[i].x; out[3 * i + 1] = rad[i].y; out[3 * i + 2] = rad[i].z; }
  } else {
    // every path sits in S / D of one of the two parities: scatter the records back to path order
    DevBuf<float> dout;
    CUDA_TRY(ctx, dout.Alloc(16 * n));
    CUDA_TRY(ctx, cudaMemsetAsync(dout.ptr, 0, sizeof(float) * 16 * n, d.stream));
    pbr::GatherVertexKernel<<<(n32 + 255) / 256, 256, 0, d.stream>>>(d.wave, 0u, n32, dout.ptr);
    pbr::GatherVertexKernel<<<(n32 + 255) / 256, 256, 0, d.stream>>>(d.wave, 1u, n32, dout.ptr);
    CUDA_TRY(ctx, cudaMemcpyAsync(out, dout.ptr, sizeof(float) * 16 * n, cudaMemcpyDeviceToHost, d.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
    CUDA_TRY(ctx, cudaGetLastError());
    for (uint64_t i = 0; i < n; ++i) {
      float* o = out + 16 * i;
      if (face_t[2 * i] < 0.f) { for (int k = 0; k < 16; ++k) o[k] = 0.f; continue; }
      o[0] = 1.f;
      o[14] = face_t[2 * i];
      o[15] = face_t[2 * i + 1];
    }
    dout.Free();
  }
  acc.Free();
  rc = FetchStats(ctx, d);
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->stats.paths = n;
  ctx->stats.closest_rays = d.h_stats[pbr::kStatClosest];
  ctx->stats.shadow_rays = d.h_stats[pbr::kStatShadow];
  ctx->stats.sss_rays = d.h_stats[pbr::kStatSss];
  ctx->stats.sss_skipped = d.h_stats[pbr::kStatSssSkipped];
  ctx->stats.kernel_launches = tm.launches;
  dr.Free(); ds.Free();
  return rc;
}

int pbrgpu_radiance(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* radiance_out) {
  return PathHook(ctx, rays, seeds, n, radiance_out, 0);
}
int pbrgpu_radiance_mega(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* radiance_out) {
  return PathHook(ctx, rays, seeds, n, radiance_out, 2);
}
int pbrgpu_shade(pbrgpu_ctx* ctx, const pbrgpu_ray* rays, const uint64_t* seeds, uint64_t n, float* out16) {
  return PathHook(ctx, rays, seeds, n, out16, 1);
}

int pbrgpu_eval_closure(pbrgpu_ctx* ctx, int op, const float* params, const float* in, uint32_t in_stride, uint64_t n,
                        float* out, uint32_t out_stride) {
  if (!ctx) return PBRGPU_ERR_INVALID;
  if (n == 0) return PBRGPU_OK;
  Device& d = ctx->devices[0];
  CUDA_TRY(ctx, cudaSetDevice(d.id));
  DevBuf<float> dp, din, dout;
  float dummy[32] = {0};
  CUDA_TRY(ctx, dp.Upload(params ? params : dummy, 32, d.stream));
  if (in_stride) CUDA_TRY(ctx, din.Upload(in, size_t(in_stride) * n, d.stream));
  CUDA_TRY(ctx, dout.Alloc(size_t(out_stride) * n));
  KatKernel<<<unsigned((n + 127) / 128), 128, 0, d.stream>>>(op, dp.ptr, din.ptr, in_stride, n, dout.ptr, out_stride);
  CUDA_TRY(ctx, cudaMemcpyAsync(out, dout.ptr, sizeof(float) * out_stride * n, cudaMemcpyDeviceToHost, d.stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(d.stream));
  CUDA_TRY(ctx, cudaGetLastError());
  dp.Free(); din.Free(); dout.Free();
  return PBRGPU_OK;
}

}  // extern "C"
