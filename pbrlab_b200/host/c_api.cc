// C door into the C++ host API (pbrlab::Scene / Render / loaders) for the Python tests, bench.py and the graft
// entry points — ctypes cannot call C++.  Nothing here adds behaviour: every function forwards to the public
// classes the reference's callers would use.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pbrgpu.h"
#include "io/curve-mesh-io.h"
#include "io/image-io.h"
#include "io/triangle-mesh-io.h"
#include "pc-common.h"
#include "render.h"
#include "scene.h"

namespace {
thread_local std::string g_error;
// RenderLayers of pbrhost_render_layer, one per scene
std::mutex g_layers_mutex;
std::map<void*, std::unique_ptr<pbrlab::RenderLayer>> g_layers;
}

extern "C" {

void pbrhost_release_layer(void* s) {
  std::lock_guard<std::mutex> lock(g_layers_mutex);
  g_layers.erase(s);
}


struct pbrhost_flat {   // views into Scene::Flat(); valid until the scene is destroyed or re-committed
  const float* verts; uint32_t nverts;
  const float* normals; uint32_t nnormals;
  const float* texcoords; uint32_t ntexcoords;
  const uint32_t *vidx, *nidx, *tidx, *tri_material, *tri_instance, *tri_geom, *tri_prim;
  uint64_t ntris;
  const float* curve_verts; uint32_t ncurve_verts;
  const uint32_t *curve_first, *curve_material, *curve_instance, *curve_geom, *curve_prim;
  uint64_t nsegs;
  const float* materials; uint32_t nmaterials;
  pbrgpu_light_tables lights;
  float bmin[3], bmax[3];
  const float* tex_pixels; uint64_t ntex_floats;
  const uint32_t* tex_desc; uint32_t ntextures;   // 4 words per texture: offset (floats), width, height, channels
};

const char* pbrhost_last_error(void) { return g_error.c_str(); }

// CreateScene(argc, argv, &scene): argv[1..] = files.  commit_to_device = 0 stops after the host-side commit.
void* pbrhost_scene_create_on(int nfiles, const char** files, int commit_to_device, const int* device_ids,
                              int n_devices);
void* pbrhost_scene_create(int nfiles, const char** files, int commit_to_device) {
  return pbrhost_scene_create_on(nfiles, files, commit_to_device, nullptr, 0);
}
// same, on explicit CUDA devices (one id per rank in a multi-process launch; several ids = one process driving
// several GPUs, summed at frame end)
void* pbrhost_scene_create_on(int nfiles, const char** files, int commit_to_device, const int* device_ids,
                              int n_devices) {
  std::vector<std::string> store;
  store.emplace_back("pbrlab");
  for (int i = 0; i < nfiles; ++i) store.emplace_back(files[i]);
  std::vector<char*> argv;
  for (auto& s : store) argv.push_back(const_cast<char*>(s.c_str()));
  pbrlab::Scene* scene = new pbrlab::Scene();
  if (device_ids && n_devices > 0) scene->SetDevices(std::vector<int>(device_ids, device_ids + n_devices));
  try {
    if (!CreateScene(int(argv.size()), argv.data(), scene, commit_to_device != 0)) {
      g_error = "CreateScene failed";
      delete scene;
      return nullptr;
    }
  } catch (const std::exception& e) {
    g_error = e.what();
    delete scene;
    return nullptr;
  }
  return scene;
}
void pbrhost_scene_destroy(void* s) {
  pbrhost_release_layer(s);
  delete static_cast<pbrlab::Scene*>(s);
}

void pbrhost_scene_flat(void* s, pbrhost_flat* o) {
  const pbrlab::FlatScene& f = static_cast<pbrlab::Scene*>(s)->Flat();
  memset(o, 0, sizeof(*o));
  o->verts = f.verts.data(); o->nverts = uint32_t(f.verts.size() / 4);
  o->normals = f.normals.data(); o->nnormals = uint32_t(f.normals.size() / 4);
  o->texcoords = f.texcoords.data(); o->ntexcoords = uint32_t(f.texcoords.size() / 2);
  o->vidx = f.vidx.data(); o->nidx = f.nidx.data(); o->tidx = f.tidx.data();
  o->tri_material = f.tri_material.data(); o->tri_instance = f.tri_instance.data();
  o->tri_geom = f.tri_geom.data(); o->tri_prim = f.tri_prim.data();
  o->ntris = f.tri_prim.size();
  o->curve_verts = f.curve_verts.data(); o->ncurve_verts = uint32_t(f.curve_verts.size() / 4);
  o->curve_first = f.curve_first.data(); o->curve_material = f.curve_material.data();
  o->curve_instance = f.curve_instance.data(); o->curve_geom = f.curve_geom.data(); o->curve_prim = f.curve_prim.data();
  o->nsegs = f.curve_prim.size();
  o->materials = f.materials.data(); o->nmaterials = uint32_t(f.materials.size() / 28);
  o->lights.num_lights = uint32_t(f.lights.light_probability.size());
  o->lights.light_probability = f.lights.light_probability.data();
  o->lights.light_cdf = f.lights.light_cdf.data();
  o->lights.light_prim_offset = f.lights.light_prim_offset.data();
  o->lights.num_light_prims = uint32_t(f.lights.prim_probability.size());
  o->lights.prim_probability = f.lights.prim_probability.data();
  o->lights.prim_cdf = f.lights.prim_cdf.data();
  o->lights.prim_area_pdf = f.lights.prim_area_pdf.data();
  o->lights.prim_emission = f.lights.prim_emission.data();
  o->lights.prim_is_emissive = f.lights.prim_is_emissive.data();
  o->lights.prim_triangle = f.light_prim_triangle.data();
  for (int k = 0; k < 3; ++k) { o->bmin[k] = f.bmin[k]; o->bmax[k] = f.bmax[k]; }
  o->tex_pixels = f.tex_pixels.data(); o->ntex_floats = f.tex_pixels.size();
  o->tex_desc = f.tex_desc.data(); o->ntextures = uint32_t(f.tex_desc.size() / 4);
}

void* pbrhost_scene_ctx(void* s) { return static_cast<pbrlab::Scene*>(s)->DeviceContext(); }

// pbrlab::Render() into a RenderLayer that lives as long as the scene, the way the reference's GUI and CLI hold one
// across frames (pc/pbrlab-gui.cc:207-238): *rgba / *count point INTO the layer (valid until the next call for this
// scene or its destruction).  Returns seconds inside Render(), < 0 on failure.
double pbrhost_render_layer(void* s, uint32_t w, uint32_t h, uint32_t spp, uint64_t seed, float** rgba, uint32_t** count) {
  try {
    pbrlab::SetRenderSeed(seed);
    std::atomic_bool cancel(false);
    std::atomic_size_t finish_pass(0);
    pbrlab::RenderLayer* layer = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_layers_mutex);
      std::unique_ptr<pbrlab::RenderLayer>& slot = g_layers[s];
      if (!slot) slot.reset(new pbrlab::RenderLayer());
      layer = slot.get();
    }
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = pbrlab::Render(*static_cast<pbrlab::Scene*>(s), w, h, spp, cancel, layer, &finish_pass);
    const auto t1 = std::chrono::steady_clock::now();
    if (!ok) { g_error = pbrlab::LastRenderError(); return -1.0; }
    if (finish_pass.load() != spp) { g_error = "finish_pass != num_sample"; return -1.0; }
    *rgba = layer->rgba.data();
    *count = layer->count.data();
    return std::chrono::duration<double>(t1 - t0).count();
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1.0;
  }
}

// pbrlab::Render() through the public C++ entry point.  Returns seconds inside Render(), < 0 on failure.
double pbrhost_render(void* s, uint32_t w, uint32_t h, uint32_t spp, uint64_t seed, float* rgba, uint32_t* count) {
  try {
    pbrlab::SetRenderSeed(seed);
    std::atomic_bool cancel(false);
    std::atomic_size_t finish_pass(0);
    pbrlab::RenderLayer layer;
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = pbrlab::Render(*static_cast<pbrlab::Scene*>(s), w, h, spp, cancel, &layer, &finish_pass);
    const auto t1 = std::chrono::steady_clock::now();
    if (!ok) { g_error = pbrlab::LastRenderError(); return -1.0; }
    if (finish_pass.load() != spp) { g_error = "finish_pass != num_sample"; return -1.0; }
    if (rgba) memcpy(rgba, layer.rgba.data(), sizeof(float) * layer.rgba.size());
    if (count) memcpy(count, layer.count.data(), sizeof(uint32_t) * layer.count.size());
    return std::chrono::duration<double>(t1 - t0).count();
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1.0;
  }
}

// Render() with the cancel flag raised from a second thread, as the GUI does (pc/glfw-window.cc:621-625): the flag goes
// up once *finish_pass has reached `cancel_at_pass` (0: before the call).  A third thread watches finish_pass the way
// the GUI's progress bar does.  out[0] = Render()'s return value, out[1] = final finish_pass, out[2] = 1 if
// finish_pass was ever seen to decrease or to exceed num_sample, out[3] = passes finished when the flag was raised.
// Returns seconds inside Render(), < 0 on an exception (which Render() must not throw).
double pbrhost_render_cancel(void* s, uint32_t w, uint32_t h, uint32_t spp, uint64_t seed, uint32_t cancel_at_pass,
                             float* rgba, uint32_t* count, uint64_t* out) {
  try {
    pbrlab::SetRenderSeed(seed);
    std::atomic_bool cancel(cancel_at_pass == 0);
    std::atomic_size_t finish_pass(0);
    std::atomic_bool done(false), bad(false);
    std::atomic_size_t raised_at(0);
    pbrlab::RenderLayer layer;
    std::thread gui([&]() {
      size_t last = 0;
      bool started = false;
      while (!done.load()) {
        const size_t p = finish_pass.load();
        // Render() resets *finish_pass to 0 once, at its start (render.cc:213); after that it may only grow
        if (p > spp || (started && p < last)) bad = true;
        if (p > 0) started = true;
        last = p;
        if (!cancel.load() && p >= cancel_at_pass) { raised_at = p; cancel = true; }
        std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    });
    const auto t0 = std::chrono::steady_clock::now();
    const bool ok = pbrlab::Render(*static_cast<pbrlab::Scene*>(s), w, h, spp, cancel, &layer, &finish_pass);
    const auto t1 = std::chrono::steady_clock::now();
    done = true;
    gui.join();
    out[0] = ok ? 1 : 0;
    out[1] = finish_pass.load();
    out[2] = bad.load() ? 1 : 0;
    out[3] = raised_at.load();
    if (rgba) memcpy(rgba, layer.rgba.data(), sizeof(float) * layer.rgba.size());
    if (count) memcpy(count, layer.count.data(), sizeof(uint32_t) * layer.count.size());
    return std::chrono::duration<double>(t1 - t0).count();
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1.0;
  }
}

// the CLI's output stage on the device: rgba8 [w*h*4] of the frame the last pbrhost_render left on the GPU
int pbrhost_resolve_srgb8(void* s, uint32_t w, uint32_t h, uint8_t* rgba8) {
  pbrgpu_ctx* ctx = static_cast<pbrlab::Scene*>(s)->DeviceContext();
  if (!ctx) { g_error = "scene is not committed to a device"; return 1; }
  const int rc = pbrgpu_resolve_srgb8(ctx, w, h, rgba8);
  if (rc != PBRGPU_OK) g_error = pbrgpu_last_error(ctx);
  return rc;
}

// io::WritePNG8 (the encoder behind the CLI's rgba.png): pixels are 8-bit, row-major, `channels` interleaved
int pbrhost_write_png8(const char* path, const uint8_t* pixels, uint32_t w, uint32_t h, uint32_t channels) {
  return pbrlab::io::WritePNG8(path, pixels, w, h, channels) ? 1 : 0;
}

// Scene::TraceFirstHit1 / AnyHit1 through the C++ API (single-ray entry points of the reference)
int pbrhost_trace1(void* s, const float* ray8, float* out6, uint32_t* ids3) {
  try {
    pbrlab::Ray r;
    r.ray_org = pbrlab::float3(ray8[0], ray8[1], ray8[2]); r.min_t = ray8[3];
    r.ray_dir = pbrlab::float3(ray8[4], ray8[5], ray8[6]); r.max_t = ray8[7];
    const pbrlab::TraceResult tr = static_cast<pbrlab::Scene*>(s)->TraceFirstHit1(r);
    out6[0] = tr.t; out6[1] = tr.u; out6[2] = tr.v;
    out6[3] = tr.normal_g[0]; out6[4] = tr.normal_g[1]; out6[5] = tr.normal_g[2];
    ids3[0] = tr.instance_id; ids3[1] = tr.geom_id; ids3[2] = tr.prim_id;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}
int pbrhost_anyhit1(void* s, const float* ray8) {
  try {
    pbrlab::Ray r;
    r.ray_org = pbrlab::float3(ray8[0], ray8[1], ray8[2]); r.min_t = ray8[3];
    r.ray_dir = pbrlab::float3(ray8[4], ray8[5], ray8[6]); r.max_t = ray8[7];
    return static_cast<pbrlab::Scene*>(s)->AnyHit1(r) ? 1 : 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
}

// Text formatting for the synthetic OBJ generators (pbrlab_b200/scenes.py): appends n rows to `path`.
//   kind 0: "v %.6f %.6f %.6f"   kind 1: "vn %.5f %.5f %.5f"   (d: n x 3 doubles)
//   kind 2: "f %d//%d %d//%d %d//%d"                           (idx: n x 6 int64)
// Same text as numpy.savetxt with those formats (both round correctly), produced by all host threads: the 20 M-triangle
// C5 file is 1.5 GB of text and took minutes through savetxt.
int pbrhost_append_rows(const char* path, int kind, const double* d, const int64_t* idx, uint64_t n) {
  const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<std::string> parts(nt);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    th.emplace_back([&, t]() {
      const uint64_t b = n * t / nt, e = n * (t + 1) / nt;
      std::string& out = parts[t];
      out.reserve(size_t(e - b) * (kind == 2 ? 56 : 36));
      char buf[160];
      for (uint64_t i = b; i < e; ++i) {
        int len;
        if (kind == 0) len = snprintf(buf, sizeof(buf), "v %.6f %.6f %.6f\n", d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        else if (kind == 1) len = snprintf(buf, sizeof(buf), "vn %.5f %.5f %.5f\n", d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        else len = snprintf(buf, sizeof(buf), "f %lld//%lld %lld//%lld %lld//%lld\n", (long long)idx[6 * i],
                            (long long)idx[6 * i + 1], (long long)idx[6 * i + 2], (long long)idx[6 * i + 3],
                            (long long)idx[6 * i + 4], (long long)idx[6 * i + 5]);
        out.append(buf, size_t(len));
      }
    });
  }
  for (auto& t : th) t.join();
  FILE* f = fopen(path, "ab");
  if (!f) { g_error = std::string("cannot append to ") + path; return 0; }
  bool ok = true;
  for (const auto& p : parts) ok = ok && fwrite(p.data(), 1, p.size(), f) == p.size();
  ok = (fclose(f) == 0) && ok;
  return ok ? 1 : 0;
}

// loader-level access (CyHair -> Bezier), two-call pattern: vt == nullptr returns the sizes
int pbrhost_hair_load(const char* path, float* vt, uint64_t* nfloats, uint32_t* idx, uint64_t* nidx) {
  std::vector<float> v;
  std::vector<uint32_t> ind;
  const bool ok = pbrlab::io::LoadCurveMeshAsCubicBezierCurve(path, false, &v, &ind);
  *nfloats = v.size();
  *nidx = ind.size();
  if (vt) memcpy(vt, v.data(), sizeof(float) * v.size());
  if (idx) memcpy(idx, ind.data(), sizeof(uint32_t) * ind.size());
  return ok ? 1 : 0;
}

}  // extern "C"
