#include "render.h"

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <exception>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/pbrgpu.h"

namespace pbrlab {

static std::atomic<uint64_t> g_render_seed(1234567890ull);   // the reference's constant (src/render.cc:215)
void SetRenderSeed(uint64_t seed) { g_render_seed = seed; }
uint64_t GetRenderSeed(void) { return g_render_seed.load(); }

static std::mutex g_error_mutex;
static std::string g_last_error;
static void SetLastError(const std::string& msg) {
  std::lock_guard<std::mutex> lock(g_error_mutex);
  g_last_error = msg;
}
std::string LastRenderError(void) {
  std::lock_guard<std::mutex> lock(g_error_mutex);
  return g_last_error;
}

bool Render(const Scene& scene, const uint32_t width, const uint32_t height, const uint32_t num_sample,
            const std::atomic_bool& cancel_render_flag, RenderLayer* layer, std::atomic_size_t* finish_pass) {
  // The reference's Render() never throws; a GUI render thread calls it in a loop.  Device failures
  // are reported as `false` + LastRenderError() and leave a cleared layer.
  layer->Resize(width, height);   // PrepareRendering (src/render.cc:99-100)
  layer->Clear();
  *finish_pass = 0;
  pbrgpu_ctx* ctx = scene.DeviceContext();
  if (!ctx) {
    SetLastError("pbrlab::Render: scene was not committed to a device (CommitScene)");
    fprintf(stderr, "%s\n", LastRenderError().c_str());
    return false;
  }
  try {
    scene.SyncMaterialsToDevice();  // materials are live-editable between calls (pc/pbrlab-gui.cc:207-238)
  } catch (const std::exception& e) {   // e.g. an edit that makes a texture id invalid: report, do not unwind
    SetLastError(std::string("pbrlab::Render: ") + e.what());
    fprintf(stderr, "%s\n", LastRenderError().c_str());
    return false;
  }

  // The C ABI polls a plain int and reports progress through a plain size_t; a watcher mirrors the caller's atomics.
  volatile int cancel_int = cancel_render_flag.load() ? 1 : 0;
  size_t progress = 0;
  bool done = false;
  std::mutex done_mutex;
  std::condition_variable done_cv;
  std::thread watcher([&]() {
    std::unique_lock<std::mutex> lock(done_mutex);
    while (!done) {
      if (cancel_render_flag.load()) cancel_int = 1;
      const size_t p = *const_cast<volatile size_t*>(&progress);
      if (p > finish_pass->load()) finish_pass->store(p);
      done_cv.wait_for(lock, std::chrono::milliseconds(1));   // (woken at once when the frame is done)
    }
  });
  const int rc = pbrgpu_render(ctx, width, height, num_sample, g_render_seed.load(), 0, 1, &cancel_int,
                               layer->rgba.data(), layer->count.data(), &progress);
  {
    std::lock_guard<std::mutex> lock(done_mutex);
    done = true;
  }
  done_cv.notify_one();
  watcher.join();
  if (progress > finish_pass->load()) finish_pass->store(progress);
  if (rc != PBRGPU_OK) {
    SetLastError(std::string("pbrlab::Render: ") + pbrgpu_last_error(ctx));
    fprintf(stderr, "%s\n", LastRenderError().c_str());
    layer->Clear();
    return false;
  }
  return true;
}

}  // namespace pbrlab
