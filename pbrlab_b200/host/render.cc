#include "render.h"

#include <chrono>
#include <stdexcept>
#include <string>
#include <thread>

#include "../../include/pbrgpu.h"

namespace pbrlab {

static std::atomic<uint64_t> g_render_seed(1234567890ull);   // the reference's constant (src/render.cc:215)
void SetRenderSeed(uint64_t seed) { g_render_seed = seed; }
uint64_t GetRenderSeed(void) { return g_render_seed.load(); }

bool Render(const Scene& scene, const uint32_t width, const uint32_t height, const uint32_t num_sample,
            const std::atomic_bool& cancel_render_flag, RenderLayer* layer, std::atomic_size_t* finish_pass) {
  pbrgpu_ctx* ctx = scene.DeviceContext();
  if (!ctx) throw std::runtime_error("pbrlab::Render: scene was not committed to a device (CommitScene)");
  layer->Resize(width, height);   // PrepareRendering (src/render.cc:99-100)
  layer->Clear();
  *finish_pass = 0;
  scene.SyncMaterialsToDevice();  // materials are live-editable between calls (pc/pbrlab-gui.cc:207-238)

  // The C ABI polls a plain int and reports progress through a plain size_t; a watcher mirrors the caller's atomics.
  volatile int cancel_int = cancel_render_flag.load() ? 1 : 0;
  size_t progress = 0;
  std::atomic_bool done(false);
  std::thread watcher([&]() {
    while (!done.load()) {
      if (cancel_render_flag.load()) cancel_int = 1;
      const size_t p = *const_cast<volatile size_t*>(&progress);
      if (p > finish_pass->load()) finish_pass->store(p);
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
  });
  const int rc = pbrgpu_render(ctx, width, height, num_sample, g_render_seed.load(), 0, 1, &cancel_int,
                               layer->rgba.data(), layer->count.data(), &progress);
  done = true;
  watcher.join();
  if (progress > finish_pass->load()) finish_pass->store(progress);
  if (rc != PBRGPU_OK) throw std::runtime_error(std::string("pbrlab::Render: ") + pbrgpu_last_error(ctx));
  return true;
}

}  // namespace pbrlab
