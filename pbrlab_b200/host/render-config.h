// RenderConfig kept as the reference declares it (reference src/render-config.h:9-18).  `thread` is as unused
// here as it is there (src/render.cc:203-204 never consults it): the GPU backend has no host worker pool.
#ifndef PBRLAB_B200_RENDER_CONFIG_H_
#define PBRLAB_B200_RENDER_CONFIG_H_
#include <cstdint>
#include <string>
#include <vector>
namespace pbrlab {
struct RenderConfig {
  std::vector<std::string> scene_filepaths;
  uint32_t width = 512;
  uint32_t height = 512;
  uint32_t max_pass = 32;
  int thread = -1;
};
}  // namespace pbrlab
#endif  // PBRLAB_B200_RENDER_CONFIG_H_
