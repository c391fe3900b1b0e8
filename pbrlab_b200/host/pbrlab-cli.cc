// Command-line front-end: the reference's pbrlab-cli (reference pc/pbrlab-cli.cc:16-60) with the three values it
// hard-codes (512 x 512, 32 spp) exposed as flags, plus a raw dump for the parity tests.
//   pbrlab-cli [--width W] [--height H] [--spp N] [--seed S] [--gpus N] [--raw out.bin] [--png out.png] [--ppm out.ppm] files...
// Writes ./rgba.png like the reference: colour = rgba / count -> LinerToSrgb -> 8 bit (pc/pbrlab-cli.cc:47-57), the
// output stage running on the device (pbrgpu_resolve_srgb8) so only 4 bytes per pixel come back for the file.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/pbrgpu.h"
#include "pc-common.h"
#include "render.h"

#include "io/image-io.h"

int main(int argc, char** argv) {
  uint32_t width = 512, height = 512, spp = 32;
  uint64_t seed = 1234567890ull;
  int gpus = 1;
  std::string raw_path, ppm_path, png_path = "rgba.png";
  std::vector<char*> files;
  files.push_back(argv[0]);
  for (int i = 1; i < argc; ++i) {
    const std::string a(argv[i]);
    auto next = [&](const char* what) -> const char* {
      if (i + 1 >= argc) { std::cerr << "missing value for " << what << std::endl; exit(EXIT_FAILURE); }
      return argv[++i];
    };
    if (a == "--width") width = uint32_t(atoi(next("--width")));
    else if (a == "--height") height = uint32_t(atoi(next("--height")));
    else if (a == "--spp") spp = uint32_t(atoi(next("--spp")));
    else if (a == "--seed") seed = strtoull(next("--seed"), nullptr, 10);
    else if (a == "--gpus") gpus = atoi(next("--gpus"));
    else if (a == "--raw") raw_path = next("--raw");
    else if (a == "--ppm") ppm_path = next("--ppm");
    else if (a == "--png") png_path = next("--png");
    else files.push_back(argv[i]);
  }
  if (files.size() < 2) {
    std::cerr << "not specified obj filename" << std::endl;
    return EXIT_FAILURE;
  }

  pbrlab::Scene scene;
  if (gpus > 1) {
    std::vector<int> ids;
    for (int g = 0; g < gpus; ++g) ids.push_back(g);
    scene.SetDevices(ids);
  }
  const auto t0 = std::chrono::steady_clock::now();
  try {
    if (!CreateScene(int(files.size()), files.data(), &scene)) return EXIT_FAILURE;
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  const auto t1 = std::chrono::steady_clock::now();

  std::atomic_bool cancel_render_flag(false);
  std::atomic_size_t finish_pass(0);
  pbrlab::RenderLayer layer;
  pbrlab::SetRenderSeed(seed);
  try {
    pbrlab::Render(scene, width, height, spp, cancel_render_flag, &layer, &finish_pass);
  } catch (const std::exception& e) {
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  }
  const auto t2 = std::chrono::steady_clock::now();
  const double load_s = std::chrono::duration<double>(t1 - t0).count();
  const double render_s = std::chrono::duration<double>(t2 - t1).count();
  pbrgpu_stats st;
  pbrgpu_get_stats(scene.DeviceContext(), &st);
  const double rays = double(st.closest_rays + st.shadow_rays + st.sss_rays);
  printf("load %.3f s, render %.3f s, %.2f Msamples/s, %.2f Mrays/s (%.3f rays/sample), %llu kernel launches\n",
         load_s, render_s, double(st.paths) / render_s * 1e-6, rays / render_s * 1e-6,
         st.paths ? rays / double(st.paths) : 0.0, static_cast<unsigned long long>(st.kernel_launches));

  if (!raw_path.empty()) {
    FILE* fp = fopen(raw_path.c_str(), "wb");
    if (fp) {
      fwrite(&width, 4, 1, fp);
      fwrite(&height, 4, 1, fp);
      fwrite(layer.rgba.data(), sizeof(float), layer.rgba.size(), fp);
      fwrite(layer.count.data(), sizeof(uint32_t), layer.count.size(), fp);
      fclose(fp);
    }
  }
  // color = rgba / count -> sRGB -> 8 bit (reference pc/pbrlab-cli.cc:47-57, src/io/image-io.cc WritePNG), on the device
  std::vector<unsigned char> rgba8(size_t(width) * height * 4);
  if (pbrgpu_resolve_srgb8(scene.DeviceContext(), width, height, rgba8.data()) != PBRGPU_OK) {
    std::cerr << pbrgpu_last_error(scene.DeviceContext()) << std::endl;
    return EXIT_FAILURE;
  }
  if (!png_path.empty() && !pbrlab::io::WritePNG8(png_path, rgba8.data(), width, height, 4)) {
    std::cerr << "faild save image" << std::endl;
    return EXIT_FAILURE;
  }
  if (!ppm_path.empty()) {
    FILE* fp = fopen(ppm_path.c_str(), "wb");
    if (!fp) return EXIT_FAILURE;
    fprintf(fp, "P6\n%u %u\n255\n", width, height);
    std::vector<unsigned char> row(size_t(width) * 3);
    for (uint32_t y = 0; y < height; ++y) {
      for (uint32_t x = 0; x < width; ++x)
        for (int c = 0; c < 3; ++c) row[x * 3 + c] = rgba8[(size_t(y) * width + x) * 4 + c];
      fwrite(row.data(), 1, row.size(), fp);
    }
    fclose(fp);
  }
  return EXIT_SUCCESS;
}
