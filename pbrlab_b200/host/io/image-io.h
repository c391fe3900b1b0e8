// Image files for the texture path and the CLI output, with the reference's entry points
// (reference src/io/image-io.h:22-64, src/io/image-io.cc:96-215; the reference drives stb_image / stb_image_write /
// tinyexr).  Our own decoders, behaviour kept where a texture's VALUES depend on it:
//   * 8-bit files become floats as value / 255 (image-io.cc:150-153); channel count is the file's own (req_comp = 0);
//   * PNG: non-interlaced, colour types 0/2/3/4/6, bit depths 1..16 (16-bit samples keep their high byte, low-bit
//     greys are spread to 0..255, a tRNS chunk adds an alpha channel) — what stb_image does for the same files;
//   * binary PGM / PPM (P5 / P6, maxval <= 255);
//   * .hdr / .exr / JPEG / interlaced PNG: not decoded — LoadImageFromFile returns false, which the OBJ loader treats
//     like any unreadable texture file (texture id stays -1, reference src/io/triangle-mesh-io.cc:80-92).
//   * WritePNG quantises floats as (unsigned char)Clamp(v * 256, 0, 255) (image-io.cc:200-206).
#ifndef PBRLAB_B200_IMAGE_IO_H_
#define PBRLAB_B200_IMAGE_IO_H_
#include <cstddef>
#include <string>
#include <vector>

namespace pbrlab {
namespace io {
bool LoadImageFromFile(const std::string& filename, const std::string& asset_path, std::vector<float>* pixels,
                       size_t* width, size_t* height, size_t* channels);
bool WritePNG(const std::string& filename, const std::string& asset_path, const std::vector<float>& pixels,
              const size_t width, const size_t height, const size_t channels);
bool WritePNG8(const std::string& path, const unsigned char* pixels, size_t width, size_t height, size_t channels);
}  // namespace io

// reference src/image-utils.cc:7-38
float SrgbToLiner(const float c_srgb);
float LinerTosRGB(const float c_liner);
void SrgbToLiner(const std::vector<float>& src, const size_t width, const size_t height, const size_t channels,
                 std::vector<float>* out);
void LinerToSrgb(const std::vector<float>& src, const size_t width, const size_t height, const size_t channels,
                 std::vector<float>* out);
}  // namespace pbrlab
#endif  // PBRLAB_B200_IMAGE_IO_H_
