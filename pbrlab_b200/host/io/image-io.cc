#include "image-io.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>

namespace pbrlab {

float SrgbToLiner(const float c) {   // image-utils.cc:7-19
  if (c <= 0.04045f) return c / 12.92f;
  return std::pow((c + 0.055f) / (1.0f + 0.055f), 2.4f);
}
float LinerTosRGB(const float c) {   // image-utils.cc:24-36
  if (c <= 0.0031308f) return 12.92f * c;
  return std::pow((1.0f + 0.055f) * c, float(1.0 / 2.4)) - 0.055f;
}
void SrgbToLiner(const std::vector<float>& src, const size_t width, const size_t height, const size_t channels,
                 std::vector<float>* out) {
  std::vector<float> tmp(width * height * channels);
  for (size_t i = 0; i < width * height; ++i)
    for (size_t k = 0; k < channels; ++k)
      tmp[i * channels + k] = (k < 3) ? SrgbToLiner(src[i * channels + k]) : src[i * channels + k];
  out->swap(tmp);
}
void LinerToSrgb(const std::vector<float>& src, const size_t width, const size_t height, const size_t channels,
                 std::vector<float>* out) {
  std::vector<float> tmp(width * height * channels);
  for (size_t i = 0; i < width * height; ++i)
    for (size_t k = 0; k < channels; ++k)
      tmp[i * channels + k] = (k < 3) ? LinerTosRGB(src[i * channels + k]) : src[i * channels + k];
  out->swap(tmp);
}

namespace io {
namespace {

bool ReadAll(const std::string& path, std::vector<unsigned char>* out) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return false;
  fseek(fp, 0, SEEK_END);
  const long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (n <= 0) { fclose(fp); return false; }
  out->resize(size_t(n));
  const size_t got = fread(out->data(), 1, size_t(n), fp);
  fclose(fp);
  return got == size_t(n);
}

std::string JoinPath(const std::string& dir, const std::string& name) {
  if (dir.empty() || (!name.empty() && name[0] == '/')) return name;   // fs::path(a) / b with an absolute b is b
  return dir.back() == '/' ? dir + name : dir + "/" + name;
}

uint32_t Be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int Paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  if (pa <= pb && pa <= pc) return a;
  return pb <= pc ? b : c;
}

// -> 8-bit samples, `channels` per pixel
bool DecodePng(const std::vector<unsigned char>& f, std::vector<unsigned char>* out, size_t* w, size_t* h, size_t* ch) {
  static const unsigned char kSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (f.size() < 8 + 25 || memcmp(f.data(), kSig, 8) != 0) return false;
  size_t pos = 8;
  uint32_t width = 0, height = 0;
  int depth = 0, ctype = -1, interlace = 0;
  std::vector<unsigned char> idat, plte, trns;
  bool have_trns = false;
  while (pos + 12 <= f.size()) {
    const uint32_t len = Be32(&f[pos]);
    const unsigned char* type = &f[pos + 4];
    if (pos + 12 + size_t(len) > f.size()) return false;
    const unsigned char* data = &f[pos + 8];
    if (!memcmp(type, "IHDR", 4)) {
      if (len < 13) return false;
      width = Be32(data); height = Be32(data + 4);
      depth = data[8]; ctype = data[9]; interlace = data[12];
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(data, data + len);
    } else if (!memcmp(type, "tRNS", 4)) {
      trns.assign(data, data + len);
      have_trns = true;
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + size_t(len);
  }
  if (width == 0 || height == 0 || interlace != 0) return false;
  int src_ch;
  switch (ctype) {
    case 0: src_ch = 1; break;
    case 2: src_ch = 3; break;
    case 3: src_ch = 1; break;
    case 4: src_ch = 2; break;
    case 6: src_ch = 4; break;
    default: return false;
  }
  if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) return false;
  if (ctype == 3 && (depth == 16 || plte.empty())) return false;
  const size_t bpp_bits = size_t(src_ch) * depth;
  const size_t stride = (size_t(width) * bpp_bits + 7) / 8;
  const size_t bpp = std::max<size_t>(1, bpp_bits / 8);
  std::vector<unsigned char> raw((stride + 1) * height);
  uLongf raw_len = uLongf(raw.size());
  if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size()) return false;
  // un-filter in place
  std::vector<unsigned char> img(stride * height);
  for (size_t y = 0; y < height; ++y) {
    const unsigned char ft = raw[y * (stride + 1)];
    const unsigned char* in = &raw[y * (stride + 1) + 1];
    unsigned char* cur = &img[y * stride];
    const unsigned char* up = y ? &img[(y - 1) * stride] : nullptr;
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
      int v = in[x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += Paeth(a, b, c); break;
        default: return false;
      }
      cur[x] = static_cast<unsigned char>(v);
    }
  }
  // expand to 8-bit samples
  const bool key = have_trns && (ctype == 0 || ctype == 2);
  const bool pal_alpha = have_trns && ctype == 3;
  const int out_ch = (ctype == 3) ? (pal_alpha ? 4 : 3) : src_ch + (key ? 1 : 0);
  out->assign(size_t(width) * height * out_ch, 255);
  for (size_t y = 0; y < height; ++y) {
    const unsigned char* row = &img[y * stride];
    for (size_t x = 0; x < width; ++x) {
      unsigned s8[4] = {0, 0, 0, 0};
      uint32_t s16[4] = {0, 0, 0, 0};
      for (int c = 0; c < src_ch; ++c) {
        const size_t i = x * src_ch + c;
        if (depth == 8) { s8[c] = row[i]; s16[c] = row[i]; }
        else if (depth == 16) { s8[c] = row[2 * i]; s16[c] = (uint32_t(row[2 * i]) << 8) | row[2 * i + 1]; }
        else {
          const size_t bit = i * depth;
          const unsigned v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
          s16[c] = v;
          s8[c] = (ctype == 3) ? v : v * (255u / ((1u << depth) - 1u));   // greys are spread to 0..255
        }
      }
      unsigned char* o = &(*out)[(y * width + x) * out_ch];
      if (ctype == 3) {
        const unsigned idx = s8[0];
        for (int c = 0; c < 3; ++c) o[c] = (size_t(idx) * 3 + c < plte.size()) ? plte[idx * 3 + c] : 0;
        if (pal_alpha) o[3] = idx < trns.size() ? trns[idx] : 255;
      } else {
        for (int c = 0; c < src_ch; ++c) o[c] = static_cast<unsigned char>(s8[c]);
        if (key) {
          bool match = trns.size() >= size_t(2 * src_ch);
          for (int c = 0; match && c < src_ch; ++c) {
            const uint32_t k = (uint32_t(trns[2 * c]) << 8) | trns[2 * c + 1];
            match = (k == s16[c]);
          }
          o[src_ch] = match ? 0 : 255;
        }
      }
    }
  }
  *w = width; *h = height; *ch = size_t(out_ch);
  return true;
}

bool DecodePnm(const std::vector<unsigned char>& f, std::vector<unsigned char>* out, size_t* w, size_t* h, size_t* ch) {
  if (f.size() < 7 || f[0] != 'P' || (f[1] != '5' && f[1] != '6')) return false;
  size_t pos = 2;
  auto next_int = [&](long* v) {
    for (;;) {
      while (pos < f.size() && isspace(f[pos])) ++pos;
      if (pos < f.size() && f[pos] == '#') { while (pos < f.size() && f[pos] != '\n') ++pos; continue; }
      break;
    }
    if (pos >= f.size() || !isdigit(f[pos])) return false;
    long x = 0;
    while (pos < f.size() && isdigit(f[pos])) x = x * 10 + (f[pos++] - '0');
    *v = x;
    return true;
  };
  long W, H, M;
  if (!next_int(&W) || !next_int(&H) || !next_int(&M) || W <= 0 || H <= 0 || M <= 0 || M > 255) return false;
  ++pos;   // the single whitespace after maxval
  const size_t c = (f[1] == '6') ? 3 : 1, n = size_t(W) * H * c;
  if (pos + n > f.size()) return false;
  out->assign(f.begin() + pos, f.begin() + pos + n);
  *w = size_t(W); *h = size_t(H); *ch = c;
  return true;
}

}  // namespace

bool LoadImageFromFile(const std::string& filename, const std::string& asset_path, std::vector<float>* pixels,
                       size_t* width, size_t* height, size_t* channels) {
  if (!pixels || !width || !height || !channels) return false;
  const std::string path = JoinPath(asset_path, filename);
  std::vector<unsigned char> file, px;
  if (!ReadAll(path, &file)) {
    std::cerr << "warning : cannot read image [" << path << "]" << std::endl;
    return false;
  }
  size_t w = 0, h = 0, c = 0;
  if (!DecodePng(file, &px, &w, &h, &c) && !DecodePnm(file, &px, &w, &h, &c)) {
    std::cerr << "warning : image [" << path << "] is in a format this loader does not decode "
              << "(PNG non-interlaced, binary PGM/PPM)" << std::endl;
    return false;
  }
  pixels->resize(px.size());
  for (size_t i = 0; i < px.size(); ++i) (*pixels)[i] = float(px[i]) / float(255);   // image-io.cc:150-153
  *width = w; *height = h; *channels = c;
  return w != 0 && h != 0 && c != 0;
}

namespace {
void PutChunk(std::vector<unsigned char>* out, const char* type, const unsigned char* data, size_t len) {
  const uint32_t n = uint32_t(len);
  const unsigned char hdr[8] = {static_cast<unsigned char>(n >> 24), static_cast<unsigned char>(n >> 16),
                                static_cast<unsigned char>(n >> 8), static_cast<unsigned char>(n),
                                static_cast<unsigned char>(type[0]), static_cast<unsigned char>(type[1]),
                                static_cast<unsigned char>(type[2]), static_cast<unsigned char>(type[3])};
  out->insert(out->end(), hdr, hdr + 8);
  if (len) out->insert(out->end(), data, data + len);
  uLong crc = crc32(0L, hdr + 4, 4);
  if (len) crc = crc32(crc, data, uInt(len));
  const unsigned char tail[4] = {static_cast<unsigned char>(crc >> 24), static_cast<unsigned char>(crc >> 16),
                                 static_cast<unsigned char>(crc >> 8), static_cast<unsigned char>(crc)};
  out->insert(out->end(), tail, tail + 4);
}
}  // namespace

bool WritePNG8(const std::string& path, const unsigned char* pixels, size_t width, size_t height, size_t channels) {
  if (!pixels || width == 0 || height == 0 || channels < 1 || channels > 4) return false;
  static const int kType[5] = {0, 0, 4, 2, 6};
  std::vector<unsigned char> raw((width * channels + 1) * height);
  for (size_t y = 0; y < height; ++y) {
    raw[y * (width * channels + 1)] = 0;   // filter: none
    memcpy(&raw[y * (width * channels + 1) + 1], pixels + y * width * channels, width * channels);
  }
  uLongf clen = compressBound(uLong(raw.size()));
  std::vector<unsigned char> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), uLong(raw.size()), 6) != Z_OK) return false;
  std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  unsigned char ihdr[13];
  const uint32_t w = uint32_t(width), h = uint32_t(height);
  ihdr[0] = w >> 24; ihdr[1] = w >> 16; ihdr[2] = w >> 8; ihdr[3] = w;
  ihdr[4] = h >> 24; ihdr[5] = h >> 16; ihdr[6] = h >> 8; ihdr[7] = h;
  ihdr[8] = 8; ihdr[9] = static_cast<unsigned char>(kType[channels]); ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
  PutChunk(&out, "IHDR", ihdr, 13);
  PutChunk(&out, "IDAT", comp.data(), clen);
  PutChunk(&out, "IEND", nullptr, 0);
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return false;
  const bool ok = fwrite(out.data(), 1, out.size(), fp) == out.size();
  fclose(fp);
  return ok;
}

bool WritePNG(const std::string& filename, const std::string& asset_path, const std::vector<float>& pixels,
              const size_t width, const size_t height, const size_t channels) {
  const std::string path = JoinPath(asset_path, filename);
  const size_t dot = path.find_last_of('.');
  if (dot == std::string::npos || path.substr(dot) != ".png") {
    std::cerr << "warning! the file extension is not \"png\"" << std::endl;
    return false;
  }
  if (pixels.empty() || pixels.size() != width * height * channels) {
    std::cerr << "the image data is broken" << std::endl;
    return false;
  }
  std::vector<unsigned char> p8(pixels.size());
  for (size_t i = 0; i < pixels.size(); ++i)
    p8[i] = static_cast<unsigned char>(std::max(0.0f, std::min(255.0f, pixels[i] * 256.0f)));
  if (!WritePNG8(path, p8.data(), width, height, channels)) {
    std::cerr << "faild save image" << std::endl;
    return false;
  }
  return true;
}

}  // namespace io
}  // namespace pbrlab
