#include "triangle-mesh-io.h"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <functional>
#include <thread>

#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <thread>

#include <sys/stat.h>

#include "image-io.h"

namespace pbrlab {
namespace io {
namespace {

bool ReadFile(const std::string& path, std::string* out) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return false;
  fseek(fp, 0, SEEK_END);
  const long n = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  out->resize(size_t(n));
  const size_t got = n > 0 ? fread(&(*out)[0], 1, size_t(n), fp) : 0;
  fclose(fp);
  out->push_back('\n');
  return got == size_t(n);
}

inline bool IsSpace(char c) { return c == ' ' || c == '\t'; }
inline bool IsEol(char c) { return c == '\n' || c == '\r' || c == '\0'; }
inline void SkipSpace(const char** p) { while (IsSpace(**p)) ++*p; }

// decimal -> float through double, like the reference's loader (real_t = float, parsed as double)
inline float ParseFloat(const char** p) {
  SkipSpace(p);
  char* end = nullptr;
  const double d = strtod(*p, &end);
  if (end == *p) {   // not a number: skip the token
    while (!IsSpace(**p) && !IsEol(**p)) ++*p;
    return 0.f;
  }
  *p = end;
  return static_cast<float>(d);
}

inline bool ParseInt(const char** p, int* v) {
  char* end = nullptr;
  const long x = strtol(*p, &end, 10);
  if (end == *p) return false;
  *p = end;
  *v = int(x);
  return true;
}

// OBJ index -> zero based (negative = relative to the elements read so far); 0 is invalid
inline bool FixIndex(int idx, int n, int* out) {
  if (idx > 0) { *out = idx - 1; return true; }
  if (idx == 0) return false;
  *out = n + idx;
  return true;
}

struct Corner { int v, vt, vn; };

std::string RestOfLine(const char* p) {
  const char* e = p;
  while (!IsEol(*e)) ++e;
  while (e > p && IsSpace(e[-1])) --e;
  return std::string(p, size_t(e - p));
}

struct RawMaterial {
  std::string name;
  std::map<std::string, std::string> keys;   // first occurrence wins
};

void LoadMtl(const std::string& path, std::vector<RawMaterial>* out, std::map<std::string, int>* index) {
  std::string text;
  if (!ReadFile(path, &text)) {
    std::cerr << "warning : material file [" << path << "] not found" << std::endl;
    return;
  }
  const char* p = text.c_str();
  RawMaterial cur;
  bool have = false;
  while (*p) {
    SkipSpace(&p);
    const char* line = p;
    while (!IsEol(*p)) ++p;
    const char* eol = p;
    while (*p == '\n' || *p == '\r') ++p;
    if (line == eol || *line == '#') continue;
    const char* sp = line;
    while (sp < eol && !IsSpace(*sp)) ++sp;
    const std::string key(line, size_t(sp - line));
    const char* val = sp;
    while (val < eol && IsSpace(*val)) ++val;
    const char* vend = eol;
    while (vend > val && IsSpace(vend[-1])) --vend;
    const std::string value(val, size_t(vend - val));
    if (key == "newmtl") {
      if (have) {
        index->insert(std::make_pair(cur.name, int(out->size())));
        out->push_back(cur);
      }
      cur = RawMaterial();
      cur.name = value;
      have = true;
      continue;
    }
    if (!have) continue;
    cur.keys.insert(std::make_pair(key, value));
  }
  if (have) {
    index->insert(std::make_pair(cur.name, int(out->size())));
    out->push_back(cur);
  }
}

bool GetFloat(const RawMaterial& m, const char* key, float* out) {
  auto it = m.keys.find(key);
  if (it == m.keys.end()) return false;
  *out = float(atof(it->second.c_str()));   // reference: std::atof (triangle-mesh-io.cc:35-38)
  return true;
}
bool GetFloat3(const RawMaterial& m, const char* key, float3* out) {
  auto it = m.keys.find(key);
  if (it == m.keys.end()) return false;
  double x = 0.0, y = 0.0, z = 0.0;
  sscanf(it->second.c_str(), "%lf %lf %lf", &x, &y, &z);   // reference :40-53
  *out = float3(float(x), float(y), float(z));
  return true;
}

// tinyobj::ParseTextureNameAndOption (reference src/io/tiny_obj_loader.h:1243-1318): options first, the file name
// is everything from the first non-option token to the end of the line (it may contain spaces).  Only -colorspace
// matters to the caller; the other options are skipped with their arguments.
void ParseTextureNameAndOption(const std::string& value, std::string* texname, std::string* colorspace) {
  texname->clear();
  colorspace->clear();
  const char* t = value.c_str();
  auto skip_space = [&]() { while (*t == ' ' || *t == '\t') ++t; };
  auto skip_token = [&]() { skip_space(); while (*t && *t != ' ' && *t != '\t' && *t != '\r') ++t; };
  auto opt = [&](const char* name) {
    const size_t n = strlen(name);
    if (strncmp(t, name, n) == 0 && (t[n] == ' ' || t[n] == '\t')) { t += n; return true; }
    return false;
  };
  auto skip_reals = [&](int max_count) {   // parseReal / parseReal2 / parseReal3: numbers until something else
    for (int k = 0; k < max_count; ++k) {
      skip_space();
      char* end = nullptr;
      strtod(t, &end);
      if (end == t) break;
      t = end;
    }
  };
  while (*t && *t != '\r' && *t != '\n') {
    skip_space();
    if (opt("-blendu") || opt("-blendv") || opt("-clamp") || opt("-type") || opt("-texres") || opt("-imfchan")) skip_token();
    else if (opt("-boost") || opt("-bm")) skip_reals(1);
    else if (opt("-mm")) skip_reals(2);
    else if (opt("-o") || opt("-s") || opt("-t")) skip_reals(3);
    else if (opt("-colorspace")) {
      skip_space();
      const char* b = t;
      skip_token();
      colorspace->assign(b, t);
    } else {
      *texname = std::string(t);
      while (!texname->empty() && (texname->back() == '\r' || texname->back() == '\n')) texname->pop_back();
      break;
    }
  }
}

bool IsHdr(const std::string& path) {   // reference triangle-mesh-io.cc:106-114
  const size_t dot = path.find_last_of('.');
  if (dot == std::string::npos) return false;
  std::string e = path.substr(dot);
  std::transform(e.begin(), e.end(), e.begin(), [](unsigned char c) { return char(tolower(c)); });
  return e == ".exr" || e == ".hdr";
}

// LoadTextureFromTinyObjMaterial + LoadTexture (reference triangle-mesh-io.cc:80-141): sRGB files are converted to
// linear at load time unless the file is HDR or `-colorspace` names something other than sRGB; a file that cannot
// be read leaves the id at -1.
void LoadMaterialTexture(const RawMaterial& m, const char* key, const std::string& base_dir, uint32_t* tex_id,
                         std::vector<Texture>* textures) {
  auto it = m.keys.find(key);
  if (it == m.keys.end() || !textures) return;
  std::string texname, colorspace;
  ParseTextureNameAndOption(it->second, &texname, &colorspace);
  const bool degamma = (colorspace.empty() || colorspace == "sRGB") && !IsHdr(texname);
  std::vector<float> pixels;
  size_t w = 0, h = 0, c = 0;
  if (!LoadImageFromFile(texname, base_dir, &pixels, &w, &h, &c)) { *tex_id = uint32_t(-1); return; }
  if (degamma) SrgbToLiner(pixels, w, h, c, &pixels);
  *tex_id = uint32_t(textures->size());
  textures->emplace_back(pixels, uint32_t(w), uint32_t(h), uint32_t(c), texname);
  std::cout << "Loaded texture for " << key << " : " << texname << std::endl;
}

MaterialParameter ToPrincipled(const RawMaterial& m, const std::string& base_dir, std::vector<Texture>* textures) {
  CyclesPrincipledBsdfParameter p;
  GetFloat3(m, "base_color", &p.base_color);
  LoadMaterialTexture(m, "map_base_color", base_dir, &p.base_color_tex_id, textures);
  GetFloat(m, "subsurface", &p.subsurface);
  GetFloat3(m, "subsurface_radius", &p.subsurface_radius);
  GetFloat3(m, "subsurface_color", &p.subsurface_color);
  LoadMaterialTexture(m, "map_subsurface_color", base_dir, &p.subsurface_color_tex_id, textures);
  GetFloat(m, "metallic", &p.metallic);
  GetFloat(m, "specular", &p.specular);
  GetFloat(m, "specular_tint", &p.specular_tint);
  GetFloat(m, "roughness", &p.roughness);
  GetFloat(m, "anisotropic", &p.anisotropic);
  GetFloat(m, "anisotropic_rotation", &p.anisotropic_rotation);
  GetFloat(m, "sheen", &p.sheen);
  GetFloat(m, "sheen_tint", &p.sheen_tint);
  GetFloat(m, "clearcoat", &p.clearcoat);
  GetFloat(m, "clearcoat_roughness", &p.clearcoat_roughness);
  GetFloat(m, "ior", &p.ior);
  GetFloat(m, "transmission", &p.transmission);
  GetFloat(m, "transmission_roughness", &p.transmission_roughness);
  p.name = m.name;
  return MaterialParameter(p);
}

struct ShapeBuild {
  std::string name;
  std::vector<uint32_t> v, vn, vt, mat;
};

// ---- binary cache of the parsed geometry (at 20 M triangles the 1.5 GB of OBJ text takes 13 s to
// parse).  Opt-in (PBRLAB_SCENE_CACHE=1): `<file>.pbrcache` next to the OBJ holds the attribute pools, the per-shape
// index arrays and the names of the material libraries; it is used only while the OBJ and every MTL it names still
// have the size and modification time recorded in it.  Materials and textures are always read from their files.
struct FileStamp { uint64_t size = 0; int64_t mtime = 0; bool ok = false; };
FileStamp StampOf(const std::string& path) {
  FileStamp st;
  struct stat sb;
  if (stat(path.c_str(), &sb) == 0) { st.size = uint64_t(sb.st_size); st.mtime = int64_t(sb.st_mtime); st.ok = true; }
  return st;
}
const char kCacheMagic[8] = {'P', 'B', 'R', 'C', 'A', 'C', 'H', '2'};
template <class T> bool PutVec(FILE* fp, const std::vector<T>& v) {
  const uint64_t n = v.size();
  return fwrite(&n, 8, 1, fp) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, fp) == n);
}
template <class T> bool GetVec(FILE* fp, std::vector<T>* v) {
  uint64_t n = 0;
  if (fread(&n, 8, 1, fp) != 1 || n > (uint64_t(1) << 36) / sizeof(T)) return false;
  v->resize(size_t(n));
  return n == 0 || fread(v->data(), sizeof(T), n, fp) == n;
}
bool PutStr(FILE* fp, const std::string& s) { return PutVec(fp, std::vector<char>(s.begin(), s.end())); }
bool GetStr(FILE* fp, std::string* s) {
  std::vector<char> v;
  if (!GetVec(fp, &v)) return false;
  s->assign(v.begin(), v.end());
  return true;
}
bool PutStamp(FILE* fp, const FileStamp& st) { return fwrite(&st.size, 8, 1, fp) == 1 && fwrite(&st.mtime, 8, 1, fp) == 1; }
bool SameStamp(FILE* fp, const FileStamp& now) {
  uint64_t size = 0; int64_t mtime = 0;
  return fread(&size, 8, 1, fp) == 1 && fread(&mtime, 8, 1, fp) == 1 && now.ok && size == now.size && mtime == now.mtime;
}

bool WriteObjCache(const std::string& obj, const std::string& base_dir, const std::vector<std::string>& mtllibs,
                   const Attribute& attr, const std::vector<ShapeBuild>& shapes) {
  const std::string tmp = obj + ".pbrcache.tmp";
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) return false;
  bool ok = fwrite(kCacheMagic, 8, 1, fp) == 1 && PutStamp(fp, StampOf(obj));
  const uint64_t nlib = mtllibs.size();
  ok = ok && fwrite(&nlib, 8, 1, fp) == 1;
  for (const std::string& m : mtllibs)
    ok = ok && PutStr(fp, m) && PutStamp(fp, StampOf(base_dir.empty() ? m : base_dir + "/" + m));
  ok = ok && PutVec(fp, attr.vertices) && PutVec(fp, attr.normals) && PutVec(fp, attr.texcoords);
  const uint64_t ns = shapes.size();
  ok = ok && fwrite(&ns, 8, 1, fp) == 1;
  for (const ShapeBuild& sh : shapes)
    ok = ok && PutStr(fp, sh.name) && PutVec(fp, sh.v) && PutVec(fp, sh.vn) && PutVec(fp, sh.vt) && PutVec(fp, sh.mat);
  ok = (fclose(fp) == 0) && ok;
  if (ok) ok = rename(tmp.c_str(), (obj + ".pbrcache").c_str()) == 0;
  if (!ok) remove(tmp.c_str());
  return ok;
}

bool ReadObjCache(const std::string& obj, const std::string& base_dir, std::vector<std::string>* mtllibs, Attribute* attr,
                  std::vector<ShapeBuild>* shapes) {
  FILE* fp = fopen((obj + ".pbrcache").c_str(), "rb");
  if (!fp) return false;
  char magic[8];
  bool ok = fread(magic, 8, 1, fp) == 1 && memcmp(magic, kCacheMagic, 8) == 0 && SameStamp(fp, StampOf(obj));
  uint64_t nlib = 0;
  ok = ok && fread(&nlib, 8, 1, fp) == 1 && nlib < 4096;
  for (uint64_t i = 0; ok && i < nlib; ++i) {
    std::string m;
    ok = GetStr(fp, &m) && SameStamp(fp, StampOf(base_dir.empty() ? m : base_dir + "/" + m));
    if (ok) mtllibs->push_back(m);
  }
  ok = ok && GetVec(fp, &attr->vertices) && GetVec(fp, &attr->normals) && GetVec(fp, &attr->texcoords);
  uint64_t ns = 0;
  ok = ok && fread(&ns, 8, 1, fp) == 1 && ns < (uint64_t(1) << 24);
  for (uint64_t i = 0; ok && i < ns; ++i) {
    ShapeBuild sh;
    ok = GetStr(fp, &sh.name) && GetVec(fp, &sh.v) && GetVec(fp, &sh.vn) && GetVec(fp, &sh.vt) && GetVec(fp, &sh.mat);
    if (ok) shapes->push_back(std::move(sh));
  }
  fclose(fp);
  if (!ok) { mtllibs->clear(); shapes->clear(); *attr = Attribute(); }
  return ok;
}

}  // namespace

bool LoadTriangleMeshFromObj(const std::string& filename, std::vector<TriangleMesh>* meshes,
                             std::vector<MaterialParameter>* material_params, std::vector<Texture>* textures) {
  std::shared_ptr<Attribute> attr(new Attribute());
  std::vector<RawMaterial> raw_materials;
  std::map<std::string, int> material_index;
  std::vector<ShapeBuild> shapes;
  std::vector<std::string> mtllibs;
  const char* cache_env = getenv("PBRLAB_SCENE_CACHE");
  const bool use_cache = cache_env && cache_env[0] == '1';
  {
    const size_t sl = filename.find_last_of('/');
    const std::string dir = (sl == std::string::npos) ? std::string("") : filename.substr(0, sl);
    if (use_cache && ReadObjCache(filename, dir, &mtllibs, attr.get(), &shapes)) {
      std::cerr << "Load obj file [" << filename << "] from its binary cache" << std::endl;
      for (const std::string& m : mtllibs) LoadMtl(dir.empty() ? m : dir + "/" + m, &raw_materials, &material_index);
      meshes->clear();
      for (const ShapeBuild& sb : shapes) meshes->emplace_back(sb.name, attr, sb.v, sb.vn, sb.vt, sb.mat);
      for (const RawMaterial& m : raw_materials) material_params->push_back(ToPrincipled(m, dir, textures));
      return true;
    }
  }
  const bool verbose = getenv("PBRLAB_VERBOSE_LOAD") != nullptr;
  const auto tl0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (verbose) std::cerr << "  obj load: " << what << " at " << std::chrono::duration<double>(std::chrono::steady_clock::now() - tl0).count() << " s" << std::endl;
  };
  std::string text;
  if (!ReadFile(filename, &text)) {
    std::cerr << "error : cannot open [" << filename << "]" << std::endl;
    return false;
  }
  lap("file read");
  const size_t slash = filename.find_last_of('/');
  const std::string base_dir = (slash == std::string::npos) ? std::string("") : filename.substr(0, slash);
  std::cerr << "base dir : " << base_dir << std::endl;

  // ---- parse on all host threads (the 1.5 GB of text of the 20 M-triangle scene took 13 s on one).
  // The file is cut into chunks at line ends.  (1) every chunk counts its v / vn / vt lines: prefix sums give each
  // chunk its place in the attribute pools and the number of elements "read so far" that relative (negative) indices
  // and the quad split need; (2) the chunks parse their attribute lines into the pools; (3) the chunks parse their
  // face lines into chunk-local triangle lists and note the lines that change the sequential state (o, g, usemtl,
  // mtllib) with the triangle count at which they occur; (4) one thread walks the chunks in file order, applies those
  // events and appends the triangle runs in between to the shapes.  Same arrays as the one-pass loader this replaces
  // (tests/test_loader.py compares them with the reference's loader element for element).
  struct Event { int kind; uint64_t tri_index; std::string text; };   // kind: 0 = o / g, 1 = usemtl, 2 = mtllib
  struct Chunk {
    const char* begin; const char* end;
    uint64_t nv = 0, nn = 0, nt = 0, nlines = 0;       // counted in (1)
    uint64_t v0 = 0, n0 = 0, t0 = 0, line0 = 0;        // elements / lines before this chunk
    std::vector<uint32_t> v, vn, vt;                   // 3 per triangle
    std::vector<Event> events;
    uint64_t bad_line = 0;                             // first `f' line that failed to parse (global number), 0 = none
  };
  const char* const text_begin = text.c_str();
  const char* const text_end = text_begin + text.size();
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t want_chunks = text.size() < (size_t(4) << 20) ? 1 : size_t(hw) * 4;
  std::vector<Chunk> chunks;
  {
    const size_t step = std::max<size_t>(text.size() / want_chunks, 1);
    const char* b = text_begin;
    while (b < text_end) {
      const char* e = (size_t(text_end - b) <= step) ? text_end : b + step;
      while (e < text_end && e[-1] != '\n') ++e;   // cut after a line end
      Chunk c;
      c.begin = b; c.end = e;
      chunks.push_back(std::move(c));
      b = e;
    }
  }
  auto parallel_chunks = [&](const std::function<void(Chunk&)>& fn) {
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    const unsigned nth = unsigned(std::min<size_t>(hw, chunks.size()));
    for (unsigned t = 1; t < nth; ++t)
      th.emplace_back([&]() { for (size_t i; (i = next.fetch_add(1)) < chunks.size();) fn(chunks[i]); });
    for (size_t i; (i = next.fetch_add(1)) < chunks.size();) fn(chunks[i]);
    for (auto& t : th) t.join();
  };
  // a line starts after leading blanks; returns the start of the next line
  auto next_line = [](const char* p, const char* end) {
    while (p < end && !IsEol(*p)) ++p;
    while (p < end && *p == '\r') ++p;
    if (p < end && (*p == '\n' || *p == '\0')) ++p;
    return p;
  };
  auto line_kind = [](const char* line) {   // 1 v, 2 vn, 3 vt, 4 f, 5 o/g, 6 usemtl, 7 mtllib, 0 other
    const char c0 = line[0], c1 = c0 ? line[1] : 0;
    if (c0 == 'v' && IsSpace(c1)) return 1;
    if (c0 == 'v' && c1 == 'n' && IsSpace(line[2])) return 2;
    if (c0 == 'v' && c1 == 't' && IsSpace(line[2])) return 3;
    if (c0 == 'f' && IsSpace(c1)) return 4;
    if ((c0 == 'o' || c0 == 'g') && IsSpace(c1)) return 5;
    if (strncmp(line, "usemtl", 6) == 0) return 6;
    if (strncmp(line, "mtllib", 6) == 0 && IsSpace(line[6])) return 7;
    return 0;
  };
  // (1) count
  parallel_chunks([&](Chunk& c) {
    const char* p = c.begin;
    while (p < c.end) {
      ++c.nlines;
      SkipSpace(&p);
      const int k = line_kind(p);
      if (k == 1) ++c.nv; else if (k == 2) ++c.nn; else if (k == 3) ++c.nt;
      p = next_line(p, c.end);
    }
  });
  {
    uint64_t v = 0, n = 0, t = 0, l = 0;
    for (Chunk& c : chunks) { c.v0 = v; c.n0 = n; c.t0 = t; c.line0 = l; v += c.nv; n += c.nn; t += c.nt; l += c.nlines; }
    attr->vertices.resize(size_t(v) * 4);
    attr->normals.resize(size_t(n) * 4);
    attr->texcoords.resize(size_t(t) * 2);
  }
  lap("counted");
  // (2) attributes
  parallel_chunks([&](Chunk& c) {
    const char* p = c.begin;
    float* V = attr->vertices.data() + size_t(c.v0) * 4;
    float* N = attr->normals.data() + size_t(c.n0) * 4;
    float* T = attr->texcoords.data() + size_t(c.t0) * 2;
    while (p < c.end) {
      SkipSpace(&p);
      const int k = line_kind(p);
      if (k == 1) {
        p += 2;
        V[0] = ParseFloat(&p); V[1] = ParseFloat(&p); V[2] = ParseFloat(&p); V[3] = 1.0f;
        V += 4;
      } else if (k == 2) {
        p += 3;
        N[0] = ParseFloat(&p); N[1] = ParseFloat(&p); N[2] = ParseFloat(&p); N[3] = 1.0f;
        N += 4;
      } else if (k == 3) {
        p += 3;
        const float u = ParseFloat(&p), v = ParseFloat(&p);
        T[0] = u;
        T[1] = 1.f - v;   // reference triangle-mesh-io.cc:286
        T += 2;
      }
      p = next_line(p, c.end);
    }
  });
  lap("attributes parsed");
  // (3) faces and state-changing lines
  parallel_chunks([&](Chunk& c) {
    const char* p = c.begin;
    uint64_t nv = c.v0, nn = c.n0, nt = c.t0, line_no = c.line0;
    std::vector<Corner> face;
    const float* V = attr->vertices.data();
    auto emit = [&](const Corner& a, const Corner& b, const Corner& cc) {
      c.v.push_back(uint32_t(a.v)); c.v.push_back(uint32_t(b.v)); c.v.push_back(uint32_t(cc.v));
      c.vn.push_back(uint32_t(a.vn)); c.vn.push_back(uint32_t(b.vn)); c.vn.push_back(uint32_t(cc.vn));
      c.vt.push_back(uint32_t(a.vt)); c.vt.push_back(uint32_t(b.vt)); c.vt.push_back(uint32_t(cc.vt));
    };
    while (p < c.end) {
      ++line_no;
      SkipSpace(&p);
      const char* line = p;
      const int k = line_kind(line);
      if (k == 1) ++nv;
      else if (k == 2) ++nn;
      else if (k == 3) ++nt;
      else if (k == 4 && c.bad_line == 0) {
        p += 2;
        face.clear();
        bool ok = true;
        for (;;) {
          SkipSpace(&p);
          if (IsEol(*p)) break;
          Corner cr = {-1, -1, -1};
          int raw;
          if (!ParseInt(&p, &raw) || !FixIndex(raw, int(nv), &cr.v)) { ok = false; break; }
          if (*p == '/') {
            ++p;
            if (*p == '/') {           // v//vn
              ++p;
              if (ParseInt(&p, &raw) && !FixIndex(raw, int(nn), &cr.vn)) { ok = false; break; }
            } else {
              if (ParseInt(&p, &raw) && !FixIndex(raw, int(nt), &cr.vt)) { ok = false; break; }
              if (*p == '/') {
                ++p;
                if (ParseInt(&p, &raw) && !FixIndex(raw, int(nn), &cr.vn)) { ok = false; break; }
              }
            }
          }
          face.push_back(cr);
        }
        if (!ok) {
          c.bad_line = line_no;
        } else {
          const size_t n = face.size();
          if (n == 3) {
            emit(face[0], face[1], face[2]);
          } else if (n == 4) {
            // split along the shorter diagonal (reference src/io/tiny_obj_loader.h:1519-1575)
            auto d2 = [&](int a, int b) {
              const float dx = V[4 * b] - V[4 * a], dy = V[4 * b + 1] - V[4 * a + 1], dz = V[4 * b + 2] - V[4 * a + 2];
              return dx * dx + dy * dy + dz * dz;
            };
            if (d2(face[0].v, face[2].v) < d2(face[1].v, face[3].v)) {
              emit(face[0], face[1], face[2]);
              emit(face[0], face[2], face[3]);
            } else {
              emit(face[0], face[1], face[3]);
              emit(face[1], face[2], face[3]);
            }
          } else if (n > 4) {
            for (size_t q = 1; q + 1 < n; ++q) emit(face[0], face[q], face[q + 1]);
          }
        }
      } else if (k == 5) {
        c.events.push_back({0, uint64_t(c.v.size() / 3), RestOfLine(line + 2)});
      } else if (k == 6) {
        const char* q = line + 6;
        SkipSpace(&q);
        c.events.push_back({1, uint64_t(c.v.size() / 3), RestOfLine(q)});
      } else if (k == 7) {
        const char* q = line + 7;
        SkipSpace(&q);
        c.events.push_back({2, uint64_t(c.v.size() / 3), RestOfLine(q)});
      }
      p = next_line(p, c.end);
    }
  });
  lap("faces parsed");
  // (4) the sequential state, in file order
  ShapeBuild cur;
  int cur_material = -1;
  auto flush = [&]() {
    if (!cur.v.empty()) shapes.push_back(std::move(cur));
    cur = ShapeBuild();
  };
  auto append_run = [&](const Chunk& c, uint64_t t0, uint64_t t1) {
    if (t1 <= t0) return;
    cur.v.insert(cur.v.end(), c.v.begin() + 3 * t0, c.v.begin() + 3 * t1);
    cur.vn.insert(cur.vn.end(), c.vn.begin() + 3 * t0, c.vn.begin() + 3 * t1);
    cur.vt.insert(cur.vt.end(), c.vt.begin() + 3 * t0, c.vt.begin() + 3 * t1);
    cur.mat.insert(cur.mat.end(), size_t(t1 - t0), uint32_t(cur_material));
  };
  for (Chunk& c : chunks) {
    uint64_t pos = 0;
    // a parse error stops the load at that line, like the one-pass loader (everything before it is irrelevant then)
    if (c.bad_line) {
      std::cerr << "error : Failed to parse `f' line " << c.bad_line << std::endl;
      return false;
    }
    for (const Event& ev : c.events) {
      append_run(c, pos, ev.tri_index);
      pos = ev.tri_index;
      if (ev.kind == 0) {
        const std::string name = ev.text;
        if (!cur.v.empty()) shapes.push_back(std::move(cur));
        cur = ShapeBuild();
        cur.name = name;
      } else if (ev.kind == 1) {
        auto it = material_index.find(ev.text);
        cur_material = (it == material_index.end()) ? -1 : it->second;
      } else {
        mtllibs.push_back(ev.text);
        LoadMtl(base_dir.empty() ? ev.text : base_dir + "/" + ev.text, &raw_materials, &material_index);
      }
    }
    append_run(c, pos, c.v.size() / 3);
    std::vector<uint32_t>().swap(c.v); std::vector<uint32_t>().swap(c.vn); std::vector<uint32_t>().swap(c.vt);
  }
  flush();
  lap("shapes assembled");
  if (use_cache) WriteObjCache(filename, base_dir, mtllibs, *attr, shapes);

  meshes->clear();
  for (const ShapeBuild& s : shapes) meshes->emplace_back(s.name, attr, s.v, s.vn, s.vt, s.mat);
  for (const RawMaterial& m : raw_materials) material_params->push_back(ToPrincipled(m, base_dir, textures));
  lap("meshes built");
  return true;
}

}  // namespace io
}  // namespace pbrlab
