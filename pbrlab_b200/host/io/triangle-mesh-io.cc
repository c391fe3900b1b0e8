#include "triangle-mesh-io.h"

#include <algorithm>
#include <cctype>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>

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
  std::string text;
  if (!ReadFile(filename, &text)) {
    std::cerr << "error : cannot open [" << filename << "]" << std::endl;
    return false;
  }
  const size_t slash = filename.find_last_of('/');
  const std::string base_dir = (slash == std::string::npos) ? std::string("") : filename.substr(0, slash);
  std::cerr << "base dir : " << base_dir << std::endl;

  ShapeBuild cur;
  int cur_material = -1;
  std::vector<Corner> face;

  auto flush = [&]() {
    if (!cur.v.empty()) shapes.push_back(cur);
    const std::string keep = cur.name;
    cur = ShapeBuild();
    cur.name = keep;
  };

  const char* p = text.c_str();
  size_t line_no = 0;
  while (*p) {
    ++line_no;
    SkipSpace(&p);
    const char* line = p;
    const char c0 = line[0], c1 = c0 ? line[1] : 0;
    if (c0 == 'v' && IsSpace(c1)) {
      p += 2;
      const float x = ParseFloat(&p), y = ParseFloat(&p), z = ParseFloat(&p);
      attr->vertices.push_back(x); attr->vertices.push_back(y); attr->vertices.push_back(z);
      attr->vertices.push_back(1.0f);
    } else if (c0 == 'v' && c1 == 'n' && IsSpace(line[2])) {
      p += 3;
      const float x = ParseFloat(&p), y = ParseFloat(&p), z = ParseFloat(&p);
      attr->normals.push_back(x); attr->normals.push_back(y); attr->normals.push_back(z);
      attr->normals.push_back(1.0f);
    } else if (c0 == 'v' && c1 == 't' && IsSpace(line[2])) {
      p += 3;
      const float u = ParseFloat(&p), v = ParseFloat(&p);
      attr->texcoords.push_back(u);
      attr->texcoords.push_back(1.f - v);   // reference triangle-mesh-io.cc:286
    } else if (c0 == 'f' && IsSpace(c1)) {
      p += 2;
      face.clear();
      const int nv = int(attr->vertices.size() / 4), nn = int(attr->normals.size() / 4),
                nt = int(attr->texcoords.size() / 2);
      for (;;) {
        SkipSpace(&p);
        if (IsEol(*p)) break;
        Corner c = {-1, -1, -1};
        int raw;
        if (!ParseInt(&p, &raw) || !FixIndex(raw, nv, &c.v)) {
          std::cerr << "error : Failed to parse `f' line " << line_no << std::endl;
          return false;
        }
        if (*p == '/') {
          ++p;
          if (*p == '/') {           // v//vn
            ++p;
            if (ParseInt(&p, &raw) && !FixIndex(raw, nn, &c.vn)) return false;
          } else {
            if (ParseInt(&p, &raw) && !FixIndex(raw, nt, &c.vt)) return false;
            if (*p == '/') {
              ++p;
              if (ParseInt(&p, &raw) && !FixIndex(raw, nn, &c.vn)) return false;
            }
          }
        }
        face.push_back(c);
      }
      auto emit = [&](const Corner& a, const Corner& b, const Corner& c) {
        cur.v.push_back(uint32_t(a.v)); cur.v.push_back(uint32_t(b.v)); cur.v.push_back(uint32_t(c.v));
        cur.vn.push_back(uint32_t(a.vn)); cur.vn.push_back(uint32_t(b.vn)); cur.vn.push_back(uint32_t(c.vn));
        cur.vt.push_back(uint32_t(a.vt)); cur.vt.push_back(uint32_t(b.vt)); cur.vt.push_back(uint32_t(c.vt));
        cur.mat.push_back(uint32_t(cur_material));
      };
      const size_t n = face.size();
      if (n == 3) {
        emit(face[0], face[1], face[2]);
      } else if (n == 4) {
        // split along the shorter diagonal (reference src/io/tiny_obj_loader.h:1519-1575)
        const float* V = attr->vertices.data();
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
        for (size_t k = 1; k + 1 < n; ++k) emit(face[0], face[k], face[k + 1]);
      }
    } else if (c0 == 'o' && IsSpace(c1)) {
      flush();
      cur.name = RestOfLine(line + 2);
    } else if (c0 == 'g' && IsSpace(c1)) {
      flush();
      cur.name = RestOfLine(line + 2);
    } else if (strncmp(line, "usemtl", 6) == 0) {
      const char* q = line + 6;
      SkipSpace(&q);
      const std::string name = RestOfLine(q);
      auto it = material_index.find(name);
      cur_material = (it == material_index.end()) ? -1 : it->second;
    } else if (strncmp(line, "mtllib", 6) == 0 && IsSpace(line[6])) {
      const char* q = line + 7;
      SkipSpace(&q);
      const std::string name = RestOfLine(q);
      mtllibs.push_back(name);
      LoadMtl(base_dir.empty() ? name : base_dir + "/" + name, &raw_materials, &material_index);
    }
    while (!IsEol(*p)) ++p;
    while (*p == '\r') ++p;
    if (*p == '\n') ++p;
  }
  flush();
  if (use_cache) WriteObjCache(filename, base_dir, mtllibs, *attr, shapes);

  meshes->clear();
  for (const ShapeBuild& s : shapes) meshes->emplace_back(s.name, attr, s.v, s.vn, s.vt, s.mat);
  for (const RawMaterial& m : raw_materials) material_params->push_back(ToPrincipled(m, base_dir, textures));
  return true;
}

}  // namespace io
}  // namespace pbrlab
