// texture.cpp -- host half of the texture operands (see texture.h).  The per-hit evaluation is device code
// (pbrt-v1_b200/csrc/pb_texture.h); what runs here happens once per scene, exactly as in the reference's
// constructors: reading the image (core/exrio.cpp:40-73), resampling it to a power of two with Lanczos weights
// and box-filtering the pyramid (core/mipmap.h:89-165), filling the EWA weight table (:167-176).
#include "texture.h"
// The lookup arithmetic is the device's stage code (csrc/pb_texture.h).  This host copy runs inside constructors
// whose results the reference computes with glibc's float libm (logf in MIPMap::Lookup), so it is compiled against
// the float functions rather than the device's double-rounded ones.
#define PB_HOST_LIBM_FLOAT 1
#include "../csrc/pb_texture.h"
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ------------------------------------------------------------------ OpenEXR input
// Scan-line files, NO_COMPRESSION, channels of type HALF or FLOAT.  The reference reads R, G and B through a
// HALF frame buffer (exrio.cpp:49-58): FLOAT channels are narrowed to half on the way, absent channels read 0.
namespace {
float HalfBitsToFloat(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, mant = h & 0x3ffu, bits;
  if (exp == 0) {
    if (mant == 0) bits = sign;
    else {                              // subnormal half: renormalise
      int e = -1;
      do { ++e; mant <<= 1; } while (!(mant & 0x400u));
      bits = sign | (uint32_t)(127 - 15 - e) << 23 | (mant & 0x3ffu) << 13;
    }
  } else if (exp == 31) bits = sign | 0x7f800000u | mant << 13;
  else bits = sign | (exp + 127 - 15) << 23 | mant << 13;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
uint16_t FloatToHalfBits(float f) {   // round to nearest even, as half.h
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  int32_t exp = (int32_t)((x >> 23) & 0xffu) - 127 + 15;
  uint32_t mant = x & 0x7fffffu;
  if (((x >> 23) & 0xffu) == 0xffu) return (uint16_t)(sign | 0x7c00u | (mant ? 0x200u | (mant >> 13) : 0u));
  if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
  if (exp <= 0) {
    if (exp < -10) return (uint16_t)sign;
    mant |= 0x800000u;
    uint32_t shift = (uint32_t)(14 - exp), half = mant >> shift, rem = mant & ((1u << shift) - 1u), mid = 1u << (shift - 1);
    if (rem > mid || (rem == mid && (half & 1u))) ++half;
    return (uint16_t)(sign | half);
  }
  uint32_t half = ((uint32_t)exp << 10) | (mant >> 13), rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ++half;
  return (uint16_t)(sign | half);
}
struct ExrChannel { string name; int type; };
bool ReadAll(const string &name, std::vector<unsigned char> *buf) {
  FILE *f = fopen(name.c_str(), "rb");
  if (!f) return false;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf->insert(buf->end(), tmp, tmp + n);
  fclose(f);
  return true;
}
}  // namespace

Spectrum *ReadImage(const string &name, int *width, int *height) {
  std::vector<unsigned char> buf;
  const char *why = NULL;
  if (!ReadAll(name, &buf)) why = strerror(errno);
  std::vector<ExrChannel> channels;
  int32_t dw[4] = {0, 0, -1, -1};
  int compression = -1;
  size_t pos = 8;
  if (!why) {
    uint32_t magic = 0, version = 0;
    if (buf.size() >= 8) { memcpy(&magic, &buf[0], 4); memcpy(&version, &buf[4], 4); }
    if (magic != 20000630u) why = "not an OpenEXR file";
    else if (version & 0x1a00u) why = "tiled, deep or multi-part files are not supported";
  }
  while (!why) {   // header attributes: name\0 type\0 size value, ended by an empty name
    size_t e = pos;
    while (e < buf.size() && buf[e]) ++e;
    if (e >= buf.size()) { why = "truncated header"; break; }
    string attr((const char *)&buf[pos], e - pos);
    pos = e + 1;
    if (attr == "") break;
    e = pos;
    while (e < buf.size() && buf[e]) ++e;
    if (e + 5 > buf.size()) { why = "truncated header"; break; }
    pos = e + 1;
    int32_t size;
    memcpy(&size, &buf[pos], 4);
    pos += 4;
    if (size < 0 || pos + (size_t)size > buf.size()) { why = "truncated header"; break; }
    if (attr == "channels") {
      // channel list: name\0, then 16 bytes (pixel type, pLinear + reserved, x / y sampling); every read stays
      // inside the attribute's `size` bytes
      size_t p = pos;
      const size_t end = pos + (size_t)size;
      while (p < end && buf[p]) {
        ExrChannel ch;
        while (p < end && buf[p]) ch.name += (char)buf[p++];
        if (p >= end || end - (p + 1) < 16) { why = "truncated channel list"; break; }
        ++p;
        int32_t type;
        memcpy(&type, &buf[p], 4);
        ch.type = type;
        p += 16;
        channels.push_back(ch);
      }
      if (why) break;
    } else if (attr == "compression") compression = buf[pos];
    else if (attr == "dataWindow" && size == 16) memcpy(dw, &buf[pos], 16);
    pos += (size_t)size;
  }
  if (!why && compression != 0) why = "only uncompressed files are supported";
  if (!why && (dw[2] < dw[0] || dw[3] < dw[1])) why = "no data window";
  if (why) {
    Error("Unable to read image file \"%s\": %s", name.c_str(), why);
    return NULL;
  }
  // (sizes in 64 bits: a hostile data window must not wrap the pixel count)
  const int64_t w64 = (int64_t)dw[2] - dw[0] + 1, h64 = (int64_t)dw[3] - dw[1] + 1;
  if (w64 < 1 || h64 < 1 || w64 > (1 << 20) || h64 > (1 << 20) || (uint64_t)(w64 * h64) > (uint64_t)buf.size()) {
    // an uncompressed file holds at least 2 bytes per pixel and channel: more pixels than bytes cannot be real
    Error("Unable to read image file \"%s\": data window %lld x %lld does not fit the file", name.c_str(),
          (long long)w64, (long long)h64);
    return NULL;
  }
  int w = (int)w64, h = (int)h64;
  size_t rowBytes = 0;
  for (size_t c = 0; c < channels.size(); ++c) {
    if (channels[c].type != 1 && channels[c].type != 2) {
      Error("Unable to read image file \"%s\": UINT channels are not supported", name.c_str());
      return NULL;
    }
    rowBytes += (size_t)w * (channels[c].type == 1 ? 2 : 4);
  }
  // the scan-line offset table and every scan line it points at must lie inside the file (OpenEXR throws on a
  // truncated file and the reference's ReadImage returns NULL, core/exrio.cpp:40-73)
  if (pos > buf.size() || (buf.size() - pos) / 8 < (size_t)h || rowBytes > buf.size()) {
    Error("Unable to read image file \"%s\": truncated scan-line offset table", name.c_str());
    return NULL;
  }
  for (int row = 0; row < h; ++row) {
    uint64_t off;
    memcpy(&off, &buf[pos + 8 * (size_t)row], 8);
    if (off > buf.size() || buf.size() - off < 8 + rowBytes) {
      Error("Unable to read image file \"%s\": scan line %d lies outside the file", name.c_str(), row);
      return NULL;
    }
  }
  Spectrum *ret = new Spectrum[(size_t)w * h];
  for (size_t i = 0; i < (size_t)w * h; ++i) ret[i] = Spectrum(0.f);
  for (int row = 0; row < h; ++row) {
    uint64_t off;
    memcpy(&off, &buf[pos + 8 * (size_t)row], 8);
    int32_t y;
    memcpy(&y, &buf[off], 4);
    if (y < dw[1] || y > dw[3]) continue;
    const unsigned char *p = &buf[off + 8];
    for (size_t c = 0; c < channels.size(); ++c) {
      int comp = channels[c].name == "R" ? 0 : channels[c].name == "G" ? 1 : channels[c].name == "B" ? 2 : -1;
      size_t bytes = channels[c].type == 1 ? 2 : 4;
      if (comp >= 0)
        for (int x = 0; x < w; ++x) {
          uint16_t hb;
          if (bytes == 2) memcpy(&hb, p + 2 * (size_t)x, 2);
          else { float f; memcpy(&f, p + 4 * (size_t)x, 4); hb = FloatToHalfBits(f); }
          ret[(size_t)(y - dw[1]) * w + x].c[comp] = HalfBitsToFloat(hb);
        }
      p += bytes * (size_t)w;
    }
  }
  *width = w;
  *height = h;
  return ret;
}

// ------------------------------------------------------------------ MIPMap construction
namespace {
// Lanczos, core/texture.cpp:241-249
float Lanczos(float x, float tau = 2) {
  x = fabsf(x);
  if (x < 1e-5) return 1;
  if (x > 1.) return 0;
  x *= M_PI;
  float s = sinf(x * tau) / (x * tau);
  float lanczos = sinf(x) / x;
  return s * lanczos;
}
inline int Mod(int a, int b) {
  int n = int(a / b);
  a -= n * b;
  if (a < 0) a += b;
  return a;
}
inline bool IsPowerOf2(int v) { return (v & (v - 1)) == 0; }
inline uint32_t RoundUpPow2(uint32_t v) {
  v--;
  v |= v >> 1; v |= v >> 2; v |= v >> 4; v |= v >> 8; v |= v >> 16;
  return v + 1;
}
struct ResampleWeight { int firstTexel; float weight[4]; };
// MIPMap::resampleWeights, core/mipmap.h:46-66
std::vector<ResampleWeight> ResampleWeights(int oldres, int newres) {
  std::vector<ResampleWeight> wt(newres);
  float filterwidth = 2.f;
  for (int i = 0; i < newres; ++i) {
    float center = (i + .5f) * oldres / newres;
    wt[i].firstTexel = Floor2Int((center - filterwidth) + 0.5f);
    for (int j = 0; j < 4; ++j) {
      float pos = wt[i].firstTexel + j + .5f;
      wt[i].weight[j] = Lanczos((pos - center) / filterwidth);
    }
    float invSumWts = 1.f / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
    for (int j = 0; j < 4; ++j) wt[i].weight[j] *= invSumWts;
  }
  return wt;
}
int WrapIndex(int i, int res, uint32_t wrap) {
  if (wrap == PBRT_B200_TEXWRAP_REPEAT) return Mod(i, res);
  if (wrap == PBRT_B200_TEXWRAP_CLAMP) return Clamp(i, 0, res - 1);
  return i;
}
// One image plane of `ch` interleaved floats.  Every operation of the MIPMap<T> constructor acts on the
// components of T independently (T = float, or Spectrum whose operators are component-wise), so the loops
// below run the same float expression per component in the same order.
struct Plane {
  int w, h, ch;
  std::vector<float> v;
  float &at(int s, int t, int c) { return v[((size_t)t * w + s) * ch + c]; }
  float get(int s, int t, int c) const { return v[((size_t)t * w + s) * ch + c]; }
};
float LevelTexel(const Plane &l, int s, int t, int c, uint32_t wrap) {   // MIPMap::texel, mipmap.h:178-199
  if (wrap == PBRT_B200_TEXWRAP_REPEAT) { s = Mod(s, l.w); t = Mod(t, l.h); }
  else if (wrap == PBRT_B200_TEXWRAP_CLAMP) { s = Clamp(s, 0, l.w - 1); t = Clamp(t, 0, l.h - 1); }
  else if (s < 0 || s >= l.w || t < 0 || t >= l.h) return 0.f;
  return l.get(s, t, c);
}
// mipmap.h:89-165; appends the levels to the store and returns (first level, level count)
std::pair<uint32_t, uint32_t> BuildPyramid(TextureStore &st, Plane img, uint32_t wrap) {
  int sres = img.w, tres = img.h, ch = img.ch;
  if (!IsPowerOf2(sres) || !IsPowerOf2(tres)) {
    int sPow2 = (int)RoundUpPow2((uint32_t)sres), tPow2 = (int)RoundUpPow2((uint32_t)tres);
    std::vector<ResampleWeight> sWeights = ResampleWeights(sres, sPow2);
    Plane res;
    res.w = sPow2; res.h = tPow2; res.ch = ch;
    res.v.assign((size_t)sPow2 * tPow2 * ch, 0.f);
    for (int t = 0; t < tres; ++t)
      for (int s = 0; s < sPow2; ++s)
        for (int c = 0; c < ch; ++c) {
          float acc = 0.f;
          for (int j = 0; j < 4; ++j) {
            int origS = WrapIndex(sWeights[s].firstTexel + j, sres, wrap);
            if (origS >= 0 && origS < sres) acc += sWeights[s].weight[j] * img.get(origS, t, c);
          }
          res.at(s, t, c) = acc;
        }
    std::vector<ResampleWeight> tWeights = ResampleWeights(tres, tPow2);
    std::vector<float> work(tPow2);
    for (int s = 0; s < sPow2; ++s)
      for (int c = 0; c < ch; ++c) {
        for (int t = 0; t < tPow2; ++t) {
          work[t] = 0.f;
          for (int j = 0; j < 4; ++j) {
            int offset = WrapIndex(tWeights[t].firstTexel + j, tres, wrap);
            if (offset >= 0 && offset < tres) work[t] += tWeights[t].weight[j] * res.get(s, offset, c);
          }
        }
        for (int t = 0; t < tPow2; ++t) res.at(s, t, c) = Clamp(work[t], 0.f, INFINITY);
      }
    img = res;
    sres = sPow2;
    tres = tPow2;
  }
  int nLevels = 1 + Log2Int(float(max(sres, tres)));
  uint32_t first = (uint32_t)st.levels.size();
  Plane prev = img;
  for (int i = 0; i < nLevels; ++i) {
    Plane cur;
    if (i == 0) cur = prev;
    else {
      cur.w = max(1, prev.w / 2); cur.h = max(1, prev.h / 2); cur.ch = ch;
      cur.v.resize((size_t)cur.w * cur.h * ch);
      for (int t = 0; t < cur.h; ++t)
        for (int s = 0; s < cur.w; ++s)
          for (int c = 0; c < ch; ++c)
            cur.at(s, t, c) = .25f * (LevelTexel(prev, 2 * s, 2 * t, c, wrap) + LevelTexel(prev, 2 * s + 1, 2 * t, c, wrap) +
                                      LevelTexel(prev, 2 * s, 2 * t + 1, c, wrap) + LevelTexel(prev, 2 * s + 1, 2 * t + 1, c, wrap));
    }
    PbrtB200TexLevel l;
    l.offset = st.texels.size(); l.width = (uint32_t)cur.w; l.height = (uint32_t)cur.h;
    st.levels.push_back(l);
    st.texels.insert(st.texels.end(), cur.v.begin(), cur.v.end());
    prev = cur;
  }
  return std::make_pair(first, (uint32_t)nLevels);
}
}  // namespace

TextureStore::TextureStore() {
  for (int i = 0; i < PBRT_B200_EWA_LUT_SIZE; ++i) {   // mipmap.h:167-176
    float alpha = 2;
    float r2 = float(i) / float(PBRT_B200_EWA_LUT_SIZE - 1);
    ewaLut[i] = expf(-alpha * r2) - expf(-alpha);
  }
}
int TextureStore::AddConstant(const float v[3], int channels) {
  PbrtB200Texture t;
  memset(&t, 0, sizeof(t));
  t.kind = PBRT_B200_TEX_CONSTANT;
  t.channels = (uint32_t)channels;
  for (int i = 0; i < 3; ++i) { t.value[i] = v[i]; t.arg[i] = -1; }
  nodes.push_back(t);
  return (int)nodes.size() - 1;
}
int TextureStore::AddScale(int a, int b, int channels) {
  PbrtB200Texture t;
  memset(&t, 0, sizeof(t));
  t.kind = PBRT_B200_TEX_SCALE;
  t.channels = (uint32_t)channels;
  t.arg[0] = a; t.arg[1] = b; t.arg[2] = -1;
  nodes.push_back(t);
  return (int)nodes.size() - 1;
}
int TextureStore::AddMix(int a, int b, int amount, int channels) {
  PbrtB200Texture t;
  memset(&t, 0, sizeof(t));
  t.kind = PBRT_B200_TEX_MIX;
  t.channels = (uint32_t)channels;
  t.arg[0] = a; t.arg[1] = b; t.arg[2] = amount;
  nodes.push_back(t);
  return (int)nodes.size() - 1;
}
int TextureStore::Depth(int node) const {
  const PbrtB200Texture &t = nodes[node];
  if (t.kind != PBRT_B200_TEX_SCALE && t.kind != PBRT_B200_TEX_MIX && t.kind != PBRT_B200_TEX_CHECKERBOARD &&
      t.kind != PBRT_B200_TEX_DOTS) return 0;
  int d = 0;
  for (int i = 0; i < 3; ++i)
    if (t.arg[i] >= 0) d = max(d, Depth(t.arg[i]));
  return d + 1;
}
// the "mapping" block every 2D texture plugin starts with (e.g. textures/imagemap.cpp:133-151)
static void ParseMapping2D(const Transform &tex2world, const TextureParams &tp, PbrtB200Texture *t) {
  string type = tp.FindString("mapping");
  t->mapping = PBRT_B200_TEXMAP_UV;
  t->su = t->sv = 1.f;
  if (type == "" || type == "uv") {
    t->su = tp.FindFloat("uscale", 1.);
    t->sv = tp.FindFloat("vscale", 1.);
    t->du = tp.FindFloat("udelta", 0.);
    t->dv = tp.FindFloat("vdelta", 0.);
  } else if (type == "spherical" || type == "cylindrical") {
    t->mapping = type == "spherical" ? PBRT_B200_TEXMAP_SPHERICAL : PBRT_B200_TEXMAP_CYLINDRICAL;
    Transform w2t = tex2world.GetInverse();
    const Matrix4x4 &M = w2t.GetMatrix();
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) t->world_to_texture[4 * i + j] = M.m[i][j];
  } else if (type == "planar") {
    t->mapping = PBRT_B200_TEXMAP_PLANAR;
    Vector v1 = tp.FindVector("v1", Vector(1, 0, 0)), v2 = tp.FindVector("v2", Vector(0, 1, 0));
    t->v1[0] = v1.x; t->v1[1] = v1.y; t->v1[2] = v1.z;
    t->v2[0] = v2.x; t->v2[1] = v2.y; t->v2[2] = v2.z;
    t->du = tp.FindFloat("udelta", 0.f);
    t->dv = tp.FindFloat("vdelta", 0.f);
  } else
    Error("2D texture mapping \"%s\" unknown", type.c_str());
}
int TextureStore::AddProcedural(const string &texname, const Transform &tex2world, const TextureParams &tp,
                                int channels, int arg0, int arg1) {
  PbrtB200Texture t;
  memset(&t, 0, sizeof(t));
  t.channels = (uint32_t)channels;
  t.arg[0] = arg0; t.arg[1] = arg1; t.arg[2] = -1;
  if (texname == "uv" || texname == "bilerp" || texname == "checkerboard" || texname == "dots") {
    if (texname == "uv" && channels != 3) return -1;   // uv.cpp:49-52: no float version
    if (texname == "checkerboard") {
      int dim = tp.FindInt("dimension", 2);
      if (dim != 2) {
        Error("checkerboard: only the 2D texture is supported (the 3D one supersamples with random numbers)");
        return -1;
      }
    }
    ParseMapping2D(tex2world, tp, &t);
    if (texname == "uv") t.kind = PBRT_B200_TEX_UV;
    else if (texname == "bilerp") {
      t.kind = PBRT_B200_TEX_BILERP;
      const char *names[4] = {"v00", "v01", "v10", "v11"};
      for (int k = 0; k < 4; ++k) {
        float dflt = (k & 1) ? 1.f : 0.f;   // bilerp.cpp:77-79,105-107
        if (channels == 3) {
          Spectrum v = tp.FindSpectrum(names[k], Spectrum(dflt));
          for (int c = 0; c < 3; ++c) t.param[3 * k + c] = v.c[c];
        } else {
          float v = tp.FindFloat(names[k], dflt);
          for (int c = 0; c < 3; ++c) t.param[3 * k + c] = v;
        }
      }
    } else if (texname == "checkerboard") {
      t.kind = PBRT_B200_TEX_CHECKERBOARD;
      string aamode = tp.FindString("aamode");
      if (aamode == "" || aamode == "closedform") t.iparam[0] = PBRT_B200_CHECKER_CLOSEDFORM;
      else if (aamode == "none") t.iparam[0] = PBRT_B200_CHECKER_NONE;
      else {
        Error("checkerboard: anti-aliasing mode \"%s\" draws random numbers inside the texture and is not "
              "supported (use closedform or none)", aamode.c_str());
        return -1;
      }
    } else
      t.kind = PBRT_B200_TEX_DOTS;
  } else {
    if (texname == "marble" && channels != 3) return -1;   // marble.cpp:79-82
    t.kind = texname == "fbm" ? PBRT_B200_TEX_FBM : texname == "wrinkled" ? PBRT_B200_TEX_WRINKLED
           : texname == "windy" ? PBRT_B200_TEX_WINDY : PBRT_B200_TEX_MARBLE;
    // `new IdentityMapping3D(tex2world)`: the plugins hand the texture-to-world transform to a mapping that
    // applies it as world-to-texture (fbm.cpp:56, core/texture.h:96-97); kept as is
    const Matrix4x4 &M = tex2world.GetMatrix();
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) t.world_to_texture[4 * i + j] = M.m[i][j];
    if (t.kind != PBRT_B200_TEX_WINDY) {
      t.iparam[0] = tp.FindInt("octaves", 8);
      t.param[0] = tp.FindFloat("roughness", .5f);
    }
    if (t.kind == PBRT_B200_TEX_MARBLE) {
      t.param[1] = tp.FindFloat("scale", 1.f);
      t.param[2] = tp.FindFloat("variation", .2f);
    }
  }
  nodes.push_back(t);
  return (int)nodes.size() - 1;
}
// CreateFloatTexture / CreateSpectrumTexture, textures/imagemap.cpp:130-195
int TextureStore::AddImageMap(const Transform &tex2world, const TextureParams &tp, int channels) {
  PbrtB200Texture t;
  memset(&t, 0, sizeof(t));
  t.kind = PBRT_B200_TEX_IMAGEMAP;
  t.channels = (uint32_t)channels;
  for (int i = 0; i < 3; ++i) t.arg[i] = -1;
  ParseMapping2D(tex2world, tp, &t);
  t.max_anisotropy = tp.FindFloat("maxanisotropy", 8.f);
  t.trilinear = tp.FindBool("trilinear", false) ? 1u : 0u;
  string wrap = tp.FindString("wrap");
  t.wrap = PBRT_B200_TEXWRAP_REPEAT;
  if (wrap == "black") t.wrap = PBRT_B200_TEXWRAP_BLACK;
  else if (wrap == "clamp") t.wrap = PBRT_B200_TEXWRAP_CLAMP;
  string filename = tp.FindString("filename");
  // ImageTexture::GetTexture: the pyramid depends on the file, the wrap mode (it enters the resampling and
  // the box filter) and the texel type; filter and anisotropy only matter at lookup time
  char key[64];
  snprintf(key, sizeof(key), "|%u|%d", t.wrap, channels);
  std::map<string, std::pair<uint32_t, uint32_t> >::iterator it = pyramids.find(filename + key);
  if (it == pyramids.end()) {
    int width = 0, height = 0;
    Spectrum *texels = ReadImage(ResolveDataFile(filename), &width, &height);
    Plane img;
    std::pair<uint32_t, uint32_t> pyr;
    if (texels) {
      imageSizes[filename] = std::make_pair(width, height);
      img.w = width; img.h = height; img.ch = channels;
      img.v.resize((size_t)width * height * channels);
      for (int i = 0; i < width * height; ++i) {
        if (channels == 3) for (int c = 0; c < 3; ++c) img.v[3 * (size_t)i + c] = texels[i].c[c];
        else {   // convert(Spectrum -> float) = y(), core/color.h:141-146
          const float YWeight[3] = {0.212671f, 0.715160f, 0.072169f};
          float v = 0.;
          for (int c = 0; c < 3; ++c) v += YWeight[c] * texels[i].c[c];
          img.v[i] = v;
        }
      }
      delete[] texels;
      pyr = BuildPyramid(*this, img, t.wrap);
    } else {   // one-valued MIPMap built with the constructor's defaults (imagemap.cpp:111-117, mipmap.h:39-40)
      t.wrap = PBRT_B200_TEXWRAP_REPEAT; t.trilinear = 0; t.max_anisotropy = 8.f;
      img.w = img.h = 1; img.ch = channels;
      img.v.assign(channels, 1.f);
      pyr = BuildPyramid(*this, img, PBRT_B200_TEXWRAP_REPEAT);
    }
    it = pyramids.insert(std::make_pair(filename + key, pyr)).first;
  }
  t.first_level = it->second.first;
  t.n_levels = it->second.second;
  nodes.push_back(t);
  return (int)nodes.size() - 1;
}
int TextureStore::AddRadianceMap(const string &filename) {
  PbrtB200Texture t;
  memset(&t, 0, sizeof(t));
  t.kind = PBRT_B200_TEX_IMAGEMAP;
  t.channels = 3;
  for (int i = 0; i < 3; ++i) t.arg[i] = -1;
  t.mapping = PBRT_B200_TEXMAP_UV;
  t.su = t.sv = 1.f;
  t.wrap = PBRT_B200_TEXWRAP_REPEAT;
  t.max_anisotropy = 8.f;
  string key = filename + "|0|3";   // same pyramid as a repeat-wrapped color image map of the file
  std::map<string, std::pair<uint32_t, uint32_t> >::iterator it = pyramids.find(key);
  if (it == pyramids.end()) {
    int width = 0, height = 0;
    Spectrum *texels = ReadImage(ResolveDataFile(filename), &width, &height);
    if (!texels) return -1;
    imageSizes[filename] = std::make_pair(width, height);
    Plane img;
    img.w = width; img.h = height; img.ch = 3;
    img.v.resize((size_t)width * height * 3);
    for (int i = 0; i < width * height; ++i)
      for (int c = 0; c < 3; ++c) img.v[3 * (size_t)i + c] = texels[i].c[c];
    delete[] texels;
    it = pyramids.insert(std::make_pair(key, BuildPyramid(*this, img, PBRT_B200_TEXWRAP_REPEAT))).first;
  }
  t.first_level = it->second.first;
  t.n_levels = it->second.second;
  nodes.push_back(t);
  return (int)nodes.size() - 1;
}
void TextureStore::LookupWidth(int node, float s, float t, float width, float rgb[3]) const {
  SceneView sv;
  memset(&sv, 0, sizeof(sv));
  sv.textures = &nodes[0];
  sv.texLevels = &levels[0];
  sv.texels = &texels[0];
  sv.ewaLut = ewaLut;
  Spec v = mip_lookup_width(sv, nodes[node], s, t, width);
  for (int c = 0; c < 3; ++c) rgb[c] = v.c[c];
}
