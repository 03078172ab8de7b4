// film.cpp -- pixel filters, the "image" film's host half, and EXR output.
//
// Filters: filters/{box,triangle,gaussian,mitchell,sinc}.cpp -- only ever
// sampled into the 16x16 table at film construction (film/image.cpp:89-101),
// which is what gets uploaded to the device.
// ImageFilm: crop/pixel-extent arithmetic of film/image.cpp:69-88,148-156; the
// accumulate (AddSample) and resolve (WriteImage) loops run on the GPU.
// EXR: the reference calls OpenEXR's RgbaOutputFile (core/exrio.cpp:75-96),
// which is not available here, so a minimal scan-line writer produces the same
// kind of file: half RGBA, display window = full resolution, data window = crop.
#include "core.h"
#include <errno.h>

namespace {

class BoxFilter : public Filter {
 public:
  BoxFilter(float xw, float yw) : Filter(xw, yw) {}
  float Evaluate(float, float) const { return 1.; }
};
class TriangleFilter : public Filter {
 public:
  TriangleFilter(float xw, float yw) : Filter(xw, yw) {}
  float Evaluate(float x, float y) const {
    return max(0.f, xWidth - fabsf(x)) * max(0.f, yWidth - fabsf(y));
  }
};
class GaussianFilter : public Filter {
 public:
  GaussianFilter(float xw, float yw, float a) : Filter(xw, yw) {
    alpha = a;
    expX = expf(-alpha * xWidth * xWidth);
    expY = expf(-alpha * yWidth * yWidth);
  }
  float Evaluate(float x, float y) const { return Gaussian(x, expX) * Gaussian(y, expY); }
 private:
  float Gaussian(float d, float expv) const {
    return max(0.f, float(expf(-alpha * d * d) - expv));
  }
  float alpha, expX, expY;
};
class MitchellFilter : public Filter {
 public:
  MitchellFilter(float b, float c, float xw, float yw) : Filter(xw, yw), B(b), C(c) {}
  float Evaluate(float x, float y) const {
    return Mitchell1D(x * invXWidth) * Mitchell1D(y * invYWidth);
  }
 private:
  // filters/mitchell.cpp:36-46
  float Mitchell1D(float x) const {
    x = fabsf(2.f * x);
    if (x > 1.f)
      return ((-B - 6 * C) * x * x * x + (6 * B + 30 * C) * x * x + (-12 * B - 48 * C) * x +
              (8 * B + 24 * C)) * (1.f / 6.f);
    else
      return ((12 - 9 * B - 6 * C) * x * x * x + (-18 + 12 * B + 6 * C) * x * x + (6 - 2 * B)) *
             (1.f / 6.f);
  }
  float B, C;
};
class LanczosSincFilter : public Filter {
 public:
  LanczosSincFilter(float xw, float yw, float t) : Filter(xw, yw), tau(t) {}
  float Evaluate(float x, float y) const { return Sinc1D(x * invXWidth) * Sinc1D(y * invYWidth); }
 private:
  float Sinc1D(float x) const {
    x = fabsf(x);
    if (x < 1e-5) return 1.f;
    if (x > 1.) return 0.f;
    x *= M_PI;
    float sinc = sinf(x * tau) / (x * tau);
    float lanczos = sinf(x) / x;
    return sinc * lanczos;
  }
  float tau;
};

#define FILTER_TABLE_SIZE 16

class ImageFilm : public Film {
 public:
  ImageFilm(int xres, int yres, Filter *filt, const float crop[4], const string &fn, bool premult,
            int wf)
      : Film(xres, yres), filter(filt), filename(fn), premultiplyAlpha(premult) {
    memcpy(cropWindow, crop, 4 * sizeof(float));
    if (wf > 0)
      Warning("\"writefrequency\" is ignored: the image is written once, after the GPU render.");
    // film/image.cpp:79-85
    xPixelStart = Ceil2Int(xResolution * cropWindow[0]);
    xPixelCount = max(1, Ceil2Int(xResolution * cropWindow[1]) - xPixelStart);
    yPixelStart = Ceil2Int(yResolution * cropWindow[2]);
    yPixelCount = max(1, Ceil2Int(yResolution * cropWindow[3]) - yPixelStart);
    // film/image.cpp:89-101
    float *ftp = filterTable;
    for (int y = 0; y < FILTER_TABLE_SIZE; ++y) {
      float fy = ((float)y + .5f) * filter->yWidth / FILTER_TABLE_SIZE;
      for (int x = 0; x < FILTER_TABLE_SIZE; ++x) {
        float fx = ((float)x + .5f) * filter->xWidth / FILTER_TABLE_SIZE;
        *ftp++ = filter->Evaluate(fx, fy);
      }
    }
  }
  ~ImageFilm() { delete filter; }
  // film/image.cpp:148-156
  void GetSampleExtent(int *xstart, int *xend, int *ystart, int *yend) const {
    *xstart = Floor2Int(xPixelStart + .5f - filter->xWidth);
    *xend = Floor2Int(xPixelStart + .5f + xPixelCount + filter->xWidth);
    *ystart = Floor2Int(yPixelStart + .5f - filter->yWidth);
    *yend = Floor2Int(yPixelStart + .5f + yPixelCount + filter->yWidth);
  }
  void Flatten(PbrtB200Film *out) const {
    out->x_resolution = xResolution; out->y_resolution = yResolution;
    out->x_pixel_start = xPixelStart; out->y_pixel_start = yPixelStart;
    out->x_pixel_count = xPixelCount; out->y_pixel_count = yPixelCount;
    out->filter_x_width = filter->xWidth; out->filter_y_width = filter->yWidth;
    out->filter_inv_x_width = filter->invXWidth; out->filter_inv_y_width = filter->invYWidth;
    memcpy(out->filter_table, filterTable, sizeof(filterTable));
    out->premultiply_alpha = premultiplyAlpha ? 1 : 0;
  }
  void WriteImage(const float *rgba) {
    int nPix = xPixelCount * yPixelCount;
    vector<float> rgb(3 * nPix), alpha(nPix);
    for (int i = 0; i < nPix; ++i) {
      rgb[3 * i] = rgba[4 * i]; rgb[3 * i + 1] = rgba[4 * i + 1]; rgb[3 * i + 2] = rgba[4 * i + 2];
      alpha[i] = rgba[4 * i + 3];
    }
    WriteRGBAImage(filename, &rgb[0], &alpha[0], xPixelCount, yPixelCount, xResolution,
                   yResolution, xPixelStart, yPixelStart);
  }
 private:
  Filter *filter;
  string filename;
  bool premultiplyAlpha;
  float cropWindow[4];
  int xPixelStart, yPixelStart, xPixelCount, yPixelCount;
  float filterTable[FILTER_TABLE_SIZE * FILTER_TABLE_SIZE];
};

// IEEE binary32 -> binary16, round to nearest even (what half.h's converter
// produces for finite values; overflow goes to infinity).
uint16_t FloatToHalf(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
  uint32_t mant = x & 0x7fffffu;
  if (((x >> 23) & 0xff) == 0xff) {                 // inf / nan
    if (mant == 0) return (uint16_t)(sign | 0x7c00u);
    uint32_t m = mant >> 13;
    return (uint16_t)(sign | 0x7c00u | m | (m == 0));
  }
  if (exp <= 0) {                                    // half subnormal or zero
    if (exp < -10) return (uint16_t)sign;
    mant |= 0x800000u;
    int shift = 14 - exp;
    uint32_t half = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half & 1u))) ++half;
    return (uint16_t)(sign | half);
  }
  uint32_t half = ((uint32_t)exp << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1u))) ++half;   // may carry into exponent
  if (half >= 0x7c00u) return (uint16_t)(sign | 0x7c00u);
  return (uint16_t)(sign | half);
}

void PutAttr(vector<unsigned char> &o, const char *name, const char *type, const void *data,
             uint32_t size) {
  o.insert(o.end(), name, name + strlen(name) + 1);
  o.insert(o.end(), type, type + strlen(type) + 1);
  const unsigned char *s = (const unsigned char *)&size;
  o.insert(o.end(), s, s + 4);
  const unsigned char *d = (const unsigned char *)data;
  o.insert(o.end(), d, d + size);
}

}  // namespace

Filter *CreateNamedFilter(const string &name, const ParamSet &ps) {
  if (name == "box")
    return new BoxFilter(ps.FindOneFloat("xwidth", .5f), ps.FindOneFloat("ywidth", .5f));
  if (name == "triangle")
    return new TriangleFilter(ps.FindOneFloat("xwidth", 2.), ps.FindOneFloat("ywidth", 2.));
  if (name == "gaussian") {
    float xw = ps.FindOneFloat("xwidth", 2.), yw = ps.FindOneFloat("ywidth", 2.);
    return new GaussianFilter(xw, yw, ps.FindOneFloat("alpha", 2.f));
  }
  if (name == "mitchell") {
    float xw = ps.FindOneFloat("xwidth", 2.), yw = ps.FindOneFloat("ywidth", 2.);
    float B = ps.FindOneFloat("B", 1.f / 3.f), C = ps.FindOneFloat("C", 1.f / 3.f);
    return new MitchellFilter(B, C, xw, yw);
  }
  if (name == "sinc") {
    float xw = ps.FindOneFloat("xwidth", 4.), yw = ps.FindOneFloat("ywidth", 4.);
    return new LanczosSincFilter(xw, yw, ps.FindOneFloat("tau", 3.f));
  }
  Error("Pixel filter plugin \"%s\" unknown.", name.c_str());
  return NULL;
}

// film/image.cpp:213-232
Film *CreateImageFilm(const ParamSet &params, Filter *filter) {
  string filename = params.FindOneString("filename", "pbrt.exr");
  bool premultiplyAlpha = params.FindOneBool("premultiplyalpha", true);
  int xres = params.FindOneInt("xresolution", 640);
  int yres = params.FindOneInt("yresolution", 480);
  float crop[4] = {0, 1, 0, 1};
  int cwi;
  const float *cr = params.FindFloat("cropwindow", &cwi);
  if (cr && cwi == 4) {
    crop[0] = Clamp(min(cr[0], cr[1]), 0., 1.);
    crop[1] = Clamp(max(cr[0], cr[1]), 0., 1.);
    crop[2] = Clamp(min(cr[2], cr[3]), 0., 1.);
    crop[3] = Clamp(max(cr[2], cr[3]), 0., 1.);
  }
  int writeFrequency = params.FindOneInt("writefrequency", -1);
  return new ImageFilm(xres, yres, filter, crop, filename, premultiplyAlpha, writeFrequency);
}

// Scan-line OpenEXR 2 file, uncompressed, channels A,B,G,R of type HALF.
void WriteRGBAImage(const string &name, const float *pixels, const float *alpha, int xRes,
                    int yRes, int totalXRes, int totalYRes, int xOffset, int yOffset) {
  vector<unsigned char> hdr;
  const uint32_t magic = 20000630u, version = 2u;
  hdr.insert(hdr.end(), (const unsigned char *)&magic, (const unsigned char *)&magic + 4);
  hdr.insert(hdr.end(), (const unsigned char *)&version, (const unsigned char *)&version + 4);
  {
    vector<unsigned char> ch;
    const char *names[4] = {"A", "B", "G", "R"};
    for (int i = 0; i < 4; ++i) {
      ch.push_back((unsigned char)names[i][0]); ch.push_back(0);
      int32_t rec[4] = {1 /*HALF*/, 0 /*pLinear + reserved*/, 1, 1};
      const unsigned char *r = (const unsigned char *)rec;
      ch.insert(ch.end(), r, r + 16);
    }
    ch.push_back(0);
    PutAttr(hdr, "channels", "chlist", &ch[0], (uint32_t)ch.size());
  }
  unsigned char comp = 0;
  PutAttr(hdr, "compression", "compression", &comp, 1);
  int32_t dataW[4] = {xOffset, yOffset, xOffset + xRes - 1, yOffset + yRes - 1};
  PutAttr(hdr, "dataWindow", "box2i", dataW, 16);
  int32_t dispW[4] = {0, 0, totalXRes - 1, totalYRes - 1};
  PutAttr(hdr, "displayWindow", "box2i", dispW, 16);
  unsigned char lineOrder = 0;
  PutAttr(hdr, "lineOrder", "lineOrder", &lineOrder, 1);
  float par = 1.f;
  PutAttr(hdr, "pixelAspectRatio", "float", &par, 4);
  float swc[2] = {0.f, 0.f};
  PutAttr(hdr, "screenWindowCenter", "v2f", swc, 8);
  float sww = 1.f;
  PutAttr(hdr, "screenWindowWidth", "float", &sww, 4);
  hdr.push_back(0);

  FILE *f = fopen(name.c_str(), "wb");
  if (!f) {
    Error("Unable to write image file \"%s\": %s", name.c_str(), strerror(errno));
    return;
  }
  fwrite(&hdr[0], 1, hdr.size(), f);
  uint64_t lineBytes = 8 + (uint64_t)xRes * 4 * 2;
  uint64_t base = hdr.size() + (uint64_t)yRes * 8;
  for (int y = 0; y < yRes; ++y) {
    uint64_t off = base + (uint64_t)y * lineBytes;
    fwrite(&off, 8, 1, f);
  }
  vector<uint16_t> line((size_t)xRes * 4);
  for (int y = 0; y < yRes; ++y) {
    int32_t yy = yOffset + y, sz = xRes * 4 * 2;
    fwrite(&yy, 4, 1, f);
    fwrite(&sz, 4, 1, f);
    for (int x = 0; x < xRes; ++x) {
      int i = y * xRes + x;
      line[x] = FloatToHalf(alpha ? alpha[i] : 1.f);
      line[xRes + x] = FloatToHalf(pixels[3 * i + 2]);
      line[2 * xRes + x] = FloatToHalf(pixels[3 * i + 1]);
      line[3 * xRes + x] = FloatToHalf(pixels[3 * i]);
    }
    fwrite(&line[0], 2, line.size(), f);
  }
  fclose(f);
}
