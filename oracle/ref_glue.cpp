// ref_glue.cpp -- the pieces linked into the compiled reference (oracle/_ref)
// in place of what cannot be built here.  Compiled against the REFERENCE's
// headers by oracle/build_ref.sh; never part of the product.
//
//  * WriteRGBAImage / ReadImage: core/exrio.cpp needs OpenEXR (absent).  The
//    replacement dumps exactly the float arrays ImageFilm::WriteImage hands
//    over (film/image.cpp:204-207) to "<name>.raw": a 32-byte header
//    ("PBRTRAW1", xRes, yRes, totalX, totalY, xOff, yOff as int32) followed by
//    xRes*yRes RGBA float32 quadruples.
//  * keyed RNG (only when built with -DPBRT_REF_KEYED): genrand_int32 /
//    genrand_real2 replacements + the reseed hooks that build_ref.sh splices
//    into copies of samplers/stratified.cpp and core/scene.cpp.
//  * optional per-sample dump (env PBRT_REF_DUMP_SAMPLES=<file>): imageX,
//    imageY, L.rgb (after the guards of scene.cpp:60-74), alpha per sample.
#include "pbrt.h"
#include "color.h"
#include "sampling.h"
#include "keyed_rng.h"
#include <errno.h>
#include <math.h>
#include <string.h>
#include <string>
#include <vector>

COREDLL void WriteRGBAImage(const string &name, float *pixels, float *alpha, int xRes, int yRes,
                            int totalXRes, int totalYRes, int xOffset, int yOffset) {
  string fn = name + ".raw";
  FILE *f = fopen(fn.c_str(), "wb");
  if (!f) { Error("Unable to write image file \"%s\"", fn.c_str()); return; }
  int hdr[6] = {xRes, yRes, totalXRes, totalYRes, xOffset, yOffset};
  fwrite("PBRTRAW1", 1, 8, f);
  fwrite(hdr, 4, 6, f);
  for (int i = 0; i < xRes * yRes; ++i) {
    float px[4] = {pixels[3 * i], pixels[3 * i + 1], pixels[3 * i + 2], alpha ? alpha[i] : 1.f};
    fwrite(px, 4, 4, f);
  }
  fclose(f);
}
// ReadImage without OpenEXR: uncompressed scan-line files with HALF (or FLOAT, narrowed to half as the
// reference's HALF frame buffer does) channels -- the files tests/parity_util.py::write_exr produces.
static float RefHalf(unsigned short h) {
  int e = (h >> 10) & 31, m = h & 1023;
  double v = e == 0 ? ldexp((double)m, -24) : e == 31 ? (m ? NAN : INFINITY) : ldexp((double)(m + 1024), e - 25);
  return (float)((h & 0x8000) ? -v : v);
}
static unsigned short RefToHalf(float f) {   // nearest even; texture fixtures are HALF, this path is for FLOAT files
  if (f != f) return 0x7e00;
  unsigned short sign = (f < 0 || (f == 0 && 1 / f < 0)) ? 0x8000 : 0;
  double a = fabs((double)f);
  if (a >= 65520.0) return sign | 0x7c00;
  if (a < ldexp(1.0, -25)) return sign;
  int e;
  frexp(a, &e);                       // a = m * 2^e, m in [0.5, 1)
  int he = e + 14;                    // biased half exponent of the leading bit
  double scaled = he <= 0 ? ldexp(a, 24) : ldexp(a, 11 - e);   // subnormal: units of 2^-24; normal: 11-bit significand
  double r = nearbyint(scaled);       // default rounding mode: to nearest, ties to even
  unsigned int bits = he <= 0 ? (unsigned int)r : (unsigned int)(((he - 1) << 10) + (unsigned int)r);
  return sign | (unsigned short)bits;
}
COREDLL Spectrum *ReadImage(const string &name, int *width, int *height) {
  FILE *f = fopen(name.c_str(), "rb");
  if (!f) {
    Error("Unable to read image file \"%s\": %s", name.c_str(), strerror(errno));
    return NULL;
  }
  std::vector<unsigned char> d;
  int ch;
  while ((ch = fgetc(f)) != EOF) d.push_back((unsigned char)ch);
  fclose(f);
  struct Chan { std::string name; int type; };
  std::vector<Chan> chans;
  int win[4] = {0, 0, -1, -1}, comp = -1;
  size_t p = 8;
  bool ok = d.size() > 8 && d[0] == 0x76 && d[1] == 0x2f && d[2] == 0x31 && d[3] == 0x01;
  while (ok) {
    std::string an, at;
    while (p < d.size() && d[p]) an += (char)d[p++];
    ++p;
    if (an.empty()) break;
    while (p < d.size() && d[p]) at += (char)d[p++];
    ++p;
    if (p + 4 > d.size()) { ok = false; break; }
    int sz;
    memcpy(&sz, &d[p], 4);
    p += 4;
    if (sz < 0 || p + sz > d.size()) { ok = false; break; }
    if (an == "channels") {
      size_t q = p;
      while (d[q]) {
        Chan c;
        while (d[q]) c.name += (char)d[q++];
        ++q;
        memcpy(&c.type, &d[q], 4);
        q += 16;
        chans.push_back(c);
      }
    } else if (an == "compression") comp = d[p];
    else if (an == "dataWindow") memcpy(win, &d[p], 16);
    p += sz;
  }
  if (!ok || comp != 0 || win[2] < win[0] || win[3] < win[1]) {
    Error("Unable to read image file \"%s\": not an uncompressed scan-line OpenEXR file", name.c_str());
    return NULL;
  }
  int w = win[2] - win[0] + 1, h = win[3] - win[1] + 1;
  std::vector<float> rgb(3 * w * h, 0.f);
  for (int row = 0; row < h; ++row) {
    unsigned long long off;
    memcpy(&off, &d[p + 8 * row], 8);
    int y;
    memcpy(&y, &d[off], 4);
    const unsigned char *src = &d[off + 8];
    for (size_t c = 0; c < chans.size(); ++c) {
      int k = chans[c].name == "R" ? 0 : chans[c].name == "G" ? 1 : chans[c].name == "B" ? 2 : -1;
      int bytes = chans[c].type == 1 ? 2 : 4;
      for (int x = 0; x < w && k >= 0; ++x) {
        unsigned short hb;
        if (bytes == 2) memcpy(&hb, src + 2 * x, 2);
        else { float v; memcpy(&v, src + 4 * x, 4); hb = RefToHalf(v); }
        rgb[3 * ((y - win[1]) * w + x) + k] = RefHalf(hb);
      }
      src += bytes * w;
    }
  }
  Spectrum *ret = new Spectrum[w * h];
  for (int i = 0; i < w * h; ++i) ret[i] = Spectrum(&rgb[3 * i]);   // exrio.cpp:63-66
  *width = w;
  *height = h;
  return ret;
}

// incremented by the kd-tree plugin built from oracle/build_ref.sh's counted copy
COREDLL unsigned long long gRefIntersectCalls = 0, gRefIntersectPCalls = 0;

static FILE *gDump = NULL;
static unsigned long long gSampleCounter = 0;
static double gRenderSeconds = 0;

#include <sys/time.h>
static double Now() {
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}
static double gRenderStart = 0;

COREDLL void pbrtRefBeginRender() {
  gSampleCounter = 0;
  gRefIntersectCalls = gRefIntersectPCalls = 0;
  const char *fn = getenv("PBRT_REF_DUMP_SAMPLES");
  gDump = fn ? fopen(fn, "wb") : NULL;
  gRenderStart = Now();
}
COREDLL void pbrtRefEndRender() {
  gRenderSeconds = Now() - gRenderStart;
  if (gDump) fclose(gDump);
  gDump = NULL;
  // timing of the sample loop alone (parse, kd build and image write excluded)
  fprintf(stderr, "PBRT_REF_RENDER_SECONDS %.6f SAMPLES %llu\n", gRenderSeconds, gSampleCounter);
  fprintf(stderr, "PBRT_REF_RAYS INTERSECT %llu INTERSECTP %llu\n", gRefIntersectCalls, gRefIntersectPCalls);
}
// Called by the counted kdtree.cpp copy when a KdTreeAccel has been built (see build_ref.sh).  Prints the
// exact node / leaf-reference counts (the reference's own statistics round them to 3-4 digits,
// core/util.cpp:263-276) and, with PBRT_REF_DUMP_KD=<file>, appends the tree in the flat form of
// include/pbrt_b200.h: per node {word0 = the reference's flags|split|nPrims word, word1 = above child |
// primitive number | offset into the reference list}, then the leaves' reference lists in node order.
COREDLL void pbrtRefKdBuilt(const void *nodes, int nNodes, const void *mailboxBase, int mailboxStride, int nPrims,
                            const float *bounds) {
  const unsigned char *nb = (const unsigned char *)nodes;
  unsigned long long nRefs = 0, nLeaves = 0;
  for (int i = 0; i < nNodes; ++i) {
    unsigned int w0;
    memcpy(&w0, nb + 16 * (size_t)i, 4);
    if ((w0 & 3u) == 3u) { ++nLeaves; if ((w0 >> 2) > 1) nRefs += w0 >> 2; }
  }
  fprintf(stderr, "PBRT_REF_KD NODES %d LEAVES %llu REFS %llu PRIMS %d\n", nNodes, nLeaves, nRefs, nPrims);
  const char *fn = getenv("PBRT_REF_DUMP_KD");
  if (!fn) return;
  FILE *f = fopen(fn, "ab");
  if (!f) return;
  unsigned int hdr[4] = {0x4b445452u, (unsigned int)nNodes, (unsigned int)nRefs, (unsigned int)nPrims};
  fwrite(hdr, 4, 4, f);
  fwrite(bounds, 4, 6, f);
  unsigned int *words = new unsigned int[2 * (size_t)nNodes];
  unsigned int *refs = new unsigned int[nRefs ? nRefs : 1];
  unsigned long long off = 0;
  for (int i = 0; i < nNodes; ++i) {
    unsigned int w0;
    memcpy(&w0, nb + 16 * (size_t)i, 4);
    words[2 * i] = w0;
    if ((w0 & 3u) != 3u) {
      memcpy(&words[2 * i + 1], nb + 16 * (size_t)i + 8, 4);   // aboveChild
      continue;
    }
    unsigned int np = w0 >> 2;
    const unsigned char *ptr;
    memcpy(&ptr, nb + 16 * (size_t)i + 8, sizeof(ptr));
    if (np == 0) words[2 * i + 1] = 0;
    else if (np == 1) words[2 * i + 1] = (unsigned int)((ptr - (const unsigned char *)mailboxBase) / mailboxStride);
    else {
      words[2 * i + 1] = (unsigned int)off;
      const unsigned char *const *list = (const unsigned char *const *)ptr;
      for (unsigned int k = 0; k < np; ++k)
        refs[off++] = (unsigned int)((list[k] - (const unsigned char *)mailboxBase) / mailboxStride);
    }
  }
  fwrite(words, 4, 2 * (size_t)nNodes, f);
  fwrite(refs, 4, nRefs, f);
  delete[] words;
  delete[] refs;
  fclose(f);
}
COREDLL void pbrtRefDumpSample(const Sample *sample, const Spectrum &Ls, float alpha) {
  if (!gDump) return;
  float xyz[3];
  // Spectrum keeps c[] private; recover RGB exactly through operator== is not
  // possible, so rely on the layout (3 floats first): core/color.h:199.
  const float *c = (const float *)&Ls;
  float rec[6] = {sample->imageX, sample->imageY, c[0], c[1], c[2], alpha};
  (void)xyz;
  fwrite(rec, 4, 6, gDump);
}

#ifdef PBRT_REF_KEYED
static uint64_t gKey = 0, gCtr = 0;
COREDLL void pbrtReseed(int x, int y, int s, int stream) {
  gKey = pbrt_keyed_key(x, y, s, stream);
  gCtr = 0;
}
COREDLL unsigned long genrand_int32(void) { return (unsigned long)pbrt_keyed_u32(gKey, ++gCtr); }
COREDLL float genrand_real2(void) { return (RandomUInt() & 0xffffff) / float(1 << 24); }
// Called right before Li for every image sample (core/scene.cpp:57).  The
// (pixel, sample-in-pixel) pair is derived from the running sample count the
// same way the sampler walks pixels (samplers/stratified.cpp:89-97).
// bestcandidate sampler: the sampler announces each sample (table placement corner, table entry) right
// before it generates the sample's integrator patterns; that also keys the following Li.
static bool gBcSample = false;
static int gBcX = 0, gBcY = 0, gBcOff = 0;
COREDLL void pbrtRefBestCandidateSample(int cx, int cy, int off) {
  gBcSample = true; gBcX = cx; gBcY = cy; gBcOff = off;
  gKey = pbrt_keyed_key_bc(cx, cy, off, 2);
  gCtr = 0;
}
COREDLL void pbrtRefBeforeLi(const Sampler *sampler) {
  if (gBcSample) {
    gKey = pbrt_keyed_key_bc(gBcX, gBcY, gBcOff, 1);
    gCtr = 0;
    ++gSampleCounter;
    return;
  }
  int spp = sampler->samplesPerPixel;
  int w = sampler->xPixelEnd - sampler->xPixelStart;
  unsigned long long pix = gSampleCounter / spp;
  pbrtReseed(sampler->xPixelStart + (int)(pix % w), sampler->yPixelStart + (int)(pix / w),
             (int)(gSampleCounter % spp), 1);
  ++gSampleCounter;
}
#else
COREDLL void pbrtReseed(int, int, int, int) {}
COREDLL void pbrtRefBestCandidateSample(int, int, int) {}
COREDLL void pbrtRefBeforeLi(const Sampler *) { ++gSampleCounter; }
#endif
