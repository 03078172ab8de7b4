// gen_bestcandidate_table.cpp -- generates the 4096 x 5 sample table of the "bestcandidate" sampler.
//
// pbrt-v1 ships this table as data (samplers/sampledata.cpp, "Automatically generated 64x64 sample
// table (Mar 11 2003)"); the generator it came from is tools/samplepat.cpp.  This program restates
// that procedure -- best-candidate dart throwing on the unit torus for the image positions
// (samplepat.cpp:127-178), stratified time values reordered so that nearby image samples get distant
// times (:57-96), best-candidate lens positions redistributed the same way (:179-226) -- and
// reproduces the shipped table value for value when the random numbers are the standard MT19937
// genrand_real2 (u32 / 2^32, seed 5489) that the 2003 build evidently used (today's
// core/util.cpp:377-380 masks to 24 bits and yields a different table).  The table is written the
// way the reference consumes it: each value printed with %10.10f and read back as a float.
//
// Output: raw little-endian float32[4096][5] (imageX, imageY, time, lensU, lensV).  Takes ~13 minutes
// on one core (500 * n candidates for the n-th dart, twice); the result is committed as
// pbrt-v1_b200/host/bestcandidate_table.bin and checked against the reference's literals by
// tests/test_oracle.py::test_bestcandidate_table_matches_reference.
//
// build: g++ -O2 -msse2 -mfpmath=sse -ffp-contract=off tools/gen_bestcandidate_table.cpp -o /tmp/gen_bc
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace {
// MT19937 (Matsumoto & Nishimura 2002), init_genrand(5489)
struct MT {
  uint32_t mt[624];
  int mti;
  explicit MT(uint32_t s) {
    mt[0] = s;
    for (mti = 1; mti < 624; ++mti) mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
  }
  uint32_t u32() {
    if (mti >= 624) {
      for (int k = 0; k < 624; ++k) {
        uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  float real2() { return (float)(u32() * (1.0 / 4294967296.0)); }
};

const int N = 4096;
const int G = 40;   // BC_GRID_SIZE, core/pbrt.h:154
typedef std::vector<int> Grid[G][G];
inline int cell(float v) { return (int)(v * G); }
inline float wrapDist(float a, float b) {
  float d = fabsf(a - b);
  if (d < 0.5f) return d;
  return 1.f - std::max(a, b) + std::min(a, b);
}
template <class F>
inline void neighbours(const Grid &grid, int gu, int gv, F f) {
  for (int du = -1; du <= 1; ++du)
    for (int dv = -1; dv <= 1; ++dv) {
      int u = gu + du, v = gv + dv;
      if (u < 0) u += G;
      if (u >= G) u -= G;
      if (v < 0) v += G;
      if (v >= G) v -= G;
      const std::vector<int> &c = grid[u][v];
      for (size_t g = 0; g < c.size(); ++g) f(c[g]);
    }
}
void darts(MT &rng, float (*table)[2], Grid *grid) {
  table[0][0] = rng.real2();
  table[0][1] = rng.real2();
  (*grid)[cell(table[0][0])][cell(table[0][1])].push_back(0);
  for (int cur = 1; cur < N; ++cur) {
    float maxDist2 = 0.f;
    const long nCand = 500L * cur;
    for (long c = 0; c < nCand; ++c) {
      float cx = rng.real2(), cy = rng.real2();
      float dist2 = INFINITY;
      neighbours(*grid, cell(cx), cell(cy), [&](int s) {
        float xd = wrapDist(cx, table[s][0]), yd = wrapDist(cy, table[s][1]);
        dist2 = std::min(dist2, xd * xd + yd * yd);
      });
      if (dist2 > maxDist2) { maxDist2 = dist2; table[cur][0] = cx; table[cur][1] = cy; }
    }
    (*grid)[cell(table[cur][0])][cell(table[cur][1])].push_back(cur);
    if ((cur & 255) == 0) fprintf(stderr, "  dart %d / %d\n", cur, N);
  }
}
}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s out.bin\n", argv[0]); return 1; }
  static float image[N][2], lens[N][2], timeS[N];
  static Grid pixelGrid, lensGrid;
  MT rng(5489u);
  darts(rng, image, &pixelGrid);
  // the 2003 (x87) build never rounded this expression's RandomFloat() to float: (i + u32 / 2^32) / 4096
  // is exact in extended precision and rounds once, when it is stored
  for (int i = 0; i < N; ++i) timeS[i] = (float)(((double)i + rng.u32() * (1.0 / 4294967296.0)) / (double)N);
  for (int cur = 1; cur < N; ++cur) {
    int best = -1;
    float maxMin = 0.f;
    for (int t = cur; t < N; ++t) {
      float minDelta = INFINITY;
      neighbours(pixelGrid, cell(image[cur][0]), cell(image[cur][1]), [&](int o) {
        if (o < cur) minDelta = std::min(minDelta, wrapDist(timeS[o], timeS[t]));
      });
      if (minDelta > maxMin) { maxMin = minDelta; best = t; }
    }
    if (best < 0) { fprintf(stderr, "time pass: no candidate\n"); return 2; }
    std::swap(timeS[best], timeS[cur]);
  }
  darts(rng, lens, &lensGrid);
  for (int cur = 1; cur < N; ++cur) {
    int best = -1;
    float maxMin = 0.f;
    for (int s = cur; s < N; ++s) {
      float minD2 = INFINITY;
      neighbours(pixelGrid, cell(image[cur][0]), cell(image[cur][1]), [&](int o) {
        if (o < cur) {
          float dx = wrapDist(lens[o][0], lens[s][0]), dy = wrapDist(lens[o][1], lens[s][1]);
          minD2 = std::min(dx * dx + dy * dy, minD2);
        }
      });
      if (minD2 > maxMin) { maxMin = minD2; best = s; }
    }
    if (best < 0) { fprintf(stderr, "lens pass: no candidate\n"); return 2; }
    std::swap(lens[best][0], lens[cur][0]);
    std::swap(lens[best][1], lens[cur][1]);
  }
  FILE *f = fopen(argv[1], "wb");
  if (!f) { perror(argv[1]); return 1; }
  for (int i = 0; i < N; ++i) {
    float row[5] = {image[i][0], image[i][1], timeS[i], lens[i][0], lens[i][1]};
    for (int k = 0; k < 5; ++k) {   // through the decimal form the reference compiles
      char buf[64];
      snprintf(buf, sizeof(buf), "%10.10f", row[k]);
      row[k] = strtof(buf, NULL);
    }
    fwrite(row, sizeof(float), 5, f);
  }
  fclose(f);
  return 0;
}
