// pb_rng.h -- the per-sample keyed counter RNG and the stratified sampler's
// patterns evaluated by random access.
//
// Generator: SURVEY.md App. H (normative statement: oracle/keyed_rng.h).  It
// replaces the reference's global MT19937 (core/util.cpp:318-380) by
// draw(key, ctr) = hi32(splitmix64(key + ctr * golden)), re-keyed
//   (x, y, 0xFFFF, 0)  before a pixel's camera pattern  (samplers/stratified.cpp:66-85,98-115)
//   (x, y, k, 0)       before sample k's LatinHypercube block (stratified.cpp:124-128)
//   (x, y, k, 1)       before Li of sample k             (core/scene.cpp:57)
// Because the generator is counter based, draw number n of any stream can be
// evaluated directly, so one thread per image sample can reproduce the value
// the reference's *sequential* sampler code hands to that sample: the
// stratified jitters are plain indexed draws, and a shuffled entry is found
// by walking the swap sequence backwards (core/sampling.cpp:91-113).
#ifndef PB_RNG_H
#define PB_RNG_H
#include "pb_math.h"

#ifdef PB_MIX_NOINLINE   // experiment: one out-of-line copy of the mixer instead of ~20 inlined ones per shading kernel
PB_NOINLINE
#else
PB_HD
#endif
uint64_t pb_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
PB_HD uint64_t pb_key(int x, int y, int s, int stream) {
  uint64_t k = ((uint64_t)(uint32_t)x << 32) ^ ((uint64_t)(uint32_t)y << 12) ^
               ((uint64_t)(uint32_t)s << 2) ^ (uint64_t)(uint32_t)stream;
  return pb_mix64(k);
}
// bestcandidate sampler: key of entry `off` of the table placement whose corner pixel is (cx, cy)
// (oracle/keyed_rng.h: the 12-bit entry number is kept out of the bits y occupies)
PB_HD uint64_t pb_key_bc(int cx, int cy, int off, int stream) {
  return pb_key(cx ^ ((off >> 10) << 20), cy, off & 1023, stream);
}
// ctr is 1-based (first draw after a re-key has ctr == 1)
PB_HD uint32_t pb_u32(uint64_t key, uint32_t ctr) {
  return (uint32_t)(pb_mix64(key + (uint64_t)ctr * 0x9E3779B97F4A7C15ULL) >> 32);
}
// RandomFloat(): (u32 & 0xffffff) / 2^24  (core/util.cpp:377-380)
PB_HD float pb_f32(uint64_t key, uint32_t ctr) {
  return (float)(pb_u32(key, ctr) & 0xffffff) / (float)(1 << 24);
}

// Sequential stream used inside Li (stream 1).
struct RngStream {
  uint64_t key;
  uint32_t ctr;   // draws made so far
  PB_HD float next_float() { return pb_f32(key, ++ctr); }
  PB_HD void skip(uint32_t n) { ctr += n; }
};

// Shuffle(samp, count, dims) (core/sampling.cpp:91-97) draws `count` uints
// starting at ctr0+1: for i in [0,count): swap(i, u_i % count).  Returns the
// index (before the shuffle) of the entry that ends at position `pos`.
PB_HD int pb_shuffle_source(uint64_t key, uint32_t ctr0, int count, int pos) {
  for (int i = count - 1; i >= 0; --i) {
    int other = (int)(pb_u32(key, ctr0 + 1 + (uint32_t)i) % (uint32_t)count);
    if (pos == i) pos = other;
    else if (pos == other) pos = i;
  }
  return pos;
}

// One integrator sample pattern requested through Sample::Add1D/Add2D
// (core/sampling.h:48-73).  `ctr0` = draws consumed by the patterns generated
// before it in the same LatinHypercube block (1D patterns first, then 2D:
// samplers/stratified.cpp:125-128).
struct SamplePattern {
  uint32_t ctr0;
  int32_t n;      // samples in the pattern
  int32_t dims;   // 1 or 2
  int32_t offset; // lowdiscrepancy: first float of the pattern in a sample's value record
};
// Where an image sample's integrator patterns come from: computed from the key (stratified) or read
// from the pixel's pre-generated (0,2)-sequence tables (lowdiscrepancy): value e of the sample is
// ld[e * stride + index].
// bestcandidate: one-sample patterns are read from the placement's LDShuffleScrambled tables (same
// record layout), larger ones are StratifiedSample1D/2D draws of the sample's stream 2 (key0).
struct PatternSource {
  uint64_t key0;
  const float *ld;
  uint32_t stride, index;
  int32_t bc;   // 0 LatinHypercube patterns (stratified) or tables (lowdiscrepancy), 1 bestcandidate, 2 random sampler
};
// LatinHypercube(samples, n, dims) (core/sampling.cpp:98-113): n*dims floats
// (sample-major), then per dimension n swap draws.  Value of sample j, dim d.
PB_NOINLINE float pb_lhs_value_general(uint64_t key, const SamplePattern &p, int j, int d) {
  float delta = 1.f / p.n;
  uint32_t c = p.ctr0 + (uint32_t)(p.n * p.dims) + (uint32_t)(d * p.n);
  int src = pb_shuffle_source(key, c, p.n, j);
  float r = pb_f32(key, p.ctr0 + 1 + (uint32_t)(p.dims * src + d));
  return (src + r) * delta;
}
PB_HD float pb_lhs_value(uint64_t key, const SamplePattern &p, int j, int d) {
  // one-sample patterns (every path-integrator pattern): (0 + r) * (1.f / 1) == r exactly, and the
  // single swap draw exchanges element 0 with itself
  if (p.n == 1) return pb_f32(key, p.ctr0 + 1 + (uint32_t)d);
  return pb_lhs_value_general(key, p, j, d);
}

// bestcandidate sampler, patterns of size != 1 (samplers/bestcandidate.cpp:152-168):
// StratifiedSample1D(n) / StratifiedSample2D(root, root) with jitter (core/sampling.cpp:72-90)
PB_HD float pb_bc_stratified_value(uint64_t key2, const SamplePattern &p, int j, int d) {
  if (p.dims == 1) {
    float invTot = 1.f / p.n;
    return (j + pb_f32(key2, p.ctr0 + 1u + (uint32_t)j)) * invTot;
  }
  int root = 1;
  while (root * root < p.n) ++root;   // n is a perfect square (BestCandidateSampler::RoundSize)
  float dxy = 1.f / root;
  int c = d == 0 ? j % root : j / root;
  return (c + pb_f32(key2, p.ctr0 + 1u + 2u * (uint32_t)j + (uint32_t)d)) * dxy;
}

// ---- lowdiscrepancy sampler (samplers/lowdiscrepancy.cpp, core/sampling.h:132-174)
#define PB_LD_MAX_PIXEL_SAMPLES 1024
#define PB_LD_MAX_PATTERN_SAMPLES 64
// core/sampling.h:137-146
PB_HD float pb_van_der_corput(uint32_t n, uint32_t scramble) {
  n = (n << 16) | (n >> 16);
  n = ((n & 0x00ff00ffu) << 8) | ((n & 0xff00ff00u) >> 8);
  n = ((n & 0x0f0f0f0fu) << 4) | ((n & 0xf0f0f0f0u) >> 4);
  n = ((n & 0x33333333u) << 2) | ((n & 0xccccccccu) >> 2);
  n = ((n & 0x55555555u) << 1) | ((n & 0xaaaaaaaau) >> 1);
  n ^= scramble;
  return (float)n / 4294967296.f;
}
// core/sampling.h:147-151
PB_HD float pb_sobol2(uint32_t n, uint32_t scramble) {
  for (uint32_t v = 1u << 31; n != 0; n >>= 1, v ^= v >> 1)
    if (n & 0x1) scramble ^= v;
  return (float)scramble / 4294967296.f;
}
// One LDShuffleScrambled1D / 2D table (core/sampling.h:158-174) for a pixel of nPixel samples,
// pattern size pat.n: scramble draws, the (0,2)-sequence points, a Shuffle inside every block of
// pat.n points and a Shuffle of the nPixel blocks.  Point e (0 <= e < pat.n * pat.dims) of the
// pixel's sample k is written to out[(pat.offset + e) * stride + base + k] -- skipped when
// pat.offset < 0 (the time table: its draws count, its values are never read).
PB_HD void pb_ld_generate_table(uint64_t key, const SamplePattern &pat, int nPixel, float *out,
                                uint32_t stride, uint32_t base) {
  if (pat.offset < 0) return;
  uint16_t blockSrc[PB_LD_MAX_PIXEL_SAMPLES];
  uint8_t inner[PB_LD_MAX_PATTERN_SAMPLES];
  int n = pat.n, dims = pat.dims;
  uint32_t ctr = pat.ctr0;
  uint32_t scr0 = pb_u32(key, ++ctr);
  uint32_t scr1 = dims == 2 ? pb_u32(key, ++ctr) : 0u;
  uint32_t ctrInner = ctr;                              // block i's Shuffle draws: ctrInner + i*n + 1 ...
  uint32_t ctrBlock = ctr + (uint32_t)(nPixel * n);     // the block Shuffle draws follow all of them
  for (int i = 0; i < nPixel; ++i) blockSrc[i] = (uint16_t)i;
  for (int i = 0; i < nPixel; ++i) {
    uint32_t other = pb_u32(key, ctrBlock + 1 + (uint32_t)i) % (uint32_t)nPixel;
    uint16_t t = blockSrc[i]; blockSrc[i] = blockSrc[other]; blockSrc[other] = t;
  }
  for (int k = 0; k < nPixel; ++k) {
    int b = blockSrc[k];
    for (int j = 0; j < n; ++j) inner[j] = (uint8_t)j;
    for (int j = 0; j < n; ++j) {
      uint32_t other = pb_u32(key, ctrInner + (uint32_t)(b * n) + 1 + (uint32_t)j) % (uint32_t)n;
      uint8_t t = inner[j]; inner[j] = inner[other]; inner[other] = t;
    }
    for (int j = 0; j < n; ++j) {
      uint32_t i = (uint32_t)(b * n + inner[j]);
      float *o = out + (size_t)(pat.offset + j * dims) * stride + base + k;
      o[0] = pb_van_der_corput(i, scr0);
      if (dims == 2) o[stride] = pb_sobol2(i, scr1);
    }
  }
}

// ---- bestcandidate sampler tables (samplers/bestcandidate.cpp:96-105): per table placement one
// LDShuffleScrambled1D/2D(4096, 1) table per one-sample pattern = scramble draw(s), the (0,2)-sequence
// points 0..4095, a Shuffle of the 4096 entries and a final one-element Shuffle (one more draw).
#define PB_BC_TABLE 4096
// The Shuffle as a permutation: perm[o] = index of the point that ends at entry o.
// ctrShuffle = draws made before the Shuffle's first.
PB_HD void pb_bc_table_perm(uint64_t key, uint32_t ctrShuffle, uint16_t *perm) {
  for (int i = 0; i < PB_BC_TABLE; ++i) perm[i] = (uint16_t)i;
  for (int i = 0; i < PB_BC_TABLE; ++i) {
    uint32_t other = pb_u32(key, ctrShuffle + 1u + (uint32_t)i) % (uint32_t)PB_BC_TABLE;
    uint16_t t = perm[i]; perm[i] = perm[other]; perm[other] = t;
  }
}
#endif
