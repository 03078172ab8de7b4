/*
 * keyed_rng.h -- the per-sample keyed counter RNG ("O-keyed", SURVEY.md 8c /
 * App. H).  NORMATIVE statement of the generator: the CUDA kernels
 * (pbrt-v1_b200/csrc/rng.cuh) implement exactly this, integer for integer.
 *
 * TEST INFRASTRUCTURE: this header belongs to the oracle; only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline leg may use oracle/.
 *
 * The reference draws every random number from one global sequential
 * MT19937 (core/util.cpp:318-380); the number of draws per image sample is
 * data dependent, so a single diverged path shifts the stream of every later
 * sample and per-pixel comparison between two implementations is meaningless.
 * The keyed variant re-seeds a counter-based generator at fixed points:
 *
 *   stream 0, s = 0xFFFF : a pixel's camera pattern (image/lens/time strata
 *                          and their shuffles, samplers/stratified.cpp:66-85,98-115)
 *                          (lowdiscrepancy sampler: ALL of the pixel's tables -- image, lens, time
 *                          and every integrator pattern -- samplers/lowdiscrepancy.cpp:93-103)
 *   stream 0, s = k      : the LatinHypercube block of the pixel's k-th sample
 *                          (samplers/stratified.cpp:124-128)
 *   stream 1, s = k      : every draw made while evaluating Li for that sample
 *                          (core/scene.cpp:57 onward)
 *
 * bestcandidate sampler (samplers/bestcandidate.cpp): its unit is a placement of the 4096-entry
 * sample table, identified by the integer pixel of its corner (cx, cy); entry `off` of it is a sample:
 *   stream 3, (cx, cy, 0)        : the placement's three sample shifts and its LDShuffleScrambled tables
 *   stream 2, bc key (cx,cy,off) : the sample's StratifiedSample1D/2D patterns (patterns of size != 1)
 *   stream 1, bc key (cx,cy,off) : every draw made while evaluating Li for that sample
 * where bc key = key(cx ^ ((off >> 10) << 20), cy, off & 1023, stream) keeps the 12-bit entry number
 * out of the bits y occupies.
 *
 * draw(ctr) = high 32 bits of splitmix64-finalize(key + ctr * golden), ctr = 1,2,...
 * RandomFloat keeps the reference's mapping (u32 & 0xffffff) / 2^24
 * (core/util.cpp:377-380).
 */
#ifndef PBRT_ORACLE_KEYED_RNG_H
#define PBRT_ORACLE_KEYED_RNG_H
#include <stdint.h>

static inline uint64_t pbrt_keyed_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline uint64_t pbrt_keyed_key(int x, int y, int s, int stream) {
  uint64_t k = ((uint64_t)(uint32_t)x << 32) ^ ((uint64_t)(uint32_t)y << 12) ^
               ((uint64_t)(uint32_t)s << 2) ^ (uint64_t)(uint32_t)stream;
  return pbrt_keyed_mix64(k);
}
static inline uint64_t pbrt_keyed_key_bc(int cx, int cy, int off, int stream) {
  return pbrt_keyed_key(cx ^ ((off >> 10) << 20), cy, off & 1023, stream);
}
/* ctr is 1-based: the first draw after a reseed uses ctr = 1 */
static inline uint32_t pbrt_keyed_u32(uint64_t key, uint64_t ctr) {
  return (uint32_t)(pbrt_keyed_mix64(key + ctr * 0x9E3779B97F4A7C15ULL) >> 32);
}
static inline float pbrt_keyed_float(uint64_t key, uint64_t ctr) {
  return (float)(pbrt_keyed_u32(key, ctr) & 0xffffff) / (float)(1 << 24);
}
#endif
