// render.cu -- libpbrt_b200.so: the CUDA render hot path behind the C ABI of
// include/pbrt_b200.h (the drop-in for Scene::Render, core/scene.cpp:32-88).
//
// Execution model: a *wavefront*.  The image samples of a pixel window are
// processed in waves of W consecutive samples (reference order: pixel-major,
// sample-minor).  Each wave runs
//
//   generate   K0  sample + camera ray for every slot          -> ray queue
//   repeat per path vertex:
//     intersect  K1  kd-tree closest hit for the ray queue      -> hit records
//     shade      K2  integrator vertex: BSDF, light sampling    -> shadow queue, probe queue,
//                                                                  next ray queue (compacted)
//     shadow     K3  kd-tree any-hit for the shadow queue       -> 1 bit per record
//     probe      K4  closest hit + "is it the sampled light"    -> 1 bit per record
//     [advance      Whitted/DirectLighting only: fold direct light, walk the specular
//                   recursion depth-first, emit the next ray]
//   finalize   K5a fold the last pending terms, sample guards   -> per-sample radiance
//   film       K5b gather-accumulate into the pixels, reference order
//   resolve    K6  RGB->XYZ->RGB, normalise, clamp, premultiply -> float4 film
//
// Queues are compacted SoA arrays filled with warp-aggregated atomics
// (ballot + one atomicAdd per warp).  Kernels are persistent: a fixed grid of
// numSMs x blocksPerSM CTAs whose warps fetch 32-entry chunks from an atomic
// work counter, so long and short rays balance across the chip.  Nothing in
// here runs on the CPU; if no CUDA device is present every entry point fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>

#define PB_BLOCK 128   // threads per CTA of every persistent kernel
#include "pb_integrator.h"
#include "pb_kd_warp.cuh"
#include "pb_camera.h"
#include "pb_film.h"
#include "pb_setup.h"
#include "pb_noise_table.h"
#include "pb_treelet.h"

// ------------------------------------------------------------------------------------------
// error plumbing
static thread_local std::string g_lastError;
static int fail(int code, const std::string &msg) {
  g_lastError = msg;
  return code;
}
#define CU_TRY(expr)                                                                     \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(PBRT_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// tunables (pbrt_b200_tunable)
static int64_t g_waveSamples = 1 << 27;   // key 0: image samples per wave, path integrator (128 Mi: up to ~43 GB of wave state on a 180 GB part, never more than half of the free memory; C5 1387 / 1274 / 1225 / 1194 / 1184 ms per frame with 16 / 32 / 64 / 128 / 256 Mi, profiles/r02k_sweep_streams_waves.log, r02l_sweep_waves.log: a wave is ~46 launches whose tails run almost empty, so fewer, larger waves waste less)
static int64_t g_countKd = 0;             // key 1: collect kd-walk counters (costs a little time)
static int64_t g_blocksPerSM = 12;        // key 2: resident CTAs per SM for the persistent kernels
static int64_t g_overlap = 3;             // key 5: 0 = one stream; 1 = a vertex's shadow + probe traversal on a second stream beside k_path_continue and the next closest-hit launch; 2 = only the MIS probes on the second stream, beside the shadow rays; 3 (default) = 1 with the probes on a third stream.  C3 is issue bound (+1 %), the big trees wait on L2 and two traversal kernels side by side fill each other's stalls: C4 237.3 -> 224.8 ms, C5 1542 -> 1396 ms per frame (profiles/r02j_sweep_streams.log)
static int64_t g_debugCounts = 0;         // key 6: print each wave's queue sizes per iteration to stderr
static int64_t g_reserved4 = 0;           // key 4: reserved (was the fused shade kernel's occupancy variant)
static int64_t g_fastPath = 1;            // key 8: k_path_fast shades flat-triangle / matte vertices in one pass (0: general three-kernel pipeline only)
static int64_t g_lanes = 1;               // key 10: waves of a frame in flight at once: 1 (default), or 2 (path integrator, frames of at least PB_LANE_MIN_SAMPLES samples; 3 = two lanes whatever the size, for tests).  Measured (profiles/r02n_sweep_lanes.log): two lanes with half-size waves C3 121.3 vs 119.8 ms, C4 211.9 vs 210.8, C5 1204.9 vs 1216.9, an 8-GPU rank's share (8.4 M samples) 16.07 vs 15.81 ms -- the three streams of one wave already fill most of what a second wave could, and persistent grids leave the other lane's CTAs little room before a launch's tail
static int64_t g_l2Persist = 0;           // key 11: percent of the L2 set aside for a persisting access window over the kd-tree nodes (0: off, the default -- measured: 25 % changes nothing, 50 / 75 % cost C4 3 / 23 % and C5 5 / 38 % because the queues' streaming traffic loses the L2 it used; profiles/r02aa_sweep_l2_persist.log)
static int64_t g_treelets = 0;            // key 9: device node order of the top-level tree: 1 = sibling pairs in line-sized treelets (pb_treelet.h), 0 = the reference's depth-first array (default: the re-ordering measured no gain, profiles/r02g_sweep_walk_variants.log)
static int64_t g_primFrames = 1;          // key 7: precompute hit-independent BSDF / light-triangle frames at scene create (0: evaluate per hit)

#ifndef PB_DIRECT_MIN_BLOCKS
#define PB_DIRECT_MIN_BLOCKS 5
#endif
#ifndef PB_CONT_MIN_BLOCKS
#define PB_CONT_MIN_BLOCKS 6
#endif
#ifndef PB_SURF_TEX_MIN_BLOCKS
#define PB_SURF_TEX_MIN_BLOCKS 6   // the texture-evaluating instances of k_path_surface
#endif
#ifndef PB_SURF_MIN_BLOCKS
#define PB_SURF_MIN_BLOCKS 6
#endif
#ifndef PB_TRAV_INST_MIN_BLOCKS
#define PB_TRAV_INST_MIN_BLOCKS 6   // the two-level walk of scenes with object instances carries more state: 80 registers
#endif
#ifndef PB_SHADOW_INST_MIN_BLOCKS
#define PB_SHADOW_INST_MIN_BLOCKS 8  // (measured on the instanced benchmark scene: closest hit 6 > 8 > 10, any hit 8 > 6 > 10)
#endif
#ifndef PB_TRAV_MIN_BLOCKS
#define PB_TRAV_MIN_BLOCKS 11  // <= 40 registers, 44 resident warps per SM for the kd walk (C3 closest hit: 11 < 10 < 9; any hit likewise)
#endif
#ifndef PB_TRAV_TRI_MIN_BLOCKS
#define PB_TRAV_TRI_MIN_BLOCKS 9   // triangle-only closest hit: 56 registers (C4 124 ms vs 140 at 10 and 133 at 8; C5 789 vs 882)
#endif

// ------------------------------------------------------------------------------------------
// device data layout of one wave
struct RayQueue {        // compacted SoA queue of rays (closest-hit kind)
  uint32_t *slot;        // owning sample slot
  float *ox, *oy, *oz, *dx, *dy, *dz, *mint, *maxt;
  int32_t *hitPrim;      // results, same index
  float *hitT, *hitB1, *hitB2;
  int32_t *hitInst;      // instance the hit was reached through (scenes with object instances only, else null)
};
struct AuxQueue {        // shadow segments / MIS probes
  uint32_t *rec;         // direct-record index (r * nSlots + slot)
  float *ox, *oy, *oz, *dx, *dy, *dz;
  int32_t *light;        // probes only
  float *maxt;           // shadow rays only: 1 - eps for a segment, inf for a distant light's ray
};
struct WaveBuffers {
  uint32_t nSlots;       // capacity
  // Path integrator: the per-slot state lives in three 16-byte records instead (null for whitted /
  // directlighting).  By the second bounce a warp's 32 vertices belong to 32 unrelated slots, so every per-slot
  // FIELD costs a 32-byte sector per lane; grouping what a kernel reads together makes that one sector per record.
  float4 *ptc;           // pathThroughput rgb | RNG stream-1 draws so far (bits)
  float4 *plf;           // L rgb | flags (bits)
  float4 *prec;          // pending direct term rgb | record flags (bits 0-3 as recFlags, bit 4 compact record)
  // per-slot state (whitted / directlighting; the path integrator uses pendT, alpha and recB of these)
  float *L[3], *T[3], *pendT[3], *alpha;
  uint32_t *ctr;         // RNG stream-1 draws so far
  uint32_t *flags;       // bits 0-7 depth, 8 specularBounce, 9 pending direct record
  float *ld;                   // lowdiscrepancy sampler: value e of slot s at ld[e * nSlots + s]
  unsigned long long *reach;   // per sample pixel of the wave: pixel footprint of its samples (k_film_reach)
  uint32_t *pixel;             // (x + 32768) | (y + 32768) << 16: sample pixel of the slot (RNG key)
  float *imageX, *imageY;      // film coordinates of the sample
  float *outL[3], *outA;       // per-sample radiance after the guards
  // direct-lighting records: directRecords per slot
  float *recA[3], *recB[3];
  uint8_t *recFlags;     // bit0 unoccluded, bit1 probe reached light, bit2 hasShadow, bit3 hasProbe
  // Whitted / DirectLighting recursion frames: nFrames per slot, 21 floats + 1 word each
  float *frame;          // [field][frame][slot]
  uint32_t *frameFlags;  // stage | reflValid<<8 | transValid<<9
  uint32_t *nRecs;       // records written by the last wd vertex
  int32_t nFrames;
  RayQueue q[2];
  AuxQueue shadowQ, probeQ;
  uint32_t *subset;      // path: queue indices k_path_fast leaves to the general shading kernels
  float *surf[15];       // BSDF frame records of the vertices being shaded (path), by queue position
  uint32_t *surfIds;
  uint32_t *surfQueueIdx;
  // scenes with texture operands only: the camera ray's differentials (rx.o, rx.d, ry.o, ry.d: value e of slot
  // s at camDiff[e * nSlots + s]) and, per hit record, the material with its operands evaluated at the hit
  float *camDiff;
  PbrtB200Material *surfMat;
  float *frameDiff;      // whitted / directlighting: the two child-ray differentials of each frame, [24][frame][slot]
  // counters (device): [0..] see enum below
  uint32_t *counts;
  unsigned long long *stats;
};
// device statistics: general counters, then 5 per traversal kind (rays, interior, leaves, refs, tests)
enum { ST_SHADED = 0, ST_NAN, ST_NEG, ST_INF, ST_CAMERA /* bestcandidate: table entries inside the extent */, ST_KIND_BASE };
enum { TK_MAIN = 0, TK_SHADOW, TK_PROBE, TK_COUNT };
enum { TF_RAYS = 0, TF_INTERIOR, TF_LEAVES, TF_REFS, TF_TESTS, TF_COUNT };
#define ST_COUNT (ST_KIND_BASE + TK_COUNT * TF_COUNT)
#define ST_K(kind, field) (ST_KIND_BASE + (kind) * TF_COUNT + (field))
// kernel classes for per-class timing (pbrt_b200_kernel_stats)
enum { KC_GENERATE = 0, KC_INTERSECT, KC_SHADE, KC_SHADE_DIRECT, KC_SHADE_CONT, KC_SHADOW, KC_PROBE, KC_ADVANCE,
       KC_FINALIZE, KC_FILM, KC_SHADE_FAST, KC_COUNT };
static const char *const kKernelClassNames[KC_COUNT] = {"k_generate", "k_intersect", "k_shade_surface", "k_shade_direct",
                                                        "k_shade_continue", "k_shadow", "k_probe", "k_advance_wd",
                                                        "k_finalize", "k_film", "k_shade_fast"};
// counts[] layout: per iteration it: [4*it+0] main queue size, +1 shadow queue size,
// +2 probe queue size, +3 rays of the main queue that hit something (path integrator); work counters follow at PB_WORK_BASE.
#define PB_MAX_ITERS 1024
#define PB_WORK_BASE (4 * PB_MAX_ITERS)
#define PB_COUNTS_TOTAL (PB_WORK_BASE + 10 * PB_MAX_ITERS)

// A wave is a run of consecutive image samples (reference order) that may span several pixel
// windows: each part is a range of sample pixels of one window's sample extent.
#define PB_MAX_PARTS 32
struct WavePart {
  int xs, xe, ys, ye;        // sample extent of the part's window (film/image.cpp:148-156)
  uint32_t firstPixel;       // first sample pixel (raster index inside the extent)
  uint32_t nPixels;
  uint32_t slotBase;         // wave slot of the part's first sample (multiple of spp)
  int32_t window;            // index into the handle's films
  // bestcandidate sampler: the unit is a placement of the 4096-entry sample table instead of a
  // pixel (firstPixel / nPixels count placements); corners of the window's placements, scan order
  const float2 *corners;
  int bcCols;                // placements per row when every row holds the same number, else 0
  // tile mode (pbrt_b200_render_tiles): the window is a TILE of the scene's film window.  Its sample extent holds
  // every sample pixel of the frame's extent that could reach the tile, and a sample is traced only if its filter
  // footprint (film/image.cpp:106-111) really touches a pixel of [tx0, tx1) x [ty0, ty1)
  int tile, tx0, ty0, tx1, ty1;
};
#define PB_BC_INVALID_COORD -1.0e6f   // film coordinate of a table entry outside the sampler's extent
struct WaveDesc {
  int spp;                   // samples per unit (pixel; bestcandidate: PB_BC_TABLE per placement)
  int bc;                    // bestcandidate sampler
  int tiles;                 // tile mode (see WavePart)
  uint32_t nSlots;
  int nParts;
  WavePart parts[PB_MAX_PARTS];
};

// ------------------------------------------------------------------------------------------
// persistent-kernel helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
// Each warp claims 32 consecutive work items at a time.
__device__ __forceinline__ uint32_t warp_claim(uint32_t *counter) {
  uint32_t base = 0;
  if (lane_id() == 0) base = atomicAdd(counter, 32u);
  return __shfl_sync(0xffffffffu, base, 0);
}
// Same, claiming `count` (a multiple of 32) items per atomic: the shading kernels walk their claim in
// 32-item steps, so the atomic's round trip is paid once per 128 items.
#define PB_CLAIM 128u
__device__ __forceinline__ uint32_t warp_claim_n(uint32_t *counter, uint32_t count) {
  uint32_t base = 0;
  if (lane_id() == 0) base = atomicAdd(counter, count);
  return __shfl_sync(0xffffffffu, base, 0);
}
// Warp-aggregated queue push: returns this lane's index or 0xffffffff.
__device__ __forceinline__ uint32_t warp_push(uint32_t *count, bool pred) {
  uint32_t mask = __ballot_sync(0xffffffffu, pred);
  if (mask == 0) return 0xffffffffu;
  uint32_t base = 0;
  int leader = __ffs(mask) - 1;
  if ((int)lane_id() == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return pred ? base + __popc(mask & ((1u << lane_id()) - 1u)) : 0xffffffffu;
}
__device__ __forceinline__ void warp_stat_add(unsigned long long *stat, uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane_id() == 0 && v) atomicAdd(stat, (unsigned long long)v);
}

__device__ __forceinline__ Ray queue_ray(const RayQueue &q, uint32_t i) {
  Ray r;
  r.o = v3(q.ox[i], q.oy[i], q.oz[i]);
  r.d = v3(q.dx[i], q.dy[i], q.dz[i]);
  r.mint = q.mint[i]; r.maxt = q.maxt[i];
  return r;
}
__device__ __forceinline__ void queue_put(const RayQueue &q, uint32_t i, uint32_t slot, const Ray &r) {
  q.slot[i] = slot;
  q.ox[i] = r.o.x; q.oy[i] = r.o.y; q.oz[i] = r.o.z;
  q.dx[i] = r.d.x; q.dy[i] = r.d.y; q.dz[i] = r.d.z;
  q.mint[i] = r.mint; q.maxt[i] = r.maxt;
}
__device__ __forceinline__ void aux_put(const AuxQueue &q, uint32_t i, uint32_t rec, const Ray &r, int light) {
  q.rec[i] = rec;
  q.ox[i] = r.o.x; q.oy[i] = r.o.y; q.oz[i] = r.o.z;
  q.dx[i] = r.d.x; q.dy[i] = r.d.y; q.dz[i] = r.d.z;
  if (q.light) q.light[i] = light;
  if (q.maxt) q.maxt[i] = r.maxt;
}

// camera-ray differentials of a slot (textured scenes), SoA over the wave
__device__ __forceinline__ void cam_diff_store(const SceneView &sv, const WaveBuffers &wb, uint32_t slot, const CameraSample &cs) {
  RayDiff rd;
  camera_ray_differentials(sv, cs, &rd);
  const float v[12] = {rd.rxo.x, rd.rxo.y, rd.rxo.z, rd.rxd.x, rd.rxd.y, rd.rxd.z,
                       rd.ryo.x, rd.ryo.y, rd.ryo.z, rd.ryd.x, rd.ryd.y, rd.ryd.z};
#pragma unroll
  for (int e = 0; e < 12; ++e) wb.camDiff[(size_t)e * wb.nSlots + slot] = v[e];
}
__device__ __forceinline__ void cam_diff_load(const WaveBuffers &wb, uint32_t slot, RayDiff *rd) {
  float v[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) v[e] = wb.camDiff[(size_t)e * wb.nSlots + slot];
  rd->rxo = v3(v[0], v[1], v[2]); rd->rxd = v3(v[3], v[4], v[5]);
  rd->ryo = v3(v[6], v[7], v[8]); rd->ryd = v3(v[9], v[10], v[11]);
}
// ------------------------------------------------------------------------------------------
// K0a (lowdiscrepancy sampler only): the (0,2)-sequence tables of every sample pixel of the wave
// (samplers/lowdiscrepancy.cpp:93-103), one thread per (sample pixel, table).
__global__ void k_ld_tables(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveDesc wd) {
  uint32_t nPix = wd.nSlots / (uint32_t)wd.spp;
  uint32_t nTab = 2u + (uint32_t)sv.nPatterns;
  for (uint64_t id = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; id < (uint64_t)nPix * nTab;
       id += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t t = (uint32_t)(id / nPix), p = (uint32_t)(id % nPix);
    uint32_t slot0 = p * (uint32_t)wd.spp;
    uint32_t px = wb.pixel[slot0];
    uint64_t key = pb_key((int)(px & 0xffffu) - 32768, (int)(px >> 16) - 32768, 0xFFFF, 0);
    const SamplePattern &pat = t < 2 ? sv.ldCamera[t] : sv.patterns[t - 2];
    pb_ld_generate_table(key, pat, wd.spp, wb.ld, wb.nSlots, slot0);
  }
}
// pixel coordinates of every slot of the wave (the RNG keys derive from them)
__global__ void k_slot_pixels(const WaveBuffers wb, const WaveDesc wd) {
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < wd.nSlots; slot += gridDim.x * blockDim.x) {
    int part = 0;
    while (part + 1 < wd.nParts && slot >= wd.parts[part + 1].slotBase) ++part;
    const WavePart &wp = wd.parts[part];
    uint32_t pix = wp.firstPixel + (slot - wp.slotBase) / wd.spp;
    int ew = wp.xe - wp.xs;
    int x = wp.xs + (int)(pix % ew), y = wp.ys + (int)(pix / ew);
    wb.pixel[slot] = (uint32_t)(x + 32768) | ((uint32_t)(y + 32768) << 16);
  }
}

// L = 0, pathThroughput = 1, no draws, depth 0
__device__ __forceinline__ void wave_state_init(const WaveBuffers &wb, uint32_t slot) {
  if (wb.ptc) {
    wb.ptc[slot] = make_float4(1.f, 1.f, 1.f, __uint_as_float(0u));
    wb.plf[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(0u));
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { wb.L[c][slot] = 0.f; wb.T[c][slot] = 1.f; }
  wb.ctr[slot] = 0u;
  wb.flags[slot] = 0u;
}
// K0: sample generation + camera rays (samplers/stratified.cpp:87-131, cameras/perspective.cpp:51-82)
__global__ void __launch_bounds__(PB_BLOCK)
k_generate(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveDesc wd) {
  uint32_t nValid = 0;
  // (tile mode compacts the ray queue with warp ballots: the loop bound is rounded up so that whole warps iterate)
  const uint32_t nLoop = wd.tiles ? ((wd.nSlots + 31u) & ~31u) : wd.nSlots;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < nLoop;
       slot += gridDim.x * blockDim.x) {
    const bool inWave = slot < wd.nSlots;
    int part = 0;
    while (part + 1 < wd.nParts && slot >= wd.parts[part + 1].slotBase) ++part;
    const WavePart &wp = wd.parts[part];
    uint32_t pix = inWave ? wp.firstPixel + (slot - wp.slotBase) / wd.spp : 0u;
    int k = (int)(slot % wd.spp);
    if (wd.bc) {
      // samplers/bestcandidate.cpp:107-129: entry k of placement `pix`; entries outside the extent are
      // skipped (no ray), the others enter the ray queue compacted (nSlots is a multiple of 32, so
      // whole warps reach the ballot)
      float2 corner = wp.corners[pix];
      int cx = pb_floor2int_fin(corner.x), cy = pb_floor2int_fin(corner.y);
      wb.pixel[slot] = (uint32_t)(cx + 32768) | ((uint32_t)(cy + 32768) << 16);
      CameraSample cs;
      bool valid = bc_camera_sample(sv, corner.x, corner.y, k, pb_key(cx, cy, 0, 3), wp.xs, wp.xe, wp.ys, wp.ye, &cs);
      uint32_t qi = warp_push(&wb.counts[0], valid);
      if (valid) {
        Ray ray;
        perspective_generate_ray(sv.camera, cs, &ray, sv.xRes, sv.yRes);
        queue_put(wb.q[0], qi, slot, ray);
        if (wb.camDiff) cam_diff_store(sv, wb, slot, cs);
        ++nValid;
      }
      wb.imageX[slot] = valid ? film_image_coord(cs.imageX) : PB_BC_INVALID_COORD;
      wb.imageY[slot] = valid ? film_image_coord(cs.imageY) : PB_BC_INVALID_COORD;
      wave_state_init(wb, slot);
      wb.alpha[slot] = valid ? 0.f : -1.f;   // -1 marks a skipped entry in the per-sample dump
      continue;
    }
    int ew = wp.xe - wp.xs;
    int x = wp.xs + (int)(pix % ew), y = wp.ys + (int)(pix / ew);
    CameraSample cs;
    if (inWave) {
      wb.pixel[slot] = (uint32_t)(x + 32768) | ((uint32_t)(y + 32768) << 16);
      if (sv.sampler.kind == PBRT_B200_SAMPLER_LOWDISCREPANCY) {
        PatternSource ps = {0, wb.ld, wb.nSlots, slot, 0};
        ld_camera_sample(sv, ps, x, y, &cs);
      } else
        stratified_camera_sample(sv, x, y, k, &cs);
    }
    if (wd.tiles) {
      // tile of a larger frame: the sample is traced when ImageFilm::AddSample would touch a pixel of the tile
      bool valid = false;
      if (inWave) {
        float dImageX = film_image_coord(cs.imageX) - 0.5f, dImageY = film_image_coord(cs.imageY) - 0.5f;
        int x0 = pb_ceil2int_fin(dImageX - sv.filterXW), x1 = pb_floor2int_fin(dImageX + sv.filterXW);
        int y0 = pb_ceil2int_fin(dImageY - sv.filterYW), y1 = pb_floor2int_fin(dImageY + sv.filterYW);
        valid = x1 >= wp.tx0 && x0 < wp.tx1 && y1 >= wp.ty0 && y0 < wp.ty1;
      }
      uint32_t qi = warp_push(&wb.counts[0], valid);
      if (valid) {
        Ray ray;
        perspective_generate_ray(sv.camera, cs, &ray, sv.xRes, sv.yRes);
        queue_put(wb.q[0], qi, slot, ray);
        if (wb.camDiff) cam_diff_store(sv, wb, slot, cs);
        ++nValid;
      }
      if (inWave) {
        wb.imageX[slot] = valid ? film_image_coord(cs.imageX) : PB_BC_INVALID_COORD;
        wb.imageY[slot] = valid ? film_image_coord(cs.imageY) : PB_BC_INVALID_COORD;
        wave_state_init(wb, slot);
        wb.alpha[slot] = valid ? 0.f : -1.f;
      }
      continue;
    }
    Ray ray;
    perspective_generate_ray(sv.camera, cs, &ray, sv.xRes, sv.yRes);
    queue_put(wb.q[0], slot, slot, ray);
    if (wb.camDiff) cam_diff_store(sv, wb, slot, cs);
    wb.imageX[slot] = film_image_coord(cs.imageX);
    wb.imageY[slot] = film_image_coord(cs.imageY);
    wave_state_init(wb, slot);
    wb.alpha[slot] = 0.f;
  }
  if (wd.bc || wd.tiles) warp_stat_add(&wb.stats[ST_CAMERA], nValid);
}

// K0b (bestcandidate sampler only): the LDShuffleScrambled1D/2D(4096, 1) tables of every table
// placement of the wave (samplers/bestcandidate.cpp:96-105), one warp per (placement, one-sample
// pattern): the Shuffle's 4096 draws in parallel, its swaps by one lane in shared memory, then the
// (0,2)-sequence points written through the permutation.  Runs after k_generate (reads the
// placements' corner pixels from wb.pixel).
#define PB_BC_TABLE_WARPS 2
__global__ void __launch_bounds__(32 * PB_BC_TABLE_WARPS)
k_bc_tables(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveDesc wd) {
  __shared__ uint16_t permS[PB_BC_TABLE_WARPS][PB_BC_TABLE];
  const uint32_t lane = lane_id(), wib = threadIdx.x >> 5;
  uint16_t *perm = permS[wib];
  const uint32_t nUnits = wd.nSlots / PB_BC_TABLE;
  const uint32_t total = nUnits * (uint32_t)sv.nPatterns;
  const uint32_t nWarps = gridDim.x * PB_BC_TABLE_WARPS;
  for (uint32_t w = blockIdx.x * PB_BC_TABLE_WARPS + wib; w < total; w += nWarps) {
    const SamplePattern &p = sv.patterns[w % (uint32_t)sv.nPatterns];
    if (p.n != 1) continue;
    const uint32_t slot0 = (w / (uint32_t)sv.nPatterns) * PB_BC_TABLE;
    uint32_t px = wb.pixel[slot0];
    uint64_t key3 = pb_key((int)(px & 0xffffu) - 32768, (int)(px >> 16) - 32768, 0, 3);
    uint32_t scr0 = pb_u32(key3, p.ctr0 + 1u), scr1 = p.dims == 2 ? pb_u32(key3, p.ctr0 + 2u) : 0u;
    uint32_t ctrShuffle = p.ctr0 + (uint32_t)p.dims;
    for (uint32_t i = lane; i < PB_BC_TABLE; i += 32) perm[i] = (uint16_t)i;
    __syncwarp();
    // Shuffle (core/sampling.cpp:91-97): 32 swap partners drawn in parallel, the swaps themselves in
    // order by lane 0 (partners handed over by shuffles, so shared memory only holds the permutation)
    for (uint32_t base = 0; base < PB_BC_TABLE; base += 32) {
      uint32_t mine = pb_u32(key3, ctrShuffle + 1u + base + lane) % (uint32_t)PB_BC_TABLE;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        uint32_t o = __shfl_sync(0xffffffffu, mine, j);
        if (lane == 0) {
          uint16_t t = perm[base + j];
          perm[base + j] = perm[o]; perm[o] = t;
        }
      }
    }
    __syncwarp();
    for (uint32_t o = lane; o < PB_BC_TABLE; o += 32) {
      wb.ld[(size_t)p.offset * wb.nSlots + slot0 + o] = pb_van_der_corput(perm[o], scr0);
      if (p.dims == 2) wb.ld[(size_t)(p.offset + 1) * wb.nSlots + slot0 + o] = pb_sobol2(perm[o], scr1);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// K1: closest hit (KdTreeAccel::Intersect, accelerators/kdtree.cpp:313-403)
struct MainSource {
  RayQueue q;
  __device__ Ray load(uint32_t i) const { return queue_ray(q, i); }
};
struct MainSink {
  RayQueue q;
  __device__ void done(uint32_t i, const Ray &, const Hit &hit, bool) const {
    q.hitPrim[i] = hit.prim;
    if (hit.prim >= 0) {
      q.hitT[i] = hit.t; q.hitB1[i] = hit.b1; q.hitB2[i] = hit.b2;
      if (q.hitInst) q.hitInst[i] = hit.inst;
    }
  }
};
__device__ __forceinline__ void flush_kd_counters(unsigned long long *stats, int kind, const KdCounters &cnt) {
  warp_stat_add(&stats[ST_K(kind, TF_INTERIOR)], cnt.interior);
  warp_stat_add(&stats[ST_K(kind, TF_LEAVES)], cnt.leaves);
  warp_stat_add(&stats[ST_K(kind, TF_REFS)], cnt.refs);
  warp_stat_add(&stats[ST_K(kind, TF_TESTS)], cnt.tests);
}
#ifndef PB_DESCEND_STEPS_BIG
#define PB_DESCEND_STEPS_BIG 4
#endif
#ifndef PB_LANE_MIN_SAMPLES
#define PB_LANE_MIN_SAMPLES (1u << 21)   // frames below this many image samples render as one wave on one lane
#endif
#ifndef PB_PROBE_CTAS_BIG
#define PB_PROBE_CTAS_BIG 2   // CTAs per SM of the probe kernel on big trees
#endif
#ifndef PB_LEAN_WALK
#define PB_LEAN_WALK 1   // scenes without object instances walk with kd_warp_traverse_lean (0: the general loop, for A/B)
#endif
#ifndef PB_BIG_TREE_NODES
#define PB_BIG_TREE_NODES 100000u   // top-level trees above this many nodes use PB_DESCEND_STEPS_BIG
#endif
template <bool COUNT, int QK, int DS>
__global__ void __launch_bounds__(PB_BLOCK, QK == PB_QK_INSTANCES ? PB_TRAV_INST_MIN_BLOCKS
                                            : QK == PB_QK_TRI ? PB_TRAV_TRI_MIN_BLOCKS : PB_TRAV_MIN_BLOCKS)
k_intersect(const __grid_constant__ SceneView sv, const RayQueue q, const uint32_t *nPtr,
            uint32_t *work, unsigned long long *stats) {
  uint32_t n = *nPtr;
  KdCounters cnt = {0, 0, 0, 0};
  MainSource src = {q};
  MainSink sink = {q};
  if constexpr (PB_LEAN_WALK && QK != PB_QK_INSTANCES) kd_warp_traverse_lean<false, COUNT, QK>(sv, n, work, src, sink, cnt);
  else kd_warp_traverse<false, COUNT, QK, DS>(sv, n, work, src, sink, cnt);
  if (COUNT) flush_kd_counters(stats, TK_MAIN, cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[ST_K(TK_MAIN, TF_RAYS)], (unsigned long long)n);
}

// K3: shadow rays (KdTreeAccel::IntersectP, kdtree.cpp:404-488; light.cpp:30-36)
struct ShadowSource {
  AuxQueue q;
  __device__ Ray load(uint32_t i) const {
    Ray ray;
    ray.o = v3(q.ox[i], q.oy[i], q.oz[i]);
    ray.d = v3(q.dx[i], q.dy[i], q.dz[i]);
    ray.mint = PB_RAY_EPSILON; ray.maxt = q.maxt[i];   // light.h:78-83 (SetSegment / SetRay)
    return ray;
  }
};
// record flag bits live in a byte array (whitted / directlighting: several records per slot) or in the w word of
// the path integrator's per-slot record
__device__ __forceinline__ void rec_flag_set(uint8_t *recFlags, float4 *prec, uint32_t rec, uint32_t bit) {
  if (prec) {
    // (atomic: the shadow and the probe kernel of one vertex may run concurrently, tunable 5, and a record can
    // be waiting for both)
    atomicOr(reinterpret_cast<uint32_t *>(&prec[rec]) + 3, bit);
  } else
    recFlags[rec] |= (uint8_t)bit;
}
struct ShadowSink {
  AuxQueue q;
  uint8_t *recFlags;
  float4 *prec;
  __device__ void done(uint32_t i, const Ray &, const Hit &, bool occluded) const {
    if (!occluded) rec_flag_set(recFlags, prec, q.rec[i], 1u);
  }
};
#ifndef PB_SHADOW_TRI_MIN_BLOCKS
#define PB_SHADOW_TRI_MIN_BLOCKS PB_TRAV_MIN_BLOCKS
#endif
template <bool COUNT, int QK, int DS>
__global__ void __launch_bounds__(PB_BLOCK, QK == PB_QK_INSTANCES ? PB_SHADOW_INST_MIN_BLOCKS
                                            : QK == PB_QK_TRI ? PB_SHADOW_TRI_MIN_BLOCKS : PB_TRAV_MIN_BLOCKS)
k_shadow(const __grid_constant__ SceneView sv, const AuxQueue q, const uint32_t *nPtr, uint32_t *work,
         uint8_t *recFlags, float4 *prec, unsigned long long *stats) {
  uint32_t n = *nPtr;
  KdCounters cnt = {0, 0, 0, 0};
  ShadowSource src = {q};
  ShadowSink sink = {q, recFlags, prec};
  if constexpr (PB_LEAN_WALK && QK != PB_QK_INSTANCES) kd_warp_traverse_lean<true, COUNT, QK>(sv, n, work, src, sink, cnt);
  else kd_warp_traverse<true, COUNT, QK, DS>(sv, n, work, src, sink, cnt);
  if (COUNT) flush_kd_counters(stats, TK_SHADOW, cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[ST_K(TK_SHADOW, TF_RAYS)], (unsigned long long)n);
}

// K4: MIS probes (transport.cpp:177-190): closest hit, then "did it reach the sampled light,
// on its emitting side".
struct ProbeSource {
  AuxQueue q;
  __device__ Ray load(uint32_t i) const {
    Ray ray;
    ray.o = v3(q.ox[i], q.oy[i], q.oz[i]);
    ray.d = v3(q.dx[i], q.dy[i], q.dz[i]);
    ray.mint = PB_RAY_EPSILON; ray.maxt = PB_INFINITY;
    return ray;
  }
};
struct ProbeSink {
  const SceneView *sv;
  AuxQueue q;
  uint8_t *recFlags;
  float4 *prec;
  __device__ void done(uint32_t i, const Ray &ray, const Hit &hit, bool) const {
    if (hit_emits_from_light(*sv, ray, hit, q.light[i])) rec_flag_set(recFlags, prec, q.rec[i], 2u);
  }
};
template <bool COUNT, bool INST>
__global__ void __launch_bounds__(PB_BLOCK)
k_probe(const __grid_constant__ SceneView sv, const AuxQueue q, const uint32_t *nPtr, uint32_t *work,
        uint8_t *recFlags, float4 *prec, unsigned long long *stats) {
  uint32_t n = *nPtr;
  KdCounters cnt = {0, 0, 0, 0};
  ProbeSource src = {q};
  ProbeSink sink = {&sv, q, recFlags, prec};
  if constexpr (PB_LEAN_WALK && !INST) kd_warp_traverse_lean<false, COUNT, PB_QK_ALL>(sv, n, work, src, sink, cnt);
  else kd_warp_traverse<false, COUNT, PB_QK_INSTANCES, PB_DESCEND_STEPS>(sv, n, work, src, sink, cnt);
  if (COUNT) flush_kd_counters(stats, TK_PROBE, cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[ST_K(TK_PROBE, TF_RAYS)], (unsigned long long)n);
}

// ------------------------------------------------------------------------------------------
// per-slot state access
__device__ __forceinline__ Spec load3(float *const a[3], uint32_t i) { return spec3(a[0][i], a[1][i], a[2][i]); }
__device__ __forceinline__ void store3(float *const a[3], uint32_t i, const Spec &s) {
  a[0][i] = s.c[0]; a[1][i] = s.c[1]; a[2][i] = s.c[2];
}
// RNG key of a slot's image sample: (pixel x, pixel y, sample-in-pixel, stream)
struct WaveKeyInfo {
  const uint32_t *pixel;
  int spp;
  int bc;   // 1 bestcandidate sampler: pixel[] holds the placement's corner pixel, slot % spp the table entry; 2 random sampler
};
__device__ __forceinline__ uint64_t slot_key(const WaveKeyInfo &wd, uint32_t slot, int stream) {
  uint32_t p = wd.pixel[slot];
  int x = (int)(p & 0xffffu) - 32768, y = (int)(p >> 16) - 32768, k = (int)(slot % (uint32_t)wd.spp);
  if (wd.bc == 1) return pb_key_bc(x, y, k, stream == 0 ? 2 : stream);   // patterns: the sample's stream 2
  return pb_key(x, y, k, stream);
}
// `L += pathThroughput * UniformSampleOneLight(...)` (path.cpp:100-112) runs at EVERY vertex of the reference's loop,
// also when the estimate is black: x * 0 is 0 for a finite throughput and NaN for a NaN or infinite one, so a path
// whose throughput went bad (a degenerate BSDF) turns its sample into a NaN sample -- which Scene::Render then reports
// and sets to black (scene.cpp:58-74) -- at the next vertex.  The wavefront skips adding black estimates, so it adds
// that NaN explicitly: pb_bad_channels() of the throughput travels in bits 5-7 of a compact pending record, and a
// vertex that leaves no record applies it at once.  (Found by the 4000-seed GPU fuzz: two scenes where the twin
// counted NaN samples and the GPU kept a finite value, profiles/r02ac_fuzz_*.log.)
__device__ __forceinline__ uint32_t pb_bad_channels(const Spec &t) {
  uint32_t m = 0;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    if (!(fabsf(t.c[c]) <= 3.402823466e+38f)) m |= 1u << c;   // NaN or +-inf
  return m;
}
__device__ __forceinline__ void pb_add_bad_zero(Spec *L, uint32_t mask) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
    if (mask & (1u << c)) L->c[c] = L->c[c] + __uint_as_float(0x7fc00000u);
}
// Fold the slot's pending direct-lighting record (path integrator) into L.
__device__ __forceinline__ void path_fold_pending(const SceneView &sv, const WaveBuffers &wb,
                                                  uint32_t slot, uint32_t *flags, Spec *L) {
  if (!(*flags & (1u << 9))) return;
  *flags &= ~(1u << 9);
  const float4 r = wb.prec[slot];
  const uint32_t rf = __float_as_uint(r.w);
  if (rf & 16) {
    // compact record: rgb holds pathThroughput * (nLights * Ld), valid when the shadow ray got through; otherwise
    // the reference added pathThroughput * 0 (bits 5-7: channels where that is not 0)
    if (rf & 1) *L = *L + spec3(r.x, r.y, r.z);
    else if (rf & 0xe0u) pb_add_bad_zero(L, (rf >> 5) & 7u);
    return;
  }
  PathState tmp;
  tmp.L = *L;
  tmp.pendT = load3(wb.pendT, slot);
  path_resolve_direct(sv, tmp, spec3(r.x, r.y, r.z), load3(wb.recB, slot), (rf & 4) && (rf & 1), (rf & 8) && (rf & 2));
  *L = tmp.L;
}

// K2 (path): one iteration of PathIntegrator::Li's loop (integrators/path.cpp:65-143) as three
// kernels.  One fused kernel measured 65 % of its issue slots stalled on instruction fetch
// (8.7 k SASS instructions against a 32 KB instruction cache, profiles/r01a_*), so the vertex is
// cut where little state crosses:
//   k_path_surface   fold the previous vertex's direct light, miss handling, hit geometry ->
//                    BSDF frame record (15 floats + ids), emitted light        path.cpp:66-92
//   k_path_direct    UniformSampleOneLight / EstimateDirect -> record + shadow / probe rays  :93-112
//   k_path_continue  BSDF sampling, throughput, Russian roulette -> next ray                 :113-143
struct SurfaceRecord {       // SoA, compacted: record j belongs to the j-th ray of the queue that hit something
  float *f[15];              // p, nn, ng, sn, tn
  uint32_t *ids;             // material | (light + 1) << 16
  uint32_t *queueIdx;        // queue position of the ray (its slot and direction are read from there)
  PbrtB200Material *mat;     // textured scenes: the material record evaluated at the hit, else null
};
__device__ __forceinline__ void surface_store(const SurfaceRecord &r, uint32_t i, const Surface &sf) {
  const Bsdf &b = sf.bsdf;
  r.f[0][i] = b.p.x; r.f[1][i] = b.p.y; r.f[2][i] = b.p.z;
  r.f[3][i] = b.nn.x; r.f[4][i] = b.nn.y; r.f[5][i] = b.nn.z;
  r.f[6][i] = b.ng.x; r.f[7][i] = b.ng.y; r.f[8][i] = b.ng.z;
  r.f[9][i] = b.sn.x; r.f[10][i] = b.sn.y; r.f[11][i] = b.sn.z;
  r.f[12][i] = b.tn.x; r.f[13][i] = b.tn.y; r.f[14][i] = b.tn.z;
  r.ids[i] = sf.material | ((uint32_t)(sf.light + 1) << 16);
}
template <bool G>
__device__ __forceinline__ void surface_load(const SceneView &sv, const SurfaceRecord &r, uint32_t i, Surface *sf) {
  Bsdf &b = sf->bsdf;
  b.p = v3(r.f[0][i], r.f[1][i], r.f[2][i]);
  b.nn = v3(r.f[3][i], r.f[4][i], r.f[5][i]);
  b.ng = v3(r.f[6][i], r.f[7][i], r.f[8][i]);
  b.sn = v3(r.f[9][i], r.f[10][i], r.f[11][i]);
  b.tn = v3(r.f[12][i], r.f[13][i], r.f[14][i]);
  uint32_t ids = r.ids[i];
  sf->material = ids & 0xffffu;
  sf->light = (int32_t)(ids >> 16) - 1;
  bsdf_set_lobes<G>(r.mat ? r.mat[i] : sv.materials[sf->material], &b);
}

// The frame only; the caller adds the lobes (general Bsdf) or the reflectance (LambertBsdf).
template <class B>
__device__ __forceinline__ uint32_t surface_load_frame(const SurfaceRecord &r, uint32_t i, B *b) {
  b->p = v3(r.f[0][i], r.f[1][i], r.f[2][i]);
  b->nn = v3(r.f[3][i], r.f[4][i], r.f[5][i]);
  b->ng = v3(r.f[6][i], r.f[7][i], r.f[8][i]);
  b->sn = v3(r.f[9][i], r.f[10][i], r.f[11][i]);
  b->tn = v3(r.f[12][i], r.f[13][i], r.f[14][i]);
  return r.ids[i];
}
#ifndef PB_FUSE_DIFFUSE
#define PB_FUSE_DIFFUSE 0   // 1: k_path_direct also samples the continuation of its (diffuse-class) vertices
#endif
#ifndef PB_LAMBERT_PATH
#define PB_LAMBERT_PATH 1   // 0: every vertex goes through the general lobe table (A/B switch for the profiles)
#endif
// matte with sigma == 0: one Lambertian lobe (materials/matte.cpp:59-62)
__device__ __forceinline__ bool material_is_lambert(const PbrtB200Material &m) {
  return PB_LAMBERT_PATH && m.kind == PBRT_B200_MAT_MATTE && m.sigma == 0.f;
}

// One vertex of k_path_surface for queue entry i.  GENERAL = false: the caller knows the ray missed or hit a
// triangle that carries a precomputed frame, and the shape / shading-geometry code is compiled out.
// TEX: the scene has texture operands (ray differentials of the camera ray, operand evaluation at the hit)
template <bool INST, bool TEX, bool GENERAL>
__device__ __forceinline__ void path_surface_vertex(const SceneView &sv, const WaveBuffers &wb, const RayQueue &qin,
                                                    const SurfaceRecord &rec, uint32_t i, bool valid,
                                                    uint32_t *hitCount, uint32_t *specCount, uint32_t &shaded) {
  bool isHit = false;
  Surface sf;
  PbrtB200Material evalMat;
  if (valid) {
    uint32_t slot = qin.slot[i];
    Ray ray = queue_ray(qin, i);
    Hit hit;
    hit.prim = qin.hitPrim[i];
    hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.inst = -1;
    if (hit.prim >= 0) {
      hit.t = qin.hitT[i]; hit.b1 = qin.hitB1[i]; hit.b2 = qin.hitB2[i];
      if (qin.hitInst) hit.inst = qin.hitInst[i];
    }
    const float4 lf = wb.plf[slot];
    uint32_t flags = __float_as_uint(lf.w);
    PathState st;
    st.L = spec3(lf.x, lf.y, lf.z);
    path_fold_pending(sv, wb, slot, &flags, &st.L);
    st.depth = (int)(flags & 0xffu);
    st.specularBounce = (flags >> 8) & 1u;
    // the throughput only matters where emitted light can be added (path.cpp:74-82,90-91), alpha only
    // at the first vertex
    const bool mayEmit = st.depth == 0 || st.specularBounce;
    st.T = spec(0.f);
    if (mayEmit) { const float4 tc = wb.ptc[slot]; st.T = spec3(tc.x, tc.y, tc.z); }
    st.alpha = 0.f;
    if (TEX) {
      RayDiff rd;
      rd.has = st.depth == 0;   // path.cpp:133 continues with plain rays
      if (rd.has) cam_diff_load(wb, slot, &rd);
      isHit = path_vertex_surface<false, INST, true>(sv, st, ray, hit, &sf, &rd, &evalMat);
    } else
      isHit = path_vertex_surface<false, INST, false, !GENERAL>(sv, st, ray, hit, &sf);
    if (isHit) ++shaded;
    wb.plf[slot] = make_float4(st.L.c[0], st.L.c[1], st.L.c[2], __uint_as_float(flags));
    if (st.depth == 0) wb.alpha[slot] = st.alpha;
  }
  // The shading kernels that follow walk the hits only, and the hits are binned by material class so that
  // their warps do not mix glass / mirror vertices with diffuse ones: vertices whose BSDF has a non-specular
  // component fill the record arrays from the front (k_path_direct walks exactly these), specular-only ones
  // (glass, mirror: UniformSampleOneLight cannot contribute there) from the back.
  bool specOnly = false;
  if (isHit) {
    const int kind = TEX ? evalMat.kind : sv.materials[sf.material].kind;
    specOnly = kind == PBRT_B200_MAT_GLASS || kind == PBRT_B200_MAT_MIRROR;
  }
  uint32_t j = warp_push(hitCount, isHit && !specOnly);
  uint32_t js = warp_push(specCount, isHit && specOnly);
  if (specOnly) j = wb.nSlots - 1u - js;
  if (isHit) {
    surface_store(rec, j, sf);
    rec.queueIdx[j] = i;
    if (TEX) rec.mat[j] = evalMat;
  }
}

// K2a.  Untextured scenes: most hits are flat triangles whose frame is a 48-byte table read, a few are quadrics or
// triangles with per-vertex shading data whose geometry takes hundreds of instructions (ncu, r02a: the glass
// sphere's lanes ran that code at 1.7 active lanes per instruction and made up over half of the kernel's
// instructions).  So a warp shades the cheap vertices of its 32 rays at once and parks the expensive ones in a
// small shared-memory list, which it works off 32 at a time with every lane busy.
template <bool INST, bool TEX>
__global__ void __launch_bounds__(PB_BLOCK, TEX ? PB_SURF_TEX_MIN_BLOCKS : PB_SURF_MIN_BLOCKS)
k_path_surface(const __grid_constant__ SceneView sv, const WaveBuffers wb, const RayQueue qin,
               const SurfaceRecord rec, const uint32_t *nPtr, uint32_t *work, uint32_t *hitCount, uint32_t *specCount,
               const uint32_t *subset) {
  __shared__ uint32_t parkedAll[PB_BLOCK / 32][64];
  uint32_t *parked = parkedAll[threadIdx.x >> 5];
  uint32_t nParked = 0;   // warp uniform
  const uint32_t lane = lane_id();
  uint32_t n = *nPtr;
  uint32_t shaded = 0;
  for (;;) {
    uint32_t chunk = warp_claim_n(work, PB_CLAIM);
    if (chunk >= n) break;
    for (uint32_t base = chunk; base < chunk + PB_CLAIM && base < n; base += 32) {
      uint32_t i = base + lane;
      const bool valid = i < n;
      if (subset) i = valid ? subset[i] : 0u;   // the queue entries k_path_fast left to this pipeline
      if (TEX || !sv.primFrame) {
        path_surface_vertex<INST, TEX, true>(sv, wb, qin, rec, i, valid, hitCount, specCount, shaded);
        continue;
      }
      bool expensive = false;
      if (valid) {
        int32_t prim = qin.hitPrim[i];
        if (prim >= 0) {
          uint32_t pflags = __float_as_uint(__ldg(&sv.prims[3u * (uint32_t)prim + 2].b));
          expensive = !(pflags & PB_PF_DEV_FRAME) || (INST && qin.hitInst && qin.hitInst[i] >= 0);
        }
      }
      path_surface_vertex<INST, false, false>(sv, wb, qin, rec, i, valid && !expensive, hitCount, specCount, shaded);
      const uint32_t mask = __ballot_sync(0xffffffffu, expensive);
      if (mask) {
        if (expensive) parked[nParked + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = i;
        nParked += (uint32_t)__popc(mask);
        __syncwarp();
        if (nParked >= 32u) {
          nParked -= 32u;
          uint32_t pi = parked[nParked + lane];
          __syncwarp();
          path_surface_vertex<INST, false, true>(sv, wb, qin, rec, pi, true, hitCount, specCount, shaded);
        }
      }
    }
  }
  if (nParked) {
    uint32_t pi = lane < nParked ? parked[lane] : 0u;
    path_surface_vertex<INST, false, true>(sv, wb, qin, rec, pi, lane < nParked, hitCount, specCount, shaded);
  }
  warp_stat_add(&wb.stats[ST_SHADED], shaded);
}

// K2 fast path (untextured scenes).  Most path vertices are a ray that hit a flat triangle with a plain matte
// material: for those the whole loop body of PathIntegrator::Li (path.cpp:66-143) -- fold the previous
// vertex's direct light, emitted light, the UniformSampleOneLight record, the next direction -- runs here in
// registers, from the ray and hit records straight to the shadow / probe / next-ray queues, with no surface
// record through HBM in between (three kernels and ~470 bytes of traffic per vertex before).  Every other hit
// (quadrics, per-vertex shading normals, instances, glass / mirror / glossy / Oren-Nayar materials) is passed
// on: its queue index goes to `subset`, which the general k_path_surface / _direct / _continue pipeline then
// shades.  Misses end here.  The arithmetic is the general kernels' (same inlined functions, LambertBsdf).
#ifndef PB_FAST_PREFETCH
#define PB_FAST_PREFETCH 0   // prefetch a parked hit's state records and frame into L1 (measured: no effect, 35.3 vs
#endif                       // 35.1 ms per C3 frame, profiles/r02v_sweep_fast_prefetch.log -- the kernel waits on
                             // instruction fetch and issue, not on these loads)
#ifndef PB_FAST_MIN_BLOCKS
#define PB_FAST_MIN_BLOCKS 6   // 80 registers (r02d_sweep.log: 3 / 4 / 5 / 6 CTAs per SM: 54.3 / 43.8 / 40.9 / 38.2 ms per C3 frame)
#endif
// One plain vertex (flat triangle, matte): queue entry i, all in registers.
template <bool LD, int LK>
__device__ __forceinline__ void path_fast_vertex(const SceneView &sv, const WaveBuffers &wb, const WaveKeyInfo &wd,
                                                 const RayQueue &qin, const RayQueue &qout, uint32_t i, bool valid,
                                                 uint32_t *shadowCount, uint32_t *probeCount, uint32_t *nextCount,
                                                 uint32_t &shaded) {
  bool pushShadow = false, pushProbe = false, pushNext = false;
  uint32_t slot = 0;
  DirectRec direct;
  Ray next;
  if (valid) {
    const uint32_t prim = (uint32_t)qin.hitPrim[i];
    const float4 tail = __ldg((const float4 *)(sv.prims + 3u * prim + 2));
    slot = qin.slot[i];
    const float4 lf = wb.plf[slot];
    uint32_t flags = __float_as_uint(lf.w);
    PathState st;
    st.L = spec3(lf.x, lf.y, lf.z);
    path_fold_pending(sv, wb, slot, &flags, &st.L);
    st.depth = (int)(flags & 0xffu);
    st.specularBounce = (flags >> 8) & 1u;
    st.alpha = 0.f;
    Ray ray = queue_ray(qin, i);
    const float4 tc = wb.ptc[slot];
    st.T = spec3(tc.x, tc.y, tc.z);
    st.rng.ctr = __float_as_uint(tc.w);
    // path.cpp:83-92: alpha, the frame (k_prim_frames' table), emitted light
    if (st.depth == 0) st.alpha = 1.f;
    else st.T = st.T * spec(1.f);
    LambertSurface ls;
    const float4 *fr = sv.primFrame + 3u * prim;
    const float4 a = __ldg(fr), b = __ldg(fr + 1), c = __ldg(fr + 2);
    ls.bsdf.p = ray_at(ray, qin.hitT[i]);
    ls.bsdf.ng = v3(a.x, a.y, a.z); ls.bsdf.nn = v3(a.w, b.x, b.y);
    ls.bsdf.sn = v3(b.z, b.w, c.x); ls.bsdf.tn = v3(c.y, c.z, c.w);
    ls.material = __float_as_uint(tail.z);
    ls.light = __float_as_int(tail.w);
    const PbrtB200Material &mat = sv.materials[ls.material];
    ls.bsdf.R = spec3(mat.kd[0], mat.kd[1], mat.kd[2]) * PB_INV_PI;
    if (st.depth == 0 || st.specularBounce) {
      Spec Le = spec(0.f);   // Intersection::Le (primitive.cpp:142-145)
      if (ls.light >= 0 && dot(ls.bsdf.ng, -ray.d) > 0) Le = light_color(sv.lights[ls.light]);
      st.L = st.L + st.T * Le;
    }
    ++shaded;
    // path.cpp:93-112
    st.rng.key = slot_key(wd, slot, 1);
    PatternSource key0 = {slot_key(wd, slot, 0), LD ? wb.ld : (const float *)nullptr, wb.nSlots, slot, LD ? wd.bc : (wd.bc == 2 ? 2 : 0)};
    bool hasDirect;
    path_vertex_direct<false, LK>(sv, key0, st, ray, &ls, &direct, &hasDirect);
    const Spec pendT = st.pendT;
    // path.cpp:113-143
    pushNext = path_vertex_continue<false>(sv, key0, st, ray, ls, &next);
    flags = (uint32_t)st.depth | (st.specularBounce ? (1u << 8) : 0u);
    if (hasDirect) {
      pushShadow = direct.hasShadow;
      pushProbe = direct.hasProbe;
      if (pushProbe) {
        flags |= 1u << 9;
        store3(wb.pendT, slot, pendT);
        store3(wb.recB, slot, direct.termB);
        wb.prec[slot] = make_float4(direct.termA.c[0], direct.termA.c[1], direct.termA.c[2],
                                    __uint_as_float((direct.hasShadow ? 4u : 0u) | 8u));
      } else if (pushShadow) {
        Spec Ld = direct_value(direct.termA, direct.termB, true, false);
        Spec add = pendT * ((float)sv.nLights * Ld);
        flags |= 1u << 9;
        wb.prec[slot] = make_float4(add.c[0], add.c[1], add.c[2], __uint_as_float(4u | 16u | (pb_bad_channels(pendT) << 5)));
      } else
        pb_add_bad_zero(&st.L, pb_bad_channels(pendT));   // black estimate, no ray: L += pathThroughput * 0 now
    }
    wb.plf[slot] = make_float4(st.L.c[0], st.L.c[1], st.L.c[2], __uint_as_float(flags));
    wb.ptc[slot] = make_float4(st.T.c[0], st.T.c[1], st.T.c[2], __uint_as_float(st.rng.ctr));
    if (st.alpha != 0.f) wb.alpha[slot] = st.alpha;   // (depth 0 only; k_generate wrote 0)
  }
  uint32_t si = warp_push(shadowCount, pushShadow);
  if (pushShadow) aux_put(wb.shadowQ, si, slot, direct.shadow, 0);
  uint32_t pi = warp_push(probeCount, pushProbe);
  if (pushProbe) aux_put(wb.probeQ, pi, slot, direct.probe, direct.probeLight);
  uint32_t ni = warp_push(nextCount, pushNext);
  if (pushNext) queue_put(qout, ni, slot, next);
}

template <bool LD, int LK>
__global__ void __launch_bounds__(PB_BLOCK, PB_FAST_MIN_BLOCKS)
k_path_fast(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveKeyInfo wd, const RayQueue qin,
            const RayQueue qout, const uint32_t *nPtr, uint32_t *work, uint32_t *shadowCount, uint32_t *probeCount,
            uint32_t *nextCount, uint32_t *subset, uint32_t *subsetCount) {
  // a warp sorts its 32 rays into: missed (finished here, a few memory operations), to be passed on to the
  // general kernels, and plain hits -- the latter wait in a shared-memory list and are shaded 32 at a time, so that
  // the ~1700-instruction vertex body always runs with full warps (r02d ncu: 20 of 32 lanes active without this)
  __shared__ uint32_t parkedAll[PB_BLOCK / 32][64];
  uint32_t *parked = parkedAll[threadIdx.x >> 5];
  uint32_t nParked = 0;   // warp uniform
  const uint32_t lane = lane_id();
  const uint32_t n = *nPtr;
  uint32_t shaded = 0;
  for (;;) {
    uint32_t chunk = warp_claim_n(work, PB_CLAIM);
    if (chunk >= n) break;
    for (uint32_t base = chunk; base < chunk + PB_CLAIM && base < n; base += 32) {
      const uint32_t i = base + lane;
      const bool valid = i < n;
      bool pass = false, plain = false;
      if (valid) {
        const int32_t prim = qin.hitPrim[i];
        if (prim >= 0) {
          const float4 tail = __ldg((const float4 *)(sv.prims + 3u * (uint32_t)prim + 2));
          pass = !(__float_as_uint(tail.y) & PB_PF_DEV_FRAME) || (qin.hitInst && qin.hitInst[i] >= 0) ||
                 !material_is_lambert(sv.materials[__float_as_uint(tail.z)]);
          plain = !pass;
        } else {
          // the ray left the scene (path.cpp:69-82): fold the pending direct light, emitted light of infinite lights
          const uint32_t slot = qin.slot[i];
          const float4 lf = wb.plf[slot];
          uint32_t flags = __float_as_uint(lf.w);
          PathState st;
          st.L = spec3(lf.x, lf.y, lf.z);
          path_fold_pending(sv, wb, slot, &flags, &st.L);
          st.depth = (int)(flags & 0xffu);
          st.specularBounce = (flags >> 8) & 1u;
          st.alpha = 0.f;
          st.T = spec(0.f);
          if (st.depth == 0 || st.specularBounce) { const float4 tc = wb.ptc[slot]; st.T = spec3(tc.x, tc.y, tc.z); }
          Ray ray = queue_ray(qin, i);
          Hit miss; miss.prim = -1; miss.t = 0.f; miss.b1 = miss.b2 = 0.f; miss.inst = -1;
          Surface none;
          path_vertex_surface<false, false, false, true>(sv, st, ray, miss, &none);
          wb.plf[slot] = make_float4(st.L.c[0], st.L.c[1], st.L.c[2], __uint_as_float(flags));
          if (st.depth == 0) wb.alpha[slot] = st.alpha;
        }
      }
      uint32_t pj = warp_push(subsetCount, pass);
      if (pass) subset[pj] = i;
      const uint32_t mask = __ballot_sync(0xffffffffu, plain);
      if (mask) {
        if (plain) {
          parked[nParked + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = i;
#if PB_FAST_PREFETCH
          // the records the vertex will read, on their way while the hit waits in the list: the slot's two 16-byte
          // state records and the primitive's 48-byte frame are scattered reads
          const uint32_t pslot = qin.slot[i];
          asm volatile("prefetch.global.L1 [%0];" ::"l"(wb.plf + pslot));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(wb.ptc + pslot));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(sv.primFrame + 3u * (uint32_t)qin.hitPrim[i]));
#endif
        }
        nParked += (uint32_t)__popc(mask);
        __syncwarp();
        if (nParked >= 32u) {
          nParked -= 32u;
          const uint32_t pi = parked[nParked + lane];
          __syncwarp();
          path_fast_vertex<LD, LK>(sv, wb, wd, qin, qout, pi, true, shadowCount, probeCount, nextCount, shaded);
        }
      }
    }
  }
  if (nParked) {
    const uint32_t pi = lane < nParked ? parked[lane] : 0u;
    path_fast_vertex<LD, LK>(sv, wb, wd, qin, qout, pi, lane < nParked, shadowCount, probeCount, nextCount, shaded);
  }
  warp_stat_add(&wb.stats[ST_SHADED], shaded);
}

// LD: patterns come from the lowdiscrepancy tables; G: the scene has glossy (microfacet) materials.
// false folds that code away, so the common matte/glass scenes do not carry it.
template <bool LD, bool G, int LK>
__global__ void __launch_bounds__(PB_BLOCK, PB_DIRECT_MIN_BLOCKS)
k_path_direct(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveKeyInfo wd,
              const RayQueue qin, const SurfaceRecord rec, const uint32_t *nPtr, uint32_t *work,
              uint32_t *shadowCount, uint32_t *probeCount, const RayQueue qout, uint32_t *nextCount) {
  uint32_t n = *nPtr;
  for (;;) {
    uint32_t chunk = warp_claim_n(work, PB_CLAIM);
    if (chunk >= n) break;
    for (uint32_t base = chunk; base < chunk + PB_CLAIM && base < n; base += 32) {
      uint32_t j = base + lane_id();
      bool valid = j < n;
      uint32_t i = valid ? rec.queueIdx[j] : 0u;
      bool pushShadow = false, pushProbe = false;
      uint32_t slot = 0;
      DirectRec direct;
#if PB_FUSE_DIFFUSE
      bool pushNext = false;
      Ray next;
#endif
      if (valid) {
        slot = qin.slot[i];
        Ray ray;
        ray.o = v3(0, 0, 0);
        ray.d = v3(qin.dx[i], qin.dy[i], qin.dz[i]);   // only wo = -ray.d is needed
        ray.mint = 0.f; ray.maxt = 0.f;
        const float4 tc = wb.ptc[slot];
        uint32_t flags = __float_as_uint(wb.plf[slot].w);
        PathState st;
        st.T = spec3(tc.x, tc.y, tc.z);
        st.depth = (int)(flags & 0xffu);
        st.rng.key = slot_key(wd, slot, 1);
        st.rng.ctr = __float_as_uint(tc.w);
        PatternSource key0 = {slot_key(wd, slot, 0), LD ? wb.ld : (const float *)nullptr, wb.nSlots, slot, LD ? wd.bc : (wd.bc == 2 ? 2 : 0)};
        bool hasDirect;
        const PbrtB200Material &mat = rec.mat ? rec.mat[j] : sv.materials[rec.ids[j] & 0xffffu];
        if (material_is_lambert(mat)) {
          LambertSurface ls;
          uint32_t ids = surface_load_frame(rec, j, &ls.bsdf);
          ls.material = ids & 0xffffu; ls.light = (int32_t)(ids >> 16) - 1;
          ls.bsdf.R = spec3(mat.kd[0], mat.kd[1], mat.kd[2]) * PB_INV_PI;
          path_vertex_direct<G, LK>(sv, key0, st, ray, &ls, &direct, &hasDirect);
#if PB_FUSE_DIFFUSE
          st.specularBounce = (flags >> 8) & 1u;
          pushNext = path_vertex_continue<G>(sv, key0, st, ray, ls, &next);
#endif
        } else {
          Surface sf;
          surface_load<G>(sv, rec, j, &sf);
          path_vertex_direct<G, LK>(sv, key0, st, ray, &sf, &direct, &hasDirect);
#if PB_FUSE_DIFFUSE
          st.specularBounce = (flags >> 8) & 1u;
          pushNext = path_vertex_continue<G>(sv, key0, st, ray, sf, &next);
#endif
        }
#if PB_FUSE_DIFFUSE
        // (the continuation of path.cpp:113-143 in the same pass: throughput, draw count and depth go back together)
        wb.ptc[slot] = make_float4(st.T.c[0], st.T.c[1], st.T.c[2], __uint_as_float(st.rng.ctr));
        flags = (uint32_t)st.depth | (st.specularBounce ? (1u << 8) : 0u);
        reinterpret_cast<uint32_t *>(&wb.plf[slot])[3] = flags;
#else
        if (st.rng.ctr != __float_as_uint(tc.w)) reinterpret_cast<uint32_t *>(&wb.ptc[slot])[3] = st.rng.ctr;
#endif
        if (hasDirect) {
          pushShadow = direct.hasShadow;
          pushProbe = direct.hasProbe;
          if (pushProbe) {
            // rare: the estimator's value also depends on the MIS probe -> keep both terms pending
            reinterpret_cast<uint32_t *>(&wb.plf[slot])[3] = flags | (1u << 9);
            store3(wb.pendT, slot, st.pendT);
            store3(wb.recB, slot, direct.termB);
            wb.prec[slot] = make_float4(direct.termA.c[0], direct.termA.c[1], direct.termA.c[2],
                                        __uint_as_float((direct.hasShadow ? 4u : 0u) | 8u));
          } else if (pushShadow) {
            // common: only the shadow ray decides, so the addend of `L += pathThroughput *
            // UniformSampleOneLight(..)` (path.cpp:100-112) is known now; one 16-byte record stays pending
            Spec Ld = direct_value(direct.termA, direct.termB, true, false);
            Spec add = st.pendT * ((float)sv.nLights * Ld);
            reinterpret_cast<uint32_t *>(&wb.plf[slot])[3] = flags | (1u << 9);
            wb.prec[slot] = make_float4(add.c[0], add.c[1], add.c[2],
                                        __uint_as_float(4u | 16u | (pb_bad_channels(st.pendT) << 5)));
          } else if (const uint32_t bad = pb_bad_channels(st.pendT)) {
            // black estimate, no ray, and a throughput that is not finite: L += pathThroughput * 0 is a NaN
            // (rare: L lives in the slot's record, which this kernel otherwise only reads the flag word of)
            const float4 lf = wb.plf[slot];
            Spec L = spec3(lf.x, lf.y, lf.z);
            pb_add_bad_zero(&L, bad);
            wb.plf[slot] = make_float4(L.c[0], L.c[1], L.c[2], lf.w);
          }
        }
      }
      uint32_t si = warp_push(shadowCount, pushShadow);
      if (pushShadow) aux_put(wb.shadowQ, si, slot, direct.shadow, 0);
      uint32_t pi = warp_push(probeCount, pushProbe);
      if (pushProbe) aux_put(wb.probeQ, pi, slot, direct.probe, direct.probeLight);
#if PB_FUSE_DIFFUSE
      uint32_t ni = warp_push(nextCount, pushNext);
      if (pushNext) queue_put(qout, ni, slot, next);
#endif
  }
  }
}

template <bool LD, bool G>
__global__ void __launch_bounds__(PB_BLOCK, PB_CONT_MIN_BLOCKS)
k_path_continue(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveKeyInfo wd,
                const RayQueue qin, const RayQueue qout, const SurfaceRecord rec, const uint32_t *nPtr,
                const uint32_t *nSpecPtr, uint32_t *work, uint32_t *nextCount) {
  const uint32_t nDiff = *nPtr;
  const uint32_t n = nDiff + *nSpecPtr;
  // (PB_FUSE_DIFFUSE: k_path_direct already continued the diffuse-class vertices; only the specular-only ones are left)
  const uint32_t first = PB_FUSE_DIFFUSE ? nDiff : 0u;
  for (;;) {
    uint32_t chunk = first + warp_claim_n(work, PB_CLAIM);
    if (chunk >= n) break;
    for (uint32_t base = chunk; base < chunk + PB_CLAIM && base < n; base += 32) {
      uint32_t w = base + lane_id();
      bool valid = w < n;
      // records [0, nDiff) from the front, then the specular-only ones from the back of the arrays
      const bool specOnly = valid && w >= nDiff;
      uint32_t j = specOnly ? wb.nSlots - 1u - (w - nDiff) : w;
      uint32_t i = valid ? rec.queueIdx[j] : 0u;
      bool pushNext = false;
      uint32_t slot = 0;
      Ray next;
      if (valid) {
        slot = qin.slot[i];
        Ray ray;
        ray.o = v3(0, 0, 0);
        ray.d = v3(qin.dx[i], qin.dy[i], qin.dz[i]);
        ray.mint = 0.f; ray.maxt = 0.f;
        const float4 tc = wb.ptc[slot];
        const uint32_t flags = __float_as_uint(wb.plf[slot].w);
        PathState st;
        st.T = spec3(tc.x, tc.y, tc.z);
        st.depth = (int)(flags & 0xffu);
        st.specularBounce = (flags >> 8) & 1u;
        st.rng.key = slot_key(wd, slot, 1);
        st.rng.ctr = __float_as_uint(tc.w);
        PatternSource key0 = {slot_key(wd, slot, 0), LD ? wb.ld : (const float *)nullptr, wb.nSlots, slot, LD ? wd.bc : (wd.bc == 2 ? 2 : 0)};
        const PbrtB200Material &mat = rec.mat ? rec.mat[j] : sv.materials[rec.ids[j] & 0xffffu];
        if (!specOnly && material_is_lambert(mat)) {
          LambertSurface ls;
          uint32_t ids = surface_load_frame(rec, j, &ls.bsdf);
          ls.material = ids & 0xffffu; ls.light = (int32_t)(ids >> 16) - 1;
          ls.bsdf.R = spec3(mat.kd[0], mat.kd[1], mat.kd[2]) * PB_INV_PI;
          pushNext = path_vertex_continue<G>(sv, key0, st, ray, ls, &next);
        } else {
          Surface sf;
          surface_load<G>(sv, rec, j, &sf);
          if (specOnly) {
            path_vertex_direct_skip(sv, key0, st);   // k_path_direct never saw this vertex
            // (its estimate is black; the reference still added pathThroughput * 0 to L)
            if (sv.nLights != 0)
              if (const uint32_t bad = pb_bad_channels(st.T)) {
                const float4 lf = wb.plf[slot];
                Spec L = spec3(lf.x, lf.y, lf.z);
                pb_add_bad_zero(&L, bad);
                wb.plf[slot] = make_float4(L.c[0], L.c[1], L.c[2], lf.w);
              }
          }
          pushNext = path_vertex_continue<G>(sv, key0, st, ray, sf, &next);
        }
        wb.ptc[slot] = make_float4(st.T.c[0], st.T.c[1], st.T.c[2], __uint_as_float(st.rng.ctr));
        reinterpret_cast<uint32_t *>(&wb.plf[slot])[3] =
            (flags & (1u << 9)) | (uint32_t)st.depth | (st.specularBounce ? (1u << 8) : 0u);
      }
      uint32_t ni = warp_push(nextCount, pushNext);
      if (pushNext) queue_put(qout, ni, slot, next);
  }
  }
}

// ------------------------------------------------------------------------------------------
// Whitted / DirectLighting: frames in SoA
#define PB_FRAME_FLOATS 21
struct DevFrameIO {
  const WaveBuffers *wb;
  uint32_t slot;
  __device__ float *fp(int field, int d) const {
    return wb->frame + ((size_t)field * wb->nFrames + d) * wb->nSlots + slot;
  }
  __device__ float *dp(int field, int d) const {
    return wb->frameDiff + ((size_t)field * wb->nFrames + d) * wb->nSlots + slot;
  }
  // child 0: reflection, 1: transmission
  __device__ void load_diff(int d, int child, RayDiff *r) const {
    int k = 12 * child;
    r->has = true;
    r->rxo.x = *dp(k++, d); r->rxo.y = *dp(k++, d); r->rxo.z = *dp(k++, d);
    r->rxd.x = *dp(k++, d); r->rxd.y = *dp(k++, d); r->rxd.z = *dp(k++, d);
    r->ryo.x = *dp(k++, d); r->ryo.y = *dp(k++, d); r->ryo.z = *dp(k++, d);
    r->ryd.x = *dp(k++, d); r->ryd.y = *dp(k++, d); r->ryd.z = *dp(k++, d);
  }
  __device__ void store_diff(int d, int child, const RayDiff &r) const {
    int k = 12 * child;
    *dp(k++, d) = r.rxo.x; *dp(k++, d) = r.rxo.y; *dp(k++, d) = r.rxo.z;
    *dp(k++, d) = r.rxd.x; *dp(k++, d) = r.rxd.y; *dp(k++, d) = r.rxd.z;
    *dp(k++, d) = r.ryo.x; *dp(k++, d) = r.ryo.y; *dp(k++, d) = r.ryo.z;
    *dp(k++, d) = r.ryd.x; *dp(k++, d) = r.ryd.y; *dp(k++, d) = r.ryd.z;
  }
  __device__ WdFrame load(int d) const {
    WdFrame f;
    int k = 0;
    for (int c = 0; c < 3; ++c) f.L.c[c] = *fp(k++, d);
    f.p.x = *fp(k++, d); f.p.y = *fp(k++, d); f.p.z = *fp(k++, d);
    f.wiR.x = *fp(k++, d); f.wiR.y = *fp(k++, d); f.wiR.z = *fp(k++, d);
    f.wiT.x = *fp(k++, d); f.wiT.y = *fp(k++, d); f.wiT.z = *fp(k++, d);
    for (int c = 0; c < 3; ++c) f.fR.c[c] = *fp(k++, d);
    for (int c = 0; c < 3; ++c) f.fT.c[c] = *fp(k++, d);
    f.aR = *fp(k++, d); f.aT = *fp(k++, d);
    uint32_t fl = wb->frameFlags[(size_t)d * wb->nSlots + slot];
    f.stage = (int)(fl & 0xffu); f.reflValid = (fl >> 8) & 1u; f.transValid = (fl >> 9) & 1u;
    return f;
  }
  __device__ void store(int d, const WdFrame &f) const {
    int k = 0;
    for (int c = 0; c < 3; ++c) *fp(k++, d) = f.L.c[c];
    *fp(k++, d) = f.p.x; *fp(k++, d) = f.p.y; *fp(k++, d) = f.p.z;
    *fp(k++, d) = f.wiR.x; *fp(k++, d) = f.wiR.y; *fp(k++, d) = f.wiR.z;
    *fp(k++, d) = f.wiT.x; *fp(k++, d) = f.wiT.y; *fp(k++, d) = f.wiT.z;
    for (int c = 0; c < 3; ++c) *fp(k++, d) = f.fR.c[c];
    for (int c = 0; c < 3; ++c) *fp(k++, d) = f.fT.c[c];
    *fp(k++, d) = f.aR; *fp(k++, d) = f.aT;
    wb->frameFlags[(size_t)d * wb->nSlots + slot] =
        (uint32_t)f.stage | (f.reflValid ? (1u << 8) : 0u) | (f.transValid ? (1u << 9) : 0u);
  }
};

#define PB_MAX_WD_RECS 16   // per-thread staging of direct records in the wd kernels

// K2 (whitted / directlighting): vertex body up to the specular recursion
template <bool TEX>
__global__ void __launch_bounds__(PB_BLOCK)
k_shade_wd(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveKeyInfo wd,
           const RayQueue qin, const uint32_t *nPtr, uint32_t *work, uint32_t *shadowCount,
           uint32_t *probeCount) {
  uint32_t n = *nPtr;
  uint32_t shaded = 0;
  for (;;) {
    uint32_t chunk = warp_claim_n(work, PB_CLAIM);
    if (chunk >= n) break;
    for (uint32_t base = chunk; base < chunk + PB_CLAIM && base < n; base += 32) {
      uint32_t i = base + lane_id();
      bool valid = i < n;
      DirectRec recs[PB_MAX_WD_RECS];
      int nRecs = 0;
      uint32_t slot = 0;
      if (valid) {
        slot = qin.slot[i];
        Ray ray = queue_ray(qin, i);
        Hit hit;
        hit.prim = qin.hitPrim[i];
        hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.inst = -1;
        if (hit.prim >= 0) {
          hit.t = qin.hitT[i]; hit.b1 = qin.hitB1[i]; hit.b2 = qin.hitB2[i];
          if (qin.hitInst) hit.inst = qin.hitInst[i];
        }
        uint32_t flags = wb.flags[slot];
        int depth = (int)(flags & 0xffu);
        RngStream rng;
        rng.key = slot_key(wd, slot, 1);
        rng.ctr = wb.ctr[slot];
        PatternSource key0 = {slot_key(wd, slot, 0), wb.ld, wb.nSlots, slot, wd.bc};
        float alpha = 0.f;
        WdFrame fr;
        WdVertexOut out;
        DevFrameIO io = {&wb, slot};
        if (TEX) {
          // the ray into depth d is the reflection (parent stage 0) or transmission child of frame d - 1
          RayDiff rd;
          rd.has = true;
          if (depth == 0) cam_diff_load(wb, slot, &rd);
          else io.load_diff(depth - 1, (wb.frameFlags[(size_t)(depth - 1) * wb.nSlots + slot] & 0xffu) == 0 ? 0 : 1, &rd);
          wd_vertex<true>(sv, key0, ray, hit, rng, depth, depth == 0 ? &alpha : (float *)0, &fr, recs, &out, &rd);
        } else
          wd_vertex<false>(sv, key0, ray, hit, rng, depth, depth == 0 ? &alpha : (float *)0, &fr, recs, &out);
        if (hit.prim >= 0) ++shaded;
        if (depth == 0) wb.alpha[slot] = alpha;
        wb.ctr[slot] = rng.ctr;
        nRecs = out.nRecs;
        wb.nRecs[slot] = (uint32_t)nRecs;
        io.store(depth, fr);
        if (TEX) {
          if (fr.reflValid) io.store_diff(depth, 0, fr.dR);
          if (fr.transValid) io.store_diff(depth, 1, fr.dT);
        }
        for (int r = 0; r < nRecs; ++r) {
          size_t ri = (size_t)r * wb.nSlots + slot;
          wb.recA[0][ri] = recs[r].termA.c[0]; wb.recA[1][ri] = recs[r].termA.c[1]; wb.recA[2][ri] = recs[r].termA.c[2];
          wb.recB[0][ri] = recs[r].termB.c[0]; wb.recB[1][ri] = recs[r].termB.c[1]; wb.recB[2][ri] = recs[r].termB.c[2];
          wb.recFlags[ri] = (recs[r].hasShadow ? 4 : 0) | (recs[r].hasProbe ? 8 : 0);
        }
      }
      int maxRecs = nRecs;
  #pragma unroll
      for (int o = 16; o > 0; o >>= 1) maxRecs = max(maxRecs, __shfl_xor_sync(0xffffffffu, maxRecs, o));
      for (int r = 0; r < maxRecs; ++r) {
        bool ps = valid && r < nRecs && recs[r].hasShadow;
        bool pp = valid && r < nRecs && recs[r].hasProbe;
        uint32_t ri = (uint32_t)((size_t)r * wb.nSlots + slot);
        uint32_t si = warp_push(shadowCount, ps);
        if (ps) aux_put(wb.shadowQ, si, ri, recs[r].shadow, 0);
        uint32_t pi = warp_push(probeCount, pp);
        if (pp) aux_put(wb.probeQ, pi, ri, recs[r].probe, recs[r].probeLight);
      }
  }
  }
  warp_stat_add(&wb.stats[ST_SHADED], shaded);
}

// fold direct lighting into the frame, then continue the depth-first walk
__global__ void __launch_bounds__(PB_BLOCK)
k_advance_wd(const __grid_constant__ SceneView sv, const WaveBuffers wb, const RayQueue qin,
             const RayQueue qout, const uint32_t *nPtr, uint32_t *work, uint32_t *nextCount) {
  uint32_t n = *nPtr;
  for (;;) {
    uint32_t chunk = warp_claim_n(work, PB_CLAIM);
    if (chunk >= n) break;
    for (uint32_t base = chunk; base < chunk + PB_CLAIM && base < n; base += 32) {
      uint32_t i = base + lane_id();
      bool valid = i < n;
      bool pushNext = false;
      Ray next;
      uint32_t slot = 0;
      if (valid) {
        slot = qin.slot[i];
        uint32_t flags = wb.flags[slot];
        int depth = (int)(flags & 0xffu);
        DevFrameIO io = {&wb, slot};
        WdFrame fr = io.load(depth);
        int nRecs = (int)wb.nRecs[slot];
        DirectRec recs[PB_MAX_WD_RECS];
        uint8_t rflags[PB_MAX_WD_RECS];
        for (int r = 0; r < nRecs; ++r) {
          size_t ri = (size_t)r * wb.nSlots + slot;
          recs[r].termA = spec3(wb.recA[0][ri], wb.recA[1][ri], wb.recA[2][ri]);
          recs[r].termB = spec3(wb.recB[0][ri], wb.recB[1][ri], wb.recB[2][ri]);
          uint8_t f = wb.recFlags[ri];
          recs[r].hasShadow = (f & 4) != 0; recs[r].hasProbe = (f & 8) != 0;
          rflags[r] = f & 3;
        }
        wd_resolve(sv, &fr, recs, rflags, nRecs);
        io.store(depth, fr);
        RngStream rng;
        rng.key = 0; rng.ctr = wb.ctr[slot];
        pushNext = wd_advance(io, rng, &depth, &next);
        wb.ctr[slot] = rng.ctr;
        wb.flags[slot] = (flags & ~0xffu) | (uint32_t)depth;
      }
      uint32_t ni = warp_push(nextCount, pushNext);
      if (pushNext) queue_put(qout, ni, slot, next);
  }
  }
}

// ------------------------------------------------------------------------------------------
// K5a: end of Li -> guards of Scene::Render (core/scene.cpp:58-74)
__global__ void __launch_bounds__(PB_BLOCK)
k_finalize(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WaveDesc wd) {
  uint32_t nan = 0, neg = 0, inf = 0;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < wd.nSlots;
       slot += gridDim.x * blockDim.x) {
    Spec L;
    if ((wd.bc || wd.tiles) && wb.alpha[slot] == -1.f) {   // a table entry outside the extent / a culled apron sample: no sample
      store3(wb.outL, slot, spec(0.f));
      wb.outA[slot] = -1.f;
      continue;
    }
    if (sv.integrator.kind == PBRT_B200_INT_PATH) {
      const float4 lf = wb.plf[slot];
      uint32_t flags = __float_as_uint(lf.w);
      L = spec3(lf.x, lf.y, lf.z);
      path_fold_pending(sv, wb, slot, &flags, &L);
    } else {
      L = spec3(wb.frame[slot], wb.frame[(size_t)wb.nFrames * wb.nSlots + slot],
                wb.frame[(size_t)2 * wb.nFrames * wb.nSlots + slot]);   // frame 0, fields 0..2
    }
    Spec Ls = 1.f * L;   // rayWeight * Li
    int g = sample_guard(&Ls);
    nan += g == 1; neg += g == 2; inf += g == 3;
    store3(wb.outL, slot, Ls);
    wb.outA[slot] = wb.alpha[slot];
  }
  warp_stat_add(&wb.stats[ST_NAN], nan);
  warp_stat_add(&wb.stats[ST_NEG], neg);
  warp_stat_add(&wb.stats[ST_INF], inf);
}

// Pixel footprint of each sample pixel of the wave: [min x0, max x1] x [min y0, max y1] over its spp
// samples (x0..y1 as in ImageFilm::AddSample, film/image.cpp:106-111), packed 4 x int16.  Lets the
// gather skip whole sample pixels: with the 0.5-wide box filter a neighbouring sample pixel reaches
// this pixel only through a sample whose coordinate is exactly an integer.
__global__ void k_film_reach(const __grid_constant__ SceneView sv, const WaveBuffers wb, uint32_t nPix, int spp) {
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < nPix; p += gridDim.x * blockDim.x) {
    int x0 = 1 << 20, x1 = -(1 << 20), y0 = 1 << 20, y1 = -(1 << 20);
    uint32_t s0 = p * (uint32_t)spp;
    for (int k = 0; k < spp; ++k) {
      if (wb.imageX[s0 + k] == PB_BC_INVALID_COORD) continue;   // culled (tile mode)
      float dImageX = wb.imageX[s0 + k] - 0.5f, dImageY = wb.imageY[s0 + k] - 0.5f;
      x0 = min(x0, pb_ceil2int_fin(dImageX - sv.filterXW)); x1 = max(x1, pb_floor2int_fin(dImageX + sv.filterXW));
      y0 = min(y0, pb_ceil2int_fin(dImageY - sv.filterYW)); y1 = max(y1, pb_floor2int_fin(dImageY + sv.filterYW));
    }
    auto pack = [](int v) { return (unsigned long long)(uint16_t)(max(-32768, min(32767, v)) + 32768); };
    wb.reach[p] = pack(x0) | (pack(x1) << 16) | (pack(y0) << 32) | (pack(y1) << 48);
  }
}

// K5b: ImageFilm::AddSample (film/image.cpp:103-141) as a gather, one thread per pixel.
// Candidate sample pixels are visited in raster order and their samples in index order:
// the order in which the reference's serial loop would have splatted them.
__device__ __forceinline__ void film_gather_part(const SceneView &sv, const WaveBuffers &wb, const WavePart &wd, int spp,
                                                 const float *table, FilmPixel *film, const PbrtB200Window &win,
                                                 int rowFirst, int rowCount) {
  int total = rowCount * win.width;
  int ew = wd.xe - wd.xs;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int px = win.x0 + idx % win.width;
    int py = rowFirst + idx / win.width;
    FilmPixel *fpx = &film[(size_t)(py - win.y0) * win.width + (px - win.x0)];
    FilmPixel acc = *fpx;
    int sx0 = max(pb_floor2int_fin(px + .5f - sv.filterXW) - 1, wd.xs);
    int sx1 = min(pb_floor2int_fin(px + .5f + sv.filterXW) + 1, wd.xe - 1);
    int sy0 = max(pb_floor2int_fin(py + .5f - sv.filterYW) - 1, wd.ys);
    int sy1 = min(pb_floor2int_fin(py + .5f + sv.filterYW) + 1, wd.ye - 1);
    bool touched = false;
    for (int sy = sy0; sy <= sy1; ++sy)
      for (int sx = sx0; sx <= sx1; ++sx) {
        int64_t pix = (int64_t)(sy - wd.ys) * ew + (sx - wd.xs) - (int64_t)wd.firstPixel;
        if (pix < 0 || pix >= (int64_t)wd.nPixels) continue;
        uint32_t s0 = wd.slotBase + (uint32_t)pix * (uint32_t)spp;
        unsigned long long rc = wb.reach[s0 / (uint32_t)spp];
        if (px < (int)(rc & 0xffffu) - 32768 || px > (int)((rc >> 16) & 0xffffu) - 32768 ||
            py < (int)((rc >> 32) & 0xffffu) - 32768 || py > (int)(rc >> 48) - 32768)
          continue;
        for (int k = 0; k < spp; ++k) {
          uint32_t slot = s0 + k;
          float wt;
          if (!film_sample_weight(sv, table, wb.imageX[slot], wb.imageY[slot], px, py, &wt)) continue;
          float L[3] = {wb.outL[0][slot], wb.outL[1][slot], wb.outL[2][slot]};
          film_pixel_add(&acc, wt, L, wb.outA[slot]);
          touched = true;
        }
      }
    if (touched) *fpx = acc;
  }
}
__global__ void __launch_bounds__(PB_BLOCK)
k_film_gather(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WavePart wd, int spp,
              const float *filterTable, FilmPixel *film, PbrtB200Window win, int rowFirst, int rowCount) {
  __shared__ float table[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) table[t] = filterTable[t];
  __syncthreads();
  film_gather_part(sv, wb, wd, spp, table, film, win, rowFirst, rowCount);
}
// The same for every part of a wave in one launch (blockIdx.y = part): a multi-GPU rank's share of a frame is ~30
// small tiles in one wave, and one launch per tile made the film the only kernel class that did not shrink with the
// rank count (8 GPUs: 1.87 ms of a 17.9 ms frame share, profiles/r02s_scale_n8.json).
struct FilmTargets {
  FilmPixel *film[PB_MAX_PARTS];
  PbrtB200Window win[PB_MAX_PARTS];
  int rowFirst[PB_MAX_PARTS], rowCount[PB_MAX_PARTS];
};
__global__ void __launch_bounds__(PB_BLOCK)
k_film_gather_parts(const __grid_constant__ SceneView sv, const WaveBuffers wb, const __grid_constant__ WaveDesc wd,
                    const float *filterTable, const __grid_constant__ FilmTargets tg) {
  __shared__ float table[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) table[t] = filterTable[t];
  __syncthreads();
  const int p = blockIdx.y;
  if (tg.rowCount[p] > 0)
    film_gather_part(sv, wb, wd.parts[p], wd.spp, table, tg.film[p], tg.win[p], tg.rowFirst[p], tg.rowCount[p]);
}

// K5b for the bestcandidate sampler: the samples of a table placement are spread over its whole
// area, so a pixel visits the placements of the part in order and, for those whose (filter-grown)
// area contains it, all 4096 entries in order -- the order the reference's serial loop splats them.
// The table's entries never move inside a placement, so which of the 4096 can reach a pixel is first
// narrowed with bit masks: the unit square is cut into PB_BC_STRIPS strips per axis and
// bcMasks[axis][s + 1] has bit k set when entry k lies in a strip <= s (prefix masks, [0] = none);
// the entries whose x (y) strip is in [lo, hi] are prefix[hi + 1] & ~prefix[lo].  Survivors take the
// exact test in ascending entry order.
#define PB_BC_STRIPS 64
#define PB_BC_WORDS (PB_BC_TABLE / 32)
__global__ void __launch_bounds__(PB_BLOCK)
k_film_gather_bc(const __grid_constant__ SceneView sv, const WaveBuffers wb, const WavePart wd,
                 const float *filterTable, const uint32_t *bcMasks, FilmPixel *film, PbrtB200Window win,
                 int rowFirst, int rowCount) {
  __shared__ float table[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) table[t] = filterTable[t];
  __syncthreads();
  int total = rowCount * win.width;
  const float tw = sv.bcTableWidth;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int px = win.x0 + idx % win.width;
    int py = rowFirst + idx / win.width;
    FilmPixel *fpx = &film[(size_t)(py - win.y0) * win.width + (px - win.x0)];
    FilmPixel acc = *fpx;
    bool touched = false;
    // candidate placements: on a regular grid only the few columns / rows around the pixel
    uint32_t tBegin = 0, tEnd = wd.nPixels;
    int colLo = 0, colHi = 0x7fffffff;
    if (wd.bcCols > 0) {
      float reachX = sv.filterXW + 2.5f, reachY = sv.filterYW + 2.5f;
      colLo = max(0, (int)floorf((px - reachX - (float)wd.xs) / tw) - 1);
      colHi = min(wd.bcCols - 1, (int)floorf((px + reachX - (float)wd.xs) / tw) + 1);
      int rowLo = max(0, (int)floorf((py - reachY - (float)wd.ys) / tw) - 1);
      int rowHi = (int)floorf((py + reachY - (float)wd.ys) / tw) + 1;
      long long b = (long long)rowLo * wd.bcCols - (long long)wd.firstPixel;
      long long e = ((long long)rowHi + 1) * wd.bcCols - (long long)wd.firstPixel;
      tBegin = (uint32_t)max(0ll, min(b, (long long)wd.nPixels));
      tEnd = (uint32_t)max(0ll, min(e, (long long)wd.nPixels));
    }
    for (uint32_t t = tBegin; t < tEnd; ++t) {
      if (wd.bcCols > 0) {
        int col = (int)((wd.firstPixel + t) % (uint32_t)wd.bcCols);
        if (col < colLo) { t += (uint32_t)(colLo - col) - 1u; continue; }
        if (col > colHi) { t += (uint32_t)(wd.bcCols - col) - 1u; continue; }
      }
      float2 c = wd.corners[wd.firstPixel + t];
      // sample coordinates of the placement lie in [c, c + tw]; a sample reaches pixels within the
      // filter width of (coordinate - 0.5)
      if (px < c.x - 1.5f - sv.filterXW || px > c.x + tw + 0.5f + sv.filterXW ||
          py < c.y - 1.5f - sv.filterYW || py > c.y + tw + 0.5f + sv.filterYW)
        continue;
      uint32_t s0 = wd.slotBase + t * PB_BC_TABLE;
      // table-space interval of the sample coordinates that can reach the pixel, one strip of slack
      float invTw = 1.f / tw;
      int sxLo = (int)floorf((px + 0.5f - sv.filterXW - c.x) * invTw * PB_BC_STRIPS) - 1;
      int sxHi = (int)floorf((px + 0.5f + sv.filterXW - c.x) * invTw * PB_BC_STRIPS) + 1;
      int syLo = (int)floorf((py + 0.5f - sv.filterYW - c.y) * invTw * PB_BC_STRIPS) - 1;
      int syHi = (int)floorf((py + 0.5f + sv.filterYW - c.y) * invTw * PB_BC_STRIPS) + 1;
      if (sxHi < 0 || syHi < 0 || sxLo >= PB_BC_STRIPS || syLo >= PB_BC_STRIPS) continue;
      sxLo = max(sxLo, 0); syLo = max(syLo, 0);
      sxHi = min(sxHi, PB_BC_STRIPS - 1); syHi = min(syHi, PB_BC_STRIPS - 1);
      const uint32_t *xLoM = bcMasks + (size_t)sxLo * PB_BC_WORDS, *xHiM = bcMasks + (size_t)(sxHi + 1) * PB_BC_WORDS;
      const uint32_t *yBase = bcMasks + (size_t)(PB_BC_STRIPS + 1) * PB_BC_WORDS;
      const uint32_t *yLoM = yBase + (size_t)syLo * PB_BC_WORDS, *yHiM = yBase + (size_t)(syHi + 1) * PB_BC_WORDS;
      for (uint32_t w = 0; w < PB_BC_WORDS; ++w) {
        uint32_t m = (__ldg(xHiM + w) & ~__ldg(xLoM + w)) & (__ldg(yHiM + w) & ~__ldg(yLoM + w));
        while (m) {
          uint32_t slot = s0 + 32u * w + (uint32_t)(__ffs(m) - 1);
          m &= m - 1u;
          float wt;
          if (!film_sample_weight(sv, table, wb.imageX[slot], wb.imageY[slot], px, py, &wt)) continue;
          float L[3] = {wb.outL[0][slot], wb.outL[1][slot], wb.outL[2][slot]};
          film_pixel_add(&acc, wt, L, wb.outA[slot]);
          touched = true;
        }
      }
    }
    if (touched) *fpx = acc;
  }
}

// K6: ImageFilm::WriteImage up to the WriteRGBAImage call (film/image.cpp:157-203)
__global__ void k_film_resolve(const FilmPixel *film, float4 *rgba, int n, int premultiply) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float o[4];
    film_pixel_resolve(film[i], premultiply != 0, o);
    rgba[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// stand-alone ray kernels behind pbrt_b200_intersect / _intersect_p
__global__ void k_rays_closest(const __grid_constant__ SceneView sv, const float *rays, uint64_t n,
                               int32_t *prim, float *t, float *b1, float *b2) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Ray r;
    r.o = v3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]);
    r.d = v3(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5]);
    r.mint = rays[8 * i + 6]; r.maxt = rays[8 * i + 7];
    Hit h;
    kd_intersect<false, false>(sv, r, &h, (KdCounters *)0);
    prim[i] = h.prim;
    t[i] = h.prim >= 0 ? h.t : 0.f; b1[i] = h.prim >= 0 ? h.b1 : 0.f; b2[i] = h.prim >= 0 ? h.b2 : 0.f;
  }
}
__global__ void k_rays_any(const __grid_constant__ SceneView sv, const float *rays, uint64_t n, uint8_t *occ) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    Ray r;
    r.o = v3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]);
    r.d = v3(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5]);
    r.mint = rays[8 * i + 6]; r.maxt = rays[8 * i + 7];
    Hit h;
    occ[i] = kd_intersect<true, false>(sv, r, &h, (KdCounters *)0) ? 1 : 0;
  }
}
__global__ void k_camera_rays(const __grid_constant__ SceneView sv, const WavePart wd, int spp, uint64_t n,
                              float *rays, float *xy) {
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < n;
       s += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t pix = s / spp;
    int k = (int)(s % spp);
    int ew = wd.xe - wd.xs;
    CameraSample cs;
    stratified_camera_sample(sv, wd.xs + (int)(pix % ew), wd.ys + (int)(pix / ew), k, &cs);
    Ray r;
    perspective_generate_ray(sv.camera, cs, &r, sv.xRes, sv.yRes);
    float *o = rays + 8 * s;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
    o[6] = r.mint; o[7] = r.maxt;
    xy[2 * s] = cs.imageX; xy[2 * s + 1] = cs.imageY;
  }
}

// ------------------------------------------------------------------------------------------
// Per-scene tables (SceneView::primFrame / lightFrame), computed once at pbrt_b200_scene_create by the code a
// hit would run, so that the values are the same bits the per-hit evaluation produces.
__global__ void k_prim_frames(const __grid_constant__ SceneView sv, float4 *frames, PrimQuad *prims) {
  for (uint32_t prim = blockIdx.x * blockDim.x + threadIdx.x; prim < sv.nPrims; prim += gridDim.x * blockDim.x) {
    TriVerts tv = load_prim(sv, prim);
    if ((tv.flags & PBRT_B200_PF_KIND_MASK) != PBRT_B200_PRIM_TRIANGLE) continue;
    if (tv.flags & (PBRT_B200_PF_HAS_N | PBRT_B200_PF_HAS_S)) continue;   // shading frame interpolates per hit
    if (sv.materials[tv.material].tex[PBRT_B200_TEXSLOT_BUMP] >= 0) continue;
    Ray ray;
    ray.o = v3(0, 0, 0); ray.d = v3(0, 0, 1); ray.mint = 0.f; ray.maxt = PB_INFINITY;
    Hit hit;
    hit.prim = (int32_t)prim; hit.t = 1.f; hit.b1 = hit.b2 = 1.f / 3.f; hit.inst = -1;
    Surface sf;
    surface_at_hit<false, false, false>(sv, ray, hit, &sf);   // (launched with sv.primFrame == nullptr)
    const Bsdf &b = sf.bsdf;
    frames[3 * prim] = make_float4(b.ng.x, b.ng.y, b.ng.z, b.nn.x);
    frames[3 * prim + 1] = make_float4(b.nn.y, b.nn.z, b.sn.x, b.sn.y);
    frames[3 * prim + 2] = make_float4(b.sn.z, b.tn.x, b.tn.y, b.tn.z);
    prims[3 * prim + 2].b = __uint_as_float(tv.flags | PB_PF_DEV_FRAME);
  }
}
__global__ void k_light_frames(const __grid_constant__ SceneView sv, uint32_t nShapes, float4 *frames) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nShapes; j += gridDim.x * blockDim.x) {
    uint32_t prim = sv.lightShapes[j];
    TriVerts tv = load_prim(sv, prim);
    float4 a = make_float4(0, 0, 0, 0), b = a;
    if ((tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE) {
      Ray ray;
      ray.o = v3(0, 0, 0); ray.d = v3(0, 0, 1); ray.mint = 0.f; ray.maxt = PB_INFINITY;
      DiffGeom dg;
      tri_dg(sv, prim, tv, ray, 1.f, 1.f / 3.f, 1.f / 3.f, &dg);
      V3 ns;
      triangle_sample(tv, .5f, .5f, &ns);
      a = make_float4(dg.nn.x, dg.nn.y, dg.nn.z, 0.f);
      b = make_float4(ns.x, ns.y, ns.z, 0.f);
    }
    frames[2 * j] = a; frames[2 * j + 1] = b;
  }
}

// ------------------------------------------------------------------------------------------
// host side
struct WindowFilm {
  PbrtB200Window win;
  FilmPixel *film;     // device
  float4 *rgba;        // device
  float *sampleOut;    // device: 4 floats per sample (kept only when requested)
  uint64_t nSamples;
  float2 *corners;     // device: bestcandidate table placements of the window (scan order)
  std::vector<float2> cornersHost;
};

struct PbrtB200Handle {
  int device;
  cudaStream_t stream;
  cudaStream_t auxStream;        // shadow / probe traversal overlaps the main stream's next kernels
  cudaStream_t auxStream2;       // (tunable 5 = 3: the probes' own stream)
  cudaEvent_t evDirect, evAux, evAux2;
  int numSMs;
  SceneView sv;                 // device pointers
  std::vector<void *> allocs;   // scene arrays
  float *filterTable;
  uint32_t *bcMasks;            // bestcandidate: per-axis prefix strip masks of the sample table (k_film_gather_bc)
  bool hasSpecular;
  bool hasGlossy;
  bool textured;          // some material operand is a texture node (pb_texture.h)
  int primKinds;                // PB_QK_*: which primitive tests the traversal kernels need
  bool triLightsOnly;           // every light is an area light over triangles (kernels drop the other light code)
  int nFrames;
  uint64_t sceneBytes;
  // wave arena
  void *arena;
  size_t arenaBytes;
  WaveBuffers wb;
  // A second "lane": its own arena, streams and events, so that two waves of a frame can be in flight at once and
  // the almost empty tails of one wave's launches are filled by the other's kernels (tunable 10).  The code below
  // always works on the handle's primary fields; swap_lane() exchanges them with the parked set between waves.
  struct Lane {
    void *arena; size_t arenaBytes; WaveBuffers wb;
    cudaStream_t stream, auxStream, auxStream2;
    cudaEvent_t evDirect, evAux, evAux2;
  } parked;
  bool parkedReady;
  cudaEvent_t evFilm[2], evStart;
  std::vector<WindowFilm> films;
  bool keepSamples;
  PbrtB200Window frameWin;   // the scene film's pixel window (crop window), what tiles are tiles of
  FilmPixel *filmSlab;       // tile renders: the windows' unresolved films as one block too (one memset, one resolve)
  float4 *rgbaSlab;          // when the windows' resolved pixels were allocated as one block (tile renders): its base
  size_t rgbaSlabFloats;
  // multi-GPU frames (multi.cuh), rank 0: the assembled frame + staging for the other ranks' tile blocks, tile table
  float4 *gather; size_t gatherFloats;
  float4 *frame;             // last assembled frame (points into `gather`, or at the single window's pixels)
  void *tileTable; size_t tileTableCap;
  // per-kernel-class timing (tunable 3): events bracket every launch on h->stream
  std::vector<cudaEvent_t> evPool;
  std::vector<int> evClass;         // class of the launch bracketed by events 2i, 2i+1
  double classMs[KC_COUNT];
  uint64_t classLaunches[KC_COUNT];
  unsigned long long lastStats[ST_COUNT];
};

static int64_t g_profileKernels = 0;   // key 3
// Brackets one launch with events when profiling is on.
struct LaunchTimer {
  PbrtB200Handle *h; int idx;
  cudaStream_t s;
  LaunchTimer(PbrtB200Handle *hh, int cls, cudaStream_t ss = nullptr) : h(hh), idx(-1), s(ss ? ss : hh->stream) {
    ++h->classLaunches[cls];
    if (!g_profileKernels) return;
    idx = (int)h->evClass.size();
    h->evClass.push_back(cls);
    while ((int)h->evPool.size() < 2 * (idx + 1)) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      h->evPool.push_back(e);
    }
    cudaEventRecord(h->evPool[2 * idx], s);
  }
  ~LaunchTimer() { if (idx >= 0) cudaEventRecord(h->evPool[2 * idx + 1], s); }
};
#define TIMED(cls, launch) do { LaunchTimer lt_(h, cls); launch; } while (0)
#define TIMED_ON(cls, strm, launch) do { LaunchTimer lt_(h, cls, strm); launch; } while (0)

// CUDA events that are destroyed on every exit path
struct ScopedEvents {
  cudaEvent_t e[4];
  int n;
  ScopedEvents() : n(0) {}
  ~ScopedEvents() { for (int i = 0; i < n; ++i) cudaEventDestroy(e[i]); }
  cudaError_t create(int count) {
    for (; n < count; ++n) {
      cudaError_t err = cudaEventCreate(&e[n]);
      if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
  }
};
#define PB_MAX_DEVICES 64
// Small per-device cache of released device blocks: pbrt_b200_render_host() builds and tears down a
// handle per call (scene arrays, films), and cudaMalloc / cudaFree (which synchronises) per frame
// showed up as ~10 ms of a 176 ms end-to-end step.  A block is reused for a request of at most its
// size and at least half of it; the cache is bounded in blocks and bytes.
#include <mutex>
#include <unordered_map>
namespace {
struct PoolBlock { void *p; size_t bytes; };
std::mutex g_poolMutex;
std::vector<PoolBlock> g_pool[PB_MAX_DEVICES];
std::unordered_map<void *, size_t> g_poolSizes;     // live blocks handed out by pool_alloc
const size_t kPoolMaxBlocks = 512, kPoolMaxBytes = (size_t)2 << 30;
}  // namespace
static cudaError_t pool_alloc(int device, void **out, size_t bytes) {
  if (bytes == 0) bytes = 1;
  {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    if (device >= 0 && device < PB_MAX_DEVICES) {
      std::vector<PoolBlock> &pool = g_pool[device];
      int best = -1;
      for (int i = 0; i < (int)pool.size(); ++i)
        if (pool[i].bytes >= bytes && pool[i].bytes <= 2 * bytes + 4096 && (best < 0 || pool[i].bytes < pool[best].bytes)) best = i;
      if (best >= 0) {
        *out = pool[best].p;
        g_poolSizes[*out] = pool[best].bytes;
        pool.erase(pool.begin() + best);
        return cudaSuccess;
      }
    }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    g_poolSizes[*out] = bytes;
  }
  return e;
}
static void pool_free(int device, void *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_poolMutex);
  size_t bytes = 0;
  auto it = g_poolSizes.find(p);
  if (it != g_poolSizes.end()) { bytes = it->second; g_poolSizes.erase(it); }
  if (bytes && device >= 0 && device < PB_MAX_DEVICES) {
    std::vector<PoolBlock> &pool = g_pool[device];
    size_t total = bytes;
    for (const PoolBlock &b : pool) total += b.bytes;
    if (pool.size() < kPoolMaxBlocks && total <= kPoolMaxBytes) {
      pool.push_back({p, bytes});
      return;
    }
  }
  cudaFree(p);
}

template <class T>
static int upload(PbrtB200Handle *h, const T *src, size_t n, const T **dst) {
  *dst = nullptr;
  if (n == 0 || !src) return PBRT_B200_OK;
  void *d = nullptr;
  CU_TRY(pool_alloc(h->device, &d, n * sizeof(T)));
  h->allocs.push_back(d);
  CU_TRY(cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  h->sceneBytes += n * sizeof(T);
  *dst = (const T *)d;
  return PBRT_B200_OK;
}

// One retired wave arena per device is kept for the next handle: pbrt_b200_render_host() creates and
// destroys a handle per call, and a multi-GB cudaMalloc/cudaFree per frame would dominate it.
static void *g_cachedArena[PB_MAX_DEVICES][2];       // (two: a handle may run two wave lanes)
static size_t g_cachedArenaBytes[PB_MAX_DEVICES][2];
static void free_wave(PbrtB200Handle *h) {
  if (h->arena) {
    std::lock_guard<std::mutex> lock(g_poolMutex);   // (handles of several host threads share the per-device cache)
    int d = h->device;
    if (d >= 0 && d < PB_MAX_DEVICES) {
      // keep it in the smaller (or an empty) one of the device's two cache slots if it is larger than that one
      int slot = g_cachedArenaBytes[d][0] <= g_cachedArenaBytes[d][1] ? 0 : 1;
      if (h->arenaBytes > g_cachedArenaBytes[d][slot]) {
        if (g_cachedArena[d][slot]) cudaFree(g_cachedArena[d][slot]);
        g_cachedArena[d][slot] = h->arena;
        g_cachedArenaBytes[d][slot] = h->arenaBytes;
      } else
        cudaFree(h->arena);
    } else
      cudaFree(h->arena);
  }
  h->arena = nullptr;
  h->arenaBytes = 0;
  memset(&h->wb, 0, sizeof(h->wb));
}
// Exchanges the handle's working lane (arena, wave buffers, streams, events) with the parked one.
static void swap_lane(PbrtB200Handle *h) {
  PbrtB200Handle::Lane cur = {h->arena, h->arenaBytes, h->wb, h->stream, h->auxStream, h->auxStream2,
                              h->evDirect, h->evAux, h->evAux2};
  const PbrtB200Handle::Lane &o = h->parked;
  h->arena = o.arena; h->arenaBytes = o.arenaBytes; h->wb = o.wb;
  h->stream = o.stream; h->auxStream = o.auxStream; h->auxStream2 = o.auxStream2;
  h->evDirect = o.evDirect; h->evAux = o.evAux; h->evAux2 = o.evAux2;
  h->parked = cur;
}
static int ensure_parked_lane(PbrtB200Handle *h) {
  if (h->parkedReady) return PBRT_B200_OK;
  PbrtB200Handle::Lane &l = h->parked;
  CU_TRY(cudaStreamCreate(&l.stream));
  CU_TRY(cudaStreamCreate(&l.auxStream));
  CU_TRY(cudaStreamCreate(&l.auxStream2));
  CU_TRY(cudaEventCreateWithFlags(&l.evDirect, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&l.evAux, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&l.evAux2, cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&h->evFilm[0], cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&h->evFilm[1], cudaEventDisableTiming));
  CU_TRY(cudaEventCreateWithFlags(&h->evStart, cudaEventDisableTiming));
  h->parkedReady = true;
  return PBRT_B200_OK;
}

struct Carver {
  char *base; size_t off;
  template <class T> T *take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? (T *)(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};
static void carve_wave(Carver &c, WaveBuffers &wb, uint32_t nSlots, int directRecords, int nFrames, bool isPath,
                       int ldFloats, bool hasInstances, bool textured) {
  wb.nSlots = nSlots;
  wb.nFrames = nFrames;
  wb.ptc = isPath ? c.take<float4>(nSlots) : nullptr;
  wb.plf = isPath ? c.take<float4>(nSlots) : nullptr;
  wb.prec = isPath ? c.take<float4>(nSlots) : nullptr;
  for (int i = 0; i < 3; ++i) {
    wb.L[i] = isPath ? nullptr : c.take<float>(nSlots);
    wb.T[i] = isPath ? nullptr : c.take<float>(nSlots);
    wb.pendT[i] = c.take<float>(nSlots);
    wb.outL[i] = c.take<float>(nSlots);
    wb.recA[i] = isPath ? nullptr : c.take<float>((size_t)nSlots * directRecords);
    wb.recB[i] = c.take<float>((size_t)nSlots * directRecords);
  }
  wb.alpha = c.take<float>(nSlots); wb.outA = c.take<float>(nSlots);
  wb.ctr = isPath ? nullptr : c.take<uint32_t>(nSlots);
  wb.flags = isPath ? nullptr : c.take<uint32_t>(nSlots);
  wb.imageX = c.take<float>(nSlots); wb.imageY = c.take<float>(nSlots);
  wb.pixel = c.take<uint32_t>(nSlots);
  wb.reach = c.take<unsigned long long>(nSlots);
  wb.ld = ldFloats > 0 ? c.take<float>((size_t)ldFloats * nSlots) : nullptr;
  wb.recFlags = isPath ? nullptr : c.take<uint8_t>((size_t)nSlots * directRecords);
  if (!isPath) {
    wb.frame = c.take<float>((size_t)PB_FRAME_FLOATS * nFrames * nSlots);
    wb.frameFlags = c.take<uint32_t>((size_t)nFrames * nSlots);
    wb.nRecs = c.take<uint32_t>(nSlots);
  } else {
    wb.frame = nullptr; wb.frameFlags = nullptr; wb.nRecs = nullptr;
  }
  for (int k = 0; k < 2; ++k) {
    RayQueue &q = wb.q[k];
    q.slot = c.take<uint32_t>(nSlots);
    q.ox = c.take<float>(nSlots); q.oy = c.take<float>(nSlots); q.oz = c.take<float>(nSlots);
    q.dx = c.take<float>(nSlots); q.dy = c.take<float>(nSlots); q.dz = c.take<float>(nSlots);
    q.mint = c.take<float>(nSlots); q.maxt = c.take<float>(nSlots);
    q.hitPrim = c.take<int32_t>(nSlots);
    q.hitT = c.take<float>(nSlots); q.hitB1 = c.take<float>(nSlots); q.hitB2 = c.take<float>(nSlots);
    q.hitInst = hasInstances ? c.take<int32_t>(nSlots) : nullptr;
  }
  size_t auxN = (size_t)nSlots * directRecords;
  AuxQueue *aq[2] = {&wb.shadowQ, &wb.probeQ};
  for (int k = 0; k < 2; ++k) {
    AuxQueue &q = *aq[k];
    q.rec = c.take<uint32_t>(auxN);
    q.ox = c.take<float>(auxN); q.oy = c.take<float>(auxN); q.oz = c.take<float>(auxN);
    q.dx = c.take<float>(auxN); q.dy = c.take<float>(auxN); q.dz = c.take<float>(auxN);
    q.light = (k == 1) ? c.take<int32_t>(auxN) : nullptr;
    q.maxt = (k == 0) ? c.take<float>(auxN) : nullptr;
  }
  for (int k = 0; k < 15; ++k) wb.surf[k] = isPath ? c.take<float>(nSlots) : nullptr;
  wb.surfIds = isPath ? c.take<uint32_t>(nSlots) : nullptr;
  wb.subset = isPath ? c.take<uint32_t>(nSlots) : nullptr;
  wb.surfQueueIdx = isPath ? c.take<uint32_t>(nSlots) : nullptr;
  wb.camDiff = textured ? c.take<float>((size_t)12 * nSlots) : nullptr;
  wb.surfMat = (textured && isPath) ? c.take<PbrtB200Material>(nSlots) : nullptr;
  wb.frameDiff = (textured && !isPath) ? c.take<float>((size_t)24 * nFrames * nSlots) : nullptr;
  wb.counts = c.take<uint32_t>(PB_COUNTS_TOTAL);
  wb.stats = c.take<unsigned long long>(ST_COUNT);
}

static int ensure_wave(PbrtB200Handle *h, uint32_t nSlots) {
  if (h->arena && h->wb.nSlots >= nSlots) return PBRT_B200_OK;
  free_wave(h);
  bool isPath = h->sv.integrator.kind == PBRT_B200_INT_PATH;
  Carver sizing = {nullptr, 0};
  WaveBuffers tmp;
  carve_wave(sizing, tmp, nSlots, h->sv.directRecords, h->nFrames, isPath, h->sv.ldFloatsPerSample, h->sv.nInstances != 0, h->textured);
  size_t bytes = sizing.off + 256;
  int d = h->device;
  {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    if (d >= 0 && d < PB_MAX_DEVICES) {
      // the smallest cached arena that is large enough
      int best = -1;
      for (int k = 0; k < 2; ++k)
        if (g_cachedArena[d][k] && g_cachedArenaBytes[d][k] >= bytes &&
            (best < 0 || g_cachedArenaBytes[d][k] < g_cachedArenaBytes[d][best]))
          best = k;
      if (best >= 0) {
        h->arena = g_cachedArena[d][best];
        h->arenaBytes = g_cachedArenaBytes[d][best];
        g_cachedArena[d][best] = nullptr;
        g_cachedArenaBytes[d][best] = 0;
      }
    }
  }
  if (!h->arena) {
    CU_TRY(cudaMalloc(&h->arena, bytes));
    h->arenaBytes = bytes;
  }
  Carver c = {(char *)h->arena, 0};
  carve_wave(c, h->wb, nSlots, h->sv.directRecords, h->nFrames, isPath, h->sv.ldFloatsPerSample, h->sv.nInstances != 0, h->textured);
  return PBRT_B200_OK;
}

static void free_films(PbrtB200Handle *h) {
  for (auto &f : h->films) {
    if (!h->filmSlab) pool_free(h->device, f.film);
    if (!h->rgbaSlab) pool_free(h->device, f.rgba);
    pool_free(h->device, f.sampleOut);
    pool_free(h->device, f.corners);
  }
  h->films.clear();
  pool_free(h->device, h->rgbaSlab);
  pool_free(h->device, h->filmSlab);
  h->rgbaSlab = nullptr;
  h->filmSlab = nullptr;
  h->rgbaSlabFloats = 0;
}

static int persistent_grid(const PbrtB200Handle *h) { return h->numSMs * (int)g_blocksPerSM; }

static void launch_traversals_a(PbrtB200Handle *h, bool count, int grid, const RayQueue &q, uint32_t *nMain,
                                uint32_t *work) {
#define PB_LAUNCH_INTERSECT(C_, QK_)                                                                                             \
  do {                                                                                                                         \
    if (h->sv.nNodes > PB_BIG_TREE_NODES)                                                                                      \
      TIMED(KC_INTERSECT, (k_intersect<C_, QK_, PB_DESCEND_STEPS_BIG><<<grid, PB_BLOCK, 0, h->stream>>>(h->sv, q, nMain, work, h->wb.stats))); \
    else                                                                                                                       \
      TIMED(KC_INTERSECT, (k_intersect<C_, QK_, PB_DESCEND_STEPS><<<grid, PB_BLOCK, 0, h->stream>>>(h->sv, q, nMain, work, h->wb.stats)));     \
  } while (0)
  if (count) {   // the counting pass is not timed; it walks like the timed kernels (same counters as the twin's)
    if (h->primKinds == PB_QK_INSTANCES) PB_LAUNCH_INTERSECT(true, PB_QK_INSTANCES);
    else PB_LAUNCH_INTERSECT(true, PB_QK_ALL);
  } else if (h->primKinds == PB_QK_TRI) PB_LAUNCH_INTERSECT(false, PB_QK_TRI);
  else if (h->primKinds == PB_QK_SPHERE) PB_LAUNCH_INTERSECT(false, PB_QK_SPHERE);
  else if (h->primKinds == PB_QK_ALL) PB_LAUNCH_INTERSECT(false, PB_QK_ALL);
  else PB_LAUNCH_INTERSECT(false, PB_QK_INSTANCES);
}
static void launch_traversals_b(PbrtB200Handle *h, bool count, int grid, uint32_t *nShadow, uint32_t *nProbe,
                                uint32_t *workS, uint32_t *workP, cudaStream_t strm, cudaStream_t probeStrm) {
  WaveBuffers &wb = h->wb;
  // MIS probes are rare (the BSDF sample must point at the light): a small grid avoids paying the
  // ramp-up of the full persistent grid for a few thousand rays
  int probeGrid = std::min(grid, h->numSMs * (h->sv.nNodes > PB_BIG_TREE_NODES ? PB_PROBE_CTAS_BIG : 2));
  if (count) {
    if (h->primKinds == PB_QK_INSTANCES)
      TIMED_ON(KC_SHADOW, strm, (k_shadow<true, PB_QK_INSTANCES, PB_DESCEND_STEPS><<<grid, PB_BLOCK, 0, strm>>>(h->sv, wb.shadowQ, nShadow, workS, wb.recFlags, wb.prec, wb.stats)));
    else
      TIMED_ON(KC_SHADOW, strm, (k_shadow<true, PB_QK_ALL, PB_DESCEND_STEPS><<<grid, PB_BLOCK, 0, strm>>>(h->sv, wb.shadowQ, nShadow, workS, wb.recFlags, wb.prec, wb.stats)));
    if (h->primKinds == PB_QK_INSTANCES)
      TIMED_ON(KC_PROBE, probeStrm, (k_probe<true, true><<<probeGrid, PB_BLOCK, 0, probeStrm>>>(h->sv, wb.probeQ, nProbe, workP, wb.recFlags, wb.prec, wb.stats)));
    else
      TIMED_ON(KC_PROBE, probeStrm, (k_probe<true, false><<<probeGrid, PB_BLOCK, 0, probeStrm>>>(h->sv, wb.probeQ, nProbe, workP, wb.recFlags, wb.prec, wb.stats)));
  } else {
#define PB_LAUNCH_SHADOW(QK_)                                                                                                    \
  do {                                                                                                                         \
    if (h->sv.nNodes > PB_BIG_TREE_NODES)                                                                                      \
      TIMED_ON(KC_SHADOW, strm, (k_shadow<false, QK_, PB_DESCEND_STEPS_BIG><<<grid, PB_BLOCK, 0, strm>>>(h->sv, wb.shadowQ, nShadow, workS, wb.recFlags, wb.prec, wb.stats))); \
    else                                                                                                                       \
      TIMED_ON(KC_SHADOW, strm, (k_shadow<false, QK_, PB_DESCEND_STEPS><<<grid, PB_BLOCK, 0, strm>>>(h->sv, wb.shadowQ, nShadow, workS, wb.recFlags, wb.prec, wb.stats)));     \
  } while (0)
    if (h->primKinds == PB_QK_TRI) PB_LAUNCH_SHADOW(PB_QK_TRI);
    else if (h->primKinds == PB_QK_SPHERE) PB_LAUNCH_SHADOW(PB_QK_SPHERE);
    else if (h->primKinds == PB_QK_ALL) PB_LAUNCH_SHADOW(PB_QK_ALL);
    else PB_LAUNCH_SHADOW(PB_QK_INSTANCES);
    if (h->primKinds == PB_QK_INSTANCES)
      TIMED_ON(KC_PROBE, probeStrm, (k_probe<false, true><<<probeGrid, PB_BLOCK, 0, probeStrm>>>(h->sv, wb.probeQ, nProbe, workP, wb.recFlags, wb.prec, wb.stats)));
    else
      TIMED_ON(KC_PROBE, probeStrm, (k_probe<false, false><<<probeGrid, PB_BLOCK, 0, probeStrm>>>(h->sv, wb.probeQ, nProbe, workP, wb.recFlags, wb.prec, wb.stats)));
  }
}

// One wave; counts/work indices advance through `it`.
static int run_wave(PbrtB200Handle *h, const WaveDesc &wd, uint64_t *launches) {
  const SceneView &sv = h->sv;
  WaveBuffers &wb = h->wb;
  cudaStream_t st = h->stream;
  int grid = persistent_grid(h);
  bool count = g_countKd != 0;
  CU_TRY(cudaMemsetAsync(wb.counts, 0, PB_COUNTS_TOTAL * sizeof(uint32_t), st));
  // (bestcandidate: k_generate counts the table entries that fall inside the extent itself)
  if (!wd.bc && !wd.tiles) CU_TRY(cudaMemcpyAsync(&wb.counts[0], &wd.nSlots, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  int genGrid = (int)std::min<uint64_t>((wd.nSlots + PB_BLOCK - 1) / PB_BLOCK, (uint64_t)h->numSMs * 16);
  if (sv.sampler.kind == PBRT_B200_SAMPLER_LOWDISCREPANCY) {
    TIMED(KC_GENERATE, (k_slot_pixels<<<genGrid, PB_BLOCK, 0, st>>>(wb, wd)));
    uint64_t nThreads = (uint64_t)(wd.nSlots / wd.spp) * (2 + sv.nPatterns);
    int tg = (int)std::min<uint64_t>((nThreads + 127) / 128, (uint64_t)h->numSMs * 32);
    TIMED(KC_GENERATE, (k_ld_tables<<<tg, 128, 0, st>>>(sv, wb, wd)));
    *launches += 2;
  }
  TIMED(KC_GENERATE, (k_generate<<<genGrid, PB_BLOCK, 0, st>>>(sv, wb, wd)));
  ++*launches;
  if (wd.bc && sv.ldFloatsPerSample > 0) {
    uint64_t nWarps = (uint64_t)(wd.nSlots / PB_BC_TABLE) * sv.nPatterns;
    int tg = (int)std::min<uint64_t>((nWarps + PB_BC_TABLE_WARPS - 1) / PB_BC_TABLE_WARPS, (uint64_t)h->numSMs * 16);
    TIMED(KC_GENERATE, (k_bc_tables<<<tg, 32 * PB_BC_TABLE_WARPS, 0, st>>>(sv, wb, wd)));
    ++*launches;
  }
  uint32_t workIdx = PB_WORK_BASE;
  auto nextWork = [&]() { return wb.counts + (workIdx++); };
  bool isPath = sv.integrator.kind == PBRT_B200_INT_PATH;
  int cur = 0;
  WaveKeyInfo wk = {wb.pixel, wd.spp, wd.bc ? 1 : (sv.sampler.kind == PBRT_B200_SAMPLER_RANDOM ? 2 : 0)};
  const bool overlap3 = g_overlap == 3 && !count;      // like 1, and the probes on a third stream beside the shadow rays
  const bool overlap = g_overlap == 1 || overlap3;     // shadow + probe on the aux stream, beside the next closest-hit launch
  const bool probeAside = g_overlap == 2 && !count;    // only the probes on the aux stream, beside the shadow rays
  // (LD kernels = patterns read from per-unit tables: lowdiscrepancy and bestcandidate)
  const bool isLD = sv.sampler.kind == PBRT_B200_SAMPLER_LOWDISCREPANCY || wd.bc;
  if (isPath) {
    int iters = sv.integrator.max_depth + 1;
    if (iters < 1) iters = 1;
    if (iters > PB_MAX_ITERS - 1) return fail(PBRT_B200_ERR_UNSUPPORTED, "path maxdepth too large");
    for (int it = 0; it < iters; ++it) {
      uint32_t *nMain = &wb.counts[4 * it], *nShadow = &wb.counts[4 * it + 1], *nProbe = &wb.counts[4 * it + 2];
      uint32_t *nNext = &wb.counts[4 * (it + 1)];
      launch_traversals_a(h, count, grid, wb.q[cur], nMain, nextWork());
      // the previous vertex's shadow / probe results are folded in by k_path_surface
      if ((overlap || probeAside) && it > 0) CU_TRY(cudaStreamWaitEvent(st, h->evAux, 0));
      if (overlap3 && it > 0) CU_TRY(cudaStreamWaitEvent(st, h->evAux2, 0));
      {
        SurfaceRecord rec;
        for (int k = 0; k < 15; ++k) rec.f[k] = wb.surf[k];
        rec.ids = wb.surfIds;
        rec.queueIdx = wb.surfQueueIdx;
        rec.mat = wb.surfMat;
        uint32_t *nHits = &wb.counts[4 * it + 3];
        uint32_t *nSpec = nextWork();   // specular-only hits (second bin of the surface records)
        // untextured scenes: k_path_fast shades the plain vertices outright and leaves the rest as a subset
        const bool fast = g_fastPath && !h->textured && sv.primFrame != nullptr;
        const uint32_t *nGeneral = nMain;
        const uint32_t *subset = nullptr;
        if (fast) {
          uint32_t *nSub = nextWork();
#define PB_LAUNCH_FAST(LD_, LK_)                                                                                      \
  TIMED(KC_SHADE_FAST, (k_path_fast<LD_, LK_><<<grid, PB_BLOCK, 0, st>>>(sv, wb, wk, wb.q[cur], wb.q[cur ^ 1], nMain, \
                                                                        nextWork(), nShadow, nProbe, nNext, wb.subset, nSub)))
          if (isLD) { if (h->triLightsOnly) PB_LAUNCH_FAST(true, 0); else PB_LAUNCH_FAST(true, PB_LK_ALL); }
          else { if (h->triLightsOnly) PB_LAUNCH_FAST(false, 0); else PB_LAUNCH_FAST(false, PB_LK_ALL); }
#undef PB_LAUNCH_FAST
          nGeneral = nSub;
          subset = wb.subset;
          ++*launches;
        }
#define PB_LAUNCH_SURF(INST, TEX) TIMED(KC_SHADE, (k_path_surface<INST, TEX><<<grid, PB_BLOCK, 0, st>>>(sv, wb, wb.q[cur], rec, nGeneral, nextWork(), nHits, nSpec, subset)))
        if (h->textured) { if (sv.nInstances != 0) PB_LAUNCH_SURF(true, true); else PB_LAUNCH_SURF(false, true); }
        else { if (sv.nInstances != 0) PB_LAUNCH_SURF(true, false); else PB_LAUNCH_SURF(false, false); }
#undef PB_LAUNCH_SURF
#define PB_LAUNCH_DIRECT2(LD_, G_, LK_)                                                                             \
  TIMED(KC_SHADE_DIRECT, (k_path_direct<LD_, G_, LK_><<<grid, PB_BLOCK, 0, st>>>(sv, wb, wk, wb.q[cur], rec, nHits, \
                                                                                 nextWork(), nShadow, nProbe, wb.q[cur ^ 1], nNext)))
#define PB_LAUNCH_DIRECT(LD_, G_) do { if (h->triLightsOnly) PB_LAUNCH_DIRECT2(LD_, G_, 0); else PB_LAUNCH_DIRECT2(LD_, G_, PB_LK_ALL); } while (0)
        if (isLD) { if (h->hasGlossy) PB_LAUNCH_DIRECT(true, true); else PB_LAUNCH_DIRECT(true, false); }
        else { if (h->hasGlossy) PB_LAUNCH_DIRECT(false, true); else PB_LAUNCH_DIRECT(false, false); }
        if (overlap) {
          // shadow + probe rays of this vertex run on the aux stream, concurrently with
          // k_path_continue and the next k_intersect (they share no data)
          CU_TRY(cudaEventRecord(h->evDirect, st));
          CU_TRY(cudaStreamWaitEvent(h->auxStream, h->evDirect, 0));
          if (overlap3) CU_TRY(cudaStreamWaitEvent(h->auxStream2, h->evDirect, 0));
          launch_traversals_b(h, count, grid, nShadow, nProbe, nextWork(), nextWork(), h->auxStream,
                              overlap3 ? h->auxStream2 : h->auxStream);
          CU_TRY(cudaEventRecord(h->evAux, h->auxStream));
          if (overlap3) CU_TRY(cudaEventRecord(h->evAux2, h->auxStream2));
        }
#define PB_LAUNCH_CONT(LD_, G_)                                                                                    \
  TIMED(KC_SHADE_CONT, (k_path_continue<LD_, G_><<<grid, PB_BLOCK, 0, st>>>(sv, wb, wk, wb.q[cur], wb.q[cur ^ 1], rec, \
                                                                            nHits, nSpec, nextWork(), nNext)))
        if (isLD) { if (h->hasGlossy) PB_LAUNCH_CONT(true, true); else PB_LAUNCH_CONT(true, false); }
        else { if (h->hasGlossy) PB_LAUNCH_CONT(false, true); else PB_LAUNCH_CONT(false, false); }
      }
      if (probeAside) {
        // the few MIS probes of this vertex (a latency-bound launch that leaves most of the GPU idle) run on the aux
        // stream beside the shadow rays and the next closest-hit launch; the next vertex's shading waits for them
        CU_TRY(cudaEventRecord(h->evDirect, st));
        CU_TRY(cudaStreamWaitEvent(h->auxStream, h->evDirect, 0));
        launch_traversals_b(h, count, grid, nShadow, nProbe, nextWork(), nextWork(), st, h->auxStream);
        CU_TRY(cudaEventRecord(h->evAux, h->auxStream));
      } else if (!overlap)
        launch_traversals_b(h, count, grid, nShadow, nProbe, nextWork(), nextWork(), st, st);
      *launches += 6;
      cur ^= 1;
    }
    if (overlap || probeAside) CU_TRY(cudaStreamWaitEvent(st, h->evAux, 0));
    if (overlap3) CU_TRY(cudaStreamWaitEvent(st, h->evAux2, 0));
  } else {
    // Whitted / DirectLighting: the recursion depth is data dependent; one host readback of
    // the queue size per iteration ends the loop.
    uint32_t nMainHost = wd.nSlots;
    if (wd.bc || wd.tiles) {
      CU_TRY(cudaMemcpyAsync(&nMainHost, &wb.counts[0], sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaStreamSynchronize(st));
    }
    int it = 0;
    while (nMainHost > 0) {
      if (it >= PB_MAX_ITERS - 1) {
        // recycle the counter block
        CU_TRY(cudaMemsetAsync(wb.counts, 0, PB_COUNTS_TOTAL * sizeof(uint32_t), st));
        CU_TRY(cudaMemcpyAsync(&wb.counts[0], &nMainHost, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        it = 0;
        workIdx = PB_WORK_BASE;
      }
      uint32_t *nMain = &wb.counts[4 * it], *nShadow = &wb.counts[4 * it + 1], *nProbe = &wb.counts[4 * it + 2];
      uint32_t *nNext = &wb.counts[4 * (it + 1)];
      launch_traversals_a(h, count, grid, wb.q[cur], nMain, nextWork());
      if (h->textured) TIMED(KC_SHADE, (k_shade_wd<true><<<grid, PB_BLOCK, 0, st>>>(sv, wb, wk, wb.q[cur], nMain, nextWork(), nShadow, nProbe)));
      else TIMED(KC_SHADE, (k_shade_wd<false><<<grid, PB_BLOCK, 0, st>>>(sv, wb, wk, wb.q[cur], nMain, nextWork(), nShadow, nProbe)));
      launch_traversals_b(h, count, grid, nShadow, nProbe, nextWork(), nextWork(), st, st);
      TIMED(KC_ADVANCE, (k_advance_wd<<<grid, PB_BLOCK, 0, st>>>(sv, wb, wb.q[cur], wb.q[cur ^ 1], nMain, nextWork(), nNext)));
      *launches += 5;
      CU_TRY(cudaMemcpyAsync(&nMainHost, nNext, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CU_TRY(cudaStreamSynchronize(st));
      cur ^= 1;
      ++it;
    }
  }
  TIMED(KC_FINALIZE, (k_finalize<<<genGrid, PB_BLOCK, 0, st>>>(sv, wb, wd)));
  ++*launches;
  if (g_debugCounts && isPath) {
    std::vector<uint32_t> hc(4 * (sv.integrator.max_depth + 2));
    CU_TRY(cudaMemcpyAsync(hc.data(), wb.counts, hc.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    for (int it = 0; it <= sv.integrator.max_depth; ++it)
      fprintf(stderr, "pbrt_b200 wave: iteration %d main %u shadow %u probe %u\n", it, hc[4 * it], hc[4 * it + 1],
              hc[4 * it + 2]);
  }
  CU_TRY(cudaGetLastError());
  return PBRT_B200_OK;
}

static WavePart window_part(const SceneView &sv, const PbrtB200Window &w) {
  WavePart wp;
  memset(&wp, 0, sizeof(wp));
  film_sample_extent(sv, w.x0, w.y0, w.width, w.height, &wp.xs, &wp.xe, &wp.ys, &wp.ye);
  return wp;
}

extern "C" {

int pbrt_b200_abi_version(void) { return PBRT_B200_ABI_VERSION; }
int pbrt_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
const char *pbrt_b200_last_error(void) { return g_lastError.c_str(); }

int64_t pbrt_b200_tunable(int key, int64_t value) {
  int64_t *p = key == 0 ? &g_waveSamples : key == 1 ? &g_countKd : key == 2 ? &g_blocksPerSM
               : key == 3 ? &g_profileKernels : key == 4 ? &g_reserved4 : key == 5 ? &g_overlap : key == 6 ? &g_debugCounts
               : key == 7 ? &g_primFrames : key == 8 ? &g_fastPath : key == 9 ? &g_treelets : key == 10 ? &g_lanes : key == 11 ? &g_l2Persist : nullptr;
  if (!p) return -1;
  int64_t prev = *p;
  if (value >= 0) *p = value;
  return prev;
}

void pbrt_b200_scene_destroy(PbrtB200Handle *h) {
  if (!h) return;
  cudaSetDevice(h->device);
  // (released blocks may still be read by queued work of this handle's streams: drain them first)
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->auxStream) cudaStreamSynchronize(h->auxStream);
  if (h->auxStream2) cudaStreamSynchronize(h->auxStream2);
  free_films(h);
  free_wave(h);
  if (h->parkedReady) {
    cudaStreamSynchronize(h->parked.stream);
    cudaStreamSynchronize(h->parked.auxStream);
    cudaStreamSynchronize(h->parked.auxStream2);
    swap_lane(h);
    free_wave(h);
    swap_lane(h);
    cudaEventDestroy(h->parked.evDirect); cudaEventDestroy(h->parked.evAux); cudaEventDestroy(h->parked.evAux2);
    cudaEventDestroy(h->evFilm[0]); cudaEventDestroy(h->evFilm[1]); cudaEventDestroy(h->evStart);
    cudaStreamDestroy(h->parked.auxStream); cudaStreamDestroy(h->parked.auxStream2); cudaStreamDestroy(h->parked.stream);
  }
  for (void *p : h->allocs) pool_free(h->device, p);
  pool_free(h->device, h->filterTable);
  pool_free(h->device, h->bcMasks);
  pool_free(h->device, h->gather);
  pool_free(h->device, h->tileTable);
  for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
  if (h->evDirect) cudaEventDestroy(h->evDirect);
  if (h->evAux) cudaEventDestroy(h->evAux);
  if (h->evAux2) cudaEventDestroy(h->evAux2);
  if (h->auxStream) cudaStreamDestroy(h->auxStream);
  if (h->auxStream2) cudaStreamDestroy(h->auxStream2);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int pbrt_b200_scene_create(const PbrtB200Scene *scene, int device, PbrtB200Handle **out) {
  if (!scene || !out) return fail(PBRT_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    return fail(PBRT_B200_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  if (device < 0 || device >= nDev) return fail(PBRT_B200_ERR_INVALID, "device ordinal out of range");
  CU_TRY(cudaSetDevice(device));
  // (cudaGetDeviceProperties costs ~1 ms: asked once per device)
  static int devMajor[PB_MAX_DEVICES], devSMs[PB_MAX_DEVICES];
  static bool devKnown[PB_MAX_DEVICES];
  if (device >= PB_MAX_DEVICES) return fail(PBRT_B200_ERR_INVALID, "device ordinal too large");
  if (!devKnown[device]) {
    CU_TRY(cudaDeviceGetAttribute(&devMajor[device], cudaDevAttrComputeCapabilityMajor, device));
    CU_TRY(cudaDeviceGetAttribute(&devSMs[device], cudaDevAttrMultiProcessorCount, device));
    devKnown[device] = true;
  }
  if (devMajor[device] < 10) return fail(PBRT_B200_ERR_NO_DEVICE, "device is not sm_100 class");
  PbrtB200Handle *h = new PbrtB200Handle();
  h->device = device; h->stream = nullptr; h->auxStream = nullptr; h->auxStream2 = nullptr; h->evDirect = nullptr; h->evAux = nullptr; h->evAux2 = nullptr;
  memset(&h->parked, 0, sizeof(h->parked)); h->parkedReady = false;
  h->numSMs = devSMs[device];
  h->filterTable = nullptr; h->bcMasks = nullptr; h->arena = nullptr; h->arenaBytes = 0; h->sceneBytes = 0; h->keepSamples = false;
  memset(&h->wb, 0, sizeof(h->wb));
  memset(&h->sv, 0, sizeof(h->sv));
  memset(h->classMs, 0, sizeof(h->classMs));
  memset(h->classLaunches, 0, sizeof(h->classLaunches));
  memset(h->lastStats, 0, sizeof(h->lastStats));
  std::string err = pb_setup_view(*scene, &h->sv);
  if (!err.empty()) { delete h; return fail(PBRT_B200_ERR_INVALID, err); }
  h->frameWin.x0 = scene->film.x_pixel_start; h->frameWin.y0 = scene->film.y_pixel_start;
  h->frameWin.width = scene->film.x_pixel_count; h->frameWin.height = scene->film.y_pixel_count;
  h->rgbaSlab = nullptr; h->rgbaSlabFloats = 0; h->filmSlab = nullptr;
  h->gather = nullptr; h->gatherFloats = 0; h->frame = nullptr; h->tileTable = nullptr; h->tileTableCap = 0;
  h->hasSpecular = pb_scene_has_specular(*scene);
  h->hasGlossy = false;
  h->textured = scene->n_textures != 0;   // material operands (below) or light maps
  for (uint32_t i = 0; i < scene->n_materials; ++i) {
    int k = scene->materials[i].kind;
    if (k < PBRT_B200_MAT_MATTE || k > PBRT_B200_MAT_MEASURED) { delete h; return fail(PBRT_B200_ERR_INVALID, "unknown material kind"); }
    // plastic / uber / shinymetal / translucent: microfacet lobes and up to 4 lobes per BSDF (kernels instantiated with G = true)
    h->hasGlossy |= k >= PBRT_B200_MAT_PLASTIC;
    for (int t = 0; t < PBRT_B200_TEXSLOTS; ++t) h->textured |= scene->materials[i].tex[t] >= 0;
  }
  h->primKinds = PB_QK_TRI;
  for (uint32_t i = 0; i < scene->n_quadrics; ++i)
    h->primKinds = std::max(h->primKinds, scene->quadrics[i].kind == PBRT_B200_QUADRIC_SPHERE ? PB_QK_SPHERE : PB_QK_ALL);
  if (scene->n_instances != 0) h->primKinds = PB_QK_INSTANCES;
  h->triLightsOnly = true;
  for (uint32_t i = 0; i < scene->n_lights; ++i) {
    const PbrtB200Light &l = scene->lights[i];
    if (l.kind != PBRT_B200_LIGHT_AREA) { h->triLightsOnly = false; continue; }
    for (uint32_t k = 0; k < l.n_shapes; ++k) {
      uint32_t prim = scene->light_shapes[l.first_shape + k];
      if (prim >= scene->n_prims || (scene->prims[prim].flags & PBRT_B200_PF_KIND_MASK) != PBRT_B200_PRIM_TRIANGLE)
        h->triLightsOnly = false;
    }
  }
  bool isPath = scene->integrator.kind == PBRT_B200_INT_PATH;
  h->nFrames = isPath ? 0 : (h->hasSpecular ? std::max(0, scene->integrator.max_depth) + 1 : 1);
  if (!isPath && h->sv.directRecords > PB_MAX_WD_RECS) {
    delete h;
    return fail(PBRT_B200_ERR_UNSUPPORTED, "too many light samples per shaded point for whitted/directlighting");
  }
  if (!isPath && h->nFrames > 64) { delete h; return fail(PBRT_B200_ERR_UNSUPPORTED, "maxdepth too large"); }
  int rc = PBRT_B200_OK;
  cudaError_t ce = cudaStreamCreate(&h->stream);
  if (ce == cudaSuccess) ce = cudaStreamCreate(&h->auxStream);
  if (ce == cudaSuccess) ce = cudaStreamCreate(&h->auxStream2);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evAux2, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evDirect, cudaEventDisableTiming);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->evAux, cudaEventDisableTiming);
  if (ce != cudaSuccess) { delete h; return fail(PBRT_B200_ERR_CUDA, cudaGetErrorString(ce)); }
  SceneView &sv = h->sv;
  const PbrtB200Prim *dprims = nullptr;
#define UP(expr) do { rc = (expr); if (rc != PBRT_B200_OK) { pbrt_b200_scene_destroy(h); return rc; } } while (0)
  if (PB_LEAN_WALK && g_treelets && scene->n_instances == 0 && scene->n_nodes > 1) {
    std::vector<PbrtB200KdNode> ordered;
    pb_treelet_order(scene->nodes, scene->n_nodes, &ordered);
    UP(upload(h, ordered.data(), ordered.size(), &sv.nodes));
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {   // (`ordered` is about to go away)
      pbrt_b200_scene_destroy(h);
      return fail(PBRT_B200_ERR_CUDA, "upload of the kd-tree nodes failed");
    }
    sv.pairedNodes = 1;
  } else
    UP(upload(h, scene->nodes, scene->n_nodes, &sv.nodes));
  if (g_l2Persist > 0 && scene->n_nodes > 0) {
    // tunable 11: the kd-tree nodes as a persisting-L2 access window of this handle's streams, so that the wave's
    // streaming queue traffic does not evict them (experiment; percent of the L2 the device may set aside)
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
      const size_t setAside = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize,
                                               (size_t)prop.l2CacheSize * (size_t)std::min<int64_t>(g_l2Persist, 100) / 100);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside);
      const size_t bytes = (size_t)scene->n_nodes * sizeof(PbrtB200KdNode);
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof(attr));
      attr.accessPolicyWindow.base_ptr = (void *)sv.nodes;
      attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
      attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setAside / (double)attr.accessPolicyWindow.num_bytes);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaStreamSetAttribute(h->auxStream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaStreamSetAttribute(h->auxStream2, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }
  UP(upload(h, scene->prim_refs, scene->n_prim_refs, &sv.primRefs));
  UP(upload(h, scene->prims, scene->n_prims, &dprims));
  sv.prims = (const PrimQuad *)dprims;
  UP(upload(h, scene->prim_aux, scene->n_prims, &sv.aux));
  UP(upload(h, scene->meshes, scene->n_meshes, &sv.meshes));
  UP(upload(h, scene->quadrics, scene->n_quadrics, &sv.quadrics));
  UP(upload(h, scene->instances, scene->n_instances, &sv.instances));
  if (scene->sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE)
    UP(upload(h, scene->bestcandidate_table, (size_t)PB_BC_TABLE * 5, &sv.bcTable));
  UP(upload(h, scene->vertex_n, (size_t)scene->n_vertices * 3, &sv.vertexN));
  UP(upload(h, scene->vertex_s, (size_t)scene->n_vertices * 3, &sv.vertexS));
  UP(upload(h, scene->vertex_uv, (size_t)scene->n_vertices * 2, &sv.vertexUV));
  UP(upload(h, scene->materials, scene->n_materials, &sv.materials));
  UP(upload(h, scene->lights, scene->n_lights + scene->n_unsampled_lights, &sv.lights));
  UP(upload(h, scene->light_shapes, scene->n_light_shapes, &sv.lightShapes));
  UP(upload(h, scene->light_cdf, scene->n_light_shapes, &sv.lightCdf));
  UP(upload(h, scene->textures, scene->n_textures, &sv.textures));
  UP(upload(h, scene->tex_levels, scene->n_tex_levels, &sv.texLevels));
  UP(upload(h, scene->texels, (size_t)scene->n_texel_floats, &sv.texels));
  if (scene->n_textures) UP(upload(h, scene->ewa_weight_lut, (size_t)PBRT_B200_EWA_LUT_SIZE, &sv.ewaLut));
  UP(upload(h, scene->light_distributions, (size_t)scene->n_light_distribution_floats, &sv.lightDist));
  {
    bool noise = false;
    for (uint32_t i = 0; i < scene->n_textures; ++i) noise |= scene->textures[i].kind >= PBRT_B200_TEX_DOTS;
    if (noise) {
      unsigned char perm[2 * PB_NOISE_PERM_SIZE];
      pb_noise_perm_fill(perm);
      UP(upload(h, perm, sizeof(perm), &sv.noisePerm));
    }
  }
#undef UP
  if (scene->sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE) {
    // prefix strip masks of the table's image positions (see k_film_gather_bc)
    std::vector<uint32_t> masks((size_t)2 * (PB_BC_STRIPS + 1) * PB_BC_WORDS, 0u);
    for (int axis = 0; axis < 2; ++axis) {
      uint32_t *base = masks.data() + (size_t)axis * (PB_BC_STRIPS + 1) * PB_BC_WORDS;
      for (int k = 0; k < PB_BC_TABLE; ++k) {
        int sidx = (int)floorf(scene->bestcandidate_table[5 * k + axis] * PB_BC_STRIPS);
        sidx = std::min(std::max(sidx, 0), PB_BC_STRIPS - 1);
        for (int t = sidx + 1; t <= PB_BC_STRIPS; ++t) base[(size_t)t * PB_BC_WORDS + k / 32] |= 1u << (k % 32);
      }
    }
    ce = pool_alloc(h->device, (void **)&h->bcMasks, masks.size() * sizeof(uint32_t));
    if (ce == cudaSuccess)
      ce = cudaMemcpy(h->bcMasks, masks.data(), masks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
      pbrt_b200_scene_destroy(h);
      return fail(PBRT_B200_ERR_CUDA, cudaGetErrorString(ce));
    }
  }
  if (g_primFrames) {
    // hit-independent BSDF frames of flat triangles and the light triangles' normals (pb_scene.h)
    void *pf = nullptr, *lf = nullptr;
    ce = cudaSuccess;
    if (!h->textured && scene->n_prims) ce = pool_alloc(h->device, &pf, (size_t)scene->n_prims * 3 * sizeof(float4));
    if (ce == cudaSuccess && scene->n_light_shapes) ce = pool_alloc(h->device, &lf, (size_t)scene->n_light_shapes * 2 * sizeof(float4));
    if (pf) h->allocs.push_back(pf);
    if (lf) h->allocs.push_back(lf);
    if (ce != cudaSuccess) {
      pbrt_b200_scene_destroy(h);
      return fail(PBRT_B200_ERR_CUDA, cudaGetErrorString(ce));
    }
    if (pf) {
      int grid = (int)std::min<uint64_t>(((uint64_t)scene->n_prims + 127) / 128, (uint64_t)h->numSMs * 16);
      k_prim_frames<<<grid, 128, 0, h->stream>>>(sv, (float4 *)pf, (PrimQuad *)sv.prims);   // (sv.primFrame still null)
      sv.primFrame = (const float4 *)pf;
    }
    if (lf) {
      int grid = (int)std::min<uint64_t>(((uint64_t)scene->n_light_shapes + 127) / 128, (uint64_t)h->numSMs * 16);
      k_light_frames<<<grid, 128, 0, h->stream>>>(sv, scene->n_light_shapes, (float4 *)lf);
      sv.lightFrame = (const float4 *)lf;
    }
  }
  ce = pool_alloc(h->device, (void **)&h->filterTable, 256 * sizeof(float));
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(h->filterTable, scene->film.filter_table, 256 * sizeof(float), cudaMemcpyHostToDevice, h->stream);
  h->sceneBytes += 256 * sizeof(float);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  if (ce != cudaSuccess) {
    pbrt_b200_scene_destroy(h);
    return fail(PBRT_B200_ERR_CUDA, cudaGetErrorString(ce));
  }
  *out = h;
  return PBRT_B200_OK;
}

// The sample extent of a TILE of the frame window `frame`: every sample pixel of the frame's own extent
// (film/image.cpp:148-156) that can reach a pixel of the tile -- one more column / row towards lower coordinates
// than a crop render of the same pixels would sample (a sample whose coordinate rounds up to the next integer
// splats into the next pixel as well, image.cpp:108-111), so that the tile's pixels receive exactly the samples
// they receive in the whole frame.  k_generate culls the samples of this extent that do not reach the tile.
static WavePart tile_part(const SceneView &sv, const PbrtB200Window &frame, const PbrtB200Window &t) {
  WavePart wp;
  memset(&wp, 0, sizeof(wp));
  int fxs, fxe, fys, fye;
  film_sample_extent(sv, frame.x0, frame.y0, frame.width, frame.height, &fxs, &fxe, &fys, &fye);
  wp.xs = std::max(fxs, pb_floor2int_fin(t.x0 + .5f - sv.filterXW) - 1);
  wp.xe = std::min(fxe, pb_floor2int_fin(t.x0 + t.width - 1 + .5f + sv.filterXW) + 2);
  wp.ys = std::max(fys, pb_floor2int_fin(t.y0 + .5f - sv.filterYW) - 1);
  wp.ye = std::min(fye, pb_floor2int_fin(t.y0 + t.height - 1 + .5f + sv.filterYW) + 2);
  wp.tile = 1;
  wp.tx0 = t.x0; wp.ty0 = t.y0; wp.tx1 = t.x0 + t.width; wp.ty1 = t.y0 + t.height;
  return wp;
}

static int render_windows(PbrtB200Handle *h, const PbrtB200Window *windows, int n_windows, PbrtB200Stats *stats,
                          bool tiles);
int pbrt_b200_render(PbrtB200Handle *h, const PbrtB200Window *windows, int n_windows, PbrtB200Stats *stats) {
  return render_windows(h, windows, n_windows, stats, false);
}
int pbrt_b200_render_tiles(PbrtB200Handle *h, const PbrtB200Window *tiles, int n_tiles, PbrtB200Stats *stats) {
  if (h && h->sv.sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE)
    return fail(PBRT_B200_ERR_UNSUPPORTED, "tile mode: the bestcandidate sampler's table placements are tied to the "
                                           "window they start in (samplers/bestcandidate.cpp:81-105)");
  return render_windows(h, tiles, n_tiles, stats, true);
}
static int render_windows(PbrtB200Handle *h, const PbrtB200Window *windows, int n_windows, PbrtB200Stats *stats,
                          bool tiles) {
  if (!h || !windows || n_windows < 1) return fail(PBRT_B200_ERR_INVALID, "bad arguments");
  // PBRT_HOST_PHASES=1: host wall clock of this call's phases to stderr (where a frame's time outside kernels goes)
  static const bool hostPhases = getenv("PBRT_HOST_PHASES") != nullptr;
  auto hostNow = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double hp0 = hostPhases ? hostNow() : 0.0;
  double hp1 = 0, hp2 = 0, hp3 = 0;
  CU_TRY(cudaSetDevice(h->device));
  // keep the film buffers of the previous render when the windows are unchanged (frame loops)
  bool reuseFilms = (int)h->films.size() == n_windows && (h->rgbaSlab != nullptr) == tiles;
  for (int w = 0; reuseFilms && w < n_windows; ++w) {
    const PbrtB200Window &a = h->films[w].win, &b = windows[w];
    reuseFilms = a.x0 == b.x0 && a.y0 == b.y0 && a.width == b.width && a.height == b.height &&
                 (h->films[w].sampleOut != nullptr) == h->keepSamples;
  }
  if (!reuseFilms) free_films(h);
  const SceneView &sv = h->sv;
  uint64_t launches = 0;
  memset(h->classMs, 0, sizeof(h->classMs));
  memset(h->classLaunches, 0, sizeof(h->classLaunches));
  h->evClass.clear();
  ScopedEvents ev;
  CU_TRY(ev.create(2));
  const cudaEvent_t e0 = ev.e[0], e1 = ev.e[1];
  // per-window sample extents; size the wave arena for min(tunable, all samples)
  const bool bc = sv.sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE;
  // samples per unit of work: a pixel's samples, or the 4096 entries of one bestcandidate table placement
  const int spp = bc ? PB_BC_TABLE : sv.sampler.x_samples * sv.sampler.y_samples;
  std::vector<std::vector<float2>> placements(n_windows);
  std::vector<WavePart> ext(n_windows);
  std::vector<uint64_t> nPixW(n_windows);
  uint64_t totalSamples = 0;
  for (int w = 0; w < n_windows; ++w) {
    const PbrtB200Window &win = windows[w];
    if (win.width < 1 || win.height < 1) return fail(PBRT_B200_ERR_INVALID, "empty window");
    if (win.x0 < -16384 || win.y0 < -16384 || win.x0 + win.width > 16384 || win.y0 + win.height > 16384)
      return fail(PBRT_B200_ERR_UNSUPPORTED, "window outside +-16384 pixels");
    if (tiles) {
      const PbrtB200Window &f = h->frameWin;
      if (win.x0 < f.x0 || win.y0 < f.y0 || win.x0 + win.width > f.x0 + f.width || win.y0 + win.height > f.y0 + f.height)
        return fail(PBRT_B200_ERR_INVALID, "tile outside the scene's film window");
      ext[w] = tile_part(sv, f, win);
    } else
      ext[w] = window_part(sv, win);
    ext[w].window = w;
    nPixW[w] = (uint64_t)(ext[w].xe - ext[w].xs) * (ext[w].ye - ext[w].ys);
    if (bc) {
      std::vector<float> cx, cy;
      pb_bc_placements(ext[w].xs, ext[w].xe, ext[w].ys, ext[w].ye, sv.bcTableWidth, &cx, &cy);
      placements[w].resize(cx.size());
      for (size_t i = 0; i < cx.size(); ++i) placements[w][i] = make_float2(cx[i], cy[i]);
      nPixW[w] = cx.size();
      if (nPixW[w] == 0) return fail(PBRT_B200_ERR_INVALID, "bestcandidate sampler: window holds no table placement");
    }
    totalSamples += nPixW[w] * spp;
  }
  // Waves of equal size: the frame's units (pixels / table placements) are split over the fewest waves the
  // tunable allows, and a remainder of less than 1/16 wave is absorbed by growing the waves instead of paying a
  // whole extra round of ~40 launches for it (C3: 67.24 M samples = 4 waves of 16.81 M, not 4 x 16.78 M + 0.13 M).
  const uint64_t units = totalSamples / spp;
  // The tunable bounds a wave; so do the 32-bit record index of the whitted / directlighting queues
  // (record * nSlots + slot) and the device's memory: a wave's arena may take half of what is free now (plus
  // this device's retired arena, which ensure_wave() would reuse).  The per-slot state of whitted / directlighting
  // (recursion frames) is several times the path integrator's: they keep the 16 Mi-sample bound of round 1.
  uint64_t waveSamples = (uint64_t)g_waveSamples;
  if (h->sv.integrator.kind != PBRT_B200_INT_PATH) waveSamples = std::min<uint64_t>(waveSamples, 1ull << 24);
  waveSamples = std::min<uint64_t>(waveSamples, 0xffffffffull / (uint64_t)std::max(1, h->sv.directRecords));
  // two lanes: the frame is cut into an even number of waves (at least two), each lane's arena holds one
  const bool twoLanes = h->sv.integrator.kind == PBRT_B200_INT_PATH && !bc &&
                        (g_lanes == 3 || (g_lanes == 2 && totalSamples >= PB_LANE_MIN_SAMPLES)) && units >= 2;
  // slots of one wave when a wave may hold `ws` samples: equal waves, a remainder below 1/16 wave absorbed
  auto wave_slots = [&](uint64_t ws) {
    if (twoLanes) ws = std::max<uint64_t>(ws / 2, (uint64_t)spp);   // (the bounds are for both lanes together)
    const uint64_t unitCap = std::max<uint64_t>(1, ws / spp);
    uint64_t nWaves = std::max<uint64_t>(1, (units + unitCap - 1) / unitCap);
    if (nWaves > 1 && units - (nWaves - 1) * unitCap <= unitCap / 16) --nWaves;
    if (twoLanes) nWaves = std::max<uint64_t>(2, nWaves + (nWaves & 1));
    return ((units + nWaves - 1) / nWaves) * spp;
  };
  auto arena_bytes = [&](uint64_t slots) {
    Carver sizing = {nullptr, 0};
    WaveBuffers tmp;
    carve_wave(sizing, tmp, (uint32_t)slots, h->sv.directRecords, h->nFrames, h->sv.integrator.kind == PBRT_B200_INT_PATH,
               h->sv.ldFloatsPerSample, h->sv.nInstances != 0, h->textured);
    return sizing.off + 256;
  };
  uint64_t capSlots = wave_slots(waveSamples);
  // The device is asked for its free memory only when neither the arena at hand nor a retired one of this device
  // holds that wave: frame loops and the per-call handles of pbrt_b200_render_host never call cudaMemGetInfo, whose
  // latency next to tens of GB of allocations is erratic (0.1 ms, then 7-44 ms in a quarter of the calls:
  // profiles/r02z_host_gap.log).
  if (capSlots <= 0x7fffffffull / 4 && !(h->arena && (uint64_t)h->wb.nSlots >= capSlots)) {
    const size_t need = arena_bytes(capSlots) * (twoLanes ? 2 : 1);
    size_t reusable = h->arenaBytes;
    {
      std::lock_guard<std::mutex> lock(g_poolMutex);
      if (h->device >= 0 && h->device < PB_MAX_DEVICES)
        reusable = twoLanes ? g_cachedArenaBytes[h->device][0] + g_cachedArenaBytes[h->device][1]
                            : std::max(reusable, std::max(g_cachedArenaBytes[h->device][0], g_cachedArenaBytes[h->device][1]));
    }
    size_t freeB = 0, totalB = 0;
    if (reusable < need && cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
      const double bytesPerSlot = (double)arena_bytes(1u << 20) / (double)(1u << 20);
      const double budget = std::max(0.5 * (double)freeB, (double)reusable);
      waveSamples = std::min<uint64_t>(waveSamples, std::max<uint64_t>(1u << 20, (uint64_t)(budget / bytesPerSlot)));
      capSlots = wave_slots(waveSamples);
    }
  }
  if (capSlots > 0x7fffffffull / 4) return fail(PBRT_B200_ERR_UNSUPPORTED, "wave too large");
  int rc = ensure_wave(h, (uint32_t)capSlots);
  if (rc != PBRT_B200_OK) return rc;
  if (twoLanes) {
    rc = ensure_parked_lane(h);
    if (rc != PBRT_B200_OK) return rc;
    swap_lane(h);
    rc = ensure_wave(h, (uint32_t)capSlots);
    if (rc == PBRT_B200_OK) CU_TRY(cudaMemsetAsync(h->wb.stats, 0, ST_COUNT * sizeof(unsigned long long), h->stream));
    swap_lane(h);
    if (rc != PBRT_B200_OK) return rc;
  }
  CU_TRY(cudaMemsetAsync(h->wb.stats, 0, ST_COUNT * sizeof(unsigned long long), h->stream));
  uint64_t cameraRays = 0;
  CU_TRY(cudaEventRecord(e0, h->stream));
  for (int w = 0; w < n_windows; ++w) {
    const PbrtB200Window &win = windows[w];
    size_t nPixOut = (size_t)win.width * win.height;
    if (reuseFilms) {
      if (!h->filmSlab) CU_TRY(cudaMemsetAsync(h->films[w].film, 0, nPixOut * sizeof(FilmPixel), h->stream));
      else if (w == 0) CU_TRY(cudaMemsetAsync(h->filmSlab, 0, h->rgbaSlabFloats / 4 * sizeof(FilmPixel), h->stream));
      continue;
    }
    h->films.push_back(WindowFilm());
    WindowFilm &wf = h->films.back();   // registered first: an allocation failure below leaves it to free_films
    wf.win = win; wf.film = nullptr; wf.rgba = nullptr; wf.sampleOut = nullptr; wf.corners = nullptr;
    if (bc) {
      wf.cornersHost = placements[w];
      CU_TRY(pool_alloc(h->device, (void **)&wf.corners, placements[w].size() * sizeof(float2)));
      CU_TRY(cudaMemcpyAsync(wf.corners, wf.cornersHost.data(), placements[w].size() * sizeof(float2),
                             cudaMemcpyHostToDevice, h->stream));
    }
    if (!tiles) CU_TRY(pool_alloc(h->device, (void **)&wf.film, nPixOut * sizeof(FilmPixel)));
    if (tiles) {
      // tile renders resolve into ONE block, window after window, so that a rank's finished tiles leave in a
      // single transfer (multi.cuh)
      if (w == 0) {
        size_t total = 0;
        for (int k = 0; k < n_windows; ++k) total += (size_t)windows[k].width * windows[k].height;
        CU_TRY(pool_alloc(h->device, (void **)&h->rgbaSlab, total * sizeof(float4)));
        h->rgbaSlabFloats = total * 4;
        CU_TRY(pool_alloc(h->device, (void **)&h->filmSlab, total * sizeof(FilmPixel)));
        CU_TRY(cudaMemsetAsync(h->filmSlab, 0, total * sizeof(FilmPixel), h->stream));
      }
      size_t before = 0;
      for (int k = 0; k < w; ++k) before += (size_t)windows[k].width * windows[k].height;
      wf.rgba = h->rgbaSlab + before;
      wf.film = h->filmSlab + before;
    } else
      CU_TRY(pool_alloc(h->device, (void **)&wf.rgba, nPixOut * sizeof(float4)));
    if (!tiles) CU_TRY(cudaMemsetAsync(wf.film, 0, nPixOut * sizeof(FilmPixel), h->stream));
    wf.nSamples = nPixW[w] * spp;
    if (h->keepSamples) CU_TRY(pool_alloc(h->device, (void **)&wf.sampleOut, wf.nSamples * 4 * sizeof(float)));
  }
  if (hostPhases) hp1 = hostNow();
  // waves: consecutive samples in reference order, spanning windows when a window is small
  int curW = 0;
  uint64_t curP = 0;
  int waveIdx = 0;
  if (twoLanes) {
    // the other lane starts after the film buffers are cleared
    CU_TRY(cudaEventRecord(h->evStart, h->stream));
    CU_TRY(cudaStreamWaitEvent(h->parked.stream, h->evStart, 0));
  }
  while (curW < n_windows) {
    // odd waves run on the second lane; the film accumulates wave after wave whichever lane a wave ran on
    // (a pixel's sum takes its samples in the reference's order, film/image.cpp:103-141)
    const bool otherLane = twoLanes && (waveIdx & 1);
    if (otherLane) swap_lane(h);
    struct LaneGuard {   // back to the first lane on every way out of this iteration
      PbrtB200Handle *h; bool on;
      ~LaneGuard() { if (on) swap_lane(h); }
    } laneGuard = {h, otherLane};
    WaveDesc wd;
    memset(&wd, 0, sizeof(wd));
    wd.spp = spp;
    wd.bc = bc ? 1 : 0;
    wd.tiles = tiles ? 1 : 0;
    while (curW < n_windows && wd.nParts < PB_MAX_PARTS && wd.nSlots < capSlots) {
      uint64_t room = (capSlots - wd.nSlots) / spp;
      uint64_t take = std::min<uint64_t>(room, nPixW[curW] - curP);
      if (take == 0) break;
      WavePart &wp = wd.parts[wd.nParts++];
      wp = ext[curW];
      wp.firstPixel = (uint32_t)curP;
      wp.nPixels = (uint32_t)take;
      wp.slotBase = wd.nSlots;
      wp.corners = bc ? h->films[curW].corners : nullptr;
      wp.bcCols = 0;
      if (bc) {
        // placements per row, when the window's rows are all alike
        const std::vector<float2> &pc = h->films[curW].cornersHost;
        size_t cols = 1;
        while (cols < pc.size() && pc[cols].y == pc[0].y) ++cols;
        bool regular = pc.size() % cols == 0;
        for (size_t r = 0; regular && r < pc.size() / cols; ++r)
          regular = pc[r * cols].y == pc[r * cols + cols - 1].y && (r == 0 || pc[r * cols].y != pc[r * cols - 1].y);
        wp.bcCols = regular ? (int)cols : 0;
      }
      wd.nSlots += (uint32_t)(take * spp);
      curP += take;
      if (curP == nPixW[curW]) { ++curW; curP = 0; }
    }
    rc = run_wave(h, wd, &launches);
    if (rc != PBRT_B200_OK) return rc;
    if (!bc) {
      uint32_t nPixWave = wd.nSlots / (uint32_t)spp;
      int grid = (int)std::min<uint32_t>((nPixWave + 127) / 128, (uint32_t)h->numSMs * 16);
      TIMED(KC_FILM, (k_film_reach<<<grid, 128, 0, h->stream>>>(sv, h->wb, nPixWave, spp)));
      ++launches;
    }
    if (twoLanes && waveIdx > 0) CU_TRY(cudaStreamWaitEvent(h->stream, h->evFilm[(waveIdx - 1) & 1], 0));
    const bool gatherTogether = !bc && wd.nParts > 1;
    FilmTargets tg;
    int tgMaxTotal = 0;
    for (int pi = 0; pi < wd.nParts; ++pi) {
      const WavePart &wp = wd.parts[pi];
      WindowFilm &wf = h->films[wp.window];
      const PbrtB200Window &win = wf.win;
      int ew = wp.xe - wp.xs;
      if (bc) {
        // pixel rows the part's table placements can reach
        float yMin = 1e30f, yMax = -1e30f;
        for (uint32_t t = 0; t < wp.nPixels; ++t) {
          float y = wf.cornersHost[wp.firstPixel + t].y;
          yMin = std::min(yMin, y); yMax = std::max(yMax, y);
        }
        int rowFirst = std::max(win.y0, (int)floorf(yMin - 1.5f - sv.filterYW));
        int rowLast = std::min(win.y0 + win.height - 1, (int)ceilf(yMax + sv.bcTableWidth + 0.5f + sv.filterYW));
        if (rowLast >= rowFirst) {
          int rows = rowLast - rowFirst + 1;
          int total = rows * win.width;
          int grid = std::min((total + PB_BLOCK - 1) / PB_BLOCK, h->numSMs * 16);
          TIMED(KC_FILM, (k_film_gather_bc<<<grid, PB_BLOCK, 0, h->stream>>>(sv, h->wb, wp, h->filterTable, h->bcMasks,
                                                                             wf.film, win, rowFirst, rows)));
          ++launches;
        }
      }
      // pixel rows this part's samples can reach
      int syFirst = wp.ys + (int)(wp.firstPixel / ew), syLast = wp.ys + (int)((wp.firstPixel + wp.nPixels - 1) / ew);
      int rowFirst = std::max(win.y0, (int)floorf(syFirst - 0.5f - sv.filterYW) - 1);
      int rowLast = std::min(win.y0 + win.height - 1, (int)ceilf(syLast + 1 - 0.5f + sv.filterYW) + 1);
      if (gatherTogether) {
        tg.film[pi] = wf.film; tg.win[pi] = win; tg.rowFirst[pi] = rowFirst;
        tg.rowCount[pi] = std::max(0, rowLast - rowFirst + 1);
        tgMaxTotal = std::max(tgMaxTotal, tg.rowCount[pi] * win.width);
      } else if (!bc && rowLast >= rowFirst) {
        int rows = rowLast - rowFirst + 1;
        int total = rows * win.width;
        int grid = std::min((total + PB_BLOCK - 1) / PB_BLOCK, h->numSMs * 16);
        TIMED(KC_FILM, (k_film_gather<<<grid, PB_BLOCK, 0, h->stream>>>(sv, h->wb, wp, spp, h->filterTable, wf.film, win,
                                                                        rowFirst, rows)));
        ++launches;
      }
      if (wf.sampleOut) {
        // interleave L.rgb, alpha for the test entry point
        uint32_t nS = wp.nPixels * (uint32_t)spp;
        for (int c = 0; c < 4; ++c) {
          const float *src = (c < 3 ? h->wb.outL[c] : h->wb.outA) + wp.slotBase;
          CU_TRY(cudaMemcpy2DAsync(wf.sampleOut + (size_t)wp.firstPixel * spp * 4 + c, 4 * sizeof(float), src,
                                   sizeof(float), sizeof(float), nS, cudaMemcpyDeviceToDevice, h->stream));
        }
      }
    }
    if (gatherTogether && tgMaxTotal > 0) {
      // (parts of one wave never share a pixel of the same window twice in a launch: a window's parts are
      // consecutive pixel ranges of ONE window only when the window spans waves, and then it is a single part here)
      dim3 grid((unsigned)std::min((tgMaxTotal + PB_BLOCK - 1) / PB_BLOCK, h->numSMs * 4), (unsigned)wd.nParts);
      TIMED(KC_FILM, (k_film_gather_parts<<<grid, PB_BLOCK, 0, h->stream>>>(sv, h->wb, wd, h->filterTable, tg)));
      ++launches;
    }
    if (twoLanes) CU_TRY(cudaEventRecord(h->evFilm[waveIdx & 1], h->stream));
    cameraRays += wd.nSlots;
    ++waveIdx;
  }
  if (twoLanes && waveIdx > 0) CU_TRY(cudaStreamWaitEvent(h->stream, h->evFilm[(waveIdx - 1) & 1], 0));
  if (h->filmSlab) {
    const size_t total = h->rgbaSlabFloats / 4;
    int grid = std::min((int)((total + 255) / 256), h->numSMs * 16);
    TIMED(KC_FILM, (k_film_resolve<<<grid, 256, 0, h->stream>>>(h->filmSlab, h->rgbaSlab, (int)total, sv.premultiply)));
    ++launches;
  }
  for (int w = 0; w < n_windows && !h->filmSlab; ++w) {
    WindowFilm &wf = h->films[w];
    size_t nPixOut = (size_t)wf.win.width * wf.win.height;
    int grid = std::min((int)((nPixOut + 255) / 256), h->numSMs * 16);
    TIMED(KC_FILM, (k_film_resolve<<<grid, 256, 0, h->stream>>>(wf.film, wf.rgba, (int)nPixOut, sv.premultiply)));
    ++launches;
  }
  CU_TRY(cudaEventRecord(e1, h->stream));
  if (hostPhases) hp2 = hostNow();
  CU_TRY(cudaStreamSynchronize(h->stream));
  if (hostPhases) hp3 = hostNow();
  CU_TRY(cudaGetLastError());
  float ms = 0.f;
  CU_TRY(cudaEventElapsedTime(&ms, e0, e1));
  for (size_t i = 0; i < h->evClass.size(); ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, h->evPool[2 * i], h->evPool[2 * i + 1]) == cudaSuccess) h->classMs[h->evClass[i]] += t;
  }
  {
    unsigned long long *hs = h->lastStats;
    CU_TRY(cudaMemcpy(hs, h->wb.stats, sizeof(h->lastStats), cudaMemcpyDeviceToHost));
    if (twoLanes) {
      unsigned long long other[ST_COUNT];
      CU_TRY(cudaMemcpy(other, h->parked.wb.stats, sizeof(other), cudaMemcpyDeviceToHost));
      for (int k = 0; k < ST_COUNT; ++k) hs[k] += other[k];
    }
  }
  if (stats) {
    const unsigned long long *hs = h->lastStats;
    memset(stats, 0, sizeof(*stats));
    stats->camera_rays = (bc || tiles) ? hs[ST_CAMERA] : cameraRays;
    stats->intersect_rays = hs[ST_K(TK_MAIN, TF_RAYS)] + hs[ST_K(TK_PROBE, TF_RAYS)];
    stats->shadow_rays = hs[ST_K(TK_SHADOW, TF_RAYS)];
    stats->points_shaded = hs[ST_SHADED];
    for (int k = 0; k < TK_COUNT; ++k) {
      stats->interior_visits += hs[ST_K(k, TF_INTERIOR)]; stats->leaf_visits += hs[ST_K(k, TF_LEAVES)];
      stats->prim_refs_read += hs[ST_K(k, TF_REFS)]; stats->prim_tests += hs[ST_K(k, TF_TESTS)];
    }
    stats->nan_samples = hs[ST_NAN]; stats->negative_samples = hs[ST_NEG]; stats->infinite_samples = hs[ST_INF];
    stats->render_ms = ms;
    stats->kernel_launches = launches;
  }
  if (hostPhases)
    fprintf(stderr, "pbrt_b200 host phases: setup %.3f ms, enqueue %.3f ms, wait %.3f ms, wrap-up %.3f ms; device %.3f ms, "
                    "%llu launches\n", hp1 - hp0, hp2 - hp1, hp3 - hp2, hostNow() - hp3, ms, (unsigned long long)launches);
  return PBRT_B200_OK;
}

int pbrt_b200_kernel_stats(PbrtB200Handle *h, PbrtB200KernelStats *out, int max_out, int *n_out) {
  if (!h || !out || !n_out) return fail(PBRT_B200_ERR_INVALID, "null argument");
  int n = 0;
  for (int c = 0; c < KC_COUNT && n < max_out; ++c, ++n) {
    PbrtB200KernelStats &k = out[n];
    memset(&k, 0, sizeof(k));
    strncpy(k.name, kKernelClassNames[c], sizeof(k.name) - 1);
    k.launches = h->classLaunches[c];
    k.ms = h->classMs[c];
    int tk = c == KC_INTERSECT ? TK_MAIN : c == KC_SHADOW ? TK_SHADOW : c == KC_PROBE ? TK_PROBE : -1;
    if (tk >= 0) {
      k.rays = h->lastStats[ST_K(tk, TF_RAYS)];
      k.interior_visits = h->lastStats[ST_K(tk, TF_INTERIOR)]; k.leaf_visits = h->lastStats[ST_K(tk, TF_LEAVES)];
      k.prim_refs_read = h->lastStats[ST_K(tk, TF_REFS)]; k.prim_tests = h->lastStats[ST_K(tk, TF_TESTS)];
    }
  }
  *n_out = n;
  return PBRT_B200_OK;
}

int pbrt_b200_film_download(PbrtB200Handle *h, int w, float *rgba, float *raw) {
  if (!h || w < 0 || w >= (int)h->films.size()) return fail(PBRT_B200_ERR_INVALID, "no such window");
  CU_TRY(cudaSetDevice(h->device));
  const WindowFilm &f = h->films[w];
  size_t n = (size_t)f.win.width * f.win.height;
  if (rgba) CU_TRY(cudaMemcpy(rgba, f.rgba, n * sizeof(float4), cudaMemcpyDeviceToHost));
  if (raw) CU_TRY(cudaMemcpy(raw, f.film, n * sizeof(FilmPixel), cudaMemcpyDeviceToHost));
  return PBRT_B200_OK;
}

int pbrt_b200_film_device_ptr(PbrtB200Handle *h, int w, void **dptr, uint64_t *n_floats) {
  if (!h || w < 0 || w >= (int)h->films.size()) return fail(PBRT_B200_ERR_INVALID, "no such window");
  *dptr = h->films[w].rgba;
  *n_floats = (uint64_t)h->films[w].win.width * h->films[w].win.height * 4;
  return PBRT_B200_OK;
}

int pbrt_b200_render_host(const PbrtB200Scene *scene, int device, const PbrtB200Window *windows,
                          int n_windows, float *const *rgba_out, float *const *raw_out,
                          PbrtB200Stats *stats) {
  PbrtB200Handle *h = nullptr;
  int nDev = 0;
  if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    return fail(PBRT_B200_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  CU_TRY(cudaSetDevice(device));
  ScopedEvents ev;
  CU_TRY(ev.create(4));
  const cudaEvent_t e0 = ev.e[0], e1 = ev.e[1], e2 = ev.e[2], e3 = ev.e[3];
  CU_TRY(cudaEventRecord(e0, 0));
  int rc = pbrt_b200_scene_create(scene, device, &h);
  if (rc != PBRT_B200_OK) return rc;
  struct HandleGuard {
    PbrtB200Handle *&h;
    ~HandleGuard() { if (h) pbrt_b200_scene_destroy(h); }
  } guard = {h};
  CU_TRY(cudaEventRecord(e1, 0));
  PbrtB200Stats st;
  rc = pbrt_b200_render(h, windows, n_windows, &st);
  if (rc != PBRT_B200_OK) return rc;
  CU_TRY(cudaEventRecord(e2, 0));
  uint64_t d2h = 0;
  for (int w = 0; w < n_windows; ++w) {
    rc = pbrt_b200_film_download(h, w, rgba_out ? rgba_out[w] : nullptr, raw_out ? raw_out[w] : nullptr);
    if (rc != PBRT_B200_OK) return rc;
    size_t n = (size_t)windows[w].width * windows[w].height;
    if (rgba_out && rgba_out[w]) d2h += n * 16;
    if (raw_out && raw_out[w]) d2h += n * 20;
  }
  CU_TRY(cudaEventRecord(e3, 0));
  CU_TRY(cudaEventSynchronize(e3));
  float up = 0, down = 0;
  cudaEventElapsedTime(&up, e0, e1);
  cudaEventElapsedTime(&down, e2, e3);
  st.upload_ms = up; st.download_ms = down;
  st.h2d_bytes = h->sceneBytes; st.d2h_bytes = d2h;
  if (stats) *stats = st;
  pbrt_b200_scene_destroy(h);
  h = nullptr;
  return PBRT_B200_OK;
}

int pbrt_b200_intersect(PbrtB200Handle *h, const float *rays, uint64_t n, int32_t *hit_prim, float *hit_t,
                        float *hit_b1, float *hit_b2) {
  if (!h) return fail(PBRT_B200_ERR_INVALID, "null handle");
  if (n == 0) return PBRT_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  float *dr; int32_t *dp; float *dt, *d1, *d2;
  CU_TRY(cudaMalloc((void **)&dr, n * 8 * sizeof(float)));
  CU_TRY(cudaMalloc((void **)&dp, n * sizeof(int32_t)));
  CU_TRY(cudaMalloc((void **)&dt, n * sizeof(float)));
  CU_TRY(cudaMalloc((void **)&d1, n * sizeof(float)));
  CU_TRY(cudaMalloc((void **)&d2, n * sizeof(float)));
  CU_TRY(cudaMemcpy(dr, rays, n * 8 * sizeof(float), cudaMemcpyHostToDevice));
  int grid = (int)std::min<uint64_t>((n + 127) / 128, (uint64_t)h->numSMs * 16);
  k_rays_closest<<<grid, 128, 0, h->stream>>>(h->sv, dr, n, dp, dt, d1, d2);
  CU_TRY(cudaStreamSynchronize(h->stream));
  CU_TRY(cudaMemcpy(hit_prim, dp, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(hit_t, dt, n * sizeof(float), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(hit_b1, d1, n * sizeof(float), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(hit_b2, d2, n * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(dp); cudaFree(dt); cudaFree(d1); cudaFree(d2);
  return PBRT_B200_OK;
}

int pbrt_b200_intersect_p(PbrtB200Handle *h, const float *rays, uint64_t n, uint8_t *occluded) {
  if (!h) return fail(PBRT_B200_ERR_INVALID, "null handle");
  if (n == 0) return PBRT_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  float *dr; uint8_t *docc;
  CU_TRY(cudaMalloc((void **)&dr, n * 8 * sizeof(float)));
  CU_TRY(cudaMalloc((void **)&docc, n));
  CU_TRY(cudaMemcpy(dr, rays, n * 8 * sizeof(float), cudaMemcpyHostToDevice));
  int grid = (int)std::min<uint64_t>((n + 127) / 128, (uint64_t)h->numSMs * 16);
  k_rays_any<<<grid, 128, 0, h->stream>>>(h->sv, dr, n, docc);
  CU_TRY(cudaStreamSynchronize(h->stream));
  CU_TRY(cudaMemcpy(occluded, docc, n, cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(docc);
  return PBRT_B200_OK;
}

int pbrt_b200_camera_rays(PbrtB200Handle *h, const PbrtB200Window *w, uint64_t n, float *rays, float *image_xy) {
  if (!h || !w) return fail(PBRT_B200_ERR_INVALID, "null argument");
  if (n == 0) return PBRT_B200_OK;
  CU_TRY(cudaSetDevice(h->device));
  if (h->sv.sampler.kind != PBRT_B200_SAMPLER_STRATIFIED)
    return fail(PBRT_B200_ERR_UNSUPPORTED, "pbrt_b200_camera_rays: stratified sampler only (test entry point)");
  WavePart wd = window_part(h->sv, *w);
  int spp = h->sv.sampler.x_samples * h->sv.sampler.y_samples;
  uint64_t total = (uint64_t)(wd.xe - wd.xs) * (wd.ye - wd.ys) * spp;
  if (n > total) return fail(PBRT_B200_ERR_INVALID, "more rays requested than the window has samples");
  float *dr, *dxy;
  CU_TRY(cudaMalloc((void **)&dr, n * 8 * sizeof(float)));
  CU_TRY(cudaMalloc((void **)&dxy, n * 2 * sizeof(float)));
  int grid = (int)std::min<uint64_t>((n + 127) / 128, (uint64_t)h->numSMs * 16);
  k_camera_rays<<<grid, 128, 0, h->stream>>>(h->sv, wd, spp, n, dr, dxy);
  CU_TRY(cudaStreamSynchronize(h->stream));
  CU_TRY(cudaMemcpy(rays, dr, n * 8 * sizeof(float), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(image_xy, dxy, n * 2 * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(dr); cudaFree(dxy);
  return PBRT_B200_OK;
}

uint64_t pbrt_b200_sample_slots(PbrtB200Handle *h, int w) {
  if (!h || w < 0 || w >= (int)h->films.size()) return 0;
  return h->films[w].nSamples;
}

int pbrt_b200_sample_radiance(PbrtB200Handle *h, int w, uint64_t n, float *ls) {
  if (!h) return fail(PBRT_B200_ERR_INVALID, "null handle");
  if (ls == nullptr) {   // arm: keep per-sample radiance on the next render
    h->keepSamples = n != 0;
    return PBRT_B200_OK;
  }
  if (w < 0 || w >= (int)h->films.size() || !h->films[w].sampleOut)
    return fail(PBRT_B200_ERR_INVALID, "per-sample radiance was not kept: call with ls == NULL, n = 1 before rendering");
  if (n > h->films[w].nSamples) return fail(PBRT_B200_ERR_INVALID, "more samples requested than rendered");
  CU_TRY(cudaSetDevice(h->device));
  CU_TRY(cudaMemcpy(ls, h->films[w].sampleOut, n * 4 * sizeof(float), cudaMemcpyDeviceToHost));
  return PBRT_B200_OK;
}

}  // extern "C"

#include "multi.cuh"
