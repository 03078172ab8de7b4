// twin.cpp -- CPU restatement of the render hot path, sample by sample.
//
// TEST INFRASTRUCTURE (oracle/): only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg may load this.  The product (libpbrt_b200.so)
// never links or calls it and has no CPU fallback.
//
// What it is: a scalar, single-threaded driver that walks image samples in the
// reference's order (core/scene.cpp:43-84) and, for each, runs sample
// generation -> camera ray -> kd-tree closest hit -> integrator vertex ->
// shadow / MIS-probe rays -> film splat, using the per-stage arithmetic of
// pbrt-v1_b200/csrc/pb_*.h (each function there cites the reference lines it
// restates) compiled for the host with gcc -ffp-contract=off.  It exists so
// that the stage arithmetic can be pinned against the compiled reference
// (oracle/_ref, built by oracle/build_ref.sh) in the GPU-less build container,
// and so that GPU results can be compared stage by stage (rays, hits,
// per-sample radiance) on the GPU box, where /root/reference does not exist.
//
// Pinning: tests/test_oracle.py renders the parity scenes with
// oracle/_ref/pbrt_ref_keyed and compares images and per-sample radiance with
// this twin; committed fixtures from the same runs live in tests/golden/.
#include "../pbrt-v1_b200/csrc/pb_integrator.h"
#include "../pbrt-v1_b200/csrc/pb_camera.h"
#include "../pbrt-v1_b200/csrc/pb_film.h"
#include "../pbrt-v1_b200/csrc/pb_setup.h"
#include "../pbrt-v1_b200/csrc/pb_noise_table.h"
#include "../pbrt-v1_b200/csrc/pb_treelet.h"
#include <vector>
#include <string>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>

namespace {
std::string gError;

bool build_view(const PbrtB200Scene &s, SceneView *sv) {
  memset(sv, 0, sizeof(*sv));
  std::string err = pb_setup_view(s, sv);
  if (!err.empty()) { gError = err; return false; }
  sv->nodes = s.nodes; sv->primRefs = s.prim_refs;
  sv->prims = (const PrimQuad *)s.prims; sv->aux = s.prim_aux;
  sv->meshes = s.meshes; sv->quadrics = s.quadrics; sv->instances = s.instances;
  sv->bcTable = s.bestcandidate_table;
  sv->textures = s.textures; sv->texLevels = s.tex_levels; sv->texels = s.texels; sv->ewaLut = s.ewa_weight_lut;
  sv->lightDist = s.light_distributions;
  static unsigned char noisePerm[2 * PB_NOISE_PERM_SIZE];
  pb_noise_perm_fill(noisePerm);
  sv->noisePerm = noisePerm;
  sv->vertexN = s.vertex_n; sv->vertexS = s.vertex_s; sv->vertexUV = s.vertex_uv;
  sv->materials = s.materials; sv->lights = s.lights;
  sv->lightShapes = s.light_shapes; sv->lightCdf = s.light_cdf;
  return true;
}

struct Counters {
  uint64_t intersect, shadow, shaded;
  KdCounters kd;
  uint64_t nan, neg, inf;
};

struct VecFrameIO {
  std::vector<WdFrame> *frames;
  WdFrame load(int d) { return (*frames)[d]; }
  void store(int d, const WdFrame &f) {
    if ((int)frames->size() <= d) frames->resize(d + 1);
    (*frames)[d] = f;
  }
};

void add_counts(Counters &c, const KdCounters &k) {
  c.kd.interior += k.interior; c.kd.leaves += k.leaves; c.kd.refs += k.refs; c.kd.tests += k.tests;
}

bool trace_closest(const SceneView &sv, const Ray &ray, Hit *hit, Counters &c) {
  KdCounters k = {0, 0, 0, 0};
  ++c.intersect;
  bool r = kd_intersect<false, true>(sv, ray, hit, &k);
  add_counts(c, k);
  return r;
}
bool trace_any(const SceneView &sv, const Ray &ray, Counters &c) {
  KdCounters k = {0, 0, 0, 0};
  Hit h;
  ++c.shadow;
  bool r = kd_intersect<true, true>(sv, ray, &h, &k);
  add_counts(c, k);
  return r;
}

// resolve the two ray queries of a DirectRec; returns flag bits (1 unoccluded, 2 probe hit)
uint8_t trace_direct(const SceneView &sv, const DirectRec &rec, Counters &c) {
  uint8_t f = 0;
  if (rec.hasShadow && !trace_any(sv, rec.shadow, c)) f |= 1;
  if (rec.hasProbe) {
    Hit h;
    trace_closest(sv, rec.probe, &h, c);
    if (hit_emits_from_light(sv, rec.probe, h, rec.probeLight)) f |= 2;
  }
  return f;
}

Spec li_path(const SceneView &sv, const PatternSource &key0, uint64_t key1, const Ray &camRay, const RayDiff &camDiff,
             float *alpha, Counters &c) {
  PathState st;
  st.L = spec(0.f); st.T = spec(1.f); st.pendT = spec(0.f);
  st.alpha = 0.f; st.depth = 0; st.specularBounce = false;
  st.rng.key = key1; st.rng.ctr = 0;
  Ray ray = camRay;
  for (;;) {
    Hit hit;
    trace_closest(sv, ray, &hit, c);
    PathVertexOut out;
    // only the camera ray has differentials: path.cpp:133 continues with `RayDifferential(p, wi)`
    path_vertex<true>(sv, key0, st, ray, hit, &out, st.depth == 0 ? &camDiff : (const RayDiff *)0);
    if (hit.prim >= 0) ++c.shaded;
    if (out.hasDirect) {
      uint8_t f = trace_direct(sv, out.direct, c);
      path_resolve_direct(sv, st, out.direct.termA, out.direct.termB, (f & 1) != 0, (f & 2) != 0);
    }
    if (!out.hasNext) break;
    ray = out.next;
  }
  *alpha = st.alpha;
  return st.L;
}

Spec li_wd(const SceneView &sv, const PatternSource &key0, uint64_t key1, const Ray &camRay, const RayDiff &camDiff,
           float *alpha, Counters &c) {
  RngStream rng; rng.key = key1; rng.ctr = 0;
  std::vector<WdFrame> frames(1);
  VecFrameIO io = {&frames};
  std::vector<DirectRec> recs(wd_max_records(sv));
  std::vector<uint8_t> flags(recs.size());
  Ray ray = camRay;
  int depth = 0;
  *alpha = 0.f;
  for (;;) {
    Hit hit;
    trace_closest(sv, ray, &hit, c);
    WdFrame fr;
    WdVertexOut out;
    // the ray into depth d is the reflection (parent stage 0) or transmission child of frame d - 1
    const RayDiff *rd = depth == 0 ? &camDiff : (frames[depth - 1].stage == 0 ? &frames[depth - 1].dR : &frames[depth - 1].dT);
    wd_vertex<true>(sv, key0, ray, hit, rng, depth, depth == 0 ? alpha : (float *)0, &fr, recs.data(), &out, rd);
    if (hit.prim >= 0) ++c.shaded;
    for (int r = 0; r < out.nRecs; ++r) flags[r] = trace_direct(sv, recs[r], c);
    wd_resolve(sv, &fr, recs.data(), flags.data(), out.nRecs);
    io.store(depth, fr);
    if (!wd_advance(io, rng, &depth, &ray)) break;
  }
  return frames[0].L;
}

}  // namespace

extern "C" {

const char *pbrt_twin_last_error(void) { return gError.c_str(); }

// Render one pixel window exactly like a reference crop-window render.
//   rgba     width*height*4   what ImageFilm::WriteImage hands to WriteRGBAImage
//   raw      width*height*5   L.rgb, alpha, weightSum (may be NULL)
//   samples  nSamples*6       imageX, imageY, L.rgb (after guards), alpha, reference order (may be NULL)
// `frame` != NULL: `win` is a TILE of the frame window (include/pbrt_b200.h, pbrt_b200_render_tiles): the extent
// holds every sample pixel of the frame's extent that can reach the tile, and samples whose filter footprint
// misses the tile are skipped before they are traced.
static int twin_render_impl(const PbrtB200Scene *scene, const PbrtB200Window *win, const PbrtB200Window *frame,
                            float *rgba, float *raw, float *samples, PbrtB200Stats *stats) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  int xs, xe, ys, ye;
  film_sample_extent(sv, win->x0, win->y0, win->width, win->height, &xs, &xe, &ys, &ye);
  if (frame) {
    if (sv.sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE) return PBRT_B200_ERR_UNSUPPORTED;
    int fxs, fxe, fys, fye;
    film_sample_extent(sv, frame->x0, frame->y0, frame->width, frame->height, &fxs, &fxe, &fys, &fye);
    xs = pb_imax(fxs, pb_floor2int(win->x0 + .5f - sv.filterXW) - 1);
    xe = pb_imin(fxe, pb_floor2int(win->x0 + win->width - 1 + .5f + sv.filterXW) + 2);
    ys = pb_imax(fys, pb_floor2int(win->y0 + .5f - sv.filterYW) - 1);
    ye = pb_imin(fye, pb_floor2int(win->y0 + win->height - 1 + .5f + sv.filterYW) + 2);
  }
  int spp = sv.sampler.x_samples * sv.sampler.y_samples;
  std::vector<FilmPixel> film((size_t)win->width * win->height);
  memset(film.data(), 0, film.size() * sizeof(FilmPixel));
  Counters c;
  memset(&c, 0, sizeof(c));
  uint64_t nSamples = 0;
  const bool ld = sv.sampler.kind == PBRT_B200_SAMPLER_LOWDISCREPANCY;
  std::vector<float> ldTables(ld ? (size_t)sv.ldFloatsPerSample * spp : 0);
  auto t0 = std::chrono::steady_clock::now();
  // everything after the sampler: camera ray, Li, guards, film (core/scene.cpp:43-78)
  auto take_sample = [&](CameraSample cs, const PatternSource &key0, uint64_t key1) {
    if (frame) {
      float dX = film_image_coord(cs.imageX) - 0.5f, dY = film_image_coord(cs.imageY) - 0.5f;
      if (pb_floor2int(dX + sv.filterXW) < win->x0 || pb_ceil2int(dX - sv.filterXW) >= win->x0 + win->width ||
          pb_floor2int(dY + sv.filterYW) < win->y0 || pb_ceil2int(dY - sv.filterYW) >= win->y0 + win->height)
        return;
    }
    Ray ray;
    perspective_generate_ray(sv.camera, cs, &ray, sv.xRes, sv.yRes);
    RayDiff camDiff;
    camera_ray_differentials(sv, cs, &camDiff);
    float alpha = 0.f;
    Spec Ls = (sv.integrator.kind == PBRT_B200_INT_PATH)
                  ? li_path(sv, key0, key1, ray, camDiff, &alpha, c)
                  : li_wd(sv, key0, key1, ray, camDiff, &alpha, c);
    Ls = 1.f * Ls;   // rayWeight * Li, scene.cpp:58
    cs.imageX = film_image_coord(cs.imageX);
    cs.imageY = film_image_coord(cs.imageY);
    int g = sample_guard(&Ls);
    if (g == 1) ++c.nan; else if (g == 2) ++c.neg; else if (g == 3) ++c.inf;
    if (samples) {
      float *o = samples + 6 * nSamples;
      o[0] = cs.imageX; o[1] = cs.imageY; o[2] = Ls.c[0]; o[3] = Ls.c[1]; o[4] = Ls.c[2]; o[5] = alpha;
    }
    ++nSamples;
    // ImageFilm::AddSample, image.cpp:103-141
    float dImageX = cs.imageX - 0.5f, dImageY = cs.imageY - 0.5f;
    int x0 = pb_imax(pb_ceil2int(dImageX - sv.filterXW), win->x0);
    int x1 = pb_imin(pb_floor2int(dImageX + sv.filterXW), win->x0 + win->width - 1);
    int y0 = pb_imax(pb_ceil2int(dImageY - sv.filterYW), win->y0);
    int y1 = pb_imin(pb_floor2int(dImageY + sv.filterYW), win->y0 + win->height - 1);
    for (int py = y0; py <= y1; ++py)
      for (int px = x0; px <= x1; ++px) {
        float wt;
        if (!film_sample_weight(sv, scene->film.filter_table, cs.imageX, cs.imageY, px, py, &wt))
          continue;
        film_pixel_add(&film[(size_t)(py - win->y0) * win->width + (px - win->x0)], wt, Ls.c, alpha);
      }
  };
  if (sv.sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE) {
    // samplers/bestcandidate.cpp:78-171: table placements in scan order, their 4096 entries in order
    std::vector<float> cxs, cys;
    pb_bc_placements(xs, xe, ys, ye, sv.bcTableWidth, &cxs, &cys);
    std::vector<float> tables((size_t)sv.ldFloatsPerSample * PB_BC_TABLE);
    std::vector<uint16_t> perm(PB_BC_TABLE);
    for (size_t t = 0; t < cxs.size(); ++t) {
      int cx = pb_floor2int(cxs[t]), cy = pb_floor2int(cys[t]);
      uint64_t key3 = pb_key(cx, cy, 0, 3);
      for (int pi = 0; pi < sv.nPatterns; ++pi) {
        const SamplePattern &p = sv.patterns[pi];
        if (p.n != 1) continue;
        uint32_t scr0 = pb_u32(key3, p.ctr0 + 1), scr1 = p.dims == 2 ? pb_u32(key3, p.ctr0 + 2) : 0u;
        pb_bc_table_perm(key3, p.ctr0 + (uint32_t)p.dims, perm.data());
        for (int o = 0; o < PB_BC_TABLE; ++o) {
          tables[(size_t)p.offset * PB_BC_TABLE + o] = pb_van_der_corput(perm[o], scr0);
          if (p.dims == 2) tables[(size_t)(p.offset + 1) * PB_BC_TABLE + o] = pb_sobol2(perm[o], scr1);
        }
      }
      for (int o = 0; o < PB_BC_TABLE; ++o) {
        CameraSample cs;
        if (!bc_camera_sample(sv, cxs[t], cys[t], o, key3, xs, xe, ys, ye, &cs)) continue;
        PatternSource key0;
        key0.key0 = pb_key_bc(cx, cy, o, 2);
        key0.ld = tables.empty() ? (const float *)0 : tables.data();
        key0.stride = PB_BC_TABLE; key0.index = (uint32_t)o; key0.bc = 1;
        take_sample(cs, key0, pb_key_bc(cx, cy, o, 1));
      }
    }
  } else
  for (int y = ys; y < ye; ++y)
    for (int x = xs; x < xe; ++x) {
      if (ld) {
        // samplers/lowdiscrepancy.cpp:93-103: all of the pixel's tables at once
        uint64_t keyP = pb_key(x, y, 0xFFFF, 0);
        for (int t = 0; t < 3; ++t) pb_ld_generate_table(keyP, sv.ldCamera[t], spp, ldTables.data(), spp, 0);
        for (int t = 0; t < sv.nPatterns; ++t) pb_ld_generate_table(keyP, sv.patterns[t], spp, ldTables.data(), spp, 0);
      }
      for (int k = 0; k < spp; ++k) {
        PatternSource key0;
        key0.key0 = pb_key(x, y, k, 0);
        key0.ld = ld ? ldTables.data() : (const float *)0;
        key0.stride = (uint32_t)spp; key0.index = (uint32_t)k;
        key0.bc = sv.sampler.kind == PBRT_B200_SAMPLER_RANDOM ? 2 : 0;
        CameraSample cs;
        if (ld) ld_camera_sample(sv, key0, x, y, &cs);
        else stratified_camera_sample(sv, x, y, k, &cs);
        take_sample(cs, key0, pb_key(x, y, k, 1));
      }
    }
  auto t1 = std::chrono::steady_clock::now();
  for (size_t i = 0; i < film.size(); ++i) {
    film_pixel_resolve(film[i], sv.premultiply != 0, rgba + 4 * i);
    if (raw) {
      raw[5 * i] = film[i].L[0]; raw[5 * i + 1] = film[i].L[1]; raw[5 * i + 2] = film[i].L[2];
      raw[5 * i + 3] = film[i].alpha; raw[5 * i + 4] = film[i].weightSum;
    }
  }
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->camera_rays = nSamples;
    stats->intersect_rays = c.intersect; stats->shadow_rays = c.shadow; stats->points_shaded = c.shaded;
    stats->interior_visits = c.kd.interior; stats->leaf_visits = c.kd.leaves;
    stats->prim_refs_read = c.kd.refs; stats->prim_tests = c.kd.tests;
    stats->nan_samples = c.nan; stats->negative_samples = c.neg; stats->infinite_samples = c.inf;
    stats->render_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  }
  return PBRT_B200_OK;
}

int pbrt_twin_render(const PbrtB200Scene *scene, const PbrtB200Window *win, float *rgba, float *raw,
                     float *samples, PbrtB200Stats *stats) {
  return twin_render_impl(scene, win, (const PbrtB200Window *)0, rgba, raw, samples, stats);
}
int pbrt_twin_render_tile(const PbrtB200Scene *scene, const PbrtB200Window *frame, const PbrtB200Window *tile,
                          float *rgba, float *raw, PbrtB200Stats *stats) {
  return twin_render_impl(scene, tile, frame, rgba, raw, (float *)0, stats);
}

uint64_t pbrt_twin_sample_count(const PbrtB200Scene *scene, const PbrtB200Window *win) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return 0;
  int xs, xe, ys, ye;
  film_sample_extent(sv, win->x0, win->y0, win->width, win->height, &xs, &xe, &ys, &ye);
  if (sv.sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE) {
    // table entries that fall inside the sampler's extent
    std::vector<float> cxs, cys;
    pb_bc_placements(xs, xe, ys, ye, sv.bcTableWidth, &cxs, &cys);
    uint64_t n = 0;
    for (size_t t = 0; t < cxs.size(); ++t)
      for (int o = 0; o < PB_BC_TABLE; ++o) {
        CameraSample cs;
        if (bc_camera_sample(sv, cxs[t], cys[t], o, 0, xs, xe, ys, ye, &cs)) ++n;
      }
    return n;
  }
  return (uint64_t)(xe - xs) * (ye - ys) * sv.sampler.x_samples * sv.sampler.y_samples;
}

// KdTreeAccel::Intersect / IntersectP for explicit rays (8 floats each).
int pbrt_twin_intersect(const PbrtB200Scene *scene, const float *rays, uint64_t n, int32_t *prim,
                        float *t, float *b1, float *b2) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    Ray r;
    r.o = v3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]);
    r.d = v3(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5]);
    r.mint = rays[8 * i + 6]; r.maxt = rays[8 * i + 7];
    Hit h;
    kd_intersect<false, false>(sv, r, &h, (KdCounters *)0);
    prim[i] = h.prim;
    t[i] = h.prim >= 0 ? h.t : 0.f; b1[i] = h.prim >= 0 ? h.b1 : 0.f; b2[i] = h.prim >= 0 ? h.b2 : 0.f;
  }
  return PBRT_B200_OK;
}
// The same walk over the device-side node order (csrc/pb_treelet.h): checks, on the CPU, that re-ordering the
// nodes changes no hit and no counter.  counters[4] = interior visits, leaf visits, references read, tests (summed).
int pbrt_twin_intersect_treelets(const PbrtB200Scene *scene, const float *rays, uint64_t n, int32_t *prim,
                                 float *t, float *b1, float *b2, uint64_t *counters, int paired) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  std::vector<PbrtB200KdNode> ordered;
  if (paired) {
    if (scene->n_instances != 0 || scene->n_nodes < 2) return PBRT_B200_ERR_UNSUPPORTED;
    pb_treelet_order(scene->nodes, scene->n_nodes, &ordered);
    sv.nodes = ordered.data();
    sv.pairedNodes = 1;
  }
  KdCounters k = {0, 0, 0, 0};
  for (int c = 0; c < 4; ++c) counters[c] = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Ray r;
    r.o = v3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]);
    r.d = v3(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5]);
    r.mint = rays[8 * i + 6]; r.maxt = rays[8 * i + 7];
    Hit h;
    k.interior = k.leaves = k.refs = k.tests = 0;
    kd_intersect<false, true>(sv, r, &h, &k);
    counters[0] += k.interior; counters[1] += k.leaves; counters[2] += k.refs; counters[3] += k.tests;
    prim[i] = h.prim;
    t[i] = h.prim >= 0 ? h.t : 0.f; b1[i] = h.prim >= 0 ? h.b1 : 0.f; b2[i] = h.prim >= 0 ? h.b2 : 0.f;
  }
  return PBRT_B200_OK;
}
int pbrt_twin_intersect_p(const PbrtB200Scene *scene, const float *rays, uint64_t n, uint8_t *occ) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  for (uint64_t i = 0; i < n; ++i) {
    Ray r;
    r.o = v3(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2]);
    r.d = v3(rays[8 * i + 3], rays[8 * i + 4], rays[8 * i + 5]);
    r.mint = rays[8 * i + 6]; r.maxt = rays[8 * i + 7];
    Hit h;
    occ[i] = kd_intersect<true, false>(sv, r, &h, (KdCounters *)0) ? 1 : 0;
  }
  return PBRT_B200_OK;
}
// Camera rays of a window in reference order: 8 floats per ray, 2 per image_xy.
int pbrt_twin_camera_rays(const PbrtB200Scene *scene, const PbrtB200Window *win, uint64_t n,
                          float *rays, float *imageXY) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  int xs, xe, ys, ye;
  film_sample_extent(sv, win->x0, win->y0, win->width, win->height, &xs, &xe, &ys, &ye);
  int spp = sv.sampler.x_samples * sv.sampler.y_samples;
  uint64_t i = 0;
  for (int y = ys; y < ye && i < n; ++y)
    for (int x = xs; x < xe && i < n; ++x)
      for (int k = 0; k < spp && i < n; ++k, ++i) {
        CameraSample cs;
        if (sv.sampler.kind == PBRT_B200_SAMPLER_LOWDISCREPANCY) {
          std::vector<float> tab((size_t)sv.ldFloatsPerSample * spp);
          uint64_t keyP = pb_key(x, y, 0xFFFF, 0);
          for (int t = 0; t < 2; ++t) pb_ld_generate_table(keyP, sv.ldCamera[t], spp, tab.data(), spp, 0);
          PatternSource ps;
          ps.key0 = 0; ps.ld = tab.data(); ps.stride = (uint32_t)spp; ps.index = (uint32_t)k;
          ld_camera_sample(sv, ps, x, y, &cs);
        } else
          stratified_camera_sample(sv, x, y, k, &cs);
        Ray r;
        perspective_generate_ray(sv.camera, cs, &r, sv.xRes, sv.yRes);
        float *o = rays + 8 * i;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
        o[6] = r.mint; o[7] = r.maxt;
        imageXY[2 * i] = cs.imageX; imageXY[2 * i + 1] = cs.imageY;
      }
  return PBRT_B200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Developer aid (tools/simt_sim.py): per-ray trace of the kd walk as a string of step codes,
// I interior step, L leaf entry, R primitive reference skipped by the mailbox, T primitive test,
// P pop.  Lets warp-scheduling strategies be compared offline.
extern "C" int pbrt_twin_trace(const PbrtB200Scene *scene, const float *rays, uint64_t n, int any,
                               uint8_t *codes, uint64_t cap, uint64_t *offsets) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  uint64_t pos = 0;
  for (uint64_t r = 0; r < n; ++r) {
    offsets[r] = pos;
    Ray ray;
    ray.o = v3(rays[8 * r], rays[8 * r + 1], rays[8 * r + 2]);
    ray.d = v3(rays[8 * r + 3], rays[8 * r + 4], rays[8 * r + 5]);
    ray.mint = rays[8 * r + 6]; ray.maxt = rays[8 * r + 7];
    float tmin, tmax;
    if (sv.nNodes == 0 || !bbox_intersect_p(sv.bmin, sv.bmax, ray, &tmin, &tmax)) continue;
    V3 invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    uint32_t todoNode[PB_KD_STACK]; float todoTmin[PB_KD_STACK], todoTmax[PB_KD_STACK];
    int todoPos = 0;
    uint32_t mailbox[PB_MAILBOX];
    for (int i = 0; i < PB_MAILBOX; ++i) mailbox[i] = 0xffffffffu;
    int mbPos = 0;
    uint32_t node = 0;
    Hit hit; hit.prim = -1; hit.inst = -1;
    bool stop = false;
    auto put = [&](uint8_t c) { if (pos < cap) codes[pos++] = c; };
    while (!stop) {
      uint32_t w0, w1;
      for (;;) {
        w0 = sv.nodes[node].word0; w1 = sv.nodes[node].word1;
        if ((w0 & 3u) == 3u) break;
        put('I');
        int axis = (int)(w0 & 3u);
        float split; { uint32_t b = w0; memcpy(&split, &b, 4); }
        float o = ray.o[axis], d = ray.d[axis];
        float tplane = (split - o) * invDir[axis];
        bool belowFirst = (o < split) || (o == split && d >= 0);
        uint32_t firstChild = belowFirst ? node + 1 : w1, secondChild = belowFirst ? w1 : node + 1;
        if (tplane > tmax || tplane <= 0) node = firstChild;
        else if (tplane < tmin) node = secondChild;
        else { todoNode[todoPos] = secondChild; todoTmin[todoPos] = tplane; todoTmax[todoPos] = tmax; ++todoPos;
               node = firstChild; tmax = tplane; }
      }
      put('L');
      uint32_t nPrims = w0 >> 2;
      for (uint32_t i = 0; i < nPrims && !stop; ++i) {
        uint32_t prim = (nPrims == 1) ? w1 : sv.primRefs[w1 + i];
        bool seen = false;
        for (int k = 0; k < PB_MAILBOX; ++k) seen |= (mailbox[k] == prim);
        if (seen) { put('R'); continue; }
        mailbox[mbPos] = prim; mbPos = (mbPos + 1) & (PB_MAILBOX - 1);
        put('T');
        bool h = any ? kd_test_prim<true>(sv, prim, ray, &hit) : kd_test_prim<false>(sv, prim, ray, &hit);
        if (h && any) stop = true;
      }
      if (stop) break;
      put('P');
      if (todoPos == 0) break;
      --todoPos; node = todoNode[todoPos]; tmin = todoTmin[todoPos]; tmax = todoTmax[todoPos];
      if (!any && ray.maxt < tmin) break;
    }
  }
  offsets[n] = pos;
  return PBRT_B200_OK;
}

// Mailbox study (test infrastructure): primitive tests per ray of the closest-hit / any-hit walk under different
// "already tested" filters -- none, the shipped 2-entry ring, direct-mapped tables of 4..64 entries indexed by a
// multiplicative hash of the primitive number, and the reference's exact per-primitive mailboxes (accelerators/kdtree.cpp:373-377).
// Every filter only skips repeats, so hits are identical; out[k] = total tests with filter k over all rays.
extern "C" int pbrt_twin_mailbox_study(const PbrtB200Scene *scene, const float *rays, uint64_t n, int any,
                                       uint64_t *out /* [8]: none, ring2, dm4, dm8, dm16, dm32, dm64, exact */) {
  SceneView sv;
  if (!build_view(*scene, &sv)) return PBRT_B200_ERR_INVALID;
  for (int k = 0; k < 8; ++k) out[k] = 0;
  std::vector<uint32_t> stamp(sv.nPrims, 0xffffffffu);
  for (uint64_t r = 0; r < n; ++r) {
    Ray ray;
    ray.o = v3(rays[8 * r], rays[8 * r + 1], rays[8 * r + 2]);
    ray.d = v3(rays[8 * r + 3], rays[8 * r + 4], rays[8 * r + 5]);
    ray.mint = rays[8 * r + 6]; ray.maxt = rays[8 * r + 7];
    float tmin, tmax;
    if (sv.nNodes == 0 || !bbox_intersect_p(sv.bmin, sv.bmax, ray, &tmin, &tmax)) continue;
    V3 invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    uint32_t todoNode[PB_KD_STACK]; float todoTmin[PB_KD_STACK], todoTmax[PB_KD_STACK];
    int todoPos = 0;
    uint32_t ring[2] = {0xffffffffu, 0xffffffffu};
    int ringPos = 0;
    uint32_t dm[5][64];
    for (int k = 0; k < 5; ++k) for (int i = 0; i < 64; ++i) dm[k][i] = 0xffffffffu;
    uint32_t node = 0;
    Hit hit; hit.prim = -1; hit.inst = -1;
    bool stop = false;
    while (!stop) {
      uint32_t w0, w1;
      for (;;) {
        w0 = sv.nodes[node].word0; w1 = sv.nodes[node].word1;
        if ((w0 & 3u) == 3u) break;
        int axis = (int)(w0 & 3u);
        float split; { uint32_t b = w0; memcpy(&split, &b, 4); }
        float o = ray.o[axis], d = ray.d[axis];
        float tplane = (split - o) * invDir[axis];
        bool belowFirst = (o < split) || (o == split && d >= 0);
        uint32_t firstChild = belowFirst ? node + 1 : w1, secondChild = belowFirst ? w1 : node + 1;
        if (tplane > tmax || tplane <= 0) node = firstChild;
        else if (tplane < tmin) node = secondChild;
        else { todoNode[todoPos] = secondChild; todoTmin[todoPos] = tplane; todoTmax[todoPos] = tmax; ++todoPos;
               node = firstChild; tmax = tplane; }
      }
      uint32_t nPrims = w0 >> 2;
      for (uint32_t i = 0; i < nPrims && !stop; ++i) {
        uint32_t prim = (nPrims == 1) ? w1 : sv.primRefs[w1 + i];
        ++out[0];
        if (ring[0] != prim && ring[1] != prim) { ++out[1]; ring[ringPos] = prim; ringPos ^= 1; }
        for (int k = 0; k < 5; ++k) {
          uint32_t slot = (prim * 0x9E3779B1u) >> (30 - k);   // multiplicative hash: mesh neighbours differ by a row stride
          if (dm[k][slot] != prim) { ++out[2 + k]; dm[k][slot] = prim; }
        }
        bool fresh = stamp[prim] != (uint32_t)r;
        if (fresh) { ++out[7]; stamp[prim] = (uint32_t)r; }
        // the walk itself follows the exact filter (the reference's)
        if (!fresh) continue;
        bool h = any ? kd_test_prim<true>(sv, prim, ray, &hit) : kd_test_prim<false>(sv, prim, ray, &hit);
        if (h && any) stop = true;
      }
      if (stop) break;
      if (todoPos == 0) break;
      --todoPos; node = todoNode[todoPos]; tmin = todoTmin[todoPos]; tmax = todoTmax[todoPos];
      if (!any && ray.maxt < tmin) break;
    }
  }
  return PBRT_B200_OK;
}

// The 8-byte todo entries of the GPU walk (csrc/pb_kd_warp.cuh) rest on an invariant of KdTreeAccel::Intersect's loop
// (accelerators/kdtree.cpp:313-403): whenever an entry is popped, its stored tmin equals the tmax that is current at
// that moment, bit for bit.  This walks the reference's loop with full 12-byte entries and counts the pops where that
// is not so (returns the count; pops[0] = number of pops examined).
extern "C" uint64_t pbrt_twin_stack_invariant(const PbrtB200Scene *scene, const float *rays, uint64_t n, uint64_t *pops) {
  SceneView sv;
  *pops = 0;
  if (!build_view(*scene, &sv) || sv.nInstances != 0) return ~0ull;
  uint64_t bad = 0;
  for (uint64_t r = 0; r < n; ++r) {
    Ray ray;
    ray.o = v3(rays[8 * r], rays[8 * r + 1], rays[8 * r + 2]);
    ray.d = v3(rays[8 * r + 3], rays[8 * r + 4], rays[8 * r + 5]);
    ray.mint = rays[8 * r + 6]; ray.maxt = rays[8 * r + 7];
    float tmin, tmax;
    if (sv.nNodes == 0 || !bbox_intersect_p(sv.bmin, sv.bmax, ray, &tmin, &tmax)) continue;
    V3 invDir = v3(1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z);
    uint32_t todoNode[PB_KD_STACK]; float todoTmin[PB_KD_STACK], todoTmax[PB_KD_STACK];
    int todoPos = 0;
    uint32_t node = 0;
    Hit hit; hit.prim = -1; hit.inst = -1;
    for (;;) {
      uint32_t w0, w1;
      for (;;) {
        w0 = sv.nodes[node].word0; w1 = sv.nodes[node].word1;
        if ((w0 & 3u) == 3u) break;
        int axis = (int)(w0 & 3u);
        float split; { uint32_t b = w0; memcpy(&split, &b, 4); }
        float o = ray.o[axis], d = ray.d[axis];
        float tplane = (split - o) * invDir[axis];
        bool belowFirst = (o < split) || (o == split && d >= 0);
        uint32_t firstChild = belowFirst ? node + 1 : w1, secondChild = belowFirst ? w1 : node + 1;
        if (tplane > tmax || tplane <= 0) node = firstChild;
        else if (tplane < tmin) node = secondChild;
        else { todoNode[todoPos] = secondChild; todoTmin[todoPos] = tplane; todoTmax[todoPos] = tmax; ++todoPos;
               node = firstChild; tmax = tplane; }
      }
      uint32_t nPrims = w0 >> 2;
      for (uint32_t i = 0; i < nPrims; ++i) {
        uint32_t prim = (nPrims == 1) ? w1 : sv.primRefs[w1 + i];
        kd_test_prim<false>(sv, prim, ray, &hit);
      }
      if (todoPos == 0) break;
      --todoPos;
      ++*pops;
      uint32_t a, b;
      memcpy(&a, &todoTmin[todoPos], 4); memcpy(&b, &tmax, 4);
      if (a != b) ++bad;
      node = todoNode[todoPos]; tmin = todoTmin[todoPos]; tmax = todoTmax[todoPos];
      if (ray.maxt < tmin) break;
    }
  }
  return bad;
}
