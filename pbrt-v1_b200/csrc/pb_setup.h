// pb_setup.h -- host-side derivation of the per-render constants in SceneView
// from a flattened scene: validation, film constants and the layout of the
// integrators' sample patterns inside each sample's LatinHypercube block
// (Sample::Sample, core/sampling.cpp:41-70; request order path.cpp:47-57,
// directlighting.cpp:39-66, emission.cpp:42-46; generation order
// samplers/stratified.cpp:125-128: all 1D patterns, then all 2D patterns).
#ifndef PB_SETUP_H
#define PB_SETUP_H
#include "pb_scene.h"
#include <string.h>
#include <string>

// Fills everything except the array pointers.  Returns an empty string or an error.
static inline std::string pb_setup_view(const PbrtB200Scene &s, SceneView *sv) {
  if (s.abi_version != PBRT_B200_ABI_VERSION) return "scene abi_version mismatch";
  if (s.sampler.x_samples < 1 || s.sampler.y_samples < 1) return "sampler needs >= 1 sample per axis";
  const bool ld = s.sampler.kind == PBRT_B200_SAMPLER_LOWDISCREPANCY;
  const bool bc = s.sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE;
  const bool rnd = s.sampler.kind == PBRT_B200_SAMPLER_RANDOM;
  if (!ld && !bc && !rnd && s.sampler.kind != PBRT_B200_SAMPLER_STRATIFIED) return "unknown sampler kind";
  if (bc && (s.sampler.y_samples != 1 || s.sampler.x_samples > PB_BC_TABLE || !s.bestcandidate_table))
    return "bestcandidate sampler: pixelsamples must be in [1, 4096] and the sample table present";
  // tableWidth = SQRT_SAMPLE_TABLE_SIZE / sqrtf(pixelSamples), samplers/bestcandidate.cpp:68-69
  sv->bcTableWidth = bc ? (float)64 / sqrtf((float)s.sampler.x_samples) : 0.f;
  if (ld && (s.sampler.y_samples != 1 || (s.sampler.x_samples & (s.sampler.x_samples - 1)) != 0 ||
             s.sampler.x_samples > PB_LD_MAX_PIXEL_SAMPLES))
    return "lowdiscrepancy sampler: pixelsamples must be a power of two <= 1024";
  // The keyed RNG packs the sample index into 10 bits next to the pixel's y (pb_rng.h: pb_key); the camera
  // pattern's key uses index 0xFFFF, whose low field equals sample 1023's.  Up to 1023 samples per pixel no two
  // keys coincide; beyond that different pixels would share LatinHypercube blocks.
  if (!bc && !ld && (int64_t)s.sampler.x_samples * s.sampler.y_samples > 1023)
    return "stratified / random sampler: more than 1023 samples per pixel";
  if (s.film.x_pixel_count < 1 || s.film.y_pixel_count < 1) return "empty film";
  // (ImageFilm keeps the path depth in 8 bits of the per-sample flag word)
  if (s.integrator.kind == PBRT_B200_INT_PATH && s.integrator.max_depth > 254) return "path integrator: maxdepth > 254";
  if ((s.n_nodes && !s.nodes) || (s.n_prim_refs && !s.prim_refs) || (s.n_prims && (!s.prims || !s.prim_aux)) ||
      (s.n_materials && !s.materials) || ((s.n_lights + s.n_unsampled_lights) && !s.lights))
    return "scene array pointer is null";
  for (uint32_t i = 0; i < s.n_prims; ++i) {
    uint32_t kind = s.prims[i].flags & PBRT_B200_PF_KIND_MASK;
    if (kind != PBRT_B200_PRIM_TRIANGLE && kind != PBRT_B200_PRIM_QUADRIC && kind != PBRT_B200_PRIM_INSTANCE) return "unknown primitive kind";
    if (kind == PBRT_B200_PRIM_INSTANCE) {
      if (s.prim_aux[i].owner >= s.n_instances) return "instance index out of range";
      continue;
    }
    if (kind == PBRT_B200_PRIM_QUADRIC && (!s.quadrics || s.prim_aux[i].owner >= s.n_quadrics)) return "quadric index out of range";
    if (kind == PBRT_B200_PRIM_TRIANGLE && (s.prims[i].flags & (PBRT_B200_PF_HAS_UV | PBRT_B200_PF_HAS_N | PBRT_B200_PF_HAS_S))) {
      const PbrtB200PrimAux &a = s.prim_aux[i];
      if (a.v[0] >= s.n_vertices || a.v[1] >= s.n_vertices || a.v[2] >= s.n_vertices) return "triangle vertex index out of range";
      if ((s.prims[i].flags & (PBRT_B200_PF_HAS_N | PBRT_B200_PF_HAS_S)) && (!s.meshes || a.owner >= s.n_meshes)) return "triangle mesh index out of range";
      if (((s.prims[i].flags & PBRT_B200_PF_HAS_UV) && !s.vertex_uv) || ((s.prims[i].flags & PBRT_B200_PF_HAS_N) && !s.vertex_n) ||
          ((s.prims[i].flags & PBRT_B200_PF_HAS_S) && !s.vertex_s))
        return "per-vertex array missing";
    }
    if (s.prims[i].material >= s.n_materials) return "primitive material index out of range";
    if (s.prims[i].light >= (int32_t)(s.n_lights + s.n_unsampled_lights)) return "primitive light index out of range";
  }
  // texture operands: indices in range, operands before their users, image maps with complete pyramids
  sv->nTextures = s.n_textures;
  for (uint32_t i = 0; i < s.n_materials; ++i)
    for (int k = 0; k < PBRT_B200_TEXSLOTS; ++k)
      if (s.materials[i].tex[k] >= (int32_t)s.n_textures) return "material texture index out of range";
  std::string texDepth(s.n_textures, '\0');   // operators above the deepest leaf, per node (operands come first)
  for (uint32_t i = 0; i < s.n_textures; ++i) {
    const PbrtB200Texture &t = s.textures[i];
    if (t.kind > PBRT_B200_TEX_MARBLE || (t.channels != 1 && t.channels != 3)) return "unknown texture kind";
    if ((t.kind == PBRT_B200_TEX_UV || t.kind == PBRT_B200_TEX_BILERP || t.kind == PBRT_B200_TEX_CHECKERBOARD ||
         t.kind == PBRT_B200_TEX_DOTS) && t.mapping > PBRT_B200_TEXMAP_PLANAR) return "unknown texture mapping or wrap mode";
    if (t.kind == PBRT_B200_TEX_CHECKERBOARD && t.iparam[0] != PBRT_B200_CHECKER_NONE && t.iparam[0] != PBRT_B200_CHECKER_CLOSEDFORM)
      return "checkerboard: unsupported anti-aliasing mode";
    if (t.kind == PBRT_B200_TEX_SCALE || t.kind == PBRT_B200_TEX_MIX || t.kind == PBRT_B200_TEX_CHECKERBOARD || t.kind == PBRT_B200_TEX_DOTS)
      for (int k = 0; k < (t.kind == PBRT_B200_TEX_MIX ? 3 : 2); ++k)
      {
        if (t.arg[k] < 0 || t.arg[k] >= (int32_t)i) return "texture operand must precede its user";
        if (texDepth[t.arg[k]] + 1 > texDepth[i]) texDepth[i] = (char)(texDepth[t.arg[k]] + 1);
        if (texDepth[i] > 3) return "texture expression more than 3 operators deep";   // PB_TEX_MAX_DEPTH
      }
    if (t.kind == PBRT_B200_TEX_IMAGEMAP) {
      if (t.n_levels < 1 || (uint64_t)t.first_level + t.n_levels > s.n_tex_levels || !s.ewa_weight_lut) return "image map without a pyramid";
      if (t.mapping > PBRT_B200_TEXMAP_PLANAR || t.wrap > PBRT_B200_TEXWRAP_CLAMP) return "unknown texture mapping or wrap mode";
      for (uint32_t l = 0; l < t.n_levels; ++l) {
        const PbrtB200TexLevel &lv = s.tex_levels[t.first_level + l];
        if (lv.width < 1 || lv.height < 1 || lv.offset + (uint64_t)lv.width * lv.height * t.channels > s.n_texel_floats)
          return "texture level outside the texel array";
      }
    }
  }
  if (s.n_top_nodes > s.n_nodes) return "n_top_nodes > n_nodes";
  // kd-trees: every interior node's above child and every leaf's primitives must stay inside the tree they
  // belong to (the top-level tree, or an instanced object's, whose indices are relative to its offsets)
  {
    auto check_tree = [&](uint32_t nodeBase, uint32_t nNodes, uint32_t refBase, uint32_t refLimit, uint32_t primBase,
                          uint32_t primLimit) -> const char * {
      for (uint32_t i = 0; i < nNodes; ++i) {
        const PbrtB200KdNode &n = s.nodes[nodeBase + i];
        if ((n.word0 & 3u) != 3u) {
          if (n.word1 <= i || n.word1 >= nNodes || i + 1 >= nNodes) return "kd node child index out of range";
        } else {
          uint32_t np = n.word0 >> 2;
          if (np == 1) { if ((uint64_t)primBase + n.word1 >= primLimit) return "kd leaf primitive out of range"; }
          else if (np > 1) {
            if ((uint64_t)refBase + n.word1 + np > refLimit) return "kd leaf reference list out of range";
            for (uint32_t k = 0; k < np; ++k)
              if ((uint64_t)primBase + s.prim_refs[refBase + n.word1 + k] >= primLimit) return "kd leaf primitive out of range";
          }
        }
      }
      return (const char *)0;
    };
    if (const char *e = check_tree(0, s.n_top_nodes, 0, s.n_prim_refs, 0, s.n_prims)) return e;
    for (uint32_t i = 0; i < s.n_instances; ++i) {
      const PbrtB200Instance &in = s.instances[i];
      if ((uint64_t)in.node_base + in.n_nodes > s.n_nodes || in.ref_base > s.n_prim_refs || in.prim_base >= s.n_prims)
        return "instance offsets out of range";
      if (const char *e = check_tree(in.node_base, in.n_nodes, in.ref_base, s.n_prim_refs, in.prim_base, s.n_prims)) return e;
    }
  }
  for (uint32_t i = 0; i < s.n_instances; ++i) {
    const PbrtB200Instance &in = s.instances[i];
    if ((uint64_t)in.node_base + in.n_nodes > s.n_nodes || in.ref_base > s.n_prim_refs || in.prim_base >= s.n_prims)
      return "instance offsets out of range";
  }
  sv->nNodes = s.n_top_nodes; sv->nPrims = s.n_prims; sv->nLights = s.n_lights;
  sv->nInstances = s.n_instances;
  sv->nInfiniteLights = 0;
  for (uint32_t i = 0; i < s.n_lights; ++i) {
    if (s.lights[i].kind < PBRT_B200_LIGHT_POINT || s.lights[i].kind > PBRT_B200_LIGHT_PROJECTION) return "unknown light kind";
    if (s.lights[i].kind == PBRT_B200_LIGHT_INFINITE) ++sv->nInfiniteLights;
    if (s.lights[i].kind == PBRT_B200_LIGHT_AREA) {
      // AreaLight over a ShapeSet (lights/area.cpp, core/shape.h:108-156): its shape list and area CDF
      const PbrtB200Light &l = s.lights[i];
      if (l.n_shapes < 1 || !s.light_shapes || !s.light_cdf || (uint64_t)l.first_shape + l.n_shapes > s.n_light_shapes)
        return "area light shape list out of range";
      for (uint32_t k = 0; k < l.n_shapes; ++k) {
        uint32_t prim = s.light_shapes[l.first_shape + k];
        if (prim >= s.n_prims) return "area light shape is not a primitive of the scene";
        uint32_t pk = s.prims[prim].flags & PBRT_B200_PF_KIND_MASK;
        if (pk == PBRT_B200_PRIM_INSTANCE) return "area light shape is an object instance";
        if (pk == PBRT_B200_PRIM_QUADRIC && s.prim_aux[prim].owner >= s.n_quadrics) return "area light quadric index out of range";
      }
    }
    if (s.lights[i].kind >= PBRT_B200_LIGHT_INFINITE && s.lights[i].radiance_map != 0 &&
        (s.lights[i].radiance_map > s.n_textures || s.textures[s.lights[i].radiance_map - 1].kind != PBRT_B200_TEX_IMAGEMAP ||
         s.textures[s.lights[i].radiance_map - 1].channels != 3))
      return "light map is not an image-map texture node";
    if (s.lights[i].dist_nu != 0) {
      const PbrtB200Light &l = s.lights[i];
      uint64_t need = 3ull * l.dist_nu + 2 + (uint64_t)l.dist_nu * (2ull * l.dist_nv + 1);
      if (l.kind != PBRT_B200_LIGHT_INFINITE || l.radiance_map == 0 || l.dist_nv == 0 ||
          l.dist_offset + need > s.n_light_distribution_floats)
        return "importance-sampled infinite light without complete distribution tables";
    }
  }
  for (int i = 0; i < 3; ++i) { sv->bmin[i] = s.bounds[i]; sv->bmax[i] = s.bounds[3 + i]; }
  sv->camera = s.camera; sv->sampler = s.sampler; sv->integrator = s.integrator;
  sv->xRes = s.film.x_resolution; sv->yRes = s.film.y_resolution;
  sv->filterXW = s.film.filter_x_width; sv->filterYW = s.film.filter_y_width;
  sv->filterInvXW = s.film.filter_inv_x_width; sv->filterInvYW = s.film.filter_inv_y_width;
  sv->premultiply = s.film.premultiply_alpha;
  int n1[PB_MAX_PATTERNS], n2[PB_MAX_PATTERNS];
  int c1 = 0, c2 = 0;
  int kind = s.integrator.kind;
  if (kind == PBRT_B200_INT_PATH) {
    for (int i = 0; i < 9; ++i) { n1[c1++] = 1; n2[c2++] = 1; }
    sv->directRecords = 1;
  } else if (kind == PBRT_B200_INT_DIRECT) {
    if (s.integrator.strategy == PBRT_B200_STRATEGY_ALL) {
      if (s.n_lights * 3 + 2 > PB_MAX_PATTERNS) return "too many lights for directlighting strategy all";
      int recs = 0;
      for (uint32_t i = 0; i < s.n_lights; ++i) {
        int ns = s.lights[i].n_samples < 1 ? 1 : s.lights[i].n_samples;
        if (ld) {   // Sampler::RoundSize: RoundUpPow2 (samplers/lowdiscrepancy.cpp:44-46)
          int r = 1;
          while (r < ns) r <<= 1;
          ns = r;
          if (ns > PB_LD_MAX_PATTERN_SAMPLES) return "lowdiscrepancy sampler: too many light samples";
        }
        if (bc) {   // BestCandidateSampler::RoundSize (samplers/bestcandidate.cpp:45-48)
          int root = pb_ceil2int(sqrtf((float)ns - .5f));
          ns = root * root;
        }
        n2[c2++] = ns; n2[c2++] = ns; n1[c1++] = ns;
        recs += ns;
      }
      sv->directRecords = recs;
    } else {
      n2[c2++] = 1; n1[c1++] = 1; n2[c2++] = 1; n1[c1++] = 1;
      sv->directRecords = 1;
    }
  } else if (kind == PBRT_B200_INT_WHITTED) {
    sv->directRecords = (int)s.n_lights;
  } else
    return "unknown surface integrator kind";
  if (sv->directRecords < 1) sv->directRecords = 1;
  for (int i = 0; i < s.integrator.n_volume_1d; ++i) {
    if (c1 >= PB_MAX_PATTERNS) return "too many sample patterns";
    n1[c1++] = 1;
  }
  if (c1 + c2 > PB_MAX_PATTERNS) return "too many sample patterns";
  uint32_t ctr = 0;
  int k = 0;
  int32_t off = 0;
  const uint32_t nPixel = (uint32_t)s.sampler.x_samples;
  if (ld) {
    // per pixel: image, lens, time tables first (samplers/lowdiscrepancy.cpp:96-98); a table of
    // pattern size n draws `dims` scrambles, n per block Shuffle and one per block for the final Shuffle
    for (int t = 0; t < 3; ++t) {
      SamplePattern &p = sv->ldCamera[t];
      p.ctr0 = ctr; p.n = 1; p.dims = t < 2 ? 2 : 1;
      p.offset = t < 2 ? off : -1;
      if (t < 2) off += 2;
      ctr += (uint32_t)p.dims + nPixel * 1u + nPixel;
    }
  }
  if (bc) {
    // placement stream (3): three sample shifts, then one table per one-sample pattern (1 or 2 scramble
    // draws, 4096 Shuffle draws, 1 draw of the final one-element Shuffle); sample stream (2): the
    // stratified draws of the larger patterns, 1D patterns first (samplers/bestcandidate.cpp:96-105,152-168)
    uint32_t tileCtr = 3, sampleCtr = 0;
    for (int pass = 0; pass < 2; ++pass) {
      int cnt = pass == 0 ? c1 : c2;
      const int *nn = pass == 0 ? n1 : n2;
      int dims = pass + 1;
      for (int i = 0; i < cnt; ++i, ++k) {
        SamplePattern &p = sv->patterns[k];
        p.n = nn[i]; p.dims = dims;
        if (nn[i] == 1) {
          p.ctr0 = tileCtr; p.offset = off;
          tileCtr += (uint32_t)dims + PB_BC_TABLE + 1u;
          off += dims;
        } else {
          p.ctr0 = sampleCtr; p.offset = -1;
          sampleCtr += (uint32_t)(dims * nn[i]);
        }
      }
    }
    sv->n1D = c1; sv->n2D = c2; sv->nPatterns = k;
    sv->ldFloatsPerSample = off;
    return std::string();
  }
  for (int i = 0; i < c1; ++i, ++k) {
    sv->patterns[k].ctr0 = ctr; sv->patterns[k].n = n1[i]; sv->patterns[k].dims = 1; sv->patterns[k].offset = off;
    if (rnd) { ctr += (uint32_t)n1[i]; continue; }   // one RandomFloat per value (samplers/random.cpp:103-105)
    ctr += ld ? 1u + nPixel * (uint32_t)n1[i] + nPixel : 2u * n1[i];
    off += n1[i];
  }
  for (int i = 0; i < c2; ++i, ++k) {
    sv->patterns[k].ctr0 = ctr; sv->patterns[k].n = n2[i]; sv->patterns[k].dims = 2; sv->patterns[k].offset = off;
    if (rnd) { ctr += 2u * (uint32_t)n2[i]; continue; }
    ctr += ld ? 2u + nPixel * (uint32_t)n2[i] + nPixel : 4u * n2[i];
    off += 2 * n2[i];
  }
  sv->n1D = c1; sv->n2D = c2; sv->nPatterns = k;
  sv->ldFloatsPerSample = ld ? off : 0;
  return std::string();
}
// bestcandidate sampler: corners of the table placements of a sampler over [xs,xe) x [ys,ye), in the
// order BestCandidateSampler::GetNextSample visits them (samplers/bestcandidate.cpp:66-72,81-90): the
// corner advances by float additions of tableWidth; the first row starts at (xs - tableWidth) + tableWidth,
// later rows at xs exactly.
#include <vector>
static inline void pb_bc_placements(int xs, int xe, int ys, int ye, float tableWidth, std::vector<float> *cx,
                                    std::vector<float> *cy) {
  cx->clear(); cy->clear();
  float xTableCorner = float(xs) - tableWidth;
  float yTableCorner = float(ys);
  for (;;) {
    xTableCorner += tableWidth;
    if (xTableCorner >= xe) {
      xTableCorner = float(xs);
      yTableCorner += tableWidth;
      if (yTableCorner >= ye) break;
    }
    cx->push_back(xTableCorner);
    cy->push_back(yTableCorner);
  }
}
// Do any materials carry a specular lobe (Whitted/DirectLighting then recurse)?
static inline bool pb_scene_has_specular(const PbrtB200Scene &s) {
  for (uint32_t i = 0; i < s.n_materials; ++i)
    if (s.materials[i].kind != PBRT_B200_MAT_MATTE) return true;
  return false;
}
#endif
