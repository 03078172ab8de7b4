// pb_texture.h -- ray differentials at a hit, 2D texture mappings, MIPMap lookups and the
// evaluation of a material's texture operands.
//   DifferentialGeometry::ComputeDifferentials   core/shape.cpp:52-106
//   SolveLinearSystem2x2                         core/util.cpp:99-108
//   UV / Spherical / Cylindrical / PlanarMapping2D::Map   core/texture.cpp:66-149
//   MIPMap::Lookup / triangle / EWA / texel      core/mipmap.h:176-315
//   ImageTexture::Evaluate                       textures/imagemap.cpp:122-128
//   ScaleTexture / MixTexture::Evaluate          textures/scale.cpp:37-39, textures/mix.cpp:40-44
// The pyramids themselves (resampling to a power of two, box-filtered levels, the EWA weight table) are built
// once on the host (pbrt-v1_b200/host/texture.cpp) and arrive as flat arrays.
#ifndef PB_TEXTURE_H
#define PB_TEXTURE_H
#include "pb_shapes.h"

#if defined(PB_TWIN_LIBM_FLOAT) && !defined(__CUDACC__)
PB_HD float pb_logf(float x) { return logf(x); }
#else
PB_NOINLINE float pb_logf(float x) { return (float)log((double)x); }
#endif
// Log2, core/pbrt.h:575-578: logf(x) * (1.f / logf(2.f))
PB_HD float pb_log2(float x) { return pb_logf(x) * (1.f / 0.693147182464599609375f); }

// RayDifferential's auxiliary rays (core/geometry.h:216-239)
struct RayDiff {
  bool has;
  V3 rxo, rxd, ryo, ryd;
};
// the screen-space derivatives ComputeDifferentials leaves in the DifferentialGeometry
struct DgDiff {
  V3 dpdx, dpdy;
  float dudx, dvdx, dudy, dvdy;
};
PB_HD bool solve_linear_2x2(const float A[2][2], const float B[2], float x[2]) {
  float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  if ((double)fabsf(det) < 1e-5) return false;
  float invDet = 1.0f / det;
  x[0] = (A[1][1] * B[0] - A[0][1] * B[1]) * invDet;
  x[1] = (A[0][0] * B[1] - A[1][0] * B[0]) * invDet;
  return true;
}
PB_HD void compute_differentials(const DiffGeom &dg, const RayDiff *ray, DgDiff *o) {
  if (ray && ray->has) {
    const V3 &nn = dg.nn, &p = dg.p;
    float d = -dot(nn, p);
    float tx = -(dot(nn, ray->rxo) + d) / dot(nn, ray->rxd);
    V3 px = ray->rxo + tx * ray->rxd;
    float ty = -(dot(nn, ray->ryo) + d) / dot(nn, ray->ryd);
    V3 py = ray->ryo + ty * ray->ryd;
    o->dpdx = px - p;
    o->dpdy = py - p;
    float A[2][2], Bx[2], By[2], x[2];
    int a0, a1;
    if (fabsf(nn.x) > fabsf(nn.y) && fabsf(nn.x) > fabsf(nn.z)) { a0 = 1; a1 = 2; }
    else if (fabsf(nn.y) > fabsf(nn.z)) { a0 = 0; a1 = 2; }
    else { a0 = 0; a1 = 1; }
    A[0][0] = dg.dpdu[a0]; A[0][1] = dg.dpdv[a0];
    A[1][0] = dg.dpdu[a1]; A[1][1] = dg.dpdv[a1];
    Bx[0] = px[a0] - p[a0]; Bx[1] = px[a1] - p[a1];
    By[0] = py[a0] - p[a0]; By[1] = py[a1] - p[a1];
    if (solve_linear_2x2(A, Bx, x)) { o->dudx = x[0]; o->dvdx = x[1]; }
    else { o->dudx = 1.f; o->dvdx = 0.f; }
    if (solve_linear_2x2(A, By, x)) { o->dudy = x[0]; o->dvdy = x[1]; }
    else { o->dudy = 0.f; o->dvdy = 1.f; }
  } else {
    o->dudx = o->dvdx = 0.f;
    o->dudy = o->dvdy = 0.f;
    o->dpdx = o->dpdy = v3(0, 0, 0);
  }
}

// ---------------------------------------------------------------- 2D mappings
PB_HD void tex_sphere(const PbrtB200Texture &t, const V3 &P, float *s, float *tt) {
  V3 vec = normalize(xf_point(t.world_to_texture, P) - v3(0, 0, 0));
  float theta = pb_acosf(pb_clamp(vec.z, -1.f, 1.f));   // SphericalTheta, geometry.h:403-405
  float ph = pb_atan2f(vec.y, vec.x);                  // SphericalPhi, geometry.h:406-409
  float phi = (ph < 0.f) ? ph + 2.f * PB_PI : ph;   // M_PI is a float literal in the reference (core/pbrt.h:201-204)
  *s = theta * PB_INV_PI;
  *tt = phi * PB_INV_TWOPI;
}
PB_HD void tex_cylinder(const PbrtB200Texture &t, const V3 &P, float *s, float *tt) {
  V3 vec = normalize(xf_point(t.world_to_texture, P) - v3(0, 0, 0));
  *s = (PB_PI + pb_atan2f(vec.y, vec.x)) / (2.f * PB_PI);
  *tt = vec.z;
}
PB_HD void tex_wrap_dt(float *dt) {   // texture.cpp:84-85
  if (*dt > .5f) *dt = 1.f - *dt;
  else if (*dt < -.5f) *dt = -(*dt + 1);
}
PB_HD void tex_map(const PbrtB200Texture &t, const DiffGeom &dg, const DgDiff &dd, float *s, float *tt,
                   float *dsdx, float *dtdx, float *dsdy, float *dtdy) {
  if (t.mapping == PBRT_B200_TEXMAP_UV) {
    *s = t.su * dg.u + t.du;
    *tt = t.sv * dg.v + t.dv;
    *dsdx = t.su * dd.dudx; *dtdx = t.sv * dd.dvdx;
    *dsdy = t.su * dd.dudy; *dtdy = t.sv * dd.dvdy;
  } else if (t.mapping == PBRT_B200_TEXMAP_PLANAR) {
    V3 vs = v3(t.v1[0], t.v1[1], t.v1[2]), vt = v3(t.v2[0], t.v2[1], t.v2[2]);
    V3 vec = dg.p - v3(0, 0, 0);
    *s = t.du + dot(vec, vs);
    *tt = t.dv + dot(vec, vt);
    *dsdx = dot(dd.dpdx, vs); *dtdx = dot(dd.dpdx, vt);
    *dsdy = dot(dd.dpdy, vs); *dtdy = dot(dd.dpdy, vt);
  } else {
    const bool sph = t.mapping == PBRT_B200_TEXMAP_SPHERICAL;
    const float delta = sph ? .1f : .01f;
    float sx, tx, sy, ty;
    if (sph) tex_sphere(t, dg.p, s, tt); else tex_cylinder(t, dg.p, s, tt);
    V3 px = dg.p + delta * dd.dpdx;
    if (sph) tex_sphere(t, px, &sx, &tx); else tex_cylinder(t, px, &sx, &tx);
    *dsdx = (sx - *s) / delta;
    *dtdx = (tx - *tt) / delta;
    tex_wrap_dt(dtdx);
    V3 py = dg.p + delta * dd.dpdy;
    if (sph) tex_sphere(t, py, &sy, &ty); else tex_cylinder(t, py, &sy, &ty);
    *dsdy = (sy - *s) / delta;
    *dtdy = (ty - *tt) / delta;
    tex_wrap_dt(dtdy);
  }
}

// ---------------------------------------------------------------- MIPMap
PB_HD int tex_mod(int a, int b) {   // Mod, core/pbrt.h:562-568
  int n = a / b;
  a -= n * b;
  if (a < 0) a += b;
  return a;
}
PB_HD int tex_iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// one texel as a Spec; a float map (channels == 1) fills all three lanes with the same value, which
// keeps every product and sum below lane-wise identical to the scalar template instance
PB_HD Spec mip_texel(const SceneView &sv, const PbrtB200Texture &t, int level, int s, int tt) {
  const PbrtB200TexLevel &l = sv.texLevels[t.first_level + level];
  int w = (int)l.width, h = (int)l.height;
  if (t.wrap == PBRT_B200_TEXWRAP_REPEAT) { s = tex_mod(s, w); tt = tex_mod(tt, h); }
  else if (t.wrap == PBRT_B200_TEXWRAP_CLAMP) { s = tex_iclamp(s, 0, w - 1); tt = tex_iclamp(tt, 0, h - 1); }
  else if (s < 0 || s >= w || tt < 0 || tt >= h) return spec(0.f);
  const float *px = sv.texels + l.offset + ((size_t)tt * w + s) * t.channels;
  return t.channels == 3 ? spec3(px[0], px[1], px[2]) : spec(px[0]);
}
PB_HD Spec mip_triangle(const SceneView &sv, const PbrtB200Texture &t, int level, float s, float tt) {
  level = tex_iclamp(level, 0, (int)t.n_levels - 1);
  const PbrtB200TexLevel &l = sv.texLevels[t.first_level + level];
  s = s * (int)l.width - 0.5f;
  tt = tt * (int)l.height - 0.5f;
  int s0 = pb_floor2int(s), t0 = pb_floor2int(tt);
  float ds = s - s0, dt = tt - t0;
  return (1.f - ds) * (1.f - dt) * mip_texel(sv, t, level, s0, t0) +
         (1.f - ds) * dt * mip_texel(sv, t, level, s0, t0 + 1) +
         ds * (1.f - dt) * mip_texel(sv, t, level, s0 + 1, t0) +
         ds * dt * mip_texel(sv, t, level, s0 + 1, t0 + 1);
}
PB_HD Spec mip_lookup_width(const SceneView &sv, const PbrtB200Texture &t, float s, float tt, float width) {
  int nLevels = (int)t.n_levels;
  float level = nLevels - 1 + pb_log2(pb_max(width, 1e-8f));
  if (level < 0) return mip_triangle(sv, t, 0, s, tt);
  else if (level >= nLevels - 1) return mip_texel(sv, t, nLevels - 1, 0, 0);
  int iLevel = pb_floor2int(level);
  float delta = level - iLevel;
  return (1.f - delta) * mip_triangle(sv, t, iLevel, s, tt) + delta * mip_triangle(sv, t, iLevel + 1, s, tt);
}
PB_NOINLINE Spec mip_ewa(const SceneView &sv, const PbrtB200Texture &t, float s, float tt, float ds0, float dt0,
                         float ds1, float dt1, int level) {
  int nLevels = (int)t.n_levels;
  if (level >= nLevels) return mip_texel(sv, t, nLevels - 1, 0, 0);
  const PbrtB200TexLevel &l = sv.texLevels[t.first_level + level];
  int w = (int)l.width, h = (int)l.height;
  s = s * w - 0.5f;
  tt = tt * h - 0.5f;
  ds0 *= w; dt0 *= h;
  ds1 *= w; dt1 *= h;
  float A = dt0 * dt0 + dt1 * dt1 + 1;
  float B = -2.f * (ds0 * dt0 + ds1 * dt1);
  float C = ds0 * ds0 + ds1 * ds1 + 1;
  float invF = 1.f / (A * C - B * B * 0.25f);
  A *= invF; B *= invF; C *= invF;
  float det = -B * B + 4.f * A * C;
  float invDet = 1.f / det;
  float uSqrt = sqrtf(det * C), vSqrt = sqrtf(A * det);
  int s0 = pb_ceil2int(s - 2.f * invDet * uSqrt);
  int s1 = pb_floor2int(s + 2.f * invDet * uSqrt);
  int t0 = pb_ceil2int(tt - 2.f * invDet * vSqrt);
  int t1 = pb_floor2int(tt + 2.f * invDet * vSqrt);
  Spec num = spec(0.f);
  float den = 0;
  for (int it = t0; it <= t1; ++it) {
    float t_ = it - tt;
    for (int is = s0; is <= s1; ++is) {
      float ss = is - s;
      float r2 = A * ss * ss + B * ss * t_ + C * t_ * t_;
      if (r2 < 1.f) {
        float weight = sv.ewaLut[pb_imin((int)(r2 * PBRT_B200_EWA_LUT_SIZE), PBRT_B200_EWA_LUT_SIZE - 1)];
        num = num + mip_texel(sv, t, level, is, it) * weight;
        den += weight;
      }
    }
  }
  // `num / den`: Spectrum::operator/(float) multiplies by the reciprocal (color.h:106-109), float / float divides
  return t.channels == 3 ? num / den : spec(num.c[0] / den);
}
PB_HD Spec mip_lookup(const SceneView &sv, const PbrtB200Texture &t, float s, float tt, float ds0, float dt0,
                      float ds1, float dt1) {
  if (t.trilinear)
    return mip_lookup_width(sv, t, s, tt,
                            2.f * pb_max(pb_max(fabsf(ds0), fabsf(dt0)), pb_max(fabsf(ds1), fabsf(dt1))));
  if (ds0 * ds0 + dt0 * dt0 < ds1 * ds1 + dt1 * dt1) {
    float tmp = ds0; ds0 = ds1; ds1 = tmp;
    tmp = dt0; dt0 = dt1; dt1 = tmp;
  }
  float majorLength = sqrtf(ds0 * ds0 + dt0 * dt0);
  float minorLength = sqrtf(ds1 * ds1 + dt1 * dt1);
  if (minorLength * t.max_anisotropy < majorLength && minorLength > 0.f) {
    float scale = majorLength / (minorLength * t.max_anisotropy);
    ds1 *= scale;
    dt1 *= scale;
    minorLength *= scale;
  }
  int nLevels = (int)t.n_levels;
  float lod = pb_max(0.f, nLevels - 1.f + pb_log2(minorLength));
  int ilod = pb_floor2int(lod);
  float d = lod - ilod;
  return (1.f - d) * mip_ewa(sv, t, s, tt, ds0, dt0, ds1, dt1, ilod) +
         d * mip_ewa(sv, t, s, tt, ds0, dt0, ds1, dt1, ilod + 1);
}

// ---------------------------------------------------------------- texture graphs
// ImageTexture::Evaluate, textures/imagemap.cpp:122-128
PB_NOINLINE Spec tex_image_eval(const SceneView &sv, const PbrtB200Texture &t, const DiffGeom &dg, const DgDiff &dd) {
  float s, tt, dsdx, dtdx, dsdy, dtdy;
  tex_map(t, dg, dd, &s, &tt, &dsdx, &dtdx, &dsdy, &dtdy);
#if defined(PB_TEX_DEBUG) && !defined(__CUDACC__)   // trace for diffing against an instrumented reference plugin
  fprintf(stderr, "TEX %a %a %a %a %a %a -> %a\n", s, tt, dsdx, dtdx, dsdy, dtdy, mip_lookup(sv, t, s, tt, dsdx, dtdx, dsdy, dtdy).c[0]);
#endif
  return mip_lookup(sv, t, s, tt, dsdx, dtdx, dsdy, dtdy);
}
// Operands always precede the node that uses them and the host refuses graphs more than PB_TEX_MAX_DEPTH
// operators deep, so the recursion is a fixed chain of (non-inlined) instances.
#define PB_TEX_MAX_DEPTH 3
template <int DEPTH>
PB_NOINLINE Spec tex_eval(const SceneView &sv, int32_t id, const DiffGeom &dg, const DgDiff &dd) {
  const PbrtB200Texture &t = sv.textures[id];
  if (t.kind == PBRT_B200_TEX_IMAGEMAP) return tex_image_eval(sv, t, dg, dd);
  if (t.kind == PBRT_B200_TEX_CONSTANT) return spec3(t.value[0], t.value[1], t.value[2]);
  if (DEPTH > 0) {
    Spec a = tex_eval<(DEPTH > 0 ? DEPTH - 1 : 0)>(sv, t.arg[0], dg, dd);
    Spec b = tex_eval<(DEPTH > 0 ? DEPTH - 1 : 0)>(sv, t.arg[1], dg, dd);
    if (t.kind == PBRT_B200_TEX_SCALE) return a * b;
    float amt = tex_eval<(DEPTH > 0 ? DEPTH - 1 : 0)>(sv, t.arg[2], dg, dd).c[0];
    return (1.f - amt) * a + amt * b;
  }
  return spec(0.f);
}
PB_HD float tex_clamp0(float v) { return v < 0.f ? 0.f : (v > PB_INFINITY ? PB_INFINITY : v); }   // Spectrum::Clamp(0, INFINITY)
// The operand evaluations of Material::GetBSDF (materials/*.cpp): every textured slot of the material record is
// replaced by its value at this hit, clamped as the material clamps it.
PB_NOINLINE void material_eval_textures(const SceneView &sv, const DiffGeom &dg, const DgDiff &dd,
                                        PbrtB200Material *m) {
#pragma unroll 1
  for (int i = 0; i < PBRT_B200_TEXSLOTS; ++i) {
    int32_t id = m->tex[i];
    if (id < 0 || i == PBRT_B200_TEXSLOT_BUMP) continue;   // the displacement is sampled by bump_samples()
    Spec v = tex_eval<PB_TEX_MAX_DEPTH>(sv, id, dg, dd);
    if (i <= PBRT_B200_TEXSLOT_OPACITY) {
      float *dst = i == PBRT_B200_TEXSLOT_KD ? m->kd : i == PBRT_B200_TEXSLOT_KS ? m->ks : i == PBRT_B200_TEXSLOT_KR ? m->kr
                 : i == PBRT_B200_TEXSLOT_KT ? m->kt : m->opacity;
      for (int c = 0; c < 3; ++c) dst[c] = tex_clamp0(v.c[c]);
    } else if (i == PBRT_B200_TEXSLOT_SIGMA)
      m->sigma = m->kind == PBRT_B200_MAT_MATTE ? pb_clamp(v.c[0], 0.f, 90.f) : v.c[0];   // matte.cpp:57
    else if (i == PBRT_B200_TEXSLOT_INDEX) m->index = v.c[0];
    else m->roughness = v.c[0];
  }
}
#endif
