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

#if (defined(PB_TWIN_LIBM_FLOAT) || defined(PB_HOST_LIBM_FLOAT)) && !defined(__CUDACC__)
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
        float weight = sv.ewaLut[pb_imin(pb_f2i(r2 * PBRT_B200_EWA_LUT_SIZE), PBRT_B200_EWA_LUT_SIZE - 1)];
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

// ---------------------------------------------------------------- Perlin noise (core/texture.cpp:156-238)
PB_HD float noise_grad(const unsigned char *perm, int x, int y, int z, float dx, float dy, float dz) {
  int h = perm[perm[perm[x] + y] + z];
  h &= 15;
  float u = h < 8 || h == 12 || h == 13 ? dx : dy;
  float v = h < 4 || h == 12 || h == 13 ? dy : dz;
  return ((h & 1) ? -u : u) + ((h & 2) ? -v : v);
}
PB_HD float noise_weight(float t) {
  float t3 = t * t * t;
  float t4 = t3 * t;
  return 6.f * t4 * t - 15.f * t4 + 10.f * t3;
}
PB_NOINLINE float pb_noise(const unsigned char *perm, float x, float y, float z) {
  int ix = pb_floor2int(x), iy = pb_floor2int(y), iz = pb_floor2int(z);
  float dx = x - ix, dy = y - iy, dz = z - iz;
  ix &= 255; iy &= 255; iz &= 255;
  float w000 = noise_grad(perm, ix, iy, iz, dx, dy, dz);
  float w100 = noise_grad(perm, ix + 1, iy, iz, dx - 1, dy, dz);
  float w010 = noise_grad(perm, ix, iy + 1, iz, dx, dy - 1, dz);
  float w110 = noise_grad(perm, ix + 1, iy + 1, iz, dx - 1, dy - 1, dz);
  float w001 = noise_grad(perm, ix, iy, iz + 1, dx, dy, dz - 1);
  float w101 = noise_grad(perm, ix + 1, iy, iz + 1, dx - 1, dy, dz - 1);
  float w011 = noise_grad(perm, ix, iy + 1, iz + 1, dx, dy - 1, dz - 1);
  float w111 = noise_grad(perm, ix + 1, iy + 1, iz + 1, dx - 1, dy - 1, dz - 1);
  float wx = noise_weight(dx), wy = noise_weight(dy), wz = noise_weight(dz);
  float x00 = pb_lerp(wx, w000, w100), x10 = pb_lerp(wx, w010, w110);
  float x01 = pb_lerp(wx, w001, w101), x11 = pb_lerp(wx, w011, w111);
  float y0 = pb_lerp(wy, x00, x10), y1 = pb_lerp(wy, x01, x11);
  return pb_lerp(wz, y0, y1);
}
PB_HD float pb_smoothstep(float mn, float mx, float value) {   // core/pbrt.h:660-663
  float v = pb_clamp((value - mn) / (mx - mn), 0.f, 1.f);
  return v * v * (-2.f * v + 3.f);
}
// FBm / Turbulence (TURB takes |Noise|), core/texture.cpp:203-238
template <bool TURB>
PB_NOINLINE float pb_fbm(const unsigned char *perm, const V3 &P, const V3 &dpdx, const V3 &dpdy, float omega, int maxOctaves) {
  float s2 = pb_max(length_sq(dpdx), length_sq(dpdy));
  float foctaves = pb_min((float)maxOctaves, 1.f - .5f * pb_log2(s2));
  int octaves = pb_floor2int(foctaves);
  float sum = 0.f, lambda = 1.f, o = 1.f;
  for (int i = 0; i < octaves; ++i) {
    float n = pb_noise(perm, lambda * P.x, lambda * P.y, lambda * P.z);
    sum += o * (TURB ? fabsf(n) : n);
    lambda *= 1.99f;
    o *= omega;
  }
  float partialOctave = foctaves - octaves;
  float n = pb_noise(perm, lambda * P.x, lambda * P.y, lambda * P.z);
  sum += o * pb_smoothstep(.3f, .7f, partialOctave) * (TURB ? fabsf(n) : n);
  return sum;
}
// MarbleTexture::Evaluate's colour spline, textures/marble.cpp:47-69
PB_NOINLINE Spec marble_color(float marble) {
  float t = .5f + .5f * pb_sinf(marble);
  const float c[9][3] = {{.58f, .58f, .6f}, {.58f, .58f, .6f}, {.58f, .58f, .6f}, {.5f, .5f, .5f}, {.6f, .59f, .58f},
                         {.58f, .58f, .6f}, {.58f, .58f, .6f}, {.2f, .2f, .33f}, {.58f, .58f, .6f}};
  const int nseg = 9 - 3;
  int first = pb_floor2int(t * nseg);
  t = (t * nseg - first);
  Spec c0 = spec3(c[first][0], c[first][1], c[first][2]), c1 = spec3(c[first + 1][0], c[first + 1][1], c[first + 1][2]);
  Spec c2 = spec3(c[first + 2][0], c[first + 2][1], c[first + 2][2]), c3 = spec3(c[first + 3][0], c[first + 3][1], c[first + 3][2]);
  Spec s0 = (1.f - t) * c0 + t * c1;
  Spec s1 = (1.f - t) * c1 + t * c2;
  Spec s2 = (1.f - t) * c2 + t * c3;
  s0 = (1.f - t) * s0 + t * s1;
  s1 = (1.f - t) * s1 + t * s2;
  return 1.5f * ((1.f - t) * s0 + t * s1);
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
// The procedural leaves: textures/{uv,bilerp,fbm,wrinkled,windy,marble}.cpp
PB_NOINLINE Spec tex_procedural_eval(const SceneView &sv, const PbrtB200Texture &t, const DiffGeom &dg, const DgDiff &dd) {
  if (t.kind == PBRT_B200_TEX_UV || t.kind == PBRT_B200_TEX_BILERP) {
    float s, tt, dsdx, dtdx, dsdy, dtdy;
    tex_map(t, dg, dd, &s, &tt, &dsdx, &dtdx, &dsdy, &dtdy);
    if (t.kind == PBRT_B200_TEX_UV) return spec3(s - pb_floor2int(s), tt - pb_floor2int(tt), 0.f);   // uv.cpp:39-46
    const float *v = t.param;                                                                     // bilerp.cpp:45-52
    return (1 - s) * (1 - tt) * spec3(v[0], v[1], v[2]) + (1 - s) * (tt) * spec3(v[3], v[4], v[5]) +
           (s) * (1 - tt) * spec3(v[6], v[7], v[8]) + (s) * (tt) * spec3(v[9], v[10], v[11]);
  }
  // IdentityMapping3D::Map, core/texture.cpp:150-155
  V3 P = xf_point(t.world_to_texture, dg.p);
  V3 dpdx = xf_vector(t.world_to_texture, dd.dpdx), dpdy = xf_vector(t.world_to_texture, dd.dpdy);
  const unsigned char *perm = sv.noisePerm;
  if (t.kind == PBRT_B200_TEX_FBM) return spec(pb_fbm<false>(perm, P, dpdx, dpdy, t.param[0], t.iparam[0]));
  if (t.kind == PBRT_B200_TEX_WRINKLED) return spec(pb_fbm<true>(perm, P, dpdx, dpdy, t.param[0], t.iparam[0]));
  if (t.kind == PBRT_B200_TEX_WINDY) {   // windy.cpp:38-46
    float windStrength = pb_fbm<false>(perm, .1f * P, .1f * dpdx, .1f * dpdy, .5f, 3);
    float waveHeight = pb_fbm<false>(perm, P, dpdx, dpdy, .5f, 6);
    return spec(fabsf(windStrength) * waveHeight);
  }
  // marble.cpp:42-46
  float scale = t.param[1], variation = t.param[2];
  P = v3(P.x * scale, P.y * scale, P.z * scale);
  float marble = P.y + variation * pb_fbm<false>(perm, P, scale * dpdx, scale * dpdy, t.param[0], t.iparam[0]);
  return marble_color(marble);
}
// Operands always precede the node that uses them and the host refuses graphs more than PB_TEX_MAX_DEPTH
// operators deep, so the recursion is a fixed chain of (non-inlined) instances.
#define PB_TEX_MAX_DEPTH 3
template <int DEPTH>
PB_NOINLINE Spec tex_eval(const SceneView &sv, int32_t id, const DiffGeom &dg, const DgDiff &dd) {
  const PbrtB200Texture &t = sv.textures[id];
  if (t.kind == PBRT_B200_TEX_IMAGEMAP) return tex_image_eval(sv, t, dg, dd);
  if (t.kind == PBRT_B200_TEX_CONSTANT) return spec3(t.value[0], t.value[1], t.value[2]);
  if (t.kind == PBRT_B200_TEX_UV || t.kind == PBRT_B200_TEX_BILERP || t.kind >= PBRT_B200_TEX_FBM)
    return tex_procedural_eval(sv, t, dg, dd);
  if (DEPTH > 0) {
    const int D = DEPTH > 0 ? DEPTH - 1 : 0;
    if (t.kind == PBRT_B200_TEX_SCALE) return tex_eval<D>(sv, t.arg[0], dg, dd) * tex_eval<D>(sv, t.arg[1], dg, dd);
    if (t.kind == PBRT_B200_TEX_MIX) {
      Spec a = tex_eval<D>(sv, t.arg[0], dg, dd);
      Spec b = tex_eval<D>(sv, t.arg[1], dg, dd);
      float amt = tex_eval<D>(sv, t.arg[2], dg, dd).c[0];
      return (1.f - amt) * a + amt * b;
    }
    float s, tt, dsdx, dtdx, dsdy, dtdy;
    tex_map(t, dg, dd, &s, &tt, &dsdx, &dtdx, &dsdy, &dtdy);
    if (t.kind == PBRT_B200_TEX_DOTS) {   // dots.cpp:43-61; arg[0] plays outsideDot, arg[1] insideDot
      const unsigned char *perm = sv.noisePerm;
      int sCell = pb_floor2int(s + .5f), tCell = pb_floor2int(tt + .5f);
      if (pb_noise(perm, sCell + .5f, tCell + .5f, .5f) > 0) {
        float radius = .35f;
        float maxShift = 0.5f - radius;
        float sCenter = sCell + maxShift * pb_noise(perm, sCell + 1.5f, tCell + 2.8f, .5f);
        float tCenter = tCell + maxShift * pb_noise(perm, sCell + 4.5f, tCell + 9.8f, .5f);
        float ds = s - sCenter, dt = tt - tCenter;
        if (ds * ds + dt * dt < radius * radius) return tex_eval<D>(sv, t.arg[1], dg, dd);
      }
      return tex_eval<D>(sv, t.arg[0], dg, dd);
    }
    // Checkerboard2D, checkerboard.cpp:57-87,128-131
    bool even = (pb_floor2int(s) + pb_floor2int(tt)) % 2 == 0;
    if (t.iparam[0] == PBRT_B200_CHECKER_CLOSEDFORM) {
      float ds = pb_max(fabsf(dsdx), fabsf(dsdy));
      float dt = pb_max(fabsf(dtdx), fabsf(dtdy));
      float s0 = s - ds, s1 = s + ds;
      float t0 = tt - dt, t1 = tt + dt;
      if (!(pb_floor2int(s0) == pb_floor2int(s1) && pb_floor2int(t0) == pb_floor2int(t1))) {
#define PB_BUMPINT(x) (pb_floor2int((x) / 2) + 2.f * pb_max(((x) / 2) - pb_floor2int((x) / 2) - .5f, 0.f))
        float sint = (PB_BUMPINT(s1) - PB_BUMPINT(s0)) / (2.f * ds);
        float tint = (PB_BUMPINT(t1) - PB_BUMPINT(t0)) / (2.f * dt);
#undef PB_BUMPINT
        float area2 = sint + tint - 2.f * sint * tint;
        if (ds > 1.f || dt > 1.f) area2 = .5f;
        return (1.f - area2) * tex_eval<D>(sv, t.arg[0], dg, dd) + area2 * tex_eval<D>(sv, t.arg[1], dg, dd);
      }
    }
    return tex_eval<D>(sv, even ? t.arg[0] : t.arg[1], dg, dd);
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
