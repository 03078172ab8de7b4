// pb_shapes.h -- ray/triangle and ray/sphere intersection and the
// differential geometry the reference builds at an accepted hit.
//   Triangle::Intersect / IntersectP   shapes/trianglemesh.cpp:213-314
//   Triangle::GetUVs                   shapes/trianglemesh.cpp:315-328
//   Triangle::GetShadingGeometry       shapes/trianglemesh.cpp:71-133
//   Sphere::Intersect / IntersectP     shapes/sphere.cpp:102-250
//   Quadratic                          core/pbrt.h:645-659
//   DifferentialGeometry ctor          core/shape.cpp:37-51
#ifndef PB_SHAPES_H
#define PB_SHAPES_H
#include "pb_scene.h"

struct TriVerts {
  V3 p1, p2, p3;
  uint32_t flags, material;
  int32_t light;
};
PB_HD TriVerts load_prim(const SceneView &sv, uint32_t prim) {
  const PrimQuad *q = sv.prims + 3u * prim;
#ifdef __CUDA_ARCH__
  float4 a = __ldg((const float4 *)q), b = __ldg((const float4 *)(q + 1)),
         c = __ldg((const float4 *)(q + 2));
  TriVerts t;
  t.p1 = v3(a.x, a.y, a.z); t.p2 = v3(a.w, b.x, b.y); t.p3 = v3(b.z, b.w, c.x);
  t.flags = __float_as_uint(c.y); t.material = __float_as_uint(c.z);
  t.light = __float_as_int(c.w);
  return t;
#else
  const PbrtB200Prim *p = (const PbrtB200Prim *)q;
  TriVerts t;
  t.p1 = v3(p->p[0], p->p[1], p->p[2]); t.p2 = v3(p->p[3], p->p[4], p->p[5]);
  t.p3 = v3(p->p[6], p->p[7], p->p[8]);
  t.flags = p->flags; t.material = p->material; t.light = p->light;
  return t;
#endif
}

// The part of Triangle::Intersect shared with IntersectP (trianglemesh.cpp:221-246).
PB_HD bool tri_test(const V3 &p1, const V3 &p2, const V3 &p3, const Ray &ray, float *tHit,
                    float *b1o, float *b2o) {
  V3 e1 = p2 - p1;
  V3 e2 = p3 - p1;
  V3 s1 = cross(ray.d, e2);
  float divisor = dot(s1, e1);
  if (divisor == 0.f) return false;
  float invDivisor = 1.f / divisor;
  V3 d = ray.o - p1;
  float b1 = dot(d, s1) * invDivisor;
  if (b1 < 0.f || b1 > 1.f) return false;
  V3 s2 = cross(d, e1);
  float b2 = dot(ray.d, s2) * invDivisor;
  if (b2 < 0.f || b1 + b2 > 1.f) return false;
  float t = dot(e2, s2) * invDivisor;
  if (t < ray.mint || t > ray.maxt) return false;
  *tHit = t; *b1o = b1; *b2o = b2;
  return true;
}

// core/pbrt.h:645-659
PB_HD bool quadratic(float A, float B, float C, float *t0, float *t1) {
  float discrim = B * B - 4.f * A * C;
  if (discrim < 0.f) return false;
  float rootDiscrim = sqrtf(discrim);
  float q;
  if (B < 0) q = -.5f * (B - rootDiscrim);
  else q = -.5f * (B + rootDiscrim);
  *t0 = q / A;
  *t1 = C / q;
  if (*t0 > *t1) { float tmp = *t0; *t0 = *t1; *t1 = tmp; }
  return true;
}

// Sphere::Intersect up to the accepted hit (sphere.cpp:104-143 == IntersectP
// :209-249).  Outputs object-space hit point and phi for the DG stage.
// wantPhi = false lets a pure hit query skip atan2f for an unclipped sphere: phi only feeds the
// clipping tests (always false when zmin/zmax/phimax do not cut the sphere: phi never exceeds the
// float 2*pi it is compared with) and the parametric u computed later from a fresh call.
PB_NOINLINE bool sphere_test(const PbrtB200Quadric &s, const Ray &r, float *tHit, V3 *phitO, float *phiO,
                             bool wantPhi = true) {
  Ray ray;
  ray.o = xf_point(s.w2o, r.o);
  ray.d = xf_vector(s.w2o, r.d);
  ray.mint = r.mint; ray.maxt = r.maxt;
  float radius = s.radius;
  float A = ray.d.x * ray.d.x + ray.d.y * ray.d.y + ray.d.z * ray.d.z;
  float B = 2 * (ray.d.x * ray.o.x + ray.d.y * ray.o.y + ray.d.z * ray.o.z);
  float C = ray.o.x * ray.o.x + ray.o.y * ray.o.y + ray.o.z * ray.o.z - radius * radius;
  float t0, t1;
  if (!quadratic(A, B, C, &t0, &t1)) return false;
  if (t0 > ray.maxt || t1 < ray.mint) return false;
  float thit = t0;
  if (t0 < ray.mint) {
    thit = t1;
    if (thit > ray.maxt) return false;
  }
  V3 phit = ray_at(ray, thit);
  bool clipped = s.zmin > -radius || s.zmax < radius || s.phi_max < 2.f * PB_PI;
  if (!clipped && !wantPhi) {
    *tHit = thit; *phitO = phit; *phiO = 0.f;
    return true;
  }
  float phi = pb_atan2f(phit.y, phit.x);
  if (phi < 0.f) phi += 2.f * PB_PI;
  if ((s.zmin > -radius && phit.z < s.zmin) || (s.zmax < radius && phit.z > s.zmax) ||
      phi > s.phi_max) {
    if (thit == t1) return false;
    if (t1 > ray.maxt) return false;
    thit = t1;
    phit = ray_at(ray, thit);
    phi = pb_atan2f(phit.y, phit.x);
    if (phi < 0.f) phi += 2.f * PB_PI;
    if (phit.z < s.zmin || phit.z > s.zmax || phi > s.phi_max) return false;
  }
  *tHit = thit; *phitO = phit; *phiO = phi;
  return true;
}

// Geometry at a hit, as much of DifferentialGeometry (core/shape.h:32-50) as
// radiance depends on with constant textures (the
// screen-space differentials only feed texture filtering).
struct DiffGeom {
  V3 p, nn, dpdu, dpdv, dndu, dndv;
  float u, v;
};
// core/shape.cpp:37-51
PB_HD void dg_init(DiffGeom *dg, const V3 &P, const V3 &dpdu, const V3 &dpdv, const V3 &dndu,
                   const V3 &dndv, float u, float v, uint32_t flags) {
  dg->p = P; dg->dpdu = dpdu; dg->dpdv = dpdv; dg->dndu = dndu; dg->dndv = dndv;
  dg->nn = normalize(cross(dpdu, dpdv));
  dg->u = u; dg->v = v;
  bool ro = (flags & PBRT_B200_PF_REVERSE) != 0, sh = (flags & PBRT_B200_PF_SWAPS) != 0;
  if (ro ^ sh) dg->nn = dg->nn * -1.f;
}

PB_HD void tri_uvs(const SceneView &sv, uint32_t prim, uint32_t flags, float uv[3][2]) {
  if (flags & PBRT_B200_PF_HAS_UV) {
    const PbrtB200PrimAux &a = sv.aux[prim];
    for (int i = 0; i < 3; ++i) {
      uv[i][0] = sv.vertexUV[2 * a.v[i]];
      uv[i][1] = sv.vertexUV[2 * a.v[i] + 1];
    }
  } else {
    uv[0][0] = 0.f; uv[0][1] = 0.f;
    uv[1][0] = 1.f; uv[1][1] = 0.f;
    uv[2][0] = 1.f; uv[2][1] = 1.f;
  }
}

// trianglemesh.cpp:248-277: DG of an accepted triangle hit.
PB_HD void tri_dg(const SceneView &sv, uint32_t prim, const TriVerts &tv, const Ray &ray, float t,
                  float b1, float b2, DiffGeom *dg) {
  V3 e1 = tv.p2 - tv.p1;
  V3 e2 = tv.p3 - tv.p1;
  V3 dpdu, dpdv;
  float uvs[3][2];
  tri_uvs(sv, prim, tv.flags, uvs);
  float du1 = uvs[0][0] - uvs[2][0];
  float du2 = uvs[1][0] - uvs[2][0];
  float dv1 = uvs[0][1] - uvs[2][1];
  float dv2 = uvs[1][1] - uvs[2][1];
  V3 dp1 = tv.p1 - tv.p3, dp2 = tv.p2 - tv.p3;
  float determinant = du1 * dv2 - dv1 * du2;
  if (determinant == 0.f) {
    coordinate_system(normalize(cross(e2, e1)), &dpdu, &dpdv);
  } else {
    float invdet = 1.f / determinant;
    dpdu = (dv2 * dp1 - dv1 * dp2) * invdet;
    dpdv = (-du2 * dp1 + du1 * dp2) * invdet;
  }
  float b0 = 1 - b1 - b2;
  float tu = b0 * uvs[0][0] + b1 * uvs[1][0] + b2 * uvs[2][0];
  float tv_ = b0 * uvs[0][1] + b1 * uvs[1][1] + b2 * uvs[2][1];
  dg_init(dg, ray_at(ray, t), dpdu, dpdv, v3(0, 0, 0), v3(0, 0, 0), tu, tv_, tv.flags);
}

// core/util.cpp:99-108
PB_HD bool solve2x2(const float A[2][2], const float B[2], float x[2]) {
  float det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  if (fabsf(det) < 1e-5) return false;
  float invDet = 1.0f / det;
  x[0] = (A[1][1] * B[0] - A[0][1] * B[1]) * invDet;
  x[1] = (A[0][0] * B[1] - A[1][0] * B[0]) * invDet;
  return true;
}

// trianglemesh.cpp:71-133: shading geometry from per-vertex N / S.
// Matrix4x4::Mul, core/transform.cpp:78-89 (row major 4x4)
PB_HD void mat4_mul(const float *m1, const float *m2, float *r) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      r[4 * i + j] = m1[4 * i + 0] * m2[0 * 4 + j] + m1[4 * i + 1] * m2[1 * 4 + j] + m1[4 * i + 2] * m2[2 * 4 + j] +
                     m1[4 * i + 3] * m2[3 * 4 + j];
}
// `inst`: the hit was reached through an InstancePrimitive, so the obj2world handed to
// GetShadingGeometry is (shape->WorldToObject * WorldToInstance).GetInverse() (primitive.cpp:71,132):
// m = InstanceToWorld.m * ObjectToWorld.m, inverse = WorldToObject.m * WorldToInstance.m (Transform::operator*).
template <bool INST = true>
PB_NOINLINE void tri_shading_dg(const SceneView &sv, uint32_t prim, uint32_t flags, const DiffGeom &dg,
                          DiffGeom *dgs, const PbrtB200Instance *inst = nullptr) {
  bool hasN = (flags & PBRT_B200_PF_HAS_N) != 0, hasS = (flags & PBRT_B200_PF_HAS_S) != 0;
  if (!hasN && !hasS) { *dgs = dg; return; }
  const PbrtB200PrimAux &a = sv.aux[prim];
  const PbrtB200Mesh &meshOwn = sv.meshes[a.owner];
  struct { const float *w2o, *o2w; } mesh = {meshOwn.w2o, meshOwn.o2w};
  float cw2o[INST ? 16 : 1], co2w[INST ? 16 : 1];
  if (INST && inst) {
    mat4_mul(meshOwn.w2o, inst->w2i, cw2o);
    mat4_mul(inst->i2w, meshOwn.o2w, co2w);
    mesh.w2o = cw2o; mesh.o2w = co2w;
  }
  float b[3];
  float uv[3][2];
  tri_uvs(sv, prim, flags, uv);
  float A[2][2] = {{uv[1][0] - uv[0][0], uv[2][0] - uv[0][0]},
                   {uv[1][1] - uv[0][1], uv[2][1] - uv[0][1]}};
  float C[2] = {dg.u - uv[0][0], dg.v - uv[0][1]};
  if (!solve2x2(A, C, &b[1])) {
    b[0] = b[1] = b[2] = 1.f / 3.f;
  } else
    b[0] = 1.f - b[1] - b[2];
  V3 n0, n1, n2;
  V3 ns, ss, ts;
  if (hasN) {
    n0 = v3(sv.vertexN[3 * a.v[0]], sv.vertexN[3 * a.v[0] + 1], sv.vertexN[3 * a.v[0] + 2]);
    n1 = v3(sv.vertexN[3 * a.v[1]], sv.vertexN[3 * a.v[1] + 1], sv.vertexN[3 * a.v[1] + 2]);
    n2 = v3(sv.vertexN[3 * a.v[2]], sv.vertexN[3 * a.v[2] + 1], sv.vertexN[3 * a.v[2] + 2]);
    // obj2world(Normal) uses the inverse of obj2world, i.e. WorldToObject
    ns = normalize(xf_normal(mesh.w2o, b[0] * n0 + b[1] * n1 + b[2] * n2));
  } else
    ns = dg.nn;
  if (hasS) {
    V3 s0 = v3(sv.vertexS[3 * a.v[0]], sv.vertexS[3 * a.v[0] + 1], sv.vertexS[3 * a.v[0] + 2]);
    V3 s1 = v3(sv.vertexS[3 * a.v[1]], sv.vertexS[3 * a.v[1] + 1], sv.vertexS[3 * a.v[1] + 2]);
    V3 s2 = v3(sv.vertexS[3 * a.v[2]], sv.vertexS[3 * a.v[2] + 1], sv.vertexS[3 * a.v[2] + 2]);
    ss = normalize(xf_vector(mesh.o2w, b[0] * s0 + b[1] * s1 + b[2] * s2));
  } else
    ss = normalize(dg.dpdu);
  ts = normalize(cross(ss, ns));
  ss = cross(ts, ns);
  V3 dndu, dndv;
  if (hasN) {
    float du1 = uv[0][0] - uv[2][0];
    float du2 = uv[1][0] - uv[2][0];
    float dv1 = uv[0][1] - uv[2][1];
    float dv2 = uv[1][1] - uv[2][1];
    V3 dn1 = n0 - n2;
    V3 dn2 = n1 - n2;
    float determinant = du1 * dv2 - dv1 * du2;
    if (determinant == 0)
      dndu = dndv = v3(0, 0, 0);
    else {
      float invdet = 1.f / determinant;
      dndu = (dv2 * dn1 - dv1 * dn2) * invdet;
      dndv = (-du2 * dn1 + du1 * dn2) * invdet;
    }
  } else
    dndu = dndv = v3(0, 0, 0);
  // `ObjectToWorld(dndu)` on a Vector (trianglemesh.cpp:126-127) is the
  // Triangle's own ObjectToWorld == the mesh's
  dg_init(dgs, dg.p, ss, ts, xf_vector(meshOwn.o2w, dndu), xf_vector(meshOwn.o2w, dndv), dg.u, dg.v,
          flags);
}

// sphere.cpp:144-205: DG of an accepted sphere hit (object-space phit, phi).
PB_NOINLINE void sphere_dg(const PbrtB200Quadric &s, const V3 &phit, float phi, DiffGeom *dg) {
  float radius = s.radius, phiMax = s.phi_max, thetaMin = s.theta_min, thetaMax = s.theta_max;
  float u = phi / phiMax;
  float theta = pb_acosf(pb_clamp(phit.z / radius, -1.f, 1.f));
  float v = (theta - thetaMin) / (thetaMax - thetaMin);
  float cosphi, sinphi;
  V3 dpdu, dpdv;
  float zradius = sqrtf(phit.x * phit.x + phit.y * phit.y);
  if (zradius == 0) {
    cosphi = 0;
    sinphi = 1;
    dpdv = (thetaMax - thetaMin) * v3(phit.z * cosphi, phit.z * sinphi, -radius * pb_sinf(theta));
    V3 norm = phit;
    dpdu = cross(dpdv, norm);
  } else {
    float invzradius = 1.f / zradius;
    cosphi = phit.x * invzradius;
    sinphi = phit.y * invzradius;
    dpdu = v3(-phiMax * phit.y, phiMax * phit.x, 0);
    dpdv = (thetaMax - thetaMin) * v3(phit.z * cosphi, phit.z * sinphi, -radius * pb_sinf(theta));
  }
  V3 d2Pduu = -phiMax * phiMax * v3(phit.x, phit.y, 0);
  V3 d2Pduv = (thetaMax - thetaMin) * phit.z * phiMax * v3(-sinphi, cosphi, 0.f);
  V3 d2Pdvv = -(thetaMax - thetaMin) * (thetaMax - thetaMin) * v3(phit.x, phit.y, phit.z);
  float E = dot(dpdu, dpdu);
  float F = dot(dpdu, dpdv);
  float G = dot(dpdv, dpdv);
  V3 N = normalize(cross(dpdu, dpdv));
  float e = dot(N, d2Pduu);
  float f = dot(N, d2Pduv);
  float g = dot(N, d2Pdvv);
  float invEGF2 = 1.f / (E * G - F * F);
  V3 dndu = (f * F - e * G) * invEGF2 * dpdu + (e * F - f * E) * invEGF2 * dpdv;
  V3 dndv = (g * F - f * G) * invEGF2 * dpdu + (f * F - g * E) * invEGF2 * dpdv;
  dg_init(dg, xf_point(s.o2w, phit), xf_vector(s.o2w, dpdu), xf_vector(s.o2w, dpdv),
          xf_vector(s.o2w, dndu), xf_vector(s.o2w, dndv), u, v, s.flags);
}

// ---------------------------------------------------------------- the other quadrics
// Hyperboloid inverse mapping (hyperboloid.cpp:105-112): v along the generating segment, phi from
// the rotation that takes the segment point to the hit point.
PB_HD float hyperboloid_phi(const PbrtB200Quadric &q, const V3 &phit, float *vOut) {
  float v = (phit.z - q.p1[2]) / (q.p2[2] - q.p1[2]);
  V3 pr = v3((1.f - v) * q.p1[0] + v * q.p2[0], (1.f - v) * q.p1[1] + v * q.p2[1],
             (1.f - v) * q.p1[2] + v * q.p2[2]);
  float phi = pb_atan2f(pr.x * phit.y - phit.x * pr.y, phit.x * pr.x + phit.y * pr.y);
  if (phi < 0) phi += 2 * PB_PI;
  *vOut = v;
  return phi;
}
// Cylinder / Cone / Paraboloid / Hyperboloid ::Intersect up to the accepted hit
// (cylinder.cpp:65-110, cone.cpp:54-105, paraboloid.cpp:58-105, hyperboloid.cpp:77-128; the
// IntersectP bodies are the same text) and Disk::Intersect (disk.cpp:62-84).
// Outputs the object-space hit point and phi for the DG stage.
PB_NOINLINE bool quadric_other_test(const PbrtB200Quadric &q, const Ray &r, float *tHit, V3 *phitO, float *phiO) {
  Ray ray;
  ray.o = xf_point(q.w2o, r.o);
  ray.d = xf_vector(q.w2o, r.d);
  ray.mint = r.mint; ray.maxt = r.maxt;
  const uint32_t kind = q.kind;
  if (kind == PBRT_B200_QUADRIC_DISK) {
    if ((double)fabsf(ray.d.z) < 1e-7) return false;
    float thit = (q.height - ray.o.z) / ray.d.z;
    if (thit < ray.mint || thit > ray.maxt) return false;
    V3 phit = ray_at(ray, thit);
    float dist2 = phit.x * phit.x + phit.y * phit.y;
    if (dist2 > q.radius * q.radius || dist2 < q.inner_radius * q.inner_radius) return false;
    float phi = pb_atan2f(phit.y, phit.x);
    if (phi < 0) phi = (float)((double)phi + 2. * (double)PB_PI);   // `phi += 2. * M_PI`
    if (phi > q.phi_max) return false;
    *tHit = thit; *phitO = phit; *phiO = phi;
    return true;
  }
  float A, B, C;
  float zlo = q.zmin, zhi = q.zmax;
  if (kind == PBRT_B200_QUADRIC_CYLINDER) {
    A = ray.d.x * ray.d.x + ray.d.y * ray.d.y;
    B = 2 * (ray.d.x * ray.o.x + ray.d.y * ray.o.y);
    C = ray.o.x * ray.o.x + ray.o.y * ray.o.y - q.radius * q.radius;
  } else if (kind == PBRT_B200_QUADRIC_CONE) {
    float height = q.height;
    float k = q.radius / height;
    k = k * k;
    A = ray.d.x * ray.d.x + ray.d.y * ray.d.y - k * ray.d.z * ray.d.z;
    B = 2 * (ray.d.x * ray.o.x + ray.d.y * ray.o.y - k * ray.d.z * (ray.o.z - height));
    C = ray.o.x * ray.o.x + ray.o.y * ray.o.y - k * (ray.o.z - height) * (ray.o.z - height);
    zlo = 0.f; zhi = height;
  } else if (kind == PBRT_B200_QUADRIC_PARABOLOID) {
    float k = q.zmax / (q.radius * q.radius);
    A = k * (ray.d.x * ray.d.x + ray.d.y * ray.d.y);
    B = 2 * k * (ray.d.x * ray.o.x + ray.d.y * ray.o.y) - ray.d.z;
    C = k * (ray.o.x * ray.o.x + ray.o.y * ray.o.y) - ray.o.z;
  } else {   // hyperboloid
    float a = q.a, c = q.c;
    A = a * ray.d.x * ray.d.x + a * ray.d.y * ray.d.y - c * ray.d.z * ray.d.z;
    B = 2.f * (a * ray.d.x * ray.o.x + a * ray.d.y * ray.o.y - c * ray.d.z * ray.o.z);
    C = a * ray.o.x * ray.o.x + a * ray.o.y * ray.o.y - c * ray.o.z * ray.o.z - 1;
  }
  float t0, t1;
  if (!quadratic(A, B, C, &t0, &t1)) return false;
  if (t0 > ray.maxt || t1 < ray.mint) return false;
  float thit = t0;
  if (t0 < ray.mint) {
    thit = t1;
    if (thit > ray.maxt) return false;
  }
  const bool hyper = kind == PBRT_B200_QUADRIC_HYPERBOLOID;
  V3 phit = ray_at(ray, thit);
  float phi, v;
  if (hyper) phi = hyperboloid_phi(q, phit, &v);
  else {
    phi = pb_atan2f(phit.y, phit.x);
    if (phi < 0.) phi += 2.f * PB_PI;
  }
  if (phit.z < zlo || phit.z > zhi || phi > q.phi_max) {
    if (thit == t1) return false;
    thit = t1;
    if (t1 > ray.maxt) return false;
    phit = ray_at(ray, thit);
    if (hyper) phi = hyperboloid_phi(q, phit, &v);
    else {
      phi = pb_atan2f(phit.y, phit.x);
      if (phi < 0.) phi += 2.f * PB_PI;
    }
    if (phit.z < zlo || phit.z > zhi || phi > q.phi_max) return false;
  }
  *tHit = thit; *phitO = phit; *phiO = phi;
  return true;
}
// Any quadric.  wantPhi as for sphere_test (the other kinds always need phi for their clipping test).
PB_HD bool quadric_test(const PbrtB200Quadric &q, const Ray &r, float *tHit, V3 *phitO, float *phiO,
                        bool wantPhi = true) {
  if (q.kind == PBRT_B200_QUADRIC_SPHERE) return sphere_test(q, r, tHit, phitO, phiO, wantPhi);
  return quadric_other_test(q, r, tHit, phitO, phiO);
}

// DG of an accepted hit on the other quadrics: cylinder.cpp:111-145, cone.cpp:106-147,
// paraboloid.cpp:106-152, hyperboloid.cpp:129-170 (dndu/dndv through the fundamental forms) and
// disk.cpp:85-103 (dndu = dndv = 0).
PB_NOINLINE void quadric_other_dg(const PbrtB200Quadric &q, const V3 &phit, float phi, DiffGeom *dg) {
  const uint32_t kind = q.kind;
  const float phiMax = q.phi_max;
  float u = phi / phiMax;
  float v;
  V3 dpdu = v3(-phiMax * phit.y, phiMax * phit.x, 0);
  V3 dpdv, dndu, dndv;
  if (kind == PBRT_B200_QUADRIC_DISK) {
    float dist2 = phit.x * phit.x + phit.y * phit.y;
    v = 1.f - ((sqrtf(dist2) - q.inner_radius) / (q.radius - q.inner_radius));
    float su = phiMax * PB_INV_TWOPI;
    dpdu = v3(dpdu.x * su, dpdu.y * su, dpdu.z * su);   // dpdu *= phiMax * INV_TWOPI
    dpdv = v3(-phit.x / (1 - v), -phit.y / (1 - v), 0.f);
    float sv = (q.radius - q.inner_radius) / q.radius;
    dpdv = v3(dpdv.x * sv, dpdv.y * sv, dpdv.z * sv);
    dndu = dndv = v3(0, 0, 0);
  } else {
    V3 d2Pduu = -phiMax * phiMax * v3(phit.x, phit.y, 0);
    V3 d2Pduv = v3(0, 0, 0), d2Pdvv = v3(0, 0, 0);
    if (kind == PBRT_B200_QUADRIC_CYLINDER) {
      v = (phit.z - q.zmin) / (q.zmax - q.zmin);
      dpdv = v3(0, 0, q.zmax - q.zmin);
    } else if (kind == PBRT_B200_QUADRIC_CONE) {
      v = phit.z / q.height;
      dpdv = v3(-phit.x / (1.f - v), -phit.y / (1.f - v), q.height);
      d2Pduv = phiMax / (1.f - v) * v3(-phit.y, -phit.x, 0.f);
    } else if (kind == PBRT_B200_QUADRIC_PARABOLOID) {
      float zmin = q.zmin, zmax = q.zmax;
      v = (phit.z - zmin) / (zmax - zmin);
      dpdv = (zmax - zmin) * v3(phit.x / (2.f * phit.z), phit.y / (2.f * phit.z), 1.f);
      d2Pduv = (zmax - zmin) * phiMax * v3(-phit.y / (2.f * phit.z), phit.x / (2.f * phit.z), 0);
      d2Pdvv = -(zmax - zmin) * (zmax - zmin) *
               v3(phit.x / (4.f * phit.z * phit.z), phit.y / (4.f * phit.z * phit.z), 0.f);
    } else {   // hyperboloid
      hyperboloid_phi(q, phit, &v);   // same expression the accepted hit evaluated
      float cosphi = pb_cosf(phi), sinphi = pb_sinf(phi);
      float ex = q.p2[0] - q.p1[0], ey = q.p2[1] - q.p1[1], ez = q.p2[2] - q.p1[2];
      dpdv = v3(ex * cosphi - ey * sinphi, ex * sinphi + ey * cosphi, ez);
      d2Pduv = phiMax * v3(-dpdv.y, dpdv.x, 0.f);
    }
    float E = dot(dpdu, dpdu);
    float F = dot(dpdu, dpdv);
    float G = dot(dpdv, dpdv);
    V3 N = normalize(cross(dpdu, dpdv));
    float e = dot(N, d2Pduu);
    float f = dot(N, d2Pduv);
    float g = dot(N, d2Pdvv);
    float invEGF2 = 1.f / (E * G - F * F);
    dndu = (f * F - e * G) * invEGF2 * dpdu + (e * F - f * E) * invEGF2 * dpdv;
    dndv = (g * F - f * G) * invEGF2 * dpdu + (f * F - g * E) * invEGF2 * dpdv;
  }
  dg_init(dg, xf_point(q.o2w, phit), xf_vector(q.o2w, dpdu), xf_vector(q.o2w, dpdv), xf_vector(q.o2w, dndu),
          xf_vector(q.o2w, dndv), u, v, q.flags);
}
PB_HD void quadric_dg(const PbrtB200Quadric &q, const V3 &phit, float phi, DiffGeom *dg) {
  if (q.kind == PBRT_B200_QUADRIC_SPHERE) sphere_dg(q, phit, phi, dg);
  else quadric_other_dg(q, phit, phi, dg);
}
#endif
