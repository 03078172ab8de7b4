// pb_integrator.h -- light sampling, the direct-lighting estimator and the
// per-vertex bodies of the three surface integrators.
//   PointLight::Sample_L             lights/point.cpp:55-66
//   AreaLight::Sample_L / Pdf / L    lights/area.cpp:58-72,93-104, core/light.h:93-96
//   Shape::Pdf(p,wi), ShapeSet       core/shape.h:96-121,150-156
//   Triangle::Sample                 shapes/trianglemesh.cpp:336-349
//   Sphere::Sample / Pdf             shapes/sphere.cpp:38-79
//   VisibilityTester::SetSegment     core/light.h:76-80
//   EstimateDirect                   core/transport.cpp:123-194
//   UniformSampleOneLight/AllLights  core/transport.cpp:31-70
//   PathIntegrator::Li               integrators/path.cpp:58-145
//   DirectLighting::Li               integrators/directlighting.cpp:91-194
//   WhittedIntegrator::Li            integrators/whitted.cpp:44-140
//
// Everything here is "one path vertex" worth of work: it never traces a ray
// itself.  Rays it needs (shadow segment, MIS probe, continuation) are
// returned to the caller, which queues them for the traversal kernels; the
// radiance terms that depend on those rays are returned as pending records
// and folded in, in the reference's order of additions, by resolve_*().
#ifndef PB_INTEGRATOR_H
#define PB_INTEGRATOR_H
#include "pb_bsdf.h"
#include "pb_kd.h"
#include "pb_texture.h"

// ---------------------------------------------------------------- geometry at a hit
struct Surface {
  Bsdf bsdf;
  int32_t light;      // area light of the primitive or -1
  uint32_t material;  // index into materials[] (lets a later kernel rebuild the lobes)
};
struct LambertSurface {   // same, for a vertex whose material is matte with sigma == 0 (pb_bsdf.h: LambertBsdf)
  LambertBsdf bsdf;
  int32_t light;
  uint32_t material;
};
// InstancePrimitive::Intersect's tail (core/primitive.cpp:72-83): the instanced shape's DG, computed
// from the instance-space ray, moved to world space.
PB_NOINLINE void instance_dg_to_world(const PbrtB200Instance &in, DiffGeom *dg) {
  dg->p = xf_point(in.i2w, dg->p);
  dg->nn = normalize(xf_normal(in.w2i, dg->nn));
  dg->dpdu = xf_vector(in.i2w, dg->dpdu);
  dg->dpdv = xf_vector(in.i2w, dg->dpdv);
  // DifferentialGeometry::dndu / dndv are declared Normal (core/shape.h:46): this step transforms them as
  // normals (inverse transpose), unlike the shapes' own `ObjectToWorld(dndu)` on a Vector
  dg->dndu = xf_normal(in.w2i, dg->dndu);
  dg->dndv = xf_normal(in.w2i, dg->dndv);
}
// INST = false compiles the instancing code out (kernels for scenes without object instances).
template <bool INST = true>
PB_HD void surface_dg(const SceneView &sv, const Ray &rayWorld, const Hit &hit, const TriVerts &tv,
                      DiffGeom *dg) {
  Ray ray = rayWorld;
  const bool instanced = INST && sv.nInstances != 0 && hit.inst >= 0;
  if (instanced) {   // Ray ray = WorldToInstance(r), primitive.cpp:68
    const PbrtB200Instance &in = sv.instances[hit.inst];
    ray.o = xf_point(in.w2i, rayWorld.o);
    ray.d = xf_vector(in.w2i, rayWorld.d);
  }
  if ((tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE) {
    tri_dg(sv, (uint32_t)hit.prim, tv, ray, hit.t, hit.b1, hit.b2, dg);
  } else {
    const PbrtB200Quadric &s = sv.quadrics[sv.aux[hit.prim].owner];
    Ray r = ray;
    r.maxt = hit.t;   // re-deriving the accepted root: same branch decisions as when it was found
    float t, phi; V3 phit;
    quadric_test(s, r, &t, &phit, &phi);
    quadric_dg(s, phit, phi, dg);
  }
  if (instanced) instance_dg_to_world(sv.instances[hit.inst], dg);
}
// Intersection::GetBSDF (core/primitive.cpp:126-141)
// Material::Bump's three evaluations of the displacement texture (core/material.cpp:33-56): at the point
// shifted along the shading dpdu, along dpdv, and at the point itself.  Of the shifted
// DifferentialGeometry only p, u, v and the inherited screen-space derivatives reach a texture in scope.
PB_NOINLINE void bump_samples(const SceneView &sv, int32_t node, const DiffGeom &dgs, const DgDiff &dd, BumpSamples *o) {
  DiffGeom e = dgs;
  float du = .5f * (fabsf(dd.dudx) + fabsf(dd.dudy));
  if (du == 0.f) du = .01f;
  e.p = dgs.p + du * dgs.dpdu;
  e.u = dgs.u + du;
  o->uDisplace = tex_eval<PB_TEX_MAX_DEPTH>(sv, node, e, dd).c[0];
  float dv = .5f * (fabsf(dd.dvdx) + fabsf(dd.dvdy));
  if (dv == 0.f) dv = .01f;
  e.p = dgs.p + dv * dgs.dpdv;
  e.u = dgs.u;
  e.v = dgs.v + dv;
  o->vDisplace = tex_eval<PB_TEX_MAX_DEPTH>(sv, node, e, dd).c[0];
  o->displace = tex_eval<PB_TEX_MAX_DEPTH>(sv, node, dgs, dd).c[0];
  o->du = du; o->dv = dv;
}
// TEX = true evaluates the material's texture operands at the hit (dg.ComputeDifferentials(ray) and the
// `Kd->Evaluate(dgs)` calls of Material::GetBSDF); `rd` holds the ray's differentials or is null, `evalMat`
// (optional) receives the material record with every textured operand replaced by its value here.
// What the specular-bounce transfer of ray differentials reads at a vertex (whitted.cpp:86-131): the geometric
// dg's dpdx / dpdy, and the shading dg's dndu / dndv with the (inherited) du/dx ... dv/dy.
struct SurfaceDiff {
  DgDiff dd;
  V3 dndu, dndv;
};
// FLATONLY: the caller has established that the primitive carries a precomputed frame (PB_PF_DEV_FRAME) and
// was not reached through an instance; the general evaluation is then compiled out of the call site.
template <bool LOBES = true, bool INST = true, bool TEX = false, bool FLATONLY = false>
PB_HD void surface_at_hit(const SceneView &sv, const Ray &ray, const Hit &hit, Surface *sf,
                          const RayDiff *rd = (const RayDiff *)0, PbrtB200Material *evalMat = (PbrtB200Material *)0,
                          SurfaceDiff *sd = (SurfaceDiff *)0) {
#ifdef __CUDA_ARCH__
  if (FLATONLY || (!TEX && sv.primFrame)) {
    // flat triangle: the frame was computed once by k_prim_frames with this very function (see pb_scene.h)
    const float4 tail = __ldg((const float4 *)(sv.prims + 3u * (uint32_t)hit.prim + 2));
    const uint32_t pflags = __float_as_uint(tail.y);
    if (FLATONLY || ((pflags & PB_PF_DEV_FRAME) && !(INST && hit.inst >= 0))) {
      const float4 *fr = sv.primFrame + 3u * (uint32_t)hit.prim;
      const float4 a = __ldg(fr), b = __ldg(fr + 1), c = __ldg(fr + 2);
      Bsdf &bs = sf->bsdf;
      bs.p = ray_at(ray, hit.t);
      bs.ng = v3(a.x, a.y, a.z); bs.nn = v3(a.w, b.x, b.y); bs.sn = v3(b.z, b.w, c.x); bs.tn = v3(c.y, c.z, c.w);
      sf->material = __float_as_uint(tail.z);
      sf->light = __float_as_int(tail.w);
      if (LOBES) bsdf_set_lobes<true>(sv.materials[sf->material], &bs);
      else { bs.eta = 1.f; bs.nLobes = 0; }
      return;
    }
  }
#endif
  if (FLATONLY) return;
  TriVerts tv = load_prim(sv, (uint32_t)hit.prim);
  DiffGeom dg, dgs;
  surface_dg<INST>(sv, ray, hit, tv, &dg);
  if ((tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE)
    tri_shading_dg<INST>(sv, (uint32_t)hit.prim, tv.flags, dg, &dgs,
                   (INST && sv.nInstances != 0 && hit.inst >= 0) ? &sv.instances[hit.inst] : (const PbrtB200Instance *)0);
  else
    dgs = dg;
  if (TEX) {
    // the shading geometry inherits u, v, p and the screen-space derivatives of the geometric one
    // (trianglemesh.cpp:130-132, shape.h:77-80), so the operands are evaluated from `dg`
    PbrtB200Material m = sv.materials[tv.material];
    DgDiff dd;
    compute_differentials(dg, rd, &dd);
    material_eval_textures(sv, dg, dd, &m);
    BumpSamples bump;
    const bool bumped = m.tex[PBRT_B200_TEXSLOT_BUMP] >= 0;
    if (bumped) bump_samples(sv, m.tex[PBRT_B200_TEXSLOT_BUMP], dgs, dd, &bump);
    make_bsdf<LOBES>(m, dg, dgs, tv.flags, &sf->bsdf, bumped ? &bump : (const BumpSamples *)0);
    if (evalMat) *evalMat = m;
    if (sd) { sd->dd = dd; sd->dndu = dgs.dndu; sd->dndv = dgs.dndv; }
  } else
    make_bsdf<LOBES>(sv.materials[tv.material], dg, dgs, tv.flags, &sf->bsdf);
  sf->light = tv.light;
  sf->material = tv.material;
}
// Intersection::Le (core/primitive.cpp:142-145) with AreaLight::L (light.h:93-96)
PB_HD Spec surface_le(const SceneView &sv, const Surface &sf, const V3 &w) {
  if (sf.light < 0) return spec(0.f);
  const PbrtB200Light &l = sv.lights[sf.light];
  return dot(sf.bsdf.ng, w) > 0 ? spec3(l.color[0], l.color[1], l.color[2]) : spec(0.f);
}
// `lightIsect.primitive->GetAreaLight() == light ? lightIsect.Le(-wi)` is non-black
// (transport.cpp:180-182): evaluated by the MIS-probe traversal kernel.
PB_HD bool hit_emits_from_light(const SceneView &sv, const Ray &ray, const Hit &hit, int light) {
  // the probe left the scene: `Li = light->Le(ray)` (transport.cpp:184-185), black except for an infinite light
  if (hit.prim < 0) return sv.nInfiniteLights != 0 && sv.lights[light].kind == PBRT_B200_LIGHT_INFINITE;
  if (sv.nInfiniteLights != 0 && sv.lights[light].kind == PBRT_B200_LIGHT_INFINITE) return false;
  TriVerts tv = load_prim(sv, (uint32_t)hit.prim);
  if (tv.light != light) return false;
  DiffGeom dg;
  surface_dg(sv, ray, hit, tv, &dg);
  return dot(dg.nn, -ray.d) > 0;
}

// ---------------------------------------------------------------- shape sampling
// Triangle::Sample, shapes/trianglemesh.cpp:336-349 (+ UniformSampleTriangle mc.cpp:136-141)
// `frame`: the light-shape entry's precomputed pair (SceneView::lightFrame) or null
PB_HD V3 triangle_sample(const TriVerts &tv, float u1, float u2, V3 *Ns, const float4 *frame = (const float4 *)0) {
  float su1 = sqrtf(u1);
  float b1 = 1.f - su1;
  float b2 = u2 * su1;
  V3 p = b1 * tv.p1 + b2 * tv.p2 + (1.f - b1 - b2) * tv.p3;
#ifdef __CUDA_ARCH__
  if (frame) {
    const float4 f = __ldg(frame + 1);
    *Ns = v3(f.x, f.y, f.z);
    return p;
  }
#endif
  V3 n = cross(tv.p2 - tv.p1, tv.p3 - tv.p1);
  *Ns = normalize(n);
  if (tv.flags & PBRT_B200_PF_REVERSE) *Ns = *Ns * -1.f;
  return p;
}
// core/mc.cpp:74-81
PB_HD V3 uniform_sample_sphere(float u1, float u2) {
  float z = 1.f - 2.f * u1;
  float r = sqrtf(pb_max(0.f, 1.f - z * z));
  float phi = 2.f * PB_PI * u2;
  float x = r * pb_cosf(phi);
  float y = r * pb_sinf(phi);
  return v3(x, y, z);
}
// core/mc.cpp:153-161
PB_HD V3 uniform_sample_cone(float u1, float u2, float costhetamax, const V3 &x, const V3 &y,
                             const V3 &z) {
  float costheta = pb_lerp(u1, costhetamax, 1.f);
  float sintheta = sqrtf(1.f - costheta * costheta);
  float phi = u2 * 2.f * PB_PI;
  return pb_cosf(phi) * sintheta * x + pb_sinf(phi) * sintheta * y + costheta * z;
}
// Sphere::Sample(u1,u2,ns), sphere.cpp:38-44
PB_HD V3 sphere_sample_area(const PbrtB200Quadric &s, float u1, float u2, V3 *ns) {
  V3 p = v3(0, 0, 0) + s.radius * uniform_sample_sphere(u1, u2);
  *ns = normalize(xf_normal(s.w2o, p));
  if (s.flags & PBRT_B200_PF_REVERSE) *ns = *ns * -1.f;
  return xf_point(s.o2w, p);
}
// Cylinder::Sample(u1,u2,ns), cylinder.cpp:37-44
PB_NOINLINE V3 cylinder_sample_area(const PbrtB200Quadric &s, float u1, float u2, V3 *ns) {
  float z = pb_lerp(u1, s.zmin, s.zmax);
  float t = u2 * s.phi_max;
  V3 p = v3(s.radius * pb_cosf(t), s.radius * pb_sinf(t), z);
  *ns = normalize(xf_normal(s.w2o, v3(p.x, p.y, 0.f)));
  if (s.flags & PBRT_B200_PF_REVERSE) *ns = *ns * -1.f;
  return xf_point(s.o2w, p);
}
// Disk::Sample(u1,u2,ns), disk.cpp:37-46
PB_NOINLINE V3 disk_sample_area(const PbrtB200Quadric &s, float u1, float u2, V3 *ns) {
  V3 p;
  concentric_sample_disk(u1, u2, &p.x, &p.y);
  p.x *= s.radius;
  p.y *= s.radius;
  p.z = s.height;
  *ns = normalize(xf_normal(s.w2o, v3(0.f, 0.f, 1.f)));
  if (s.flags & PBRT_B200_PF_REVERSE) *ns = *ns * -1.f;
  return xf_point(s.o2w, p);
}
// Sphere::Sample(p,u1,u2,ns), sphere.cpp:45-69
PB_NOINLINE V3 sphere_sample(const PbrtB200Quadric &s, const V3 &p, float u1, float u2, V3 *ns) {
  V3 Pcenter = xf_point(s.o2w, v3(0, 0, 0));
  V3 wc = normalize(Pcenter - p);
  V3 wcX, wcY;
  coordinate_system(wc, &wcX, &wcY);
  if (distance_sq(p, Pcenter) - s.radius * s.radius < 1e-4f) return sphere_sample_area(s, u1, u2, ns);
  float cosThetaMax = sqrtf(pb_max(0.f, 1.f - s.radius * s.radius / distance_sq(p, Pcenter)));
  Ray r;
  r.o = p;
  r.d = uniform_sample_cone(u1, u2, cosThetaMax, wcX, wcY, wc);
  r.mint = PB_RAY_EPSILON; r.maxt = PB_INFINITY;
  float thit, phi; V3 phit;
  if (!sphere_test(s, r, &thit, &phit, &phi, false)) thit = dot(Pcenter - p, normalize(r.d));
  V3 ps = ray_at(r, thit);
  *ns = normalize(ps - Pcenter);
  if (s.flags & PBRT_B200_PF_REVERSE) *ns = *ns * -1.f;
  return ps;
}

// Shape::Pdf(p, wi) (core/shape.h:96-107) over the light's own shape list:
// a ShapeSet tests every member without shrinking maxt, the last hit wins
// (core/shape.h:150-156).
// LK (light kinds the scene has; kernels for the common all-triangle-area-lights scenes compile the rest
// out): bit 0 = delta lights (point / spot / distant), bit 1 = area lights over quadrics.
#define PB_LK_DELTA 1
#define PB_LK_QUADRIC 2
#define PB_LK_INFINITE 4
#define PB_LK_ALL 7
template <int LK = PB_LK_ALL>
PB_NOINLINE float light_shapes_pdf(const SceneView &sv, const PbrtB200Light &l, const V3 &p, const V3 &wi) {
  Ray ray;
  ray.o = p; ray.d = wi; ray.mint = PB_RAY_EPSILON; ray.maxt = PB_INFINITY;
  bool anyHit = false;
  Hit hit; hit.prim = -1; hit.t = 0.f; hit.b1 = hit.b2 = 0.f; hit.inst = -1;
  TriVerts htv;
  uint32_t hitShape = 0;
  for (uint32_t i = 0; i < l.n_shapes; ++i) {
    uint32_t prim = sv.lightShapes[l.first_shape + i];
    TriVerts tv = load_prim(sv, prim);
    float t, b1 = 0.f, b2 = 0.f;
    bool ok;
    if (!(LK & PB_LK_QUADRIC) || (tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE)
      ok = tri_test(tv.p1, tv.p2, tv.p3, ray, &t, &b1, &b2);
    else {
      V3 phit; float phi;
      ok = quadric_test(sv.quadrics[sv.aux[prim].owner], ray, &t, &phit, &phi, false);
    }
    if (ok) { anyHit = true; hit.prim = (int32_t)prim; hit.t = t; hit.b1 = b1; hit.b2 = b2; htv = tv; hitShape = l.first_shape + i; }
  }
  if (!anyHit) return 0.f;
  DiffGeom dg;
#ifdef __CUDA_ARCH__
  if (sv.lightFrame && (htv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE) {
    const float4 f = __ldg(sv.lightFrame + 2u * hitShape);   // dg.nn of this triangle: hit independent (k_light_frames)
    dg.nn = v3(f.x, f.y, f.z);
  } else
#endif
  if (LK & PB_LK_QUADRIC) surface_dg(sv, ray, hit, htv, &dg);
  else tri_dg(sv, (uint32_t)hit.prim, htv, ray, hit.t, hit.b1, hit.b2, &dg);
  float ad = absdot(dg.nn, -wi);
  float pdf = distance_sq(p, ray_at(ray, hit.t)) / (ad * l.area);
  if (ad == 0.f) pdf = 0.f;
  return pdf;
}
// Light::Pdf(p, n, wi) for an area light (area.cpp:69-72): the single-sphere
// light overrides Shape::Pdf (sphere.cpp:70-79).
template <int LK = PB_LK_ALL>
PB_HD float area_light_pdf(const SceneView &sv, const PbrtB200Light &l, const V3 &p, const V3 &wi) {
  if ((LK & PB_LK_QUADRIC) && l.n_shapes == 1) {
    uint32_t prim = sv.lightShapes[l.first_shape];
    TriVerts tv = load_prim(sv, prim);
    if ((tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_QUADRIC &&
        sv.quadrics[sv.aux[prim].owner].kind == PBRT_B200_QUADRIC_SPHERE) {
      const PbrtB200Quadric &s = sv.quadrics[sv.aux[prim].owner];
      V3 Pcenter = xf_point(s.o2w, v3(0, 0, 0));
      if (distance_sq(p, Pcenter) - s.radius * s.radius < 1e-4f) return light_shapes_pdf<LK>(sv, l, p, wi);
      float cosThetaMax = sqrtf(pb_max(0.f, 1.f - s.radius * s.radius / distance_sq(p, Pcenter)));
      return 1.f / (2.f * PB_PI * (1.f - cosThetaMax));   // UniformConePdf, mc.cpp:142-144
    }
  }
  return light_shapes_pdf<LK>(sv, l, p, wi);
}
// shape->Sample(p, u1, u2, &ns) for the light's shape (area.cpp:63).  A ShapeSet
// draws one RandomFloat to pick the member by area (core/shape.h:115-121).
template <int LK = PB_LK_ALL>
PB_NOINLINE V3 area_light_sample_point(const SceneView &sv, const PbrtB200Light &l, const V3 &p, float u1,
                                 float u2, RngStream &rng, V3 *ns) {
  uint32_t sn = 0;
  if (l.n_shapes > 1) {
    float ls = rng.next_float();
    for (sn = 0; sn < l.n_shapes - 1; ++sn)
      if (ls < sv.lightCdf[l.first_shape + sn]) break;
  }
  uint32_t prim = sv.lightShapes[l.first_shape + sn];
  TriVerts tv = load_prim(sv, prim);
  if (!(LK & PB_LK_QUADRIC) || (tv.flags & PBRT_B200_PF_KIND_MASK) == PBRT_B200_PRIM_TRIANGLE)
    return triangle_sample(tv, u1, u2, ns, sv.lightFrame ? sv.lightFrame + 2u * (l.first_shape + sn) : (const float4 *)0);
  const PbrtB200Quadric &s = sv.quadrics[sv.aux[prim].owner];
  if (s.kind == PBRT_B200_QUADRIC_CYLINDER) return cylinder_sample_area(s, u1, u2, ns);   // Shape::Sample(P,..) default
  if (s.kind == PBRT_B200_QUADRIC_DISK) return disk_sample_area(s, u1, u2, ns);
  if (l.n_shapes > 1) return sphere_sample_area(s, u1, u2, ns);   // Shape::Sample(P,..) default
  return sphere_sample(s, p, u1, u2, ns);
}

// ---------------------------------------------------------------- direct lighting
// One EstimateDirect() call, split at its two ray queries.
struct DirectRec {
  Spec termA;        // f * Li * AbsDot(wi,n) [* weight] / lightPdf   -- needs the shadow ray unoccluded
  Spec termB;        // f * Li * AbsDot(wi,n) * weight / bsdfPdf       -- needs the probe to reach the light
  bool hasShadow, hasProbe;
  Ray shadow;        // VisibilityTester segment: [1e-3, 1 - 1e-3] along p2 - p1
  Ray probe;         // RayDifferential(p, wi): [1e-3, inf)
  int32_t probeLight;
};
PB_HD Spec light_color(const PbrtB200Light &l) { return spec3(l.color[0], l.color[1], l.color[2]); }

PB_HD void set_segment(Ray *r, const V3 &p1, const V3 &p2) {
  r->o = p1; r->d = p2 - p1; r->mint = PB_RAY_EPSILON; r->maxt = 1.f - PB_RAY_EPSILON;
}

// Light::Sample_L(p, &wi, &visibility) of the delta lights: point (lights/point.cpp:55-66), spot
// (lights/spot.cpp:59-76) and distant (lights/distant.cpp:60-65, whose visibility ray is unbounded).
// GonioPhotometricLight::Scale (goniometric.cpp:39-46) and ProjectionLight::Projection (projection.cpp:100-113):
// the factor on a point light's intensity towards world direction w (MIPMap::Lookup(s, t), filter width 0)
PB_NOINLINE Spec mapped_light_scale(const SceneView &sv, const PbrtB200Light &l, const V3 &w) {
  V3 wl = v3(l.w2l[0] * w.x + l.w2l[1] * w.y + l.w2l[2] * w.z, l.w2l[3] * w.x + l.w2l[4] * w.y + l.w2l[5] * w.z,
             l.w2l[6] * w.x + l.w2l[7] * w.y + l.w2l[8] * w.z);
  float s, t;
  if (l.kind == PBRT_B200_LIGHT_GONIOMETRIC) {
    V3 wp = normalize(wl);
    float tmp = wp.y; wp.y = wp.z; wp.z = tmp;
    float theta = pb_acosf(pb_clamp(wp.z, -1.f, 1.f));
    float ph = pb_atan2f(wp.y, wp.x);
    float phi = (ph < 0.f) ? ph + 2.f * PB_PI : ph;
    s = phi * PB_INV_TWOPI; t = theta * PB_INV_PI;
    if (l.radiance_map == 0) return spec(1.f);
  } else {
    if (wl.z < PB_RAY_EPSILON) return spec(0.f);   // hither
    V3 Pl = xf_point(l.projection, wl);
    if (Pl.x < l.screen[0] || Pl.x > l.screen[1] || Pl.y < l.screen[2] || Pl.y > l.screen[3]) return spec(0.f);
    if (l.radiance_map == 0) return spec(1.f);
    s = (Pl.x - l.screen[0]) / (l.screen[1] - l.screen[0]);
    t = (Pl.y - l.screen[2]) / (l.screen[3] - l.screen[2]);
  }
  return mip_lookup_width(sv, sv.textures[l.radiance_map - 1], s, t, 0.f);
}
PB_HD Spec delta_light_sample(const SceneView &sv, const PbrtB200Light &light, const V3 &p, V3 *wi, Ray *shadow) {
  if (light.kind == PBRT_B200_LIGHT_DISTANT) {
    *wi = v3(light.pos[0], light.pos[1], light.pos[2]);
    shadow->o = p; shadow->d = *wi; shadow->mint = PB_RAY_EPSILON; shadow->maxt = PB_INFINITY;   // SetRay
    return light_color(light);
  }
  V3 lightPos = v3(light.pos[0], light.pos[1], light.pos[2]);
  *wi = normalize(lightPos - p);
  set_segment(shadow, p, lightPos);
  if (light.kind == PBRT_B200_LIGHT_SPOT) {
    // SpotLight::Falloff(-wi), spot.cpp:66-76
    V3 w = -*wi;
    V3 wl = normalize(v3(light.w2l[0] * w.x + light.w2l[1] * w.y + light.w2l[2] * w.z,
                         light.w2l[3] * w.x + light.w2l[4] * w.y + light.w2l[5] * w.z,
                         light.w2l[6] * w.x + light.w2l[7] * w.y + light.w2l[8] * w.z));
    float costheta = wl.z;
    float falloff;
    if (costheta < light.cos_total_width) falloff = 0.f;
    else if (costheta > light.cos_falloff_start) falloff = 1.f;
    else {
      float delta = (costheta - light.cos_total_width) / (light.cos_falloff_start - light.cos_total_width);
      falloff = delta * delta * delta * delta;
    }
    return light_color(light) * falloff / distance_sq(lightPos, p);
  }
  if (light.kind >= PBRT_B200_LIGHT_GONIOMETRIC)   // Intensity * Scale(-wi) / DistanceSquared(lightPos, p)
    return light_color(light) * mapped_light_scale(sv, light, -*wi) / distance_sq(lightPos, p);
  return light_color(light) / distance_sq(lightPos, p);
}

// InfiniteAreaLight::Sample_L(p, n, u1, u2, ..), lights/infinite.cpp:88-109: cosine-weighted direction
// about n, the hemisphere chosen by one RandomFloat; visibility is an unbounded ray.
// InfiniteAreaLight::Le, lights/infinite.cpp:75-87: Lbase, times the radiance map looked up (MIPMap::Lookup(s, t),
// i.e. filter width 0) at the direction's spherical coordinates in light space
PB_NOINLINE Spec infinite_le(const SceneView &sv, const PbrtB200Light &l, const V3 &w) {
  Spec L = spec3(l.color[0], l.color[1], l.color[2]);
  if (l.radiance_map != 0) {
    V3 wl = v3(l.w2l[0] * w.x + l.w2l[1] * w.y + l.w2l[2] * w.z, l.w2l[3] * w.x + l.w2l[4] * w.y + l.w2l[5] * w.z,
               l.w2l[6] * w.x + l.w2l[7] * w.y + l.w2l[8] * w.z);
    V3 wh = normalize(wl);
    float ph = pb_atan2f(wh.y, wh.x);
    float phi = (ph < 0.f) ? ph + 2.f * PB_PI : ph;
    float s = phi * PB_INV_TWOPI;
    float t = pb_acosf(pb_clamp(wh.z, -1.f, 1.f)) * PB_INV_PI;
    L = L * mip_lookup_width(sv, sv.textures[l.radiance_map - 1], s, t, 0.f);
  }
  return L;
}
// ---- InfiniteAreaLightIS (lights/infinitesample.cpp): the "infinite" light with piecewise-constant
// importance sampling of its radiance map.  Tables: see PbrtB200Light::dist_offset.
struct InfDist {
  const float *uFunc, *uCdf, *vFuncInt, *vBase;
  float uFuncInt;
  int nu, nv;
  PB_HD const float *vFunc(int u) const { return vBase + (size_t)u * (2 * nv + 1); }
  PB_HD const float *vCdf(int u) const { return vBase + (size_t)u * (2 * nv + 1) + nv; }
};
PB_HD InfDist inf_dist(const SceneView &sv, const PbrtB200Light &l) {
  InfDist d;
  d.nu = (int)l.dist_nu; d.nv = (int)l.dist_nv;
  const float *b = sv.lightDist + l.dist_offset;
  d.uFunc = b; d.uCdf = b + d.nu; d.uFuncInt = b[2 * d.nu + 1];
  d.vFuncInt = b + 2 * d.nu + 2;
  d.vBase = b + 3 * d.nu + 2;
  return d;
}
// Distribution1D::Sample, infinitesample.cpp:43-52 (std::upper_bound: first cdf entry greater than u)
PB_HD float dist1d_sample(const float *func, const float *cdf, int count, float funcInt, float u, float *pdf) {
  int lo = 0, n = count + 1;
  while (n > 0) {
    int half = n >> 1;
    if (!(u < cdf[lo + half])) { lo += half + 1; n -= half + 1; }
    else n = half;
  }
  int offset = pb_imax(0, lo - 1);
  u = (u - cdf[offset]) / (cdf[offset + 1] - cdf[offset]);
  *pdf = func[offset] * (1.f / funcInt);
  return offset + u;
}
// InfiniteAreaLightIS::Sample_L(p, u1, u2, wi, pdf, vis), infinitesample.cpp:149-175
PB_NOINLINE Spec infinite_light_sample_is(const SceneView &sv, const PbrtB200Light &light, const V3 &p, float u1, float u2,
                                          V3 *wi, float *pdf, Ray *shadow) {
  InfDist d = inf_dist(sv, light);
  float pdfs[2];
  float fu = dist1d_sample(d.uFunc, d.uCdf, d.nu, d.uFuncInt, u1, &pdfs[0]);
  int u = pb_imin(pb_imax(pb_f2i(fu), 0), d.nu - 1);
  float fv = dist1d_sample(d.vFunc(u), d.vCdf(u), d.nv, d.vFuncInt[u], u2, &pdfs[1]);
  *wi = v3(0, 0, 0);
  if (pdfs[0] == 0.f || pdfs[1] == 0.f) { *pdf = 0.f; return spec(0.f); }
  float invNu = 1.f / d.nu, invNv = 1.f / d.nv;
  float theta = fv * invNv * PB_PI;
  float phi = fu * invNu * 2.f * PB_PI;
  float costheta = pb_cosf(theta), sintheta = pb_sinf(theta);
  if (sintheta == 0.f) { *pdf = 0.f; return spec(0.f); }
  float sinphi = pb_sinf(phi), cosphi = pb_cosf(phi);
  // LightToWorld(Vector): the inverse of the w2l rows held in the record is not stored; l2w travels in `projection`
  V3 wl = v3(sintheta * cosphi, sintheta * sinphi, costheta);
  *wi = xf_vector(light.projection, wl);
  *pdf = (float)((double)(pdfs[0] * pdfs[1]) / (2. * (double)PB_PI * (double)PB_PI * (double)sintheta));
  shadow->o = p; shadow->d = *wi; shadow->mint = PB_RAY_EPSILON; shadow->maxt = PB_INFINITY;   // SetRay
  return light_color(light) * mip_lookup_width(sv, sv.textures[light.radiance_map - 1], fu * invNu, fv * invNv, 0.f);
}
// InfiniteAreaLightIS::Pdf(p, w), infinitesample.cpp:176-187
PB_NOINLINE float infinite_light_pdf_is(const SceneView &sv, const PbrtB200Light &l, const V3 &w) {
  InfDist d = inf_dist(sv, l);
  V3 wi = v3(l.w2l[0] * w.x + l.w2l[1] * w.y + l.w2l[2] * w.z, l.w2l[3] * w.x + l.w2l[4] * w.y + l.w2l[5] * w.z,
             l.w2l[6] * w.x + l.w2l[7] * w.y + l.w2l[8] * w.z);
  float theta = pb_acosf(pb_clamp(wi.z, -1.f, 1.f));
  float ph = pb_atan2f(wi.y, wi.x);
  float phi = (ph < 0.f) ? ph + 2.f * PB_PI : ph;
  int u = pb_imin(pb_imax(pb_f2i(phi * PB_INV_TWOPI * d.nu), 0), d.nu - 1);
  int v = pb_imin(pb_imax(pb_f2i(theta * PB_INV_PI * d.nv), 0), d.nv - 1);
  return (d.uFunc[u] * d.vFunc(u)[v]) / (d.uFuncInt * d.vFuncInt[u]) * 1.f / (2.f * PB_PI * PB_PI * pb_sinf(theta));
}
PB_NOINLINE Spec infinite_light_sample_cosine(const SceneView &sv, const PbrtB200Light &light, const V3 &p, const V3 &n, float u1, float u2,
                                              RngStream &rng, V3 *wi, float *pdf, Ray *shadow) {
  float x, y, z;
  concentric_sample_disk(u1, u2, &x, &y);
  z = sqrtf(pb_max(0.f, 1.f - x * x - y * y));
  if (rng.next_float() < .5) z *= -1;
  *pdf = fabsf(z) * PB_INV_TWOPI;
  V3 v1, v2;
  coordinate_system(normalize(n), &v1, &v2);
  *wi = v3(v1.x * x + v2.x * y + n.x * z, v1.y * x + v2.y * y + n.y * z, v1.z * x + v2.z * y + n.z * z);
  shadow->o = p; shadow->d = *wi; shadow->mint = PB_RAY_EPSILON; shadow->maxt = PB_INFINITY;   // SetRay
  return infinite_le(sv, light, *wi);   // Le(RayDifferential(p, *wi))
}
// Sum of lights[i]->Le(ray) over the scene's lights for a ray that left the scene (black for all but
// infinite lights; adding a black spectrum changes nothing).
PB_HD Spec escaped_le(const SceneView &sv, const V3 &d) {
  Spec L = spec(0.f);
  for (uint32_t i = 0; i < sv.nLights; ++i)
    if (sv.lights[i].kind == PBRT_B200_LIGHT_INFINITE) L = L + infinite_le(sv, sv.lights[i], d);
  return L;
}

// transport.cpp:123-194.  ls1..bcs are already chosen by the caller
// (Sample arrays or RandomFloat, :131-146).
template <bool G = true, int LK = PB_LK_ALL, class BS = Bsdf>
PB_HD void estimate_direct(const SceneView &sv, int lightNum, const V3 &p, const V3 &n, const V3 &wo,
                           const BS &bsdf, float ls1, float ls2, float bs1, float bs2, float bcs,
                           RngStream &rng, DirectRec *rec) {
  const PbrtB200Light &light = sv.lights[lightNum];
  rec->hasShadow = rec->hasProbe = false;
  rec->termA = rec->termB = spec(0.f);
  rec->probeLight = lightNum;
  V3 wi;
  float lightPdf, bsdfPdf;
  Spec Li;
  const bool isInfinite = (LK & PB_LK_INFINITE) && light.kind == PBRT_B200_LIGHT_INFINITE;
  bool isDelta = (LK & PB_LK_DELTA) && light.kind != PBRT_B200_LIGHT_AREA && !isInfinite;
  if (isDelta) {
    lightPdf = 1.f;
    Li = delta_light_sample(sv, light, p, &wi, &rec->shadow);
  } else if (isInfinite) {
    if (light.dist_nu != 0) Li = infinite_light_sample_is(sv, light, p, ls1, ls2, &wi, &lightPdf, &rec->shadow);
    else Li = infinite_light_sample_cosine(sv, light, p, n, ls1, ls2, rng, &wi, &lightPdf, &rec->shadow);
  } else {
    V3 ns;
    V3 ps = area_light_sample_point<LK>(sv, light, p, ls1, ls2, rng, &ns);
    wi = normalize(ps - p);
    lightPdf = area_light_pdf<LK>(sv, light, p, wi);
    set_segment(&rec->shadow, p, ps);
    Li = dot(ns, -wi) > 0 ? light_color(light) : spec(0.f);
  }
  if (lightPdf > 0.f && !is_black(Li)) {
    Spec f = bsdf_f<G>(bsdf, wo, wi);
    if (!is_black(f)) {
      rec->hasShadow = true;
      // Li *= Transmittance (== 1 without volumes, integrators/emission.cpp:51)
      Li = Li * spec(1.f);
      if (isDelta)
        rec->termA = f * Li * absdot(wi, n) / lightPdf;
      else {
        bsdfPdf = bsdf_pdf<G>(bsdf, wo, wi);
        float weight = power_heuristic(1, lightPdf, 1, bsdfPdf);
        rec->termA = f * Li * absdot(wi, n) * weight / lightPdf;
      }
    }
  }
  if (!isDelta) {
    int sampled;
    Spec f = bsdf_sample_f<G>(bsdf, wo, &wi, bs1, bs2, bcs, &bsdfPdf, PB_ALL & ~PB_SPECULAR, &sampled);
    if (!is_black(f) && bsdfPdf > 0.f) {
      lightPdf = isInfinite ? (light.dist_nu != 0 ? infinite_light_pdf_is(sv, light, wi)
                                                   : absdot(n, wi) * PB_INV_TWOPI)   // InfiniteAreaLight::Pdf(p, n, wi), infinite.cpp:110-113
                            : area_light_pdf<LK>(sv, light, p, wi);
      if (lightPdf > 0.f) {
        float weight = power_heuristic(1, bsdfPdf, 1, lightPdf);
        rec->hasProbe = true;
        rec->probe.o = p; rec->probe.d = wi;
        rec->probe.mint = PB_RAY_EPSILON; rec->probe.maxt = PB_INFINITY;
        // the probe left the scene: light->Le(ray) (transport.cpp:184-185); constant for an area light
        Spec Le = (isInfinite ? infinite_le(sv, light, wi) : light_color(light)) * spec(1.f);   // Li *= scene->Transmittance(ray)
        rec->termB = f * Le * absdot(wi, n) * weight / bsdfPdf;
      }
    }
  }
}
// Ld of one EstimateDirect call once its rays are resolved
PB_HD Spec direct_value(const Spec &termA, const Spec &termB, bool unoccluded, bool probeHit) {
  Spec Ld = spec(0.f);
  if (unoccluded) Ld = Ld + termA;
  if (probeHit) Ld = Ld + termB;
  return Ld;
}

// ---------------------------------------------------------------- sample patterns
// Index helpers into SceneView::patterns, which the host library lays out as
// [1D patterns in Add1D order][2D patterns in Add2D order] (core/sampling.cpp:41-70;
// request order: path.cpp:47-57, directlighting.cpp:39-66, emission.cpp:42-46).
PB_HD float pat1d(const SceneView &sv, const PatternSource &key0, int idx, int j) {
  const SamplePattern &p = sv.patterns[idx];
  if (key0.ld && p.offset >= 0) return key0.ld[(size_t)(p.offset + j) * key0.stride + key0.index];
  if (key0.bc == 2) return pb_f32(key0.key0, p.ctr0 + 1u + (uint32_t)j);   // samplers/random.cpp:103-105
  if (key0.bc) return pb_bc_stratified_value(key0.key0, p, j, 0);
  return pb_lhs_value(key0.key0, p, j, 0);
}
PB_HD float pat2d(const SceneView &sv, const PatternSource &key0, int idx, int j, int d) {
  const SamplePattern &p = sv.patterns[sv.n1D + idx];
  if (key0.ld && p.offset >= 0) return key0.ld[(size_t)(p.offset + 2 * j + d) * key0.stride + key0.index];
  if (key0.bc == 2) return pb_f32(key0.key0, p.ctr0 + 1u + 2u * (uint32_t)j + (uint32_t)d);   // random.cpp:106-108
  if (key0.bc) return pb_bc_stratified_value(key0.key0, p, j, d);
  return pb_lhs_value(key0.key0, p, j, d);
}

// ---------------------------------------------------------------- path integrator
struct PathVertexOut {
  DirectRec direct;
  bool hasDirect;      // a direct-lighting record was produced (always, when the ray hit)
  bool hasNext;        // path continues along `next`
  Ray next;
};
struct PathState {
  Spec L, T;           // radiance so far, pathThroughput
  Spec pendT;          // throughput the pending direct term is weighted by
  float alpha;
  int32_t depth;       // pathLength of the vertex about to be shaded
  bool specularBounce;
  RngStream rng;       // stream 1
};
// One iteration of the `for (pathLength = 0; ; ++pathLength)` loop body,
// path.cpp:65-143, after scene->Intersect returned `hit`, in two halves so a
// caller can hand the direct-lighting rays on before the continuation is sampled.
//
// First half (path.cpp:68-112): miss handling, alpha, emitted light, BSDF, one
// UniformSampleOneLight record.  Returns false when the ray missed (path ends).
template <bool LOBES = true, bool INST = true, bool TEX = false, bool FLATONLY = false>
PB_HD bool path_vertex_surface(const SceneView &sv, PathState &st, const Ray &ray, const Hit &hit,
                               Surface *sf, const RayDiff *rd = (const RayDiff *)0,
                               PbrtB200Material *evalMat = (PbrtB200Material *)0) {
  int pathLength = st.depth;
  if (hit.prim < 0) {
    // path.cpp:69-82.  lights[i]->Le(ray) is black for point and area lights (core/light.cpp:41-43)
    if (pathLength == 0) st.alpha = 0.f;   // L != 0 ? 1 : 0 with L == 0
    // (TEX = false instances serve scenes without any texture node, hence without radiance maps: Le is Lbase,
    // and the map lookup's call stays out of those kernels)
    if (sv.nInfiniteLights != 0) {
      if (pathLength == 0) {
        for (uint32_t i = 0; i < sv.nLights; ++i)
          if (sv.lights[i].kind == PBRT_B200_LIGHT_INFINITE)
            st.L = st.L + (TEX ? infinite_le(sv, sv.lights[i], ray.d) : light_color(sv.lights[i]));
        st.alpha = is_black(st.L) ? 0.f : 1.f;
      } else if (st.specularBounce) {
        for (uint32_t i = 0; i < sv.nLights; ++i)
          if (sv.lights[i].kind == PBRT_B200_LIGHT_INFINITE)
            st.L = st.L + st.T * (TEX ? infinite_le(sv, sv.lights[i], ray.d) : light_color(sv.lights[i]));
      }
    }
    return false;
  }
  if (pathLength == 0) st.alpha = 1.f;
  else st.T = st.T * spec(1.f);            // pathThroughput *= scene->Transmittance(ray)
  surface_at_hit<LOBES, INST, TEX, FLATONLY>(sv, ray, hit, sf, rd, evalMat);
  if (pathLength == 0 || st.specularBounce) st.L = st.L + st.T * surface_le(sv, *sf, -ray.d);
  return true;
}
// path.cpp:93-112: one UniformSampleOneLight record at the vertex
template <bool G = true, int LK = PB_LK_ALL, class SF = Surface>
PB_HD void path_vertex_direct(const SceneView &sv, const PatternSource &key0, PathState &st, const Ray &ray,
                              const SF *sf, DirectRec *direct, bool *hasDirect) {
  *hasDirect = false;
  int pathLength = st.depth;
  const auto &bsdf = sf->bsdf;
  V3 p = bsdf.p, n = bsdf.nn;
  V3 wo = -ray.d;
  int nLights = (int)sv.nLights;
  if (nLights > 0) {
    // UniformSampleOneLight, transport.cpp:51-70
    float ln, ls1, ls2, bs1, bs2, bcs;
    if (pathLength < 3) {
      ln = pat1d(sv, key0, 3 * pathLength, 0);
      ls1 = pat2d(sv, key0, 3 * pathLength, 0, 0);
      ls2 = pat2d(sv, key0, 3 * pathLength, 0, 1);
      bs1 = pat2d(sv, key0, 3 * pathLength + 1, 0, 0);
      bs2 = pat2d(sv, key0, 3 * pathLength + 1, 0, 1);
      bcs = pat1d(sv, key0, 3 * pathLength + 1, 0);
    } else {
      ln = st.rng.next_float();
      ls1 = st.rng.next_float(); ls2 = st.rng.next_float();
      bs1 = st.rng.next_float(); bs2 = st.rng.next_float();
      bcs = st.rng.next_float();
    }
    int lightNum = pb_imin(pb_floor2int(ln * nLights), nLights - 1);
    estimate_direct<G, LK>(sv, lightNum, p, n, wo, bsdf, ls1, ls2, bs1, bs2, bcs, st.rng, direct);
    *hasDirect = true;
    st.pendT = st.T;
  }
  // NOTE: with no lights the reference dereferences lights[-1]; such scenes render black here.
}
// The RandomFloat() draws of path_vertex_direct at a vertex whose BSDF has no non-specular component, without
// the estimate: there bsdf_f is black for every wi and Sample_f(~BSDF_SPECULAR) matches nothing, so
// UniformSampleOneLight traces no ray and adds nothing (transport.cpp:51-70,153-155,171-174) -- but past
// depth 3 it still drew its six sample values (transport.cpp:141-146), and the chosen light's Sample_L may
// draw one more (ShapeSet::Sample, shape.h:115-121; InfiniteAreaLight::Sample_L, infinite.cpp:96).
PB_HD void path_vertex_direct_skip(const SceneView &sv, const PatternSource &key0, PathState &st) {
  int nLights = (int)sv.nLights;
  if (nLights <= 0) return;
  float ln;
  if (st.depth < 3) ln = pat1d(sv, key0, 3 * st.depth, 0);
  else { ln = st.rng.next_float(); st.rng.skip(5); }
  int lightNum = pb_imin(pb_floor2int(ln * nLights), nLights - 1);
  const PbrtB200Light &light = sv.lights[lightNum];
  if (light.kind == PBRT_B200_LIGHT_AREA) { if (light.n_shapes > 1) st.rng.skip(1); }
  else if (light.kind == PBRT_B200_LIGHT_INFINITE && light.dist_nu == 0) st.rng.skip(1);
}
template <bool TEX = false>
PB_HD bool path_vertex_begin(const SceneView &sv, const PatternSource &key0, PathState &st, const Ray &ray,
                             const Hit &hit, Surface *sf, DirectRec *direct, bool *hasDirect,
                             const RayDiff *rd = (const RayDiff *)0) {
  *hasDirect = false;
  if (!path_vertex_surface<true, true, TEX>(sv, st, ray, hit, sf, rd)) return false;
  path_vertex_direct(sv, key0, st, ray, sf, direct, hasDirect);
  return true;
}
// Second half (path.cpp:113-143): sample the BSDF for the next direction, update the
// throughput, Russian roulette, depth limit.  Returns true when the path continues along *next.
template <bool G = true, class SF = Surface>
PB_HD bool path_vertex_continue(const SceneView &sv, const PatternSource &key0, PathState &st, const Ray &ray,
                                const SF &sf, Ray *next) {
  int pathLength = st.depth;
  const auto &bsdf = sf.bsdf;
  V3 p = bsdf.p, n = bsdf.nn;
  V3 wo = -ray.d;
  float os1, os2, ocs;
  if (pathLength < 3) {
    os1 = pat2d(sv, key0, 3 * pathLength + 2, 0, 0);
    os2 = pat2d(sv, key0, 3 * pathLength + 2, 0, 1);
    ocs = pat1d(sv, key0, 3 * pathLength + 2, 0);
  } else {
    os1 = st.rng.next_float(); os2 = st.rng.next_float(); ocs = st.rng.next_float();
  }
  V3 wi;
  float pdf;
  int flags;
  Spec f = bsdf_sample_f<G>(bsdf, wo, &wi, os1, os2, ocs, &pdf, PB_ALL, &flags);
  if (is_black(f) || pdf == 0.f) return false;
  st.specularBounce = (flags & PB_SPECULAR) != 0;
  st.T = st.T * (f * absdot(wi, n) / pdf);
  next->o = p; next->d = wi; next->mint = PB_RAY_EPSILON; next->maxt = PB_INFINITY;
  if (pathLength > 3) {
    float continueProbability = .5f;
    if (st.rng.next_float() > continueProbability) return false;
    st.T = st.T * (1.f / continueProbability);   // Spectrum::operator/=(float)
  }
  if (pathLength == sv.integrator.max_depth) return false;
  st.depth = pathLength + 1;
  return true;
}
// `rd`: the differentials of `ray` (the camera ray's; path.cpp:133 continues with plain rays)
template <bool TEX = false>
PB_HD void path_vertex(const SceneView &sv, const PatternSource &key0, PathState &st, const Ray &ray,
                       const Hit &hit, PathVertexOut *out, const RayDiff *rd = (const RayDiff *)0) {
  out->hasNext = false;
  Surface sf;
  if (!path_vertex_begin<TEX>(sv, key0, st, ray, hit, &sf, &out->direct, &out->hasDirect, rd)) return;
  out->hasNext = path_vertex_continue(sv, key0, st, ray, sf, &out->next);
}
// L += pathThroughput * ((float)nLights * Ld), path.cpp:100-112 + transport.cpp:66-69
PB_HD void path_resolve_direct(const SceneView &sv, PathState &st, const Spec &termA,
                               const Spec &termB, bool unoccluded, bool probeHit) {
  Spec Ld = direct_value(termA, termB, unoccluded, probeHit);
  st.L = st.L + st.pendT * ((float)sv.nLights * Ld);
}

// ---------------------------------------------------------------- Whitted / DirectLighting
// Both integrators shade a vertex the same way -- emitted light, a sum over
// lights, then recursion into specular reflection and transmission
// (whitted.cpp:66-137, directlighting.cpp:104-183).  The recursion is run as
// an explicit depth-first walk over per-sample frames so that random numbers
// are consumed, and child radiance is folded into the parent
// (`L += Li(child) * f * AbsDot(wi,n)`), in exactly the reference's order.
struct WdFrame {
  Spec L;                 // radiance of this Li() call so far
  V3 p;                   // shading point (origin of both child rays)
  V3 wiR, wiT;            // specular reflection / transmission directions
  Spec fR, fT;            // f / pdf of each (BSDF::Sample_f convenience form, reflection.cpp:394-401)
  float aR, aT;           // AbsDot(wi, n)
  int32_t stage;          // 0: children not started, 1: reflection done, 2: complete
  bool reflValid, transValid;
  RayDiff dR, dT;         // textured scenes: the differentials of the two child rays
};
struct WdVertexOut {
  int nRecs;              // direct-lighting records written
};
// number of EstimateDirect-style records a vertex can produce
PB_HD int wd_max_records(const SceneView &sv) { return sv.directRecords; }

// Vertex body up to (not including) `if (rayDepth++ < maxDepth)`.  `recs` must hold
// wd_max_records() entries.  The frame is initialised here.
// TEX: texture operands are evaluated with `rd`, the differentials of `ray`, and the children's are derived.
template <bool TEX = false>
PB_HD void wd_vertex(const SceneView &sv, const PatternSource &key0, const Ray &ray, const Hit &hit,
                     RngStream &rng, int rayDepth, float *alpha, WdFrame *fr, DirectRec *recs,
                     WdVertexOut *out, const RayDiff *rd = (const RayDiff *)0) {
  out->nRecs = 0;
  fr->L = spec(0.f);
  fr->stage = 2;
  fr->reflValid = fr->transValid = false;
  fr->p = v3(0, 0, 0); fr->wiR = fr->wiT = v3(0, 0, 0);
  fr->fR = fr->fT = spec(0.f); fr->aR = fr->aT = 0.f;
  if (hit.prim < 0) {
    // whitted.cpp:52-59, directlighting.cpp:185-192
    if (alpha) *alpha = 0.f;
    if (sv.nInfiniteLights != 0) {
      fr->L = fr->L + escaped_le(sv, ray.d);
      if (alpha && !is_black(fr->L)) *alpha = 1.f;
    }
    return;
  }
  if (alpha) *alpha = 1.f;
  Surface sf;
  SurfaceDiff sd;
  surface_at_hit<true, true, TEX>(sv, ray, hit, &sf, rd, (PbrtB200Material *)0, &sd);
  const Bsdf &bsdf = sf.bsdf;
  V3 p = bsdf.p, n = bsdf.nn;
  V3 wo = -ray.d;
  fr->L = fr->L + surface_le(sv, sf, wo);
  int nLights = (int)sv.nLights;
  if (sv.integrator.kind == PBRT_B200_INT_WHITTED) {
    // whitted.cpp:68-76
    for (int i = 0; i < nLights; ++i) {
      const PbrtB200Light &light = sv.lights[i];
      DirectRec *rec = &recs[out->nRecs];
      rec->hasShadow = rec->hasProbe = false;
      rec->termA = rec->termB = spec(0.f);
      rec->probeLight = i;
      V3 wi; Spec Li;
      if (light.kind == PBRT_B200_LIGHT_INFINITE) {
        // InfiniteAreaLight::Sample_L(P, &wi, &vis), infinite.cpp:156-163: uniform direction from two
        // RandomFloat()s (g++ evaluates the arguments right to left: the first draw is u2), L / pdf
        float u2 = rng.next_float();
        float u1 = rng.next_float();
        if (light.dist_nu != 0) {   // InfiniteAreaLightIS::Sample_L(P, &wi, &vis), infinitesample.cpp:208-215
          float pdf;
          Spec L = infinite_light_sample_is(sv, light, p, u1, u2, &wi, &pdf, &rec->shadow);
          Li = pdf == 0.f ? spec(0.f) : L / pdf;
        } else {
        wi = uniform_sample_sphere(u1, u2);
        float pdf = 1.f / (4.f * PB_PI);   // UniformSpherePdf, mc.cpp:82-84
        rec->shadow.o = p; rec->shadow.d = wi; rec->shadow.mint = PB_RAY_EPSILON; rec->shadow.maxt = PB_INFINITY;
        Li = infinite_le(sv, light, wi) / pdf;
        }
      } else if (light.kind != PBRT_B200_LIGHT_AREA) {
        Li = delta_light_sample(sv, light, p, &wi, &rec->shadow);
      } else {
        // AreaLight::Sample_L(P, &wo, &vis), area.cpp:96-105.  g++ evaluates the two
        // RandomFloat() arguments right to left: the first draw is u2.
        float u2 = rng.next_float();
        float u1 = rng.next_float();
        V3 Ns;
        V3 Ps = area_light_sample_point(sv, light, p, u1, u2, rng, &Ns);
        wi = normalize(Ps - p);
        set_segment(&rec->shadow, p, Ps);
        float pdf = area_light_pdf(sv, light, p, wi);
        if (pdf == 0.f) Li = spec(0.f);
        else Li = (dot(Ns, -wi) > 0 ? light_color(light) : spec(0.f)) / pdf;
      }
      ++out->nRecs;
      if (is_black(Li)) continue;
      Spec f = bsdf_f(bsdf, wo, wi);
      if (!is_black(f)) {
        rec->hasShadow = true;
        rec->termA = f * Li * absdot(wi, n) * spec(1.f);
      }
    }
  } else if (nLights > 0) {
    if (sv.integrator.strategy == PBRT_B200_STRATEGY_ALL) {
      // UniformSampleAllLights, transport.cpp:31-50; patterns per light i:
      // 2D {2i: light, 2i+1: bsdf}, 1D {i: component} (directlighting.cpp:46-53)
      for (int i = 0; i < nLights; ++i) {
        int nSamples = sv.patterns[sv.n1D + 2 * i].n;
        for (int j = 0; j < nSamples; ++j) {
          float ls1 = pat2d(sv, key0, 2 * i, j, 0), ls2 = pat2d(sv, key0, 2 * i, j, 1);
          float bs1 = pat2d(sv, key0, 2 * i + 1, j, 0), bs2 = pat2d(sv, key0, 2 * i + 1, j, 1);
          float bcs = pat1d(sv, key0, i, j);
          estimate_direct(sv, i, p, n, wo, bsdf, ls1, ls2, bs1, bs2, bcs, rng, &recs[out->nRecs]);
          ++out->nRecs;
        }
      }
    } else {
      // UniformSampleOneLight; 2D {0: light, 1: bsdf}, 1D {0: lightNum, 1: component}
      // (directlighting.cpp:58-64)
      float ln = pat1d(sv, key0, 0, 0);
      int lightNum = pb_imin(pb_floor2int(ln * nLights), nLights - 1);
      float ls1 = pat2d(sv, key0, 0, 0, 0), ls2 = pat2d(sv, key0, 0, 0, 1);
      float bs1 = pat2d(sv, key0, 1, 0, 0), bs2 = pat2d(sv, key0, 1, 0, 1);
      float bcs = pat1d(sv, key0, 1, 0);
      estimate_direct(sv, lightNum, p, n, wo, bsdf, ls1, ls2, bs1, bs2, bcs, rng, &recs[0]);
      out->nRecs = 1;
    }
  }
  // `if (rayDepth++ < maxDepth)`: both specular children are fully determined by the BSDF
  // (the three RandomFloat()s each BSDF::Sample_f draws never matter: at most one lobe
  // matches REFLECTION|SPECULAR or TRANSMISSION|SPECULAR); the draws themselves are
  // accounted for by wd_advance().
  if (rayDepth < sv.integrator.max_depth) {
    fr->stage = 0;
    fr->p = p;
    float pdf; int sampled;
    V3 wi = v3(0, 0, 0);
    Spec f = bsdf_sample_f(bsdf, wo, &wi, 0.5f, 0.5f, 0.5f, &pdf, PB_REFLECTION | PB_SPECULAR, &sampled);
    if (!is_black(f) && pdf > 0.f) f = f * (1.f / pdf);
    if (!is_black(f) && absdot(wi, n) > 0.f) {
      fr->reflValid = true; fr->wiR = wi; fr->fR = f; fr->aR = absdot(wi, n);
      if (TEX) {
        // ray differential for specular reflection, whitted.cpp:86-103 / directlighting.cpp:133-150
        RayDiff &c = fr->dR;
        c.has = true;
        c.rxo = p + sd.dd.dpdx;
        c.ryo = p + sd.dd.dpdy;
        V3 dndx = sd.dndu * sd.dd.dudx + sd.dndv * sd.dd.dvdx;
        V3 dndy = sd.dndu * sd.dd.dudy + sd.dndv * sd.dd.dvdy;
        V3 dwodx = -rd->rxd - wo, dwody = -rd->ryd - wo;
        float dDNdx = dot(dwodx, n) + dot(wo, dndx);
        float dDNdy = dot(dwody, n) + dot(wo, dndy);
        c.rxd = wi - dwodx + 2.f * (dot(wo, n) * dndx + dDNdx * n);
        c.ryd = wi - dwody + 2.f * (dot(wo, n) * dndy + dDNdy * n);
      }
    }
    wi = v3(0, 0, 0);
    f = bsdf_sample_f(bsdf, wo, &wi, 0.5f, 0.5f, 0.5f, &pdf, PB_TRANSMISSION | PB_SPECULAR, &sampled);
    if (!is_black(f) && pdf > 0.f) f = f * (1.f / pdf);
    if (!is_black(f) && absdot(wi, n) > 0.f) {
      fr->transValid = true; fr->wiT = wi; fr->fT = f; fr->aT = absdot(wi, n);
      if (TEX) {
        // ray differential for specular transmission, whitted.cpp:108-131 / directlighting.cpp:156-179
        RayDiff &c = fr->dT;
        c.has = true;
        c.rxo = p + sd.dd.dpdx;
        c.ryo = p + sd.dd.dpdy;
        float eta = bsdf.eta;
        V3 w = -wo;
        if (dot(wo, n) < 0) eta = 1.f / eta;
        V3 dndx = sd.dndu * sd.dd.dudx + sd.dndv * sd.dd.dvdx;
        V3 dndy = sd.dndu * sd.dd.dudy + sd.dndv * sd.dd.dvdy;
        V3 dwodx = -rd->rxd - wo, dwody = -rd->ryd - wo;
        float dDNdx = dot(dwodx, n) + dot(wo, dndx);
        float dDNdy = dot(dwody, n) + dot(wo, dndy);
        float mu = eta * dot(w, n) - dot(wi, n);
        float dmudx = (eta - (eta * eta * dot(w, n)) / dot(wi, n)) * dDNdx;
        float dmudy = (eta - (eta * eta * dot(w, n)) / dot(wi, n)) * dDNdy;
        c.rxd = wi + eta * dwodx - (mu * dndx + dmudx * n);
        c.ryd = wi + eta * dwody - (mu * dndy + dmudy * n);
      }
    }
  }
}
// Fold the vertex's resolved direct-lighting records into its frame.
// flags[r] bit0: shadow ray unoccluded, bit1: MIS probe reached the light.
PB_HD void wd_resolve(const SceneView &sv, WdFrame *fr, const DirectRec *recs, const uint8_t *flags,
                      int nRecs) {
  if (nRecs == 0) return;
  if (sv.integrator.kind == PBRT_B200_INT_WHITTED) {
    for (int r = 0; r < nRecs; ++r)
      if (recs[r].hasShadow && (flags[r] & 1)) fr->L = fr->L + recs[r].termA;
    return;
  }
  int nLights = (int)sv.nLights;
  if (sv.integrator.strategy == PBRT_B200_STRATEGY_ALL) {
    Spec L = spec(0.f);
    int r = 0;
    for (int i = 0; i < nLights; ++i) {
      int nSamples = sv.patterns[sv.n1D + 2 * i].n;
      Spec Ld = spec(0.f);
      for (int j = 0; j < nSamples; ++j, ++r)
        Ld = Ld + direct_value(recs[r].termA, recs[r].termB, recs[r].hasShadow && (flags[r] & 1),
                               recs[r].hasProbe && (flags[r] & 2));
      L = L + Ld / (float)nSamples;
    }
    fr->L = fr->L + L;
  } else {
    Spec Ld = direct_value(recs[0].termA, recs[0].termB, recs[0].hasShadow && (flags[0] & 1),
                           recs[0].hasProbe && (flags[0] & 2));
    fr->L = fr->L + (float)nLights * Ld;
  }
}
// Depth-first continuation after a vertex's direct lighting is folded in.
// FrameIO: WdFrame load(int depth); void store(int depth, const WdFrame &).
// `depth` is the frame of the vertex just shaded (already stored).  Returns true
// with the next ray and its frame depth, or false when frame 0 is complete.
template <class FrameIO>
PB_HD bool wd_advance(FrameIO &io, RngStream &rng, int *depth, Ray *next) {
  int d = *depth;
  WdFrame fr = io.load(d);
  for (;;) {
    if (fr.stage == 0) {
      rng.skip(3);            // BSDF::Sample_f(wo,&wi,REFLECTION|SPECULAR) draws
      if (fr.reflValid) {
        io.store(d, fr);
        next->o = fr.p; next->d = fr.wiR; next->mint = PB_RAY_EPSILON; next->maxt = PB_INFINITY;
        *depth = d + 1;
        return true;
      }
      fr.stage = 1;
    }
    if (fr.stage == 1) {
      rng.skip(3);            // ... TRANSMISSION|SPECULAR draws
      if (fr.transValid) {
        fr.stage = 3;         // transmission child in flight
        io.store(d, fr);
        next->o = fr.p; next->d = fr.wiT; next->mint = PB_RAY_EPSILON; next->maxt = PB_INFINITY;
        *depth = d + 1;
        return true;
      }
      fr.stage = 2;
    }
    // frame complete: return to the parent
    if (d == 0) { io.store(0, fr); *depth = 0; return false; }
    WdFrame parent = io.load(d - 1);
    if (parent.stage == 0) {
      parent.L = parent.L + fr.L * parent.fR * parent.aR;   // L += Li(rd) * f * AbsDot(wi,n)
      parent.stage = 1;
    } else {
      parent.L = parent.L + fr.L * parent.fT * parent.aT;
      parent.stage = 2;
    }
    --d;
    fr = parent;
  }
}

// Scene::Render's guards, core/scene.cpp:60-74.  Returns 0 ok, 1 NaN, 2 negative, 3 infinite.
PB_HD int sample_guard(Spec *Ls) {
  if (spec_isnan(*Ls)) { *Ls = spec(0.f); return 1; }
  float y = spec_y(*Ls);
  if (y < -1e-5) { *Ls = spec(0.f); return 2; }
  if (isinf(y)) { *Ls = spec(0.f); return 3; }
  return 0;
}
#endif
