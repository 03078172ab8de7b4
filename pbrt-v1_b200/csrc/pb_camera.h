// pb_camera.h -- image-sample generation and camera rays.
//   StratifiedSampler ctor / GetNextSample   samplers/stratified.cpp:51-131
//   StratifiedSample1D/2D, Shuffle           core/sampling.cpp:72-97
//   PerspectiveCamera::GenerateRay           cameras/perspective.cpp:51-82
// One call produces the camera part of image sample k of pixel (x, y) without
// generating the pixel's other samples (see pb_rng.h for how).
#ifndef PB_CAMERA_H
#define PB_CAMERA_H
#include "pb_bsdf.h"
#include "pb_texture.h"

struct CameraSample {
  float imageX, imageY, lensU, lensV;
};
// samplers/lowdiscrepancy.cpp:105-110: the sample's camera values come from the pixel's tables
PB_HD void ld_camera_sample(const SceneView &sv, const PatternSource &ps, int x, int y, CameraSample *cs) {
  cs->imageX = x + ps.ld[(size_t)(sv.ldCamera[0].offset + 0) * ps.stride + ps.index];
  cs->imageY = y + ps.ld[(size_t)(sv.ldCamera[0].offset + 1) * ps.stride + ps.index];
  cs->lensU = ps.ld[(size_t)(sv.ldCamera[1].offset + 0) * ps.stride + ps.index];
  cs->lensV = ps.ld[(size_t)(sv.ldCamera[1].offset + 1) * ps.stride + ps.index];
}
// samplers/bestcandidate.cpp:107-129: entry `off` of the table placement with corner (cornerX, cornerY).
// key3 = the placement's stream (its first three draws are the time / lensU / lensV shifts).
// Returns false when the sample falls outside the sampler's extent [xs,xe) x [ys,ye) (the reference
// skips such entries).
PB_HD bool bc_camera_sample(const SceneView &sv, float cornerX, float cornerY, int off, uint64_t key3, int xs,
                            int xe, int ys, int ye, CameraSample *cs) {
  const float *e = sv.bcTable + 5 * off;
  float tw = sv.bcTableWidth;
  cs->imageX = cornerX + tw * e[0];
  cs->imageY = cornerY + tw * e[1];
  float lu = pb_f32(key3, 2u) + e[3];
  float lv = pb_f32(key3, 3u) + e[4];
  cs->lensU = lu > 1 ? lu - 1 : lu;   // WRAP
  cs->lensV = lv > 1 ? lv - 1 : lv;
  return !(cs->imageX < xs || cs->imageX >= xe || cs->imageY < ys || cs->imageY >= ye);
}
PB_HD void stratified_camera_sample(const SceneView &sv, int x, int y, int k, CameraSample *cs) {
  int nx = sv.sampler.x_samples, ny = sv.sampler.y_samples;
  int n = nx * ny;
  if (sv.sampler.kind == PBRT_B200_SAMPLER_RANDOM) {
    // samplers/random.cpp:77-100: 5n uniform draws per pixel: image (2n), lens (2n), time (n)
    uint64_t keyR = pb_key(x, y, 0xFFFF, 0);
    cs->imageX = pb_f32(keyR, 2u * k + 1u);
    cs->imageY = pb_f32(keyR, 2u * k + 2u);
    cs->imageX += x;
    cs->imageY += y;
    cs->lensU = pb_f32(keyR, 2u * n + 2u * k + 1u);
    cs->lensV = pb_f32(keyR, 2u * n + 2u * k + 2u);
    return;
  }
  bool jitter = sv.sampler.jitter != 0;
  uint64_t key = pb_key(x, y, 0xFFFF, 0);
  float dx = 1.f / nx, dy = 1.f / ny;
  // StratifiedSample2D(imageSamples): sample index = sy * nx + sx, jx drawn before jy
  int sx = k % nx, sy = k / nx;
  float jx = jitter ? pb_f32(key, 2u * k + 1u) : 0.5f;
  float jy = jitter ? pb_f32(key, 2u * k + 2u) : 0.5f;
  cs->imageX = (sx + jx) * dx;
  cs->imageY = (sy + jy) * dy;
  cs->imageX += x;
  cs->imageY += y;
  cs->lensU = cs->lensV = 0.5f;
  if (sv.camera.lens_radius > 0.f) {
    // lens samples: draws 2n+1..4n (jitter), shuffled by the n uints that follow the
    // 5n jitter draws (image 2n, lens 2n, time n)
    uint32_t shuffleCtr = jitter ? 5u * n : 0u;
    int src = pb_shuffle_source(key, shuffleCtr, n, k);
    int lx = src % nx, ly = src / nx;
    float lj1 = jitter ? pb_f32(key, 2u * n + 2u * src + 1u) : 0.5f;
    float lj2 = jitter ? pb_f32(key, 2u * n + 2u * src + 2u) : 0.5f;
    cs->lensU = (lx + lj1) * dx;
    cs->lensV = (ly + lj2) * dy;
  }
  // the time sample (shuffled likewise) only sets ray.time, which nothing in scope reads
}

// Scene::Render brackets the two differential-ray calls with ++imageX / --imageX
// (core/scene.cpp:47-53), so the coordinate ImageFilm::AddSample later sees is
// (v + 1) - 1, which is not always v in fp32.
PB_HD float film_image_coord(float v) {
  float t = v + 1.f;
  return t - 1.f;
}

// cameras/environment.cpp:49-63: latitude / longitude directions from the film position (needs the
// film resolution); cameras/orthographic.cpp:49-78 differs from the perspective camera in the ray
// direction (0,0,1) and maxt.
PB_NOINLINE void environment_generate_ray(const PbrtB200Camera &c, const CameraSample &s, int xRes, int yRes, Ray *ray) {
  ray->o = v3(c.ray_origin[0], c.ray_origin[1], c.ray_origin[2]);
  float theta = PB_PI * s.imageY / yRes;
  float phi = 2 * PB_PI * s.imageX / xRes;
  V3 dir = v3(pb_sinf(theta) * pb_cosf(phi), pb_cosf(theta), pb_sinf(theta) * pb_sinf(phi));
  ray->d = xf_vector(c.camera_to_world, dir);
  ray->mint = c.hither;
  ray->maxt = c.yon;
}
// cameras/perspective.cpp:51-82
PB_HD void perspective_generate_ray(const PbrtB200Camera &c, const CameraSample &s, Ray *ray, int xRes = 0, int yRes = 0) {
  if (c.kind == PBRT_B200_CAMERA_ENVIRONMENT) {
    environment_generate_ray(c, s, xRes, yRes, ray);
    return;
  }
  const bool ortho = c.kind == PBRT_B200_CAMERA_ORTHOGRAPHIC;
  V3 Pras = v3(s.imageX, s.imageY, 0);
  V3 Pcamera = xf_point(c.raster_to_camera, Pras);
  ray->o = Pcamera;
  ray->d = ortho ? v3(0, 0, 1) : v3(Pcamera.x, Pcamera.y, Pcamera.z);
  if (c.lens_radius > 0.f) {
    float lensU, lensV;
    concentric_sample_disk(s.lensU, s.lensV, &lensU, &lensV);
    lensU *= c.lens_radius;
    lensV *= c.lens_radius;
    float ft = (c.focal_distance - c.hither) / ray->d.z;
    V3 Pfocus = ray->o + ray->d * ft;
    ray->o.x += lensU * (c.focal_distance - c.hither) / c.focal_distance;
    ray->o.y += lensV * (c.focal_distance - c.hither) / c.focal_distance;
    ray->d = Pfocus - ray->o;
  }
  ray->d = normalize(ray->d);
  ray->mint = 0.f;
  ray->maxt = ortho ? c.yon - c.hither : (c.yon - c.hither) / ray->d.z;
  ray->o = xf_point(c.camera_to_world, ray->o);
  ray->d = xf_vector(c.camera_to_world, ray->d);
}
// Scene::Render's differential rays (core/scene.cpp:46-53): rx from imageX + 1; imageX is then restored by
// `--`, i.e. to (x + 1) - 1, before ry is generated from imageY + 1.  Both weights are 1 for the cameras in
// scope, so hasDifferentials is always set.
PB_HD void camera_ray_differentials(const SceneView &sv, const CameraSample &cs, RayDiff *rd) {
  CameraSample s = cs;
  Ray r;
  s.imageX = cs.imageX + 1.f;
  perspective_generate_ray(sv.camera, s, &r, sv.xRes, sv.yRes);
  rd->rxo = r.o; rd->rxd = r.d;
  s.imageX = s.imageX - 1.f;
  s.imageY = cs.imageY + 1.f;
  perspective_generate_ray(sv.camera, s, &r, sv.xRes, sv.yRes);
  rd->ryo = r.o; rd->ryd = r.d;
  rd->has = true;
}
#endif
