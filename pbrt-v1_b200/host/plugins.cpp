// plugins.cpp -- host halves of the in-scope material / light / integrator /
// camera / sampler plugins and the name -> factory registry that replaces
// dlopen("<name>.so") (core/dynload.cpp:41-61,272-460).  Parameter names and
// defaults are those of SURVEY.md App. B (each cited below).
#include "core.h"

Shape *CreateSphereShape(const Transform &, bool, const ParamSet &);
Shape *CreateCylinderShape(const Transform &, bool, const ParamSet &);
Shape *CreateDiskShape(const Transform &, bool, const ParamSet &);
Shape *CreateConeShape(const Transform &, bool, const ParamSet &);
Shape *CreateParaboloidShape(const Transform &, bool, const ParamSet &);
Shape *CreateHyperboloidShape(const Transform &, bool, const ParamSet &);
Shape *CreateHeightfieldShape(const Transform &, bool, const ParamSet &);
Shape *CreateLoopSubdivShape(const Transform &, bool, const ParamSet &);
Shape *CreateNURBSShape(const Transform &, bool, const ParamSet &);
Shape *CreateTriangleMeshShape(const Transform &, bool, const ParamSet &);
Filter *CreateNamedFilter(const string &name, const ParamSet &ps);
Film *CreateImageFilm(const ParamSet &params, Filter *filter);

namespace {

// ---------------------------------------------------------------- materials
class Matte : public Material {   // materials/matte.cpp:46-70
 public:
  Matte(const Spectrum &kd, float sig) : Kd(kd), sigma(sig) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_MATTE;
    Spectrum r = Kd.Clamp();                       // matte.cpp:56
    for (int i = 0; i < 3; ++i) out->kd[i] = r.c[i];
    out->sigma = Clamp(sigma, 0.f, 90.f);          // matte.cpp:57
  }
  Spectrum Kd; float sigma;
};
class Glass : public Material {   // materials/glass.cpp:46-70
 public:
  Glass(const Spectrum &r, const Spectrum &t, float ior) : Kr(r), Kt(t), index(ior) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_GLASS;
    Spectrum R = Kr.Clamp(), T = Kt.Clamp();
    for (int i = 0; i < 3; ++i) { out->kr[i] = R.c[i]; out->kt[i] = T.c[i]; }
    out->index = index;
  }
  Spectrum Kr, Kt; float index;
};
class Mirror : public Material {  // materials/mirror.cpp:42-60
 public:
  Mirror(const Spectrum &r) : Kr(r) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_MIRROR;
    Spectrum R = Kr.Clamp();
    for (int i = 0; i < 3; ++i) out->kr[i] = R.c[i];
  }
  Spectrum Kr;
};

class Plastic : public Material {   // materials/plastic.cpp:28-77
 public:
  Plastic(const Spectrum &kd, const Spectrum &ks, float rough) : Kd(kd), Ks(ks), roughness(rough) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_PLASTIC;
    Spectrum d = Kd.Clamp(), s = Ks.Clamp();
    for (int i = 0; i < 3; ++i) { out->kd[i] = d.c[i]; out->ks[i] = s.c[i]; }
    out->roughness = roughness;
  }
  Spectrum Kd, Ks; float roughness;
};
class UberMaterial : public Material {   // materials/uber.cpp:28-84
 public:
  UberMaterial(const Spectrum &kd, const Spectrum &ks, const Spectrum &kr, float rough, const Spectrum &op)
      : Kd(kd), Ks(ks), Kr(kr), opacity(op), roughness(rough) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_UBER;
    // the products with the opacity (uber.cpp:65-83) are formed on the device, per GetBSDF call
    Spectrum d = Kd.Clamp(), s = Ks.Clamp(), r = Kr.Clamp(), o = opacity.Clamp();
    for (int i = 0; i < 3; ++i) { out->kd[i] = d.c[i]; out->ks[i] = s.c[i]; out->kr[i] = r.c[i]; out->opacity[i] = o.c[i]; }
    out->roughness = roughness;
  }
  Spectrum Kd, Ks, Kr, opacity; float roughness;
};
class Translucent : public Material {   // materials/translucent.cpp:28-84
 public:
  Translucent(const Spectrum &kd, const Spectrum &ks, float rough, const Spectrum &refl, const Spectrum &trans)
      : Kd(kd), Ks(ks), reflect(refl), transmit(trans), roughness(rough) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_TRANSLUCENT;
    Spectrum d = Kd.Clamp(), s = Ks.Clamp(), r = reflect.Clamp(), t = transmit.Clamp();
    for (int i = 0; i < 3; ++i) { out->kd[i] = d.c[i]; out->ks[i] = s.c[i]; out->kr[i] = r.c[i]; out->kt[i] = t.c[i]; }
    out->roughness = roughness;
  }
  Spectrum Kd, Ks, reflect, transmit; float roughness;
};
class Substrate : public Material {   // materials/substrate.cpp:28-73
 public:
  Substrate(const Spectrum &kd, const Spectrum &ks, float u, float v) : Kd(kd), Ks(ks), nu(u), nv(v) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_SUBSTRATE;
    Spectrum d = Kd.Clamp(), s = Ks.Clamp();
    for (int i = 0; i < 3; ++i) { out->kd[i] = d.c[i]; out->ks[i] = s.c[i]; }
    out->roughness = nu;   // uroughness
    out->sigma = nv;       // vroughness
  }
  Spectrum Kd, Ks; float nu, nv;
};
class Measured : public Material {   // materials/{bluepaint,brushedmetal,clay,felt,primer,skin}.cpp
 public:
  Measured(int w) : which(w) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_MEASURED;
    out->index = (float)which;
  }
  int which;
};
class ShinyMetal : public Material {   // materials/shinymetal.cpp:28-65
 public:
  ShinyMetal(const Spectrum &ks, float rough, const Spectrum &kr) : Ks(ks), Kr(kr), roughness(rough) {}
  void Flatten(PbrtB200Material *out) const {
    InitRecord(out);
    out->kind = PBRT_B200_MAT_SHINYMETAL;
    Spectrum s = Ks.Clamp(), r = Kr.Clamp();
    for (int i = 0; i < 3; ++i) { out->ks[i] = s.c[i]; out->kr[i] = r.c[i]; }
    out->roughness = roughness;
  }
  Spectrum Ks, Kr; float roughness;
};

// ------------------------------------------------------------------- lights
class DistantLight : public Light {  // lights/distant.cpp:53-59
 public:
  DistantLight(const Transform &light2world, const Spectrum &radiance, const Vector &dir)
      : Light(light2world), L(radiance) {
    lightDir = Normalize(LightToWorld(dir));
  }
  void Flatten(FlatBuilder &, PbrtB200Light *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = PBRT_B200_LIGHT_DISTANT;
    out->n_samples = nSamples;
    for (int i = 0; i < 3; ++i) { out->color[i] = L.c[i]; out->pos[i] = lightDir[i]; }
  }
  Vector lightDir; Spectrum L;
};
// ComputeStep1dCDF, core/mc.cpp:31-42
static void ComputeStep1dCDF(const float *f, int nSteps, float *c, float *cdf) {
  int i;
  cdf[0] = 0.;
  for (i = 1; i < nSteps + 1; ++i) cdf[i] = cdf[i - 1] + f[i - 1] / nSteps;
  *c = cdf[nSteps];
  for (i = 1; i < nSteps + 1; ++i) cdf[i] /= *c;
}
class InfiniteAreaLight : public Light {   // lights/infinite.cpp:63-87; importance = lights/infinitesample.cpp:88-137
 public:
  InfiniteAreaLight(const Transform &light2world, const Spectrum &L, int ns, const string &texmap, bool importance = false)
      : Light(light2world, ns), Lbase(L), radianceMap(-1), distOffset(0), nu(0), nv(0) {
    TextureStore &store = CurrentTextureStore();
    if (texmap != "") radianceMap = store.AddRadianceMap(texmap);
    if (!importance) return;
    if (radianceMap < 0) Severe("infinitesample light: a readable \"mapname\" is required (the reference crashes without one)");
    int width = store.imageSizes[texmap].first, height = store.imageSizes[texmap].second;
    // scalar-valued image from the environment map
    float filter = 1.f / max(width, height);
    nu = width; nv = height;
    vector<float> img((size_t)width * height);
    for (int x = 0; x < nu; ++x) {
      float xp = (float)x / (float)nu;
      for (int y = 0; y < nv; ++y) {
        float yp = (float)y / (float)nv;
        float rgb[3];
        store.LookupWidth(radianceMap, xp, yp, filter, rgb);
        const float YWeight[3] = {0.212671f, 0.715160f, 0.072169f};   // Spectrum::y(), core/color.h:141-146
        float v = 0.;
        for (int c = 0; c < 3; ++c) v += YWeight[c] * rgb[c];
        img[y + x * nv] = v;
      }
    }
    vector<float> func(max(nu, nv)), sinVals(nv);
    for (int i = 0; i < nv; ++i) sinVals[i] = sin(M_PI * float(i + .5) / float(nv));
    // tables: uFunc[nu] uCdf[nu+1] uFuncInt vFuncInt[nu] { vFunc[nv] vCdf[nv+1] } x nu
    vector<float> &T = store.lightDistributions;
    distOffset = (uint32_t)T.size();
    T.resize(T.size() + 3 * (size_t)nu + 2 + (size_t)nu * (2 * (size_t)nv + 1));
    float *b = &T[distOffset];
    float *uFunc = b, *uCdf = b + nu, *uFuncInt = b + 2 * nu + 1, *vFuncInt = b + 2 * nu + 2, *vBase = b + 3 * nu + 2;
    for (int u = 0; u < nu; ++u) {
      for (int v = 0; v < nv; ++v) func[v] = img[u * nv + v] *= sinVals[v];
      float *vFunc = vBase + (size_t)u * (2 * nv + 1), *vCdf = vFunc + nv;
      memcpy(vFunc, &func[0], nv * sizeof(float));
      ComputeStep1dCDF(vFunc, nv, &vFuncInt[u], vCdf);
    }
    for (int u = 0; u < nu; ++u) func[u] = vFuncInt[u];
    memcpy(uFunc, &func[0], nu * sizeof(float));
    ComputeStep1dCDF(uFunc, nu, uFuncInt, uCdf);
  }
  void Flatten(FlatBuilder &, PbrtB200Light *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = PBRT_B200_LIGHT_INFINITE;
    out->n_samples = nSamples;
    for (int i = 0; i < 3; ++i) out->color[i] = Lbase.c[i];
    out->radiance_map = (uint32_t)(radianceMap + 1);
    Transform WorldToLight = LightToWorld.GetInverse();
    const Matrix4x4 &m = WorldToLight.GetMatrix();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out->w2l[3 * i + j] = m.m[i][j];
    out->dist_offset = distOffset; out->dist_nu = (uint32_t)nu; out->dist_nv = (uint32_t)nv;
    if (nu) {   // LightToWorld for the sampled directions (the projection matrix is unused by this light kind)
      const Matrix4x4 &l2w = LightToWorld.GetMatrix();
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out->projection[4 * i + j] = l2w.m[i][j];
    }
  }
  Spectrum Lbase;
  int radianceMap;   // texture-table entry holding the pyramid, or -1
  uint32_t distOffset;
  int nu, nv;
};
class SpotLight : public Light {     // lights/spot.cpp:51-58
 public:
  SpotLight(const Transform &light2world, const Spectrum &intensity, float width, float fall)
      : Light(light2world), Intensity(intensity), WorldToLight(light2world.GetInverse()) {
    lightPos = LightToWorld(Point(0, 0, 0));
    cosTotalWidth = cosf(Radians(width));
    cosFalloffStart = cosf(Radians(fall));
  }
  void Flatten(FlatBuilder &, PbrtB200Light *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = PBRT_B200_LIGHT_SPOT;
    out->n_samples = nSamples;
    for (int i = 0; i < 3; ++i) { out->color[i] = Intensity.c[i]; out->pos[i] = lightPos[i]; }
    out->cos_total_width = cosTotalWidth;
    out->cos_falloff_start = cosFalloffStart;
    const Matrix4x4 &m = WorldToLight.GetMatrix();
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out->w2l[3 * r + c] = m.m[r][c];
  }
  Point lightPos; Spectrum Intensity; float cosTotalWidth, cosFalloffStart; Transform WorldToLight;
};
class PointLight : public Light {  // lights/point.cpp:49-54
 public:
  PointLight(const Transform &light2world, const Spectrum &intensity)
      : Light(light2world), Intensity(intensity) {
    lightPos = LightToWorld(Point(0, 0, 0));
  }
  void Flatten(FlatBuilder &, PbrtB200Light *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = PBRT_B200_LIGHT_POINT;
    out->n_samples = nSamples;
    for (int i = 0; i < 3; ++i) { out->color[i] = Intensity.c[i]; out->pos[i] = lightPos[i]; }
  }
  Point lightPos; Spectrum Intensity;
};

// lights/goniometric.cpp:31-83 and lights/projection.cpp:31-99: point lights whose intensity is scaled by an image
// looked up at the direction towards the receiver
class MappedPointLight : public Light {
 public:
  MappedPointLight(int kind_, const Transform &light2world, const Spectrum &intensity, const string &texname, float fov)
      : Light(light2world), kind(kind_), Intensity(intensity), map(-1) {
    lightPos = LightToWorld(Point(0, 0, 0));
    TextureStore &store = CurrentTextureStore();
    if (texname != "") map = store.AddRadianceMap(texname);
    float aspect = 1.f;
    if (map >= 0) aspect = float(store.imageSizes[texname].first) / float(store.imageSizes[texname].second);
    else if (kind == PBRT_B200_LIGHT_PROJECTION)
      Warning("projection light without a readable \"mapname\": the reference reads an uninitialised image size "
              "here; a square projection window is used");
    if (aspect > 1.f) { screen[0] = -aspect; screen[1] = aspect; screen[2] = -1.f; screen[3] = 1.f; }
    else { screen[0] = -1.f; screen[1] = 1.f; screen[2] = -1.f / aspect; screen[3] = 1.f / aspect; }
    lightProjection = Perspective(fov, RAY_EPSILON, 1e30f);
  }
  void Flatten(FlatBuilder &, PbrtB200Light *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = kind;
    out->n_samples = nSamples;
    for (int i = 0; i < 3; ++i) { out->color[i] = Intensity.c[i]; out->pos[i] = lightPos[i]; }
    out->radiance_map = (uint32_t)(map + 1);
    Transform WorldToLight = LightToWorld.GetInverse();
    const Matrix4x4 &m = WorldToLight.GetMatrix();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out->w2l[3 * i + j] = m.m[i][j];
    const Matrix4x4 &pm = lightProjection.GetMatrix();
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) out->projection[4 * i + j] = pm.m[i][j];
    for (int i = 0; i < 4; ++i) out->screen[i] = screen[i];
  }
  int kind;
  Point lightPos; Spectrum Intensity;
  int map;
  Transform lightProjection;
  float screen[4];
};

// -------------------------------------------------------------- integrators
class WhittedIntegrator : public SurfaceIntegrator {  // integrators/whitted.cpp:141-145
 public:
  WhittedIntegrator(int md) : maxDepth(md) {}
  void Flatten(PbrtB200Integrator *out) const {
    out->kind = PBRT_B200_INT_WHITTED; out->max_depth = maxDepth; out->strategy = 0;
  }
  int maxDepth;
};
class DirectLighting : public SurfaceIntegrator {     // integrators/directlighting.cpp:195-208
 public:
  DirectLighting(int st, int md) : strategy(st), maxDepth(md) {}
  void Flatten(PbrtB200Integrator *out) const {
    out->kind = PBRT_B200_INT_DIRECT; out->max_depth = maxDepth; out->strategy = strategy;
  }
  int strategy, maxDepth;
};
class PathIntegrator : public SurfaceIntegrator {     // integrators/path.cpp:146-149
 public:
  PathIntegrator(int md) : maxDepth(md) {}
  void Flatten(PbrtB200Integrator *out) const {
    out->kind = PBRT_B200_INT_PATH; out->max_depth = maxDepth; out->strategy = 0;
  }
  int maxDepth;
};
// integrators/emission.cpp:42-46: no volumes in scope, but its two 1D sample
// requests shift every sample's random-number stream, so they are exported.
class EmissionIntegrator : public VolumeIntegrator {
 public:
  EmissionIntegrator(float ss) : stepSize(ss) {}
  int NumSamples1D() const { return 2; }
  float stepSize;
};

// ------------------------------------------------------------------- camera
// core/transform.cpp:177-180
static Transform Orthographic(float znear, float zfar) {
  return Scale(1.f, 1.f, 1.f / (zfar - znear)) * Translate(Vector(0.f, 0.f, -znear));
}
class PerspectiveCamera : public Camera {   // also the orthographic camera: both are ProjectiveCameras
 public:
  // core/camera.cpp:50-70 + cameras/perspective.cpp:38-50 / cameras/orthographic.cpp:38-48
  PerspectiveCamera(const Transform &world2cam, const float Screen[4], float hither, float yon,
                    float sopen, float sclose, float lensr, float focald, float fov, Film *f, bool ortho = false)
      : Camera(world2cam, hither, yon, sopen, sclose, f), orthographic(ortho) {
    LensRadius = lensr;
    FocalDistance = focald;
    Transform CameraToScreen = ortho ? Orthographic(hither, yon) : Perspective(fov, hither, yon);
    Transform ScreenToRaster =
        Scale(float(film->xResolution), float(film->yResolution), 1.f) *
        Scale(1.f / (Screen[1] - Screen[0]), 1.f / (Screen[2] - Screen[3]), 1.f) *
        Translate(Vector(-Screen[0], -Screen[3], 0.f));
    Transform RasterToScreen = ScreenToRaster.GetInverse();
    RasterToCamera = CameraToScreen.GetInverse() * RasterToScreen;
  }
  void Flatten(PbrtB200Camera *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = orthographic ? PBRT_B200_CAMERA_ORTHOGRAPHIC : PBRT_B200_CAMERA_PERSPECTIVE;
    memcpy(out->raster_to_camera, RasterToCamera.GetMatrix().m, 16 * sizeof(float));
    memcpy(out->camera_to_world, CameraToWorld.GetMatrix().m, 16 * sizeof(float));
    out->hither = ClipHither; out->yon = ClipYon;
    out->shutter_open = ShutterOpen; out->shutter_close = ShutterClose;
    out->lens_radius = LensRadius; out->focal_distance = FocalDistance;
  }
  Transform RasterToCamera;
  float LensRadius, FocalDistance;
  bool orthographic;
};
class EnvironmentCamera : public Camera {   // cameras/environment.cpp:28-63
 public:
  EnvironmentCamera(const Transform &world2cam, float hither, float yon, float sopen, float sclose, Film *f)
      : Camera(world2cam, hither, yon, sopen, sclose, f) {
    rayOrigin = CameraToWorld(Point(0, 0, 0));
  }
  void Flatten(PbrtB200Camera *out) const {
    memset(out, 0, sizeof(*out));
    out->kind = PBRT_B200_CAMERA_ENVIRONMENT;
    out->ray_origin[0] = rayOrigin.x; out->ray_origin[1] = rayOrigin.y; out->ray_origin[2] = rayOrigin.z;
    memcpy(out->camera_to_world, CameraToWorld.GetMatrix().m, 16 * sizeof(float));
    out->hither = ClipHither; out->yon = ClipYon;
    out->shutter_open = ShutterOpen; out->shutter_close = ShutterClose;
  }
  Point rayOrigin;
};

// ------------------------------------------------------------------ sampler
class StratifiedSampler : public Sampler {   // samplers/stratified.cpp:51-64
 public:
  StratifiedSampler(int xstart, int xend, int ystart, int yend, int xs, int ys, bool jitter)
      : Sampler(xstart, xend, ystart, yend, xs * ys), xPixelSamples(xs), yPixelSamples(ys),
        jitterSamples(jitter) {}
  void Flatten(PbrtB200Sampler *out) const {
    out->x_samples = xPixelSamples; out->y_samples = yPixelSamples; out->jitter = jitterSamples;
    out->kind = PBRT_B200_SAMPLER_STRATIFIED;
  }
  int xPixelSamples, yPixelSamples; bool jitterSamples;
};
// samplers/lowdiscrepancy.cpp:57-74: (0,2)-sequence sampler, pixelsamples rounded up to a power of 2
static int RoundUpPow2Int(int v) {
  u_int x = (u_int)v;
  x--; x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
  return (int)(x + 1);
}
class LDSampler : public Sampler {
 public:
  LDSampler(int xstart, int xend, int ystart, int yend, int ps)
      : Sampler(xstart, xend, ystart, yend, RoundUpPow2Int(ps)) {
    if ((ps & (ps - 1)) != 0) Warning("Pixel samples being rounded up to power of 2");
    pixelSamples = RoundUpPow2Int(ps);
  }
  void Flatten(PbrtB200Sampler *out) const {
    out->x_samples = pixelSamples; out->y_samples = 1; out->jitter = 0;
    out->kind = PBRT_B200_SAMPLER_LOWDISCREPANCY;
  }
  int pixelSamples;
};
class RandomSampler : public Sampler {   // samplers/random.cpp:28-76
 public:
  RandomSampler(int xstart, int xend, int ystart, int yend, int xs, int ys)
      : Sampler(xstart, xend, ystart, yend, xs * ys), xPixelSamples(xs), yPixelSamples(ys) {}
  void Flatten(PbrtB200Sampler *out) const {
    out->x_samples = xPixelSamples; out->y_samples = yPixelSamples; out->jitter = 0;
    out->kind = PBRT_B200_SAMPLER_RANDOM;
  }
  int xPixelSamples, yPixelSamples;
};
class BestCandidateSampler : public Sampler {   // samplers/bestcandidate.cpp:33-77
 public:
  BestCandidateSampler(int xstart, int xend, int ystart, int yend, int ps)
      : Sampler(xstart, xend, ystart, yend, ps) {}
  void Flatten(PbrtB200Sampler *out) const {
    out->x_samples = samplesPerPixel; out->y_samples = 1; out->jitter = 0;
    out->kind = PBRT_B200_SAMPLER_BESTCANDIDATE;
  }
};

}  // namespace

// BestCandidateSampler::sampleTable.  The reference compiles it in as data (samplers/sampledata.cpp);
// here it is the output of tools/gen_bestcandidate_table.cpp (our restatement of the generator the
// reference's table came from), stored next to this library as bestcandidate_table.bin.
#include <dlfcn.h>
const float *BestCandidateTable() {
  static vector<float> table;
  if (!table.empty()) return &table[0];
  string dir = ".";
  Dl_info info;
  if (dladdr((void *)&BestCandidateTable, &info) && info.dli_fname) {
    dir = info.dli_fname;
    size_t slash = dir.rfind('/');
    dir = (slash == string::npos) ? string(".") : dir.substr(0, slash);
  }
  string fn = dir + "/bestcandidate_table.bin";
  FILE *f = fopen(fn.c_str(), "rb");
  if (!f) Severe("Unable to open the bestcandidate sample table \"%s\"", fn.c_str());
  table.resize((size_t)PBRT_B200_BC_TABLE_SIZE * 5);
  size_t got = fread(&table[0], sizeof(float), table.size(), f);
  fclose(f);
  if (got != table.size()) Severe("Sample table \"%s\" is truncated", fn.c_str());
  return &table[0];
}

// core/light.h:88-116 + lights/area.cpp:28-57: a shape that cannot be
// intersected directly is refined (LIFO) into a ShapeSet.
AreaLight::AreaLight(const Transform &light2world, const Spectrum &le, int ns,
                     const Reference<Shape> &s)
    : Light(light2world, ns) {
  Lemit = le;
  if (s->CanIntersect()) {
    shapes.push_back(s);
    area = s->Area();
  } else {
    vector<Reference<Shape> > todo, done;
    todo.push_back(s);
    while (todo.size()) {
      Reference<Shape> sh = todo.back();
      todo.pop_back();
      if (sh->CanIntersect()) done.push_back(sh);
      else sh->Refine(todo);
    }
    if (done.size() == 1) {
      shapes = done;
      area = done[0]->Area();
    } else {
      if (done.size() > 16)
        Warning("Area light geometry turned into %d shapes; may be very inefficient.",
                (int)done.size());
      // ShapeSet constructor, core/shape.h:122-138
      shapes = done;
      area = 0;
      vector<float> areas;
      for (u_int i = 0; i < shapes.size(); ++i) {
        float a = shapes[i]->Area();
        area += a;
        areas.push_back(a);
      }
      float prevCDF = 0;
      for (u_int i = 0; i < shapes.size(); ++i) {
        areaCDF.push_back(prevCDF + areas[i] / area);
        prevCDF = areaCDF[i];
      }
    }
  }
}

// ------------------------------------------------------------- factories
// Operand slots a factory found to vary over the surface (TextureParams reports the texture-table entry).
struct TexSlots {
  TexSlots() { for (int i = 0; i < PBRT_B200_TEXSLOTS; ++i) t[i] = -1; }
  Material *Attach(Material *m) const { memcpy(m->tex, t, sizeof(t)); return m; }
  int32_t t[PBRT_B200_TEXSLOTS];
};
// "bumpmap" (Material::Bump, core/material.cpp:29-71): the default constant 0 leaves the slot empty (the device
// evaluates Bump with zero displacement); any other constant becomes a constant node
static void Bumpmap(const TextureParams &mp, TexSlots *ts) {
  int32_t node;
  float v = mp.GetFloatTexture("bumpmap", 0.f, &node);
  if (node < 0 && v != 0.f) {
    float c[3] = {v, v, v};
    node = CurrentTextureStore().AddConstant(c, 1);
  }
  ts->t[PBRT_B200_TEXSLOT_BUMP] = node;
}
static Material *CreateMatte(const Transform &, const TextureParams &mp) {
  TexSlots ts;
  Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KD]);
  float sigma = mp.GetFloatTexture("sigma", 0.f, &ts.t[PBRT_B200_TEXSLOT_SIGMA]);
  Bumpmap(mp, &ts);
  return ts.Attach(new Matte(Kd, sigma));
}
static Material *CreateGlass(const Transform &, const TextureParams &mp) {
  TexSlots ts;
  Spectrum Kr = mp.GetSpectrumTexture("Kr", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KR]);
  Spectrum Kt = mp.GetSpectrumTexture("Kt", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KT]);
  float index = mp.GetFloatTexture("index", 1.5f, &ts.t[PBRT_B200_TEXSLOT_INDEX]);
  Bumpmap(mp, &ts);
  return ts.Attach(new Glass(Kr, Kt, index));
}
static Material *CreatePlastic(const Transform &, const TextureParams &mp) {
  TexSlots ts;
  Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KD]);
  Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KS]);
  float roughness = mp.GetFloatTexture("roughness", .1f, &ts.t[PBRT_B200_TEXSLOT_ROUGHNESS]);
  Bumpmap(mp, &ts);
  return ts.Attach(new Plastic(Kd, Ks, roughness));
}
static Material *CreateUber(const Transform &, const TextureParams &mp) {   // materials/uber.cpp:86-95
  TexSlots ts;
  Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KD]);
  Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KS]);
  Spectrum Kr = mp.GetSpectrumTexture("Kr", Spectrum(0.f), &ts.t[PBRT_B200_TEXSLOT_KR]);
  float roughness = mp.GetFloatTexture("roughness", .1f, &ts.t[PBRT_B200_TEXSLOT_ROUGHNESS]);
  Spectrum opacity = mp.GetSpectrumTexture("opacity", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_OPACITY]);
  Bumpmap(mp, &ts);
  return ts.Attach(new UberMaterial(Kd, Ks, Kr, roughness, opacity));
}
static Material *CreateTranslucent(const Transform &, const TextureParams &mp) {   // materials/translucent.cpp:85-94
  TexSlots ts;
  Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KD]);
  Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KS]);
  Spectrum reflect = mp.GetSpectrumTexture("reflect", Spectrum(0.5f), &ts.t[PBRT_B200_TEXSLOT_KR]);
  Spectrum transmit = mp.GetSpectrumTexture("transmit", Spectrum(0.5f), &ts.t[PBRT_B200_TEXSLOT_KT]);
  float roughness = mp.GetFloatTexture("roughness", .1f, &ts.t[PBRT_B200_TEXSLOT_ROUGHNESS]);
  Bumpmap(mp, &ts);
  return ts.Attach(new Translucent(Kd, Ks, roughness, reflect, transmit));
}
static Material *CreateSubstrate(const Transform &, const TextureParams &mp) {   // materials/substrate.cpp:64-73
  TexSlots ts;
  Spectrum Kd = mp.GetSpectrumTexture("Kd", Spectrum(.5f), &ts.t[PBRT_B200_TEXSLOT_KD]);
  Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(.5f), &ts.t[PBRT_B200_TEXSLOT_KS]);
  float uroughness = mp.GetFloatTexture("uroughness", .1f, &ts.t[PBRT_B200_TEXSLOT_ROUGHNESS]);
  float vroughness = mp.GetFloatTexture("vroughness", .1f, &ts.t[PBRT_B200_TEXSLOT_SIGMA]);
  Bumpmap(mp, &ts);
  return ts.Attach(new Substrate(Kd, Ks, uroughness, vroughness));
}
static Material *CreateShinyMetal(const Transform &, const TextureParams &mp) {   // materials/shinymetal.cpp:66-73
  TexSlots ts;
  Spectrum Kr = mp.GetSpectrumTexture("Kr", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KR]);
  Spectrum Ks = mp.GetSpectrumTexture("Ks", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KS]);
  float roughness = mp.GetFloatTexture("roughness", .1f, &ts.t[PBRT_B200_TEXSLOT_ROUGHNESS]);
  Bumpmap(mp, &ts);
  return ts.Attach(new ShinyMetal(Ks, roughness, Kr));
}
static Material *CreateMeasured(int which, const TextureParams &mp) {   // e.g. materials/bluepaint.cpp:60-64
  TexSlots ts;
  Bumpmap(mp, &ts);
  return ts.Attach(new Measured(which));
}
static Material *CreateMirror(const Transform &, const TextureParams &mp) {
  TexSlots ts;
  Spectrum Kr = mp.GetSpectrumTexture("Kr", Spectrum(1.f), &ts.t[PBRT_B200_TEXSLOT_KR]);
  Bumpmap(mp, &ts);
  return ts.Attach(new Mirror(Kr));
}

static bool gWarnedSearchPath = false;
void UpdatePluginPath(const string &) {
  // Plugins are compiled in; SearchPath / PBRT_SEARCHPATH are accepted and ignored.
  (void)gWarnedSearchPath;
}

Reference<Shape> MakeShape(const string &name, const Transform &object2world,
                           bool reverseOrientation, const ParamSet &paramSet) {
  Shape *s = NULL;
  if (name == "sphere") s = CreateSphereShape(object2world, reverseOrientation, paramSet);
  else if (name == "cylinder") s = CreateCylinderShape(object2world, reverseOrientation, paramSet);
  else if (name == "disk") s = CreateDiskShape(object2world, reverseOrientation, paramSet);
  else if (name == "cone") s = CreateConeShape(object2world, reverseOrientation, paramSet);
  else if (name == "paraboloid") s = CreateParaboloidShape(object2world, reverseOrientation, paramSet);
  else if (name == "hyperboloid") s = CreateHyperboloidShape(object2world, reverseOrientation, paramSet);
  else if (name == "heightfield") s = CreateHeightfieldShape(object2world, reverseOrientation, paramSet);
  else if (name == "loopsubdiv") s = CreateLoopSubdivShape(object2world, reverseOrientation, paramSet);
  else if (name == "nurbs") s = CreateNURBSShape(object2world, reverseOrientation, paramSet);
  else if (name == "trianglemesh")
    s = CreateTriangleMeshShape(object2world, reverseOrientation, paramSet);
  else
    Error("Shape plugin \"%s\" is outside the B200 hot-path scope (supported: sphere, "
          "cylinder, disk, cone, paraboloid, hyperboloid, heightfield, loopsubdiv, nurbs, trianglemesh); shape ignored.", name.c_str());
  return s;
}
Reference<Material> MakeMaterial(const string &name, const Transform &mtl2world,
                                 const TextureParams &mp) {
  Material *m = NULL;
  if (name == "matte") m = CreateMatte(mtl2world, mp);
  else if (name == "glass") m = CreateGlass(mtl2world, mp);
  else if (name == "mirror") m = CreateMirror(mtl2world, mp);
  else if (name == "plastic") m = CreatePlastic(mtl2world, mp);
  else if (name == "uber") m = CreateUber(mtl2world, mp);
  else if (name == "shinymetal") m = CreateShinyMetal(mtl2world, mp);
  else if (name == "translucent") m = CreateTranslucent(mtl2world, mp);
  else if (name == "substrate") m = CreateSubstrate(mtl2world, mp);
  else if (name == "bluepaint") m = CreateMeasured(PBRT_B200_MEASURED_BLUEPAINT, mp);
  else if (name == "brushedmetal") m = CreateMeasured(PBRT_B200_MEASURED_BRUSHEDMETAL, mp);
  else if (name == "clay") m = CreateMeasured(PBRT_B200_MEASURED_CLAY, mp);
  else if (name == "felt") m = CreateMeasured(PBRT_B200_MEASURED_FELT, mp);
  else if (name == "primer") m = CreateMeasured(PBRT_B200_MEASURED_PRIMER, mp);
  else if (name == "skin") m = CreateMeasured(PBRT_B200_MEASURED_SKIN, mp);
  else
    Error("Material plugin \"%s\" unknown (supported: matte, glass, "
          "mirror, plastic, uber, shinymetal, translucent, substrate, bluepaint, brushedmetal, clay, felt, primer, skin).", name.c_str());
  if (m) mp.ReportUnused();   // core/dynload.cpp:299
  return m;
}
Light *MakeLight(const string &name, const Transform &light2world, const ParamSet &paramSet) {
  if (name == "point") {   // lights/point.cpp:78-84
    Spectrum I = paramSet.FindOneSpectrum("I", Spectrum(1.0));
    Point P = paramSet.FindOnePoint("from", Point(0, 0, 0));
    Transform l2w = Translate(Vector(P.x, P.y, P.z)) * light2world;
    Light *l = new PointLight(l2w, I);
    paramSet.ReportUnused();
    return l;
  }
  if (name == "goniometric" || name == "projection") {   // lights/goniometric.cpp:84-89, lights/projection.cpp:133-139
    Spectrum I = paramSet.FindOneSpectrum("I", Spectrum(1.0));
    float fov = name == "projection" ? paramSet.FindOneFloat("fov", 45.) : 45.f;
    string texname = paramSet.FindOneString("mapname", "");
    Light *l = new MappedPointLight(name == "projection" ? PBRT_B200_LIGHT_PROJECTION : PBRT_B200_LIGHT_GONIOMETRIC,
                                    light2world, I, texname, fov);
    paramSet.ReportUnused();
    return l;
  }
  if (name == "infinite" || name == "infinitesample") {   // lights/infinite.cpp:164-170, lights/infinitesample.cpp:217-224
    Spectrum L = paramSet.FindOneSpectrum("L", Spectrum(1.0));
    string texmap = paramSet.FindOneString("mapname", "");
    int nSamples = paramSet.FindOneInt("nsamples", 1);
    Light *l = new InfiniteAreaLight(light2world, L, nSamples, texmap, name == "infinitesample");
    paramSet.ReportUnused();
    return l;
  }
  if (name == "distant") {   // lights/distant.cpp:94-101
    Spectrum L = paramSet.FindOneSpectrum("L", Spectrum(1.0));
    Point from = paramSet.FindOnePoint("from", Point(0, 0, 0));
    Point to = paramSet.FindOnePoint("to", Point(0, 0, 1));
    Vector dir = from - to;
    Light *l = new DistantLight(light2world, L, dir);
    paramSet.ReportUnused();
    return l;
  }
  if (name == "spot") {      // lights/spot.cpp:100-121
    Spectrum I = paramSet.FindOneSpectrum("I", Spectrum(1.0));
    float coneangle = paramSet.FindOneFloat("coneangle", 30.);
    float conedelta = paramSet.FindOneFloat("conedeltaangle", 5.);
    Point from = paramSet.FindOnePoint("from", Point(0, 0, 0));
    Point to = paramSet.FindOnePoint("to", Point(0, 0, 1));
    Vector dir = Normalize(to - from);
    Vector du, dv;
    CoordinateSystem(dir, &du, &dv);
    Transform dirToZ = Transform(new Matrix4x4(du.x, du.y, du.z, 0., dv.x, dv.y, dv.z, 0., dir.x, dir.y, dir.z, 0.,
                                               0, 0, 0, 1.));
    Transform l2w = light2world * Translate(Vector(from.x, from.y, from.z)) * dirToZ.GetInverse();
    Light *l = new SpotLight(l2w, I, coneangle, coneangle - conedelta);
    paramSet.ReportUnused();
    return l;
  }
  return NULL;
}
AreaLight *MakeAreaLight(const string &name, const Transform &light2world,
                         const ParamSet &paramSet, const Reference<Shape> &shape) {
  if (name == "area") {    // lights/area.cpp:106-111
    Spectrum L = paramSet.FindOneSpectrum("L", Spectrum(1.0));
    int nSamples = paramSet.FindOneInt("nsamples", 1);
    AreaLight *a = new AreaLight(light2world, L, nSamples, shape);
    paramSet.ReportUnused();
    return a;
  }
  Error("Area light plugin \"%s\" unknown.", name.c_str());
  return NULL;
}
VolumeRegion *MakeVolumeRegion(const string &name, const Transform &, const ParamSet &) {
  Error("Volume plugin \"%s\": participating media are outside the B200 hot-path scope; ignored.",
        name.c_str());
  return NULL;
}
SurfaceIntegrator *MakeSurfaceIntegrator(const string &name, const ParamSet &params) {
  SurfaceIntegrator *si = NULL;
  if (name == "whitted") si = new WhittedIntegrator(params.FindOneInt("maxdepth", 5));
  else if (name == "path") si = new PathIntegrator(params.FindOneInt("maxdepth", 5));
  else if (name == "directlighting") {
    int maxDepth = params.FindOneInt("maxdepth", 5);
    int strategy;
    string st = params.FindOneString("strategy", "all");
    if (st == "one") strategy = PBRT_B200_STRATEGY_ONE;
    else if (st == "all") strategy = PBRT_B200_STRATEGY_ALL;
    else if (st == "weighted") {
      // transport.cpp:71-122 keeps running luminance averages across samples: inherently
      // serial, outside the scope table.
      Warning("Strategy \"weighted\" depends on sample order and is not supported on the GPU; "
              "using \"one\".");
      strategy = PBRT_B200_STRATEGY_ONE;
    } else {
      Warning("Strategy \"%s\" for direct lighting unknown. Using \"all\".", st.c_str());
      strategy = PBRT_B200_STRATEGY_ALL;
    }
    si = new DirectLighting(strategy, maxDepth);
  } else
    Error("Surface integrator plugin \"%s\" is outside the B200 hot-path scope (supported: "
          "whitted, directlighting, path).", name.c_str());
  if (si) params.ReportUnused();
  return si;
}
VolumeIntegrator *MakeVolumeIntegrator(const string &name, const ParamSet &params) {
  if (name != "emission")
    Warning("Volume integrator \"%s\" not supported; using \"emission\" (no volumes in scope).",
            name.c_str());
  VolumeIntegrator *vi = new EmissionIntegrator(params.FindOneFloat("stepsize", 1.f));
  params.ReportUnused();
  return vi;
}
Primitive *MakeAccelerator(const string &name, const vector<Reference<Primitive> > &prims,
                           const ParamSet &ps) {
  // "grid" (accelerators/grid.cpp) returns the same closest hits as the
  // kd-tree, so it aliases the one accelerator the device implements.
  if (name != "kdtree" && name != "grid") return NULL;
  int isectCost = ps.FindOneInt("intersectcost", 80);    // kdtree.cpp:489-498
  int travCost = ps.FindOneInt("traversalcost", 1);
  float emptyBonus = ps.FindOneFloat("emptybonus", 0.5f);
  int maxPrims = ps.FindOneInt("maxprims", 1);
  int maxDepth = ps.FindOneInt("maxdepth", -1);
  Primitive *p = new KdTreeAccel(prims, isectCost, travCost, emptyBonus, maxPrims, maxDepth);
  if (name == "kdtree") ps.ReportUnused();
  return p;
}
Camera *MakeCamera(const string &name, const ParamSet &params, const Transform &world2cam,
                   Film *film) {
  if (name != "perspective" && name != "orthographic" && name != "environment") {
    Error("Camera plugin \"%s\" unknown (supported: perspective, orthographic, environment).", name.c_str());
    return NULL;
  }
  // cameras/perspective.cpp:83-115, orthographic.cpp:79-112, environment.cpp:64-99 (same parameter block)
  float hither = max(1e-4f, params.FindOneFloat("hither", 1e-3f));
  float yon = min(params.FindOneFloat("yon", 1e30f), 1e30f);
  float shutteropen = params.FindOneFloat("shutteropen", 0.f);
  float shutterclose = params.FindOneFloat("shutterclose", 1.f);
  float lensradius = params.FindOneFloat("lensradius", 0.f);
  float focaldistance = params.FindOneFloat("focaldistance", 1e30f);
  float frame = params.FindOneFloat("frameaspectratio",
                                    float(film->xResolution) / float(film->yResolution));
  float screen[4];
  if (frame > 1.f) {
    screen[0] = -frame; screen[1] = frame; screen[2] = -1.f; screen[3] = 1.f;
  } else {
    screen[0] = -1.f; screen[1] = 1.f; screen[2] = -1.f / frame; screen[3] = 1.f / frame;
  }
  int swi;
  const float *sw = params.FindFloat("screenwindow", &swi);
  if (sw && swi == 4) memcpy(screen, sw, 4 * sizeof(float));
  Camera *c;
  if (name == "environment")
    c = new EnvironmentCamera(world2cam, hither, yon, shutteropen, shutterclose, film);
  else if (name == "orthographic")
    c = new PerspectiveCamera(world2cam, screen, hither, yon, shutteropen, shutterclose, lensradius, focaldistance,
                              0.f, film, true);
  else {
    float fov = params.FindOneFloat("fov", 90.);
    c = new PerspectiveCamera(world2cam, screen, hither, yon, shutteropen, shutterclose, lensradius, focaldistance,
                              fov, film);
  }
  params.ReportUnused();
  return c;
}
Sampler *MakeSampler(const string &name, const ParamSet &params, const Film *film) {
  if (name == "lowdiscrepancy") {   // samplers/lowdiscrepancy.cpp:129-135
    int xstart, xend, ystart, yend;
    film->GetSampleExtent(&xstart, &xend, &ystart, &yend);
    int nsamp = params.FindOneInt("pixelsamples", 4);
    Sampler *s = new LDSampler(xstart, xend, ystart, yend, nsamp < 1 ? 1 : nsamp);
    params.ReportUnused();
    return s;
  }
  if (name == "bestcandidate") {   // samplers/bestcandidate.cpp:172-178
    int xstart, xend, ystart, yend;
    film->GetSampleExtent(&xstart, &xend, &ystart, &yend);
    int nsamp = params.FindOneInt("pixelsamples", 4);
    if (nsamp < 1 || nsamp > PBRT_B200_BC_TABLE_SIZE) {
      Error("bestcandidate sampler: pixelsamples %d outside [1, %d]; using 4.", nsamp, PBRT_B200_BC_TABLE_SIZE);
      nsamp = 4;
    }
    Sampler *s = new BestCandidateSampler(xstart, xend, ystart, yend, nsamp);
    params.ReportUnused();
    return s;
  }
  if (name == "random") {   // samplers/random.cpp:113-122
    int xsamp = params.FindOneInt("xsamples", 2);
    int ysamp = params.FindOneInt("ysamples", 2);
    int xstart, xend, ystart, yend;
    film->GetSampleExtent(&xstart, &xend, &ystart, &yend);
    Sampler *s = new RandomSampler(xstart, xend, ystart, yend, xsamp < 1 ? 1 : xsamp, ysamp < 1 ? 1 : ysamp);
    params.ReportUnused();
    return s;
  }
  if (name != "stratified") {
    // the reference's MakeSampler returns NULL for a plug-in it cannot load (core/dynload.cpp:208-216) and
    // MakeScene then aborts with "missing plug-ins" (core/api.cpp:494-497): same here
    Error("Sampler plugin \"%s\" unknown (supported: stratified, lowdiscrepancy, bestcandidate, random).", name.c_str());
    return NULL;
  }
  bool jitter = params.FindOneBool("jitter", true);      // samplers/stratified.cpp:132-141
  int xstart, xend, ystart, yend;
  film->GetSampleExtent(&xstart, &xend, &ystart, &yend);
  int xsamp = params.FindOneInt("xsamples", 2);
  int ysamp = params.FindOneInt("ysamples", 2);
  Sampler *s = new StratifiedSampler(xstart, xend, ystart, yend, xsamp, ysamp, jitter);
  params.ReportUnused();
  return s;
}
Filter *MakeFilter(const string &name, const ParamSet &paramSet) {
  Filter *f = CreateNamedFilter(name, paramSet);
  if (f) paramSet.ReportUnused();
  return f;
}
Film *MakeFilm(const string &name, const ParamSet &paramSet, Filter *filter) {
  if (name != "image") {
    Error("Film plugin \"%s\" unknown (supported: image).", name.c_str());
    return NULL;
  }
  Film *f = CreateImageFilm(paramSet, filter);
  paramSet.ReportUnused();
  return f;
}
