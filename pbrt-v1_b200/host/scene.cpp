// scene.cpp -- Scene construction, export to the C-ABI arrays, and the
// drop-in Scene::Render() (reference: core/scene.cpp:32-129).
//
// Render() does not trace anything on the CPU: it flattens the scene and hands
// it to libpbrt_b200.so (dlopen'ed like the reference dlopen's its plugins,
// core/dynload.cpp:462-490).  If that library cannot be loaded the render
// aborts with Severe() -- there is deliberately no CPU fallback.
#include <chrono>
#include "core.h"
#include <dlfcn.h>
#include <sys/time.h>

Scene::Scene(Camera *cam, SurfaceIntegrator *si, VolumeIntegrator *vi, Sampler *s,
             Primitive *accel, const vector<Light *> &lts, VolumeRegion *vr) {
  lights = lts;
  aggregate = accel;
  camera = cam;
  sampler = s;
  surfaceIntegrator = si;
  volumeIntegrator = vi;
  volumeRegion = vr;
  if (lts.size() == 0)
    Warning("No light sources defined in scene; possibly rendering a black image.");
  bound = aggregate->WorldBound();
}
Scene::~Scene() {
  delete camera;
  delete sampler;
  delete surfaceIntegrator;
  delete volumeIntegrator;
  delete aggregate;
  delete volumeRegion;
  for (u_int i = 0; i < lights.size(); ++i) delete lights[i];
}

void AreaLight::Flatten(FlatBuilder &fb, PbrtB200Light *out) const {
  memset(out, 0, sizeof(*out));
  out->kind = PBRT_B200_LIGHT_AREA;
  out->n_samples = nSamples;
  for (int i = 0; i < 3; ++i) out->color[i] = Lemit.c[i];
  out->first_shape = (uint32_t)fb.lightShapes.size();
  out->n_shapes = (uint32_t)shapes.size();
  out->area = area;
  for (size_t i = 0; i < shapes.size(); ++i) {
    const void *owner; int index;
    shapes[i]->Identity(&owner, &index);
    map<std::pair<const void *, int>, uint32_t>::const_iterator it =
        fb.shapePrimIds.find(std::make_pair(owner, index));
    if (it == fb.shapePrimIds.end())
      Severe("Area light shape has no matching primitive in the accelerator");
    fb.lightShapes.push_back(it->second);
    fb.lightCdf.push_back(i < areaCDF.size() ? areaCDF[i] : 1.f);
  }
}

void Scene::Flatten(FlatBuilder &fb, PbrtB200Scene *out) const {
  memset(out, 0, sizeof(*out));
  out->abi_version = PBRT_B200_ABI_VERSION;
  KdTreeAccel *accel = dynamic_cast<KdTreeAccel *>(aggregate);
  if (!accel) Severe("Scene aggregate is not the kd-tree accelerator");
  for (size_t i = 0; i < lights.size(); ++i)
    if (lights[i]->IsAreaLight()) fb.lightIds[lights[i]] = (int32_t)i;
  // primitives in the accelerator's (refined) order; instanced objects' primitives follow
  vector<const Primitive *> all;
  for (size_t i = 0; i < accel->prims.size(); ++i) all.push_back(accel->prims[i].operator->());
  bool anyInstance = false;
  for (size_t i = 0; i < all.size(); ++i) anyInstance |= dynamic_cast<const InstancePrimitive *>(all[i]) != NULL;
  if (anyInstance) {
    fb.nodes = accel->nodes;
    fb.primRefs = accel->primRefs;
  }
  const size_t nTop = all.size();
  vector<const AreaLight *> unsampled;
  fb.prims.resize(nTop);
  fb.aux.resize(nTop);
  for (size_t i = 0; i < all.size(); ++i) {   // `all` grows while instances are expanded
    PbrtB200Prim &pr = fb.prims[i];
    if (const InstancePrimitive *ip = dynamic_cast<const InstancePrimitive *>(all[i])) {
      if (i >= nTop) Severe("Nested object instances are not possible in pbrt-v1 scenes");
      const Primitive *obj = ip->instance.operator->();
      PbrtB200Instance in;
      memset(&in, 0, sizeof(in));
      map<const void *, uint32_t>::iterator oi = fb.objectIds.find(obj);
      if (oi != fb.objectIds.end()) {
        in = fb.instances[oi->second];   // same object placed again: shares tree and primitives
      } else {
        fb.objectIds[obj] = (uint32_t)fb.instances.size();
        in.prim_base = (uint32_t)all.size();
        if (const KdTreeAccel *oa = dynamic_cast<const KdTreeAccel *>(obj)) {
          in.node_base = (uint32_t)fb.nodes.size();
          in.n_nodes = (uint32_t)oa->nodes.size();
          in.ref_base = (uint32_t)fb.primRefs.size();
          fb.nodes.insert(fb.nodes.end(), oa->nodes.begin(), oa->nodes.end());
          fb.primRefs.insert(fb.primRefs.end(), oa->primRefs.begin(), oa->primRefs.end());
          for (int k = 0; k < 3; ++k) { in.bounds[k] = oa->bounds.pMin[k]; in.bounds[3 + k] = oa->bounds.pMax[k]; }
          for (size_t k = 0; k < oa->prims.size(); ++k) all.push_back(oa->prims[k].operator->());
        } else {
          all.push_back(obj);   // a single intersectable primitive, no accelerator (api.cpp:441)
        }
        fb.prims.resize(all.size());
        fb.aux.resize(all.size());
      }
      memcpy(in.i2w, ip->InstanceToWorld.GetMatrix().m, sizeof(in.i2w));
      memcpy(in.w2i, ip->WorldToInstance.GetMatrix().m, sizeof(in.w2i));
      PbrtB200Prim &pi = fb.prims[i];   // (fb.prims may have been reallocated)
      memset(&pi, 0, sizeof(pi));
      pi.flags = PBRT_B200_PRIM_INSTANCE;
      pi.light = -1;
      memset(&fb.aux[i], 0, sizeof(fb.aux[i]));
      fb.aux[i].owner = (uint32_t)fb.instances.size();
      fb.instances.push_back(in);
      continue;
    }
    const GeometricPrimitive *gp = dynamic_cast<const GeometricPrimitive *>(all[i]);
    if (!gp) Severe("Non-geometric primitive reached the accelerator");
    gp->shape->Flatten(fb, &pr, &fb.aux[i]);
    if (i < nTop && gp->areaLight) {
      // (only emitters are looked up again, by AreaLight::Flatten: a map entry for each of a million triangles
      // was a third of a second of the load)
      const void *owner; int index;
      gp->shape->Identity(&owner, &index);
      fb.shapePrimIds[std::make_pair(owner, index)] = (uint32_t)i;
    }
    const Material *m = gp->material.operator->();
    map<const void *, uint32_t>::iterator mi = fb.materialIds.find(m);
    if (mi == fb.materialIds.end()) {
      PbrtB200Material pm;
      m->Flatten(&pm);
      fb.materialIds[m] = pr.material = (uint32_t)fb.materials.size();
      fb.materials.push_back(pm);
    } else
      pr.material = mi->second;
    pr.light = -1;
    if (gp->areaLight) {
      map<const void *, int32_t>::iterator li = fb.lightIds.find(gp->areaLight);
      if (li != fb.lightIds.end())
        pr.light = li->second;
      else if (i >= nTop) {
        // a shape inside ObjectBegin keeps its AreaLight (it emits when hit, Intersection::Le) but the
        // light never joins the scene's light vector (api.cpp:383-387): an unsampled emitter
        fb.lightIds[gp->areaLight] = pr.light = (int32_t)(lights.size() + unsampled.size());
        unsampled.push_back(gp->areaLight);
      }
    }
  }
  fb.lights.resize(lights.size() + unsampled.size());
  for (size_t i = 0; i < lights.size(); ++i) lights[i]->Flatten(fb, &fb.lights[i]);
  for (size_t i = 0; i < unsampled.size(); ++i) {
    PbrtB200Light *pl = &fb.lights[lights.size() + i];
    memset(pl, 0, sizeof(*pl));
    pl->kind = PBRT_B200_LIGHT_AREA;
    for (int c = 0; c < 3; ++c) pl->color[c] = unsampled[i]->Lemit.c[c];
  }

  const vector<PbrtB200KdNode> &nodes = anyInstance ? fb.nodes : accel->nodes;
  const vector<uint32_t> &primRefs = anyInstance ? fb.primRefs : accel->primRefs;
  out->nodes = nodes.empty() ? NULL : &nodes[0];
  out->n_nodes = (uint32_t)nodes.size();
  out->n_top_nodes = (uint32_t)accel->nodes.size();
  out->prim_refs = primRefs.empty() ? NULL : &primRefs[0];
  out->n_prim_refs = (uint32_t)primRefs.size();
  out->instances = fb.instances.empty() ? NULL : &fb.instances[0];
  out->n_instances = (uint32_t)fb.instances.size();
  for (int i = 0; i < 3; ++i) {
    out->bounds[i] = accel->bounds.pMin[i];
    out->bounds[3 + i] = accel->bounds.pMax[i];
  }
  out->prims = fb.prims.empty() ? NULL : &fb.prims[0];
  out->prim_aux = fb.aux.empty() ? NULL : &fb.aux[0];
  out->n_prims = (uint32_t)fb.prims.size();
  out->meshes = fb.meshes.empty() ? NULL : &fb.meshes[0];
  out->n_meshes = (uint32_t)fb.meshes.size();
  out->quadrics = fb.quadrics.empty() ? NULL : &fb.quadrics[0];
  out->n_quadrics = (uint32_t)fb.quadrics.size();
  out->vertex_n = fb.anyN ? &fb.vertexN[0] : NULL;
  out->vertex_s = fb.anyS ? &fb.vertexS[0] : NULL;
  out->vertex_uv = fb.anyUV ? &fb.vertexUV[0] : NULL;
  out->n_vertices = fb.nVertices;
  out->materials = fb.materials.empty() ? NULL : &fb.materials[0];
  out->n_materials = (uint32_t)fb.materials.size();
  out->lights = fb.lights.empty() ? NULL : &fb.lights[0];
  out->n_lights = (uint32_t)lights.size();
  out->n_unsampled_lights = (uint32_t)unsampled.size();
  fb.textures = textures;
  if (!fb.textures.nodes.empty()) {
    out->textures = &fb.textures.nodes[0];
    out->n_textures = (uint32_t)fb.textures.nodes.size();
    out->tex_levels = fb.textures.levels.empty() ? NULL : &fb.textures.levels[0];
    out->n_tex_levels = (uint32_t)fb.textures.levels.size();
    out->texels = fb.textures.texels.empty() ? NULL : &fb.textures.texels[0];
    out->n_texel_floats = fb.textures.texels.size();
    out->ewa_weight_lut = fb.textures.ewaLut;
  }
  if (!fb.textures.lightDistributions.empty()) {
    out->light_distributions = &fb.textures.lightDistributions[0];
    out->n_light_distribution_floats = fb.textures.lightDistributions.size();
  }
  out->light_shapes = fb.lightShapes.empty() ? NULL : &fb.lightShapes[0];
  out->light_cdf = fb.lightCdf.empty() ? NULL : &fb.lightCdf[0];
  out->n_light_shapes = (uint32_t)fb.lightShapes.size();
  camera->Flatten(&out->camera);
  sampler->Flatten(&out->sampler);
  out->bestcandidate_table = out->sampler.kind == PBRT_B200_SAMPLER_BESTCANDIDATE ? BestCandidateTable() : NULL;
  surfaceIntegrator->Flatten(&out->integrator);
  out->integrator.n_volume_1d = volumeIntegrator->NumSamples1D();
  camera->film->Flatten(&out->film);
}

// ------------------------------------------------------------ backend
namespace {

typedef int (*RenderHostFn)(const PbrtB200Scene *, int, const PbrtB200Window *, int,
                            float *const *, float *const *, PbrtB200Stats *);
typedef const char *(*LastErrorFn)(void);
typedef int (*RenderHostMultiFn)(const PbrtB200Scene *, const int *, int, float *, PbrtB200Stats *);
typedef int (*DeviceCountFn)(void);

int CudaBackend(const PbrtB200Scene *scene, const PbrtB200Window *windows, int nWindows,
                float *const *rgbaOut, PbrtB200Stats *stats) {
  static void *lib = NULL;
  static RenderHostFn renderHost = NULL;
  static LastErrorFn lastError = NULL;
  static RenderHostMultiFn renderHostMulti = NULL;
  static DeviceCountFn deviceCount = NULL;
  if (!lib) {
    vector<string> candidates;
    if (const char *env = getenv("PBRT_B200_LIB")) candidates.push_back(env);
    Dl_info info;
    if (dladdr((void *)&CudaBackend, &info) && info.dli_fname) {
      string dir(info.dli_fname);
      size_t slash = dir.rfind('/');
      dir = (slash == string::npos) ? string(".") : dir.substr(0, slash);
      candidates.push_back(dir + "/libpbrt_b200.so");
      candidates.push_back(dir + "/../csrc/libpbrt_b200.so");
    }
    candidates.push_back("libpbrt_b200.so");
    string tried;
    for (size_t i = 0; i < candidates.size() && !lib; ++i) {
      lib = dlopen(candidates[i].c_str(), RTLD_NOW | RTLD_LOCAL);
      if (!lib) tried += "\n  " + candidates[i] + ": " + dlerror();
    }
    if (!lib)
      Severe("Unable to load the CUDA render library libpbrt_b200.so (no CPU fallback exists):%s",
             tried.c_str());
    renderHost = (RenderHostFn)dlsym(lib, "pbrt_b200_render_host");
    lastError = (LastErrorFn)dlsym(lib, "pbrt_b200_last_error");
    renderHostMulti = (RenderHostMultiFn)dlsym(lib, "pbrt_b200_render_host_multi");
    deviceCount = (DeviceCountFn)dlsym(lib, "pbrt_b200_device_count");
    if (!renderHost || !lastError || !renderHostMulti || !deviceCount) Severe("libpbrt_b200.so does not export the render ABI");
  }
  // PBRT_GPUS=n: the frame is cut into tiles over the first n GPUs of this machine (or the ones listed in
  // PBRT_B200_DEVICES, comma separated) and assembled by one NCCL exchange -- no change to the scene file.
  // The reference's counterpart is n cropwindow renders pasted with tools/exrassemble.cpp.
  if (const char *g = getenv("PBRT_GPUS")) {
    int n = atoi(g);
    if (n > 1 && nWindows == 1) {
      vector<int> devices;
      if (const char *list = getenv("PBRT_B200_DEVICES")) {
        for (const char *q = list; *q;) {
          devices.push_back(atoi(q));
          while (*q && *q != ',') ++q;
          if (*q == ',') ++q;
        }
      } else
        for (int i = 0; i < n; ++i) devices.push_back(i);
      int have = deviceCount();
      if ((int)devices.size() < n || n > have) {
        Warning("PBRT_GPUS=%d but %d CUDA device(s) usable; rendering on %d", n, have, std::max(1, std::min(have, (int)devices.size())));
        n = std::max(1, std::min(have, (int)devices.size()));
      }
      if (n > 1) {
        int rc = renderHostMulti(scene, &devices[0], n, rgbaOut[0], stats);
        if (rc != PBRT_B200_OK) Severe("multi-GPU render failed (code %d): %s", rc, lastError());
        return rc;
      }
    }
  }
  int device = 0;
  if (const char *d = getenv("PBRT_B200_DEVICE")) device = atoi(d);
  else if (const char *lr = getenv("LOCAL_RANK")) device = atoi(lr);
  int rc = renderHost(scene, device, windows, nWindows, rgbaOut, NULL, stats);
  if (rc != PBRT_B200_OK) Severe("GPU render failed (code %d): %s", rc, lastError());
  return rc;
}

PbrtRenderBackend gBackend = CudaBackend;

}  // namespace

void pbrtSetRenderBackend(PbrtRenderBackend fn) { gBackend = fn ? fn : CudaBackend; }

// pbrt_host_load_scene() arms this: the next Render() exports the scene into the
// capture (deep-copying the kd arrays, which die with the Scene) and returns.
struct CapturedScene {
  CapturedScene() { memset(&scene, 0, sizeof(scene)); }
  FlatBuilder fb;
  PbrtB200Scene scene;
  vector<PbrtB200KdNode> nodes;
  vector<uint32_t> primRefs;
};
static CapturedScene *gCapture = NULL;

void Scene::Render() {
  if (gCapture) {
    CapturedScene *cap = gCapture;
    const auto f0 = std::chrono::steady_clock::now();
    Flatten(cap->fb, &cap->scene);
    const auto f1 = std::chrono::steady_clock::now();
    // the kd arrays die with the Scene: take them over when they are the accelerator's own vectors (no instancing),
    // copy them when Flatten assembled them in the builder
    KdTreeAccel *accel = dynamic_cast<KdTreeAccel *>(aggregate);
    if (accel && !accel->nodes.empty() && cap->scene.nodes == &accel->nodes[0] &&
        (accel->primRefs.empty() || cap->scene.prim_refs == &accel->primRefs[0])) {
      cap->nodes.swap(accel->nodes);
      cap->primRefs.swap(accel->primRefs);
    } else {
      cap->nodes.assign(cap->scene.nodes, cap->scene.nodes + cap->scene.n_nodes);
      cap->primRefs.assign(cap->scene.prim_refs, cap->scene.prim_refs + cap->scene.n_prim_refs);
    }
    if (getenv("PBRT_HOST_TIMING"))
      fprintf(stderr, "load: flatten %.3f s, copy of the kd arrays %.3f s\n", std::chrono::duration<double>(f1 - f0).count(),
              std::chrono::duration<double>(std::chrono::steady_clock::now() - f1).count());
    cap->scene.nodes = cap->nodes.empty() ? NULL : &cap->nodes[0];
    cap->scene.prim_refs = cap->primRefs.empty() ? NULL : &cap->primRefs[0];
    return;
  }
  FlatBuilder fb;
  PbrtB200Scene ps;
  Flatten(fb, &ps);
  PbrtB200Window win = {ps.film.x_pixel_start, ps.film.y_pixel_start, ps.film.x_pixel_count,
                        ps.film.y_pixel_count};
  vector<float> rgba((size_t)win.width * win.height * 4);
  float *outPtr = &rgba[0];
  PbrtB200Stats stats;
  memset(&stats, 0, sizeof(stats));
  int rc = gBackend(&ps, &win, 1, &outPtr, &stats);
  if (rc != 0) return;
  // the reference prints one Error per offending sample (scene.cpp:60-74)
  if (stats.nan_samples)
    Error("Not-a-number radiance value returned for %llu image sample(s).  Set to black.",
          (unsigned long long)stats.nan_samples);
  if (stats.negative_samples)
    Error("Negative luminance value returned for %llu image sample(s).  Set to black.",
          (unsigned long long)stats.negative_samples);
  if (stats.infinite_samples)
    Error("Infinite luminance value returned for %llu image sample(s).  Set to black.",
          (unsigned long long)stats.infinite_samples);
  camera->film->WriteImage(&rgba[0]);
  // StatsPrint(stdout) equivalent (api.cpp:479)
  double rays = double(stats.intersect_rays + stats.shadow_rays);
  printf("Camera:\n    Camera Rays Traced                            %llu\n",
         (unsigned long long)stats.camera_rays);
  printf("Lights:\n    Number of shadow rays traced                  %llu\n",
         (unsigned long long)stats.shadow_rays);
  printf("Shading:\n    Number of points shaded                       %llu\n",
         (unsigned long long)stats.points_shaded);
  printf("B200:\n    Rays (Intersect + IntersectP)                 %.0f\n", rays);
  printf("    Render loop (device ms)                       %.3f\n", stats.render_ms);
  if (stats.render_ms > 0)
    printf("    Mrays/s                                       %.2f\n",
           rays / (stats.render_ms * 1e3));
}

// ------------------------------------------------------ C entry points
// Used by the Python tests / bench (ctypes) to drive the host front end.
extern "C" {

// Parse `path`, flatten the scene that WorldEnd would render, and return it
// without rendering.  Returns NULL on failure.  One scene per file.
void *pbrt_host_load_scene(const char *path) {
  CapturedScene *cap = new CapturedScene;
  gCapture = cap;
  const bool timing = getenv("PBRT_HOST_TIMING") != NULL;
  const auto t0 = std::chrono::steady_clock::now();
  pbrtInit();
  bool ok = ParseFile(path);
  const auto t1 = std::chrono::steady_clock::now();
  pbrtCleanup();
  gCapture = NULL;
  if (timing)
    fprintf(stderr, "load: parse + WorldEnd (kd build, flatten) %.3f s, cleanup %.3f s\n",
            std::chrono::duration<double>(t1 - t0).count(),
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
  if (!ok || cap->scene.abi_version == 0) {
    delete cap;
    return NULL;
  }
  return cap;
}
const PbrtB200Scene *pbrt_host_scene(void *handle) {
  return handle ? &((CapturedScene *)handle)->scene : NULL;
}
void pbrt_host_free_scene(void *handle) { delete (CapturedScene *)handle; }
// Scalar facts about a captured scene for callers that treat PbrtB200Scene as
// opaque: out[0..5] = film xres, yres, crop x0, y0, width, height; out[6] =
// samples per pixel; out[7..11] = #nodes, #prim refs, #prims, #lights, #materials;
// out[12] = integrator kind.
void pbrt_host_scene_info(void *handle, int64_t *out) {
  const PbrtB200Scene &s = ((CapturedScene *)handle)->scene;
  out[0] = s.film.x_resolution; out[1] = s.film.y_resolution;
  out[2] = s.film.x_pixel_start; out[3] = s.film.y_pixel_start;
  out[4] = s.film.x_pixel_count; out[5] = s.film.y_pixel_count;
  out[6] = (int64_t)s.sampler.x_samples * s.sampler.y_samples;
  out[7] = s.n_nodes; out[8] = s.n_prim_refs; out[9] = s.n_prims; out[10] = s.n_lights;
  out[11] = s.n_materials; out[12] = s.integrator.kind;
}

// Full pipeline exactly as the `pbrt` executable runs it: parse, build,
// render on the GPU, write the EXR named by the scene's Film.
int pbrt_host_render_file(const char *path) {
  pbrtInit();
  bool ok = ParseFile(path);
  pbrtCleanup();
  return ok ? 0 : 1;
}
void pbrt_host_write_exr(const char *name, const float *rgba, int xRes, int yRes, int totalX,
                         int totalY, int xOff, int yOff) {
  size_t n = (size_t)xRes * yRes;
  vector<float> rgb(3 * n), a(n);
  for (size_t i = 0; i < n; ++i) {
    rgb[3 * i] = rgba[4 * i]; rgb[3 * i + 1] = rgba[4 * i + 1]; rgb[3 * i + 2] = rgba[4 * i + 2];
    a[i] = rgba[4 * i + 3];
  }
  WriteRGBAImage(name, &rgb[0], &a[0], xRes, yRes, totalX, totalY, xOff, yOff);
}

}  // extern "C"

