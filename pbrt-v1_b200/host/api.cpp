// api.cpp -- the api.h state machine: option/world blocks, CTM and attribute
// stacks, named coordinate systems, per-shape material/area-light
// instantiation, and WorldEnd -> MakeScene -> Scene::Render.
// Behavioural contract: core/api.cpp:39-529 (cited per function).
#include "api.h"
#include "core.h"
#include <stdarg.h>

extern int line_num;
extern string current_file;

// ------------------------------------------------------------ diagnostics
// core/util.cpp:36-97: "<Severity>: <message>" + optional "\tLine n, file f".
static void Report(const char *severity, const char *format, va_list args, bool fatal) {
  char *buf = NULL;
  if (vasprintf(&buf, format, args) < 0) buf = NULL;
  fprintf(stderr, "%s: %s\n", severity, buf ? buf : format);
  if (line_num != 0) fprintf(stderr, "\tLine %d, file %s\n", line_num, current_file.c_str());
  free(buf);
  if (fatal) abort();
}
void Info(const char *format, ...) {
  va_list a; va_start(a, format); Report("Notice", format, a, false); va_end(a);
}
void Warning(const char *format, ...) {
  va_list a; va_start(a, format); Report("Warning", format, a, false); va_end(a);
}
void Error(const char *format, ...) {
  va_list a; va_start(a, format); Report("Error", format, a, false); va_end(a);
}
void Severe(const char *format, ...) {
  va_list a; va_start(a, format); Report("Fatal Error", format, a, true); va_end(a);
}

// ------------------------------------------------------------------ state
bool ShapeCanSample(const Shape *s);   // shapes.cpp
namespace {

struct RenderOptions {          // core/api.cpp:39-88
  RenderOptions()
      : FilterName("mitchell"), FilmName("image"), SamplerName("bestcandidate"),
        AcceleratorName("kdtree"), SurfIntegratorName("directlighting"),
        VolIntegratorName("emission"), CameraName("perspective"), currentInstance(NULL) {}
  Scene *MakeScene() const;
  TextureStore textures;          // nodes referenced by this world block's materials
  string FilterName; ParamSet FilterParams;
  string FilmName; ParamSet FilmParams;
  string SamplerName; ParamSet SamplerParams;
  string AcceleratorName; ParamSet AcceleratorParams;
  string SurfIntegratorName, VolIntegratorName;
  ParamSet SurfIntegratorParams, VolIntegratorParams;
  string CameraName; ParamSet CameraParams;
  Transform WorldToCamera;
  mutable vector<Light *> lights;
  mutable vector<Reference<Primitive> > primitives;
  mutable vector<VolumeRegion *> volumeRegions;
  map<string, vector<Reference<Primitive> > > instances;
  vector<Reference<Primitive> > *currentInstance;
};

struct GraphicsState {          // core/api.cpp:89-107
  GraphicsState() : material("matte"), reverseOrientation(false) {}
  map<string, Reference<FloatTexture> > floatTextures;
  map<string, Reference<SpectrumTexture> > spectrumTextures;
  ParamSet materialParams;
  string material;
  ParamSet areaLightParams;
  string areaLight;
  bool reverseOrientation;
};

enum { STATE_UNINITIALIZED, STATE_OPTIONS_BLOCK, STATE_WORLD_BLOCK };
int currentApiState = STATE_UNINITIALIZED;
Transform curTransform;
map<string, Transform> namedCoordinateSystems;
RenderOptions *renderOptions = NULL;
GraphicsState graphicsState;
vector<GraphicsState> pushedGraphicsStates;
vector<Transform> pushedTransforms;

bool Initialized(const char *func) {
  if (currentApiState == STATE_UNINITIALIZED) {
    Error("pbrtInit() must be before calling \"%s()\". Ignoring.", func);
    return false;
  }
  return true;
}
bool InOptions(const char *func) {
  if (!Initialized(func)) return false;
  if (currentApiState == STATE_WORLD_BLOCK) {
    Error("Options cannot be set inside world block; \"%s\" not allowed.  Ignoring.", func);
    return false;
  }
  return true;
}
bool InWorld(const char *func) {
  if (!Initialized(func)) return false;
  if (currentApiState == STATE_OPTIONS_BLOCK) {
    Error("Scene description must be inside world block; \"%s\" not allowed. Ignoring.", func);
    return false;
  }
  return true;
}
// api.cpp:183-198: the 16 floats arrive column-major.
Transform FromColumnMajor(float tr[16]) {
  return Transform(new Matrix4x4(tr[0], tr[4], tr[8], tr[12], tr[1], tr[5], tr[9], tr[13], tr[2],
                                 tr[6], tr[10], tr[14], tr[3], tr[7], tr[11], tr[15]));
}

}  // namespace

void pbrtInit() {               // api.cpp:141-162
  if (currentApiState != STATE_UNINITIALIZED) Error("pbrtInit() has already been called.");
  currentApiState = STATE_OPTIONS_BLOCK;
  renderOptions = new RenderOptions;
  graphicsState = GraphicsState();
}
void pbrtCleanup() {            // api.cpp:163-173
  if (currentApiState == STATE_UNINITIALIZED) Error("pbrtCleanup() called without pbrtInit().");
  else if (currentApiState == STATE_WORLD_BLOCK)
    Error("pbrtCleanup() called while inside world block.");
  currentApiState = STATE_UNINITIALIZED;
  delete renderOptions;
  renderOptions = NULL;
}
void pbrtIdentity() { if (Initialized("Identity")) curTransform = Transform(); }
void pbrtTranslate(float dx, float dy, float dz) {
  if (Initialized("Translate")) curTransform = curTransform * Translate(Vector(dx, dy, dz));
}
void pbrtTransform(float tr[16]) {
  if (Initialized("Transform")) curTransform = FromColumnMajor(tr);
}
void pbrtConcatTransform(float tr[16]) {
  if (Initialized("ConcatTransform")) curTransform = curTransform * FromColumnMajor(tr);
}
void pbrtRotate(float angle, float dx, float dy, float dz) {
  if (Initialized("Rotate")) curTransform = curTransform * Rotate(angle, Vector(dx, dy, dz));
}
void pbrtScale(float sx, float sy, float sz) {
  if (Initialized("Scale")) curTransform = curTransform * Scale(sx, sy, sz);
}
void pbrtLookAt(float ex, float ey, float ez, float lx, float ly, float lz, float ux, float uy,
                float uz) {
  if (Initialized("LookAt"))
    curTransform = curTransform * LookAt(Point(ex, ey, ez), Point(lx, ly, lz), Vector(ux, uy, uz));
}
void pbrtCoordinateSystem(const string &name) {
  if (Initialized("CoordinateSystem")) namedCoordinateSystems[name] = curTransform;
}
void pbrtCoordSysTransform(const string &name) {
  if (!Initialized("CoordSysTransform")) return;
  if (namedCoordinateSystems.find(name) != namedCoordinateSystems.end())
    curTransform = namedCoordinateSystems[name];
}
void pbrtPixelFilter(const string &name, const ParamSet &params) {
  if (!InOptions("PixelFilter")) return;
  renderOptions->FilterName = name; renderOptions->FilterParams = params;
}
void pbrtFilm(const string &type, const ParamSet &params) {
  if (!InOptions("Film")) return;
  renderOptions->FilmParams = params; renderOptions->FilmName = type;
}
void pbrtSampler(const string &name, const ParamSet &params) {
  if (!InOptions("Sampler")) return;
  renderOptions->SamplerName = name; renderOptions->SamplerParams = params;
}
void pbrtAccelerator(const string &name, const ParamSet &params) {
  if (!InOptions("Accelerator")) return;
  renderOptions->AcceleratorName = name; renderOptions->AcceleratorParams = params;
}
void pbrtSurfaceIntegrator(const string &name, const ParamSet &params) {
  if (!InOptions("SurfaceIntegrator")) return;
  renderOptions->SurfIntegratorName = name; renderOptions->SurfIntegratorParams = params;
}
void pbrtVolumeIntegrator(const string &name, const ParamSet &params) {
  if (!InOptions("VolumeIntegrator")) return;
  renderOptions->VolIntegratorName = name; renderOptions->VolIntegratorParams = params;
}
void pbrtCamera(const string &name, const ParamSet &params) {   // api.cpp:254-262
  if (!InOptions("Camera")) return;
  renderOptions->CameraName = name;
  renderOptions->CameraParams = params;
  renderOptions->WorldToCamera = curTransform;
  namedCoordinateSystems["camera"] = curTransform.GetInverse();
}
void pbrtSearchPath(const string &path) {
  if (!InOptions("SearchPath")) return;
  UpdatePluginPath(path);
}
void pbrtWorldBegin() {          // api.cpp:268-273
  if (!InOptions("WorldBegin")) return;
  currentApiState = STATE_WORLD_BLOCK;
  curTransform = Transform();
  namedCoordinateSystems["world"] = curTransform;
}
void pbrtAttributeBegin() {
  if (!InWorld("AttributeBegin")) return;
  pushedGraphicsStates.push_back(graphicsState);
  pushedTransforms.push_back(curTransform);
}
void pbrtAttributeEnd() {
  if (!InWorld("AttributeEnd")) return;
  if (!pushedGraphicsStates.size()) {
    Error("Unmatched pbrtAttributeEnd() encountered. Ignoring it.");
    return;
  }
  graphicsState = pushedGraphicsStates.back();
  curTransform = pushedTransforms.back();
  pushedGraphicsStates.pop_back();
  pushedTransforms.pop_back();
}
void pbrtTransformBegin() {
  if (InWorld("TransformBegin")) pushedTransforms.push_back(curTransform);
}
void pbrtTransformEnd() {
  if (!InWorld("TransformEnd")) return;
  if (!pushedTransforms.size()) {
    Error("Unmatched pbrtTransformEnd() encountered. Ignoring it.");
    return;
  }
  curTransform = pushedTransforms.back();
  pushedTransforms.pop_back();
}
// api.cpp:305-332.  Texture classes in scope: "constant" (textures/constant.cpp), "imagemap"
// (textures/imagemap.cpp), and "scale" / "mix" over them (textures/scale.cpp:38-41, textures/mix.cpp:40-44).
// An expression whose operands are all constant has the same value at every point and is folded here once;
// anything that depends on an image map becomes a node of the scene's texture table (texture.h).
namespace {
int OperandNode(TextureStore &st, const Spectrum &v, int32_t node) {
  return node >= 0 ? node : st.AddConstant(v.c, 3);
}
int OperandNode(TextureStore &st, float v, int32_t node) {
  float c[3] = {v, v, v};
  return node >= 0 ? node : st.AddConstant(c, 1);
}
}  // namespace
TextureStore &CurrentTextureStore() { return renderOptions->textures; }
void pbrtTexture(const string &name, const string &type, const string &texname,
                 const ParamSet &params) {
  if (!InWorld("Texture")) return;
  TextureParams tp(params, params, graphicsState.floatTextures, graphicsState.spectrumTextures);
  if (type != "float" && type != "color") {
    Error("Texture type \"%s\" unknown.", type.c_str());
    return;
  }
  const bool procedural = texname == "uv" || texname == "bilerp" || texname == "checkerboard" || texname == "dots" ||
                          texname == "fbm" || texname == "wrinkled" || texname == "windy" || texname == "marble";
  if (texname != "constant" && texname != "scale" && texname != "mix" && texname != "imagemap" && !procedural) {
    Error("Texture plugin \"%s\" unknown; texture \"%s\" not created.", texname.c_str(), name.c_str());
    return;
  }
  TextureStore &store = renderOptions->textures;
  int node = -1;
  if (procedural) {
    // checkerboard.cpp:178-179,223-224: tex1 = 1, tex2 = 0; dots.cpp:89-90,116-117: inside = 1, outside = 0
    const bool two = texname == "checkerboard" || texname == "dots";
    const char *n0 = texname == "dots" ? "inside" : "tex1", *n1 = texname == "dots" ? "outside" : "tex2";
    int a0 = -1, a1 = -1;
    if (two && type == "float") {
      int32_t m0, m1;
      float v0 = tp.GetFloatTexture(n0, 1.f, &m0), v1 = tp.GetFloatTexture(n1, 0.f, &m1);
      a0 = OperandNode(store, v0, m0); a1 = OperandNode(store, v1, m1);
    } else if (two) {
      int32_t m0, m1;
      Spectrum v0 = tp.GetSpectrumTexture(n0, Spectrum(1.f), &m0), v1 = tp.GetSpectrumTexture(n1, Spectrum(0.f), &m1);
      a0 = OperandNode(store, v0, m0); a1 = OperandNode(store, v1, m1);
    }
    node = store.AddProcedural(texname, curTransform, tp, type == "float" ? 1 : 3, a0, a1);
    if (node < 0) {
      Error("Unable to create %s texture \"%s\" of class \"%s\".", type.c_str(), name.c_str(), texname.c_str());
      return;
    }
    if (type == "float") {
      if (graphicsState.floatTextures.find(name) != graphicsState.floatTextures.end())
        Warning("Texture \"%s\" being redefined", name.c_str());
      FloatTexture *ft = new FloatTexture;
      ft->node = node;
      graphicsState.floatTextures[name] = ft;
    } else {
      if (graphicsState.spectrumTextures.find(name) != graphicsState.spectrumTextures.end())
        Warning("Texture \"%s\" being redefined", name.c_str());
      SpectrumTexture *st = new SpectrumTexture;
      st->node = node;
      graphicsState.spectrumTextures[name] = st;
    }
    if (store.Depth(node) > 3)
      Severe("Texture \"%s\": expressions more than 3 operators deep are not supported", name.c_str());
    tp.ReportUnused();
    return;
  }
  if (type == "float") {
    if (graphicsState.floatTextures.find(name) != graphicsState.floatTextures.end())
      Warning("Texture \"%s\" being redefined", name.c_str());
    FloatTexture *ft = new FloatTexture;
    if (texname == "constant") ft->value = tp.FindFloat("value", 1.f);
    else if (texname == "imagemap") node = store.AddImageMap(curTransform, tp, 1);
    else if (texname == "scale") {
      int32_t n1, n2;
      float a = tp.GetFloatTexture("tex1", 1.f, &n1), b = tp.GetFloatTexture("tex2", 1.f, &n2);
      if (n1 < 0 && n2 < 0) ft->value = a * b;
      else node = store.AddScale(OperandNode(store, a, n1), OperandNode(store, b, n2), 1);
    } else {
      int32_t n1, n2, na;
      float t1 = tp.GetFloatTexture("tex1", 0.f, &n1), t2 = tp.GetFloatTexture("tex2", 1.f, &n2);
      float amt = tp.GetFloatTexture("amount", 0.5f, &na);
      if (n1 < 0 && n2 < 0 && na < 0) ft->value = (1.f - amt) * t1 + amt * t2;
      else node = store.AddMix(OperandNode(store, t1, n1), OperandNode(store, t2, n2), OperandNode(store, amt, na), 1);
    }
    ft->node = node;
    graphicsState.floatTextures[name] = ft;
  } else {
    if (graphicsState.spectrumTextures.find(name) != graphicsState.spectrumTextures.end())
      Warning("Texture \"%s\" being redefined", name.c_str());
    SpectrumTexture *st = new SpectrumTexture;
    if (texname == "constant") st->value = tp.FindSpectrum("value", Spectrum(1.f));
    else if (texname == "imagemap") node = store.AddImageMap(curTransform, tp, 3);
    else if (texname == "scale") {
      int32_t n1, n2;
      Spectrum a = tp.GetSpectrumTexture("tex1", Spectrum(1.f), &n1), b = tp.GetSpectrumTexture("tex2", Spectrum(1.f), &n2);
      if (n1 < 0 && n2 < 0) for (int i = 0; i < COLOR_SAMPLES; ++i) st->value.c[i] = a.c[i] * b.c[i];
      else node = store.AddScale(OperandNode(store, a, n1), OperandNode(store, b, n2), 3);
    } else {
      int32_t n1, n2, na;
      Spectrum t1 = tp.GetSpectrumTexture("tex1", Spectrum(0.f), &n1), t2 = tp.GetSpectrumTexture("tex2", Spectrum(1.f), &n2);
      float amt = tp.GetFloatTexture("amount", 0.5f, &na);
      if (n1 < 0 && n2 < 0 && na < 0) for (int i = 0; i < COLOR_SAMPLES; ++i) st->value.c[i] = t1.c[i] * (1.f - amt) + t2.c[i] * amt;
      else node = store.AddMix(OperandNode(store, t1, n1), OperandNode(store, t2, n2), OperandNode(store, amt, na), 3);
    }
    st->node = node;
    graphicsState.spectrumTextures[name] = st;
  }
  if (node >= 0 && store.Depth(node) > 3)
    Severe("Texture \"%s\": expressions more than 3 scale/mix operators deep are not supported", name.c_str());
  tp.ReportUnused();
}
void pbrtMaterial(const string &name, const ParamSet &params) {
  if (!InWorld("Material")) return;
  graphicsState.material = name;
  graphicsState.materialParams = params;
}
void pbrtLightSource(const string &name, const ParamSet &params) {
  if (!InWorld("LightSource")) return;
  Light *lt = MakeLight(name, curTransform, params);
  if (lt == NULL) Error("pbrtLightSource: light type \"%s\" unknown.", name.c_str());
  else renderOptions->lights.push_back(lt);
}
void pbrtAreaLightSource(const string &name, const ParamSet &params) {
  if (!InWorld("AreaLightSource")) return;
  graphicsState.areaLight = name;
  graphicsState.areaLightParams = params;
}
void pbrtShape(const string &name, const ParamSet &params) {   // api.cpp:354-396
  if (!InWorld("Shape")) return;
  Reference<Shape> shape = MakeShape(name, curTransform, graphicsState.reverseOrientation, params);
  if (!shape) return;
  params.ReportUnused();
  AreaLight *area = NULL;
  if (graphicsState.areaLight != "")
    area = MakeAreaLight(graphicsState.areaLight, curTransform, graphicsState.areaLightParams,
                         shape);
  TextureParams mp(params, graphicsState.materialParams, graphicsState.floatTextures,
                   graphicsState.spectrumTextures);
  Reference<Material> mtl = MakeMaterial(graphicsState.material, curTransform, mp);
  if (!mtl) mtl = MakeMaterial("matte", curTransform, mp);
  if (!mtl) Severe("Unable to create \"matte\" material?!");
  Reference<Primitive> prim = new GeometricPrimitive(shape, mtl, area);
  if (renderOptions->currentInstance) {
    if (area) Warning("Area lights not supported with object instancing");
    renderOptions->currentInstance->push_back(prim);
  } else {
    renderOptions->primitives.push_back(prim);
    if (area != NULL) {
      if (!ShapeCanSample(shape.operator->()))   // core/shape.h:85-88: what the reference dies with at the first light sample
        Severe("Unimplemented Shape::Sample method called");
      renderOptions->lights.push_back(area);
    }
  }
}
void pbrtReverseOrientation() {
  if (InWorld("ReverseOrientation"))
    graphicsState.reverseOrientation = !graphicsState.reverseOrientation;
}
void pbrtVolume(const string &name, const ParamSet &params) {
  if (!InWorld("Volume")) return;
  VolumeRegion *vr = MakeVolumeRegion(name, curTransform, params);
  if (vr) renderOptions->volumeRegions.push_back(vr);
}
void pbrtObjectBegin(const string &name) {    // api.cpp:409-419
  if (!InWorld("ObjectBegin")) return;
  pbrtAttributeBegin();
  if (renderOptions->currentInstance) Error("ObjectBegin called inside of instance definition");
  renderOptions->instances[name] = vector<Reference<Primitive> >();
  renderOptions->currentInstance = &renderOptions->instances[name];
}
void pbrtObjectEnd() {
  if (!InWorld("ObjectEnd")) return;
  if (!renderOptions->currentInstance) Error("ObjectEnd called outside of instance definition");
  renderOptions->currentInstance = NULL;
  pbrtAttributeEnd();
}
// api.cpp:428-457
void pbrtObjectInstance(const string &name) {
  if (!InWorld("ObjectInstance")) return;
  if (renderOptions->currentInstance) {
    Error("ObjectInstance can't be called inside instance definition");
    return;
  }
  if (renderOptions->instances.find(name) == renderOptions->instances.end()) {
    Error("Unable to find instance named \"%s\"", name.c_str());
    return;
  }
  vector<Reference<Primitive> > &in = renderOptions->instances[name];
  if (in.size() == 0) return;
  if (in.size() > 1 || !in[0]->CanIntersect()) {
    // Refine instance Primitives and create aggregate
    Reference<Primitive> accel =
        MakeAccelerator(renderOptions->AcceleratorName, in, renderOptions->AcceleratorParams);
    if (!accel) accel = MakeAccelerator("kdtree", in, ParamSet());
    if (!accel) Severe("Unable to find \"kdtree\" accelerator");
    in.erase(in.begin(), in.end());
    in.push_back(accel);
  }
  Reference<Primitive> prim = new InstancePrimitive(in[0], curTransform);
  renderOptions->primitives.push_back(prim);
}
void pbrtWorldEnd() {            // api.cpp:458-483
  if (!InWorld("WorldEnd")) return;
  while (pushedGraphicsStates.size()) {
    Warning("Missing end to pbrtAttributeBegin()");
    pushedGraphicsStates.pop_back();
    pushedTransforms.pop_back();
  }
  Scene *scene = renderOptions->MakeScene();
  if (scene) scene->Render();
  delete scene;
  renderOptions->textures.Clear();
  currentApiState = STATE_OPTIONS_BLOCK;
  curTransform = Transform();
  namedCoordinateSystems.erase(namedCoordinateSystems.begin(), namedCoordinateSystems.end());
}

// api.cpp:484-529: construction order Filter, Film, Camera, Sampler,
// SurfaceIntegrator, VolumeIntegrator, Accelerator.
Scene *RenderOptions::MakeScene() const {
  Filter *filter = MakeFilter(FilterName, FilterParams);
  Film *film = filter ? MakeFilm(FilmName, FilmParams, filter) : NULL;
  Camera *camera = film ? MakeCamera(CameraName, CameraParams, WorldToCamera, film) : NULL;
  Sampler *sampler = film ? MakeSampler(SamplerName, SamplerParams, film) : NULL;
  SurfaceIntegrator *surfaceIntegrator =
      MakeSurfaceIntegrator(SurfIntegratorName, SurfIntegratorParams);
  VolumeIntegrator *volumeIntegrator = MakeVolumeIntegrator(VolIntegratorName, VolIntegratorParams);
  Primitive *accelerator = MakeAccelerator(AcceleratorName, primitives, AcceleratorParams);
  if (!accelerator) {
    ParamSet ps;
    accelerator = MakeAccelerator("kdtree", primitives, ps);
  }
  if (!accelerator) Severe("Unable to find \"kdtree\" accelerator");
  VolumeRegion *volumeRegion = NULL;
  if (!camera || !sampler || !film || !accelerator || !filter || !surfaceIntegrator ||
      !volumeIntegrator) {
    Severe("Unable to create scene due to missing plug-ins");
    return NULL;
  }
  Scene *ret = new Scene(camera, surfaceIntegrator, volumeIntegrator, sampler, accelerator, lights,
                         volumeRegion);
  ret->textures = textures;
  primitives.erase(primitives.begin(), primitives.end());
  lights.erase(lights.begin(), lights.end());
  volumeRegions.erase(volumeRegions.begin(), volumeRegions.end());
  return ret;
}
