// core.h -- host-side scene objects of the B200 front end.
//
// These mirror pbrt-v1's abstract plugin kinds (core/shape.h, primitive.h,
// material.h, light.h, camera.h, sampling.h, film.h, transport.h) closely
// enough that the `Create*` factories keep their reference signatures
// (core/dynload.cpp:112-260), but none of them can trace a ray: their job is
// to validate parameters like the reference plugin does and then export
// themselves into the POD arrays of include/pbrt_b200.h ("Flatten" hooks).
#ifndef PBRT_B200_HOST_CORE_H
#define PBRT_B200_HOST_CORE_H
#include "pbrt.h"
#include "geometry.h"
#include "transform.h"
#include "paramset.h"
#include "texture.h"
#include "../../include/pbrt_b200.h"

// ---------------------------------------------------------------- shapes
class Shape : public ReferenceCounted {
 public:
  Shape(const Transform &o2w, bool ro)
      : ObjectToWorld(o2w), WorldToObject(o2w.GetInverse()), reverseOrientation(ro),
        transformSwapsHandedness(o2w.SwapsHandedness()) {}
  virtual ~Shape() {}
  virtual BBox ObjectBound() const = 0;
  virtual BBox WorldBound() const { return ObjectToWorld(ObjectBound()); }
  virtual bool CanIntersect() const { return true; }
  virtual void Refine(vector<Reference<Shape> > &) const {
    Severe("Unimplemented Shape::Refine() method called");
  }
  virtual float Area() const = 0;
  // identity of the underlying geometry (a refined Triangle is (mesh, index)): lets an area
  // light's private refinement be matched with the accelerator's primitives
  virtual void Identity(const void **owner, int *index) const { *owner = this; *index = 0; }
  // export: fill the geometry part of a primitive record
  virtual void Flatten(struct FlatBuilder &fb, PbrtB200Prim *prim, PbrtB200PrimAux *aux) const = 0;
  uint32_t OrientationFlags() const {
    return (reverseOrientation ? PBRT_B200_PF_REVERSE : 0u) |
           (transformSwapsHandedness ? PBRT_B200_PF_SWAPS : 0u);
  }
  const Transform ObjectToWorld, WorldToObject;
  const bool reverseOrientation, transformSwapsHandedness;
};

// ------------------------------------------------------------- materials
class Material : public ReferenceCounted {
 public:
  Material() { for (int i = 0; i < PBRT_B200_TEXSLOTS; ++i) tex[i] = -1; }
  virtual ~Material() {}
  virtual void Flatten(PbrtB200Material *out) const = 0;
  // texture node (TextureStore index) per operand slot, -1 where the operand is a constant
  int32_t tex[PBRT_B200_TEXSLOTS];
 protected:
  void InitRecord(PbrtB200Material *out) const {
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < PBRT_B200_TEXSLOTS; ++i) out->tex[i] = tex[i];
  }
};

// ---------------------------------------------------------------- lights
class Light {
 public:
  Light(const Transform &l2w, int ns = 1) : nSamples(max(1, ns)), LightToWorld(l2w) {}
  virtual ~Light() {}
  virtual bool IsAreaLight() const { return false; }
  virtual void Flatten(struct FlatBuilder &fb, PbrtB200Light *out) const = 0;
  const int nSamples;
 protected:
  const Transform LightToWorld;
};
class AreaLight : public Light {
 public:
  AreaLight(const Transform &light2world, const Spectrum &le, int ns, const Reference<Shape> &shape);
  bool IsAreaLight() const { return true; }
  void Flatten(struct FlatBuilder &fb, PbrtB200Light *out) const;
  Spectrum Lemit;
  vector<Reference<Shape> > shapes;   // one entry, or the ShapeSet members in ShapeSet order
  vector<float> areaCDF;
  float area;
};

// ------------------------------------------------------------ primitives
class Primitive : public ReferenceCounted {
 public:
  virtual ~Primitive() {}
  virtual BBox WorldBound() const = 0;
  virtual bool CanIntersect() const { return true; }
  virtual void Refine(vector<Reference<Primitive> > &) const {
    Severe("Unimplemented Primitive::Refinemethod called!");
  }
  void FullyRefine(vector<Reference<Primitive> > &refined) const;
};
class GeometricPrimitive : public Primitive {
 public:
  GeometricPrimitive(const Reference<Shape> &s, const Reference<Material> &m, AreaLight *a)
      : shape(s), material(m), areaLight(a) {}
  BBox WorldBound() const { return shape->WorldBound(); }
  bool CanIntersect() const { return shape->CanIntersect(); }
  void Refine(vector<Reference<Primitive> > &refined) const;
  Reference<Shape> shape;
  Reference<Material> material;
  AreaLight *areaLight;
};

// core/primitive.h:93-113: an instanced primitive (an object's accelerator, or its single
// intersectable primitive) placed by InstanceToWorld.  Intersection happens on the device.
class InstancePrimitive : public Primitive {
 public:
  InstancePrimitive(const Reference<Primitive> &i, const Transform &i2w)
      : instance(i), InstanceToWorld(i2w), WorldToInstance(i2w.GetInverse()) {}
  BBox WorldBound() const { return InstanceToWorld(instance->WorldBound()); }
  Reference<Primitive> instance;
  Transform InstanceToWorld, WorldToInstance;
};

// kd-tree accelerator (accelerators/kdtree.cpp): host-side SAH build only.
class KdTreeAccel : public Primitive {
 public:
  KdTreeAccel(const vector<Reference<Primitive> > &p, int icost, int tcost, float ebonus,
              int maxp, int maxDepth);
  BBox WorldBound() const { return bounds; }
  vector<Reference<Primitive> > prims;   // fully refined, build order
  vector<PbrtB200KdNode> nodes;
  vector<uint32_t> primRefs;
  BBox bounds;
  struct BoundEdge;   // (kdtree.cpp)
 private:
  int isectCost, traversalCost, maxPrims;
  float emptyBonus;
};

// ---------------------------------------------------- film, filter, camera
class Filter {
 public:
  virtual ~Filter() {}
  Filter(float xw, float yw) : xWidth(xw), yWidth(yw), invXWidth(1.f / xw), invYWidth(1.f / yw) {}
  virtual float Evaluate(float x, float y) const = 0;
  const float xWidth, yWidth;
  const float invXWidth, invYWidth;
};
class Film {
 public:
  Film(int xres, int yres) : xResolution(xres), yResolution(yres) {}
  virtual ~Film() {}
  virtual void GetSampleExtent(int *xstart, int *xend, int *ystart, int *yend) const = 0;
  virtual void Flatten(PbrtB200Film *out) const = 0;
  // receives what the reference's WriteImage computes (film/image.cpp:157-209)
  virtual void WriteImage(const float *rgba) = 0;
  const int xResolution, yResolution;
};
class Camera {
 public:
  Camera(const Transform &world2cam, float hither, float yon, float sopen, float sclose, Film *f)
      : WorldToCamera(world2cam), CameraToWorld(world2cam.GetInverse()), ClipHither(hither),
        ClipYon(yon), ShutterOpen(sopen), ShutterClose(sclose), film(f) {}
  virtual ~Camera() { delete film; }
  virtual void Flatten(PbrtB200Camera *out) const = 0;
  Transform WorldToCamera, CameraToWorld;
  float ClipHither, ClipYon, ShutterOpen, ShutterClose;
  Film *film;
};
class Sampler {
 public:
  Sampler(int xstart, int xend, int ystart, int yend, int spp)
      : xPixelStart(xstart), xPixelEnd(xend), yPixelStart(ystart), yPixelEnd(yend),
        samplesPerPixel(spp) {}
  virtual ~Sampler() {}
  virtual void Flatten(PbrtB200Sampler *out) const = 0;
  int TotalSamples() const {
    return samplesPerPixel * (xPixelEnd - xPixelStart) * (yPixelEnd - yPixelStart);
  }
  int xPixelStart, xPixelEnd, yPixelStart, yPixelEnd;
  int samplesPerPixel;
};
class SurfaceIntegrator {
 public:
  virtual ~SurfaceIntegrator() {}
  virtual void Flatten(PbrtB200Integrator *out) const = 0;
};
class VolumeIntegrator {
 public:
  virtual ~VolumeIntegrator() {}
  virtual int NumSamples1D() const = 0;
};
class VolumeRegion {
 public:
  virtual ~VolumeRegion() {}
};

// --------------------------------------------------------------- scene
// Collects the POD arrays while the object graph is being exported.
struct FlatBuilder {
  vector<PbrtB200Prim> prims;
  vector<PbrtB200PrimAux> aux;
  vector<PbrtB200Mesh> meshes;
  vector<PbrtB200Quadric> quadrics;
  vector<PbrtB200Instance> instances;
  vector<PbrtB200KdNode> nodes;        // top-level tree, then the instanced objects' trees (only filled when
  vector<uint32_t> primRefs;           // the scene has instances; otherwise the accelerator's own arrays are used)
  map<const void *, uint32_t> objectIds;   // instanced Primitive -> first PbrtB200Instance built for it
  vector<Reference<Primitive> > keepAlive;
  vector<float> vertexN, vertexS, vertexUV;
  uint32_t nVertices;
  bool anyN, anyS, anyUV;
  vector<PbrtB200Material> materials;
  vector<PbrtB200Light> lights;
  vector<uint32_t> lightShapes;
  vector<float> lightCdf;
  map<const void *, uint32_t> meshIds, meshVertexBase, materialIds;
  map<std::pair<const void *, int>, uint32_t> shapePrimIds;
  map<const void *, int32_t> lightIds;
  TextureStore textures;
  FlatBuilder() : nVertices(0), anyN(false), anyS(false), anyUV(false) {}
};

class Scene {
 public:
  Scene(Camera *cam, SurfaceIntegrator *si, VolumeIntegrator *vi, Sampler *s, Primitive *accel,
        const vector<Light *> &lts, VolumeRegion *vr);
  ~Scene();
  void Render();   // core/scene.h:34 -- the drop-in seam
  // Export into POD arrays; the returned struct points into `flat`.
  void Flatten(FlatBuilder &flat, PbrtB200Scene *out) const;
  Primitive *aggregate;
  vector<Light *> lights;
  Camera *camera;
  VolumeRegion *volumeRegion;
  SurfaceIntegrator *surfaceIntegrator;
  VolumeIntegrator *volumeIntegrator;
  Sampler *sampler;
  BBox bound;
  TextureStore textures;   // texture table of the world block this scene was made from
};

// ------------------------------------------------- plugin registry
// core/dynload.h:29-61: same Make* entry points; plugin names resolve through
// a static table instead of dlopen("<name>.so").
Reference<Shape> MakeShape(const string &name, const Transform &object2world,
                           bool reverseOrientation, const ParamSet &paramSet);
Reference<Material> MakeMaterial(const string &name, const Transform &mtl2world,
                                 const TextureParams &mp);
Light *MakeLight(const string &name, const Transform &light2world, const ParamSet &paramSet);
AreaLight *MakeAreaLight(const string &name, const Transform &light2world,
                         const ParamSet &paramSet, const Reference<Shape> &shape);
VolumeRegion *MakeVolumeRegion(const string &name, const Transform &volume2world,
                               const ParamSet &paramSet);
SurfaceIntegrator *MakeSurfaceIntegrator(const string &name, const ParamSet &paramSet);
VolumeIntegrator *MakeVolumeIntegrator(const string &name, const ParamSet &paramSet);
Primitive *MakeAccelerator(const string &name, const vector<Reference<Primitive> > &prims,
                           const ParamSet &paramSet);
Camera *MakeCamera(const string &name, const ParamSet &paramSet, const Transform &world2cam,
                   Film *film);
Sampler *MakeSampler(const string &name, const ParamSet &paramSet, const Film *film);
Filter *MakeFilter(const string &name, const ParamSet &paramSet);
Film *MakeFilm(const string &name, const ParamSet &paramSet, Filter *filter);
void UpdatePluginPath(const string &newpath);
const float *BestCandidateTable();   // 4096 x 5 floats (samplers/sampledata.cpp's table, regenerated)

// EXR output (core/exrio.cpp:75-96)
void WriteRGBAImage(const string &name, const float *pixels, const float *alpha, int xRes,
                    int yRes, int totalXRes, int totalYRes, int xOffset, int yOffset);

// Render backend hook: Scene::Render() calls this with the flattened scene.
// The default implementation dlopen()s libpbrt_b200.so (the CUDA library) and
// fails loudly when it is missing -- there is no CPU fallback.
typedef int (*PbrtRenderBackend)(const PbrtB200Scene *scene, const PbrtB200Window *windows,
                                 int nWindows, float *const *rgbaOut, PbrtB200Stats *stats);
void pbrtSetRenderBackend(PbrtRenderBackend fn);

#endif
