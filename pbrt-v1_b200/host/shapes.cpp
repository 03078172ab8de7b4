// shapes.cpp -- the quadric ("sphere", "cylinder", "disk", "cone", "paraboloid", "hyperboloid") and
// "trianglemesh" shape plugins, host half.
// Parameter handling follows shapes/sphere.cpp:252-263 and
// shapes/trianglemesh.cpp:350-405; the intersection code of those plugins is
// what the CUDA kernels implement (pbrt-v1_b200/csrc).
#include "core.h"

Shape *CreateTriangleMeshShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params);

namespace {

void CopyMatrix(const Matrix4x4 &m, float out[16]) { memcpy(out, m.m, 16 * sizeof(float)); }

// ------------------------------------------------------------------ quadrics
// Common export of a quadric: one PbrtB200Quadric record + a primitive of kind QUADRIC.
class Quadric : public Shape {
 public:
  Quadric(const Transform &o2w, bool ro) : Shape(o2w, ro) {}
  virtual void Fill(PbrtB200Quadric *q) const = 0;
  virtual bool CanSample() const { return false; }   // Shape::Sample is only implemented by some (core/shape.h:85-88)
  void Flatten(FlatBuilder &fb, PbrtB200Prim *prim, PbrtB200PrimAux *aux) const {
    PbrtB200Quadric q;
    memset(&q, 0, sizeof(q));
    CopyMatrix(ObjectToWorld.GetMatrix(), q.o2w);
    CopyMatrix(WorldToObject.GetMatrix(), q.w2o);
    q.flags = OrientationFlags();
    Fill(&q);
    memset(prim->p, 0, sizeof(prim->p));
    prim->flags = PBRT_B200_PRIM_QUADRIC | OrientationFlags();
    aux->v[0] = aux->v[1] = aux->v[2] = 0;
    aux->owner = (uint32_t)fb.quadrics.size();
    fb.quadrics.push_back(q);
  }
};

class Sphere : public Quadric {
 public:
  // shapes/sphere.cpp:80-92
  Sphere(const Transform &o2w, bool ro, float rad, float z0, float z1, float pm) : Quadric(o2w, ro) {
    radius = rad;
    zmin = Clamp(min(z0, z1), -radius, radius);
    zmax = Clamp(max(z0, z1), -radius, radius);
    thetaMin = acosf(Clamp(zmin / radius, -1.f, 1.f));
    thetaMax = acosf(Clamp(zmax / radius, -1.f, 1.f));
    phiMax = Radians(Clamp(pm, 0.0f, 360.0f));
  }
  BBox ObjectBound() const {
    return BBox(Point(-radius, -radius, zmin), Point(radius, radius, zmax));
  }
  float Area() const { return phiMax * radius * (zmax - zmin); }
  bool CanSample() const { return true; }
  void Fill(PbrtB200Quadric *q) const {
    q->kind = PBRT_B200_QUADRIC_SPHERE;
    q->radius = radius; q->zmin = zmin; q->zmax = zmax;
    q->theta_min = thetaMin; q->theta_max = thetaMax; q->phi_max = phiMax;
  }
 private:
  float radius, phiMax, zmin, zmax, thetaMin, thetaMax;
};

class Cylinder : public Quadric {   // shapes/cylinder.cpp:51-64,218-220
 public:
  Cylinder(const Transform &o2w, bool ro, float rad, float z0, float z1, float pm) : Quadric(o2w, ro) {
    radius = rad;
    zmin = min(z0, z1);
    zmax = max(z0, z1);
    phiMax = Radians(Clamp(pm, 0.0f, 360.0f));
  }
  BBox ObjectBound() const { return BBox(Point(-radius, -radius, zmin), Point(radius, radius, zmax)); }
  float Area() const { return (zmax - zmin) * phiMax * radius; }
  bool CanSample() const { return true; }
  void Fill(PbrtB200Quadric *q) const {
    q->kind = PBRT_B200_QUADRIC_CYLINDER;
    q->radius = radius; q->zmin = zmin; q->zmax = zmax; q->phi_max = phiMax;
  }
 private:
  float radius, zmin, zmax, phiMax;
};

class Disk : public Quadric {   // shapes/disk.cpp:49-61,147-150
 public:
  Disk(const Transform &o2w, bool ro, float ht, float r, float ri, float tmax) : Quadric(o2w, ro) {
    height = ht;
    radius = r;
    innerRadius = ri;
    phiMax = Radians(Clamp(tmax, 0.0f, 360.0f));
  }
  BBox ObjectBound() const { return BBox(Point(-radius, -radius, height), Point(radius, radius, height)); }
  float Area() const { return phiMax * 0.5f * (radius * radius - innerRadius * innerRadius); }
  bool CanSample() const { return true; }
  void Fill(PbrtB200Quadric *q) const {
    q->kind = PBRT_B200_QUADRIC_DISK;
    q->radius = radius; q->height = height; q->inner_radius = innerRadius; q->phi_max = phiMax;
  }
 private:
  float height, radius, innerRadius, phiMax;
};

class Cone : public Quadric {   // shapes/cone.cpp:41-53,182-186
 public:
  Cone(const Transform &o2w, bool ro, float ht, float rad, float tm) : Quadric(o2w, ro) {
    radius = rad;
    height = ht;
    phiMax = Radians(Clamp(tm, 0.0f, 360.0f));
  }
  BBox ObjectBound() const { return BBox(Point(-radius, -radius, 0), Point(radius, radius, height)); }
  float Area() const {
    return phiMax * height * height * sqrtf((height * height) + (radius * radius)) / (2.0f * radius);
  }
  void Fill(PbrtB200Quadric *q) const {
    q->kind = PBRT_B200_QUADRIC_CONE;
    q->radius = radius; q->height = height; q->phi_max = phiMax;
  }
 private:
  float radius, height, phiMax;
};

class Paraboloid : public Quadric {   // shapes/paraboloid.cpp:43-57,187-190
 public:
  Paraboloid(const Transform &o2w, bool ro, float rad, float z0, float z1, float tm) : Quadric(o2w, ro) {
    radius = rad;
    zmin = min(z0, z1);
    zmax = max(z0, z1);
    phiMax = Radians(Clamp(tm, 0.0f, 360.0f));
  }
  BBox ObjectBound() const { return BBox(Point(-radius, -radius, zmin), Point(radius, radius, zmax)); }
  float Area() const { return phiMax / 12.0f * (powf(1 + 4 * zmin, 1.5f) - powf(1 + 4 * zmax, 1.5f)); }
  void Fill(PbrtB200Quadric *q) const {
    q->kind = PBRT_B200_QUADRIC_PARABOLOID;
    q->radius = radius; q->zmin = zmin; q->zmax = zmax; q->phi_max = phiMax;
  }
 private:
  float radius, zmin, zmax, phiMax;
};

class Hyperboloid : public Quadric {   // shapes/hyperboloid.cpp:46-76,224-237
 public:
  Hyperboloid(const Transform &o2w, bool ro, const Point &point1, const Point &point2, float tm)
      : Quadric(o2w, ro) {
    p1 = point1;
    p2 = point2;
    phiMax = Radians(Clamp(tm, 0.0f, 360.0f));
    float rad1 = sqrtf(p1.x * p1.x + p1.y * p1.y);
    float rad2 = sqrtf(p2.x * p2.x + p2.y * p2.y);
    rmax = max(rad1, rad2);
    zmin = min(p1.z, p2.z);
    zmax = max(p1.z, p2.z);
    // implicit-form coefficients: walk along the generating line until they are finite
    if (p2.z == 0.) swap(p1, p2);
    Point pp = p1;
    float xy1, xy2;
    do {
      pp += 2.f * (p2 - p1);
      xy1 = pp.x * pp.x + pp.y * pp.y;
      xy2 = p2.x * p2.x + p2.y * p2.y;
      a = (1.f / xy1 - (pp.z * pp.z) / (xy1 * p2.z * p2.z)) / (1 - (xy2 * pp.z * pp.z) / (xy1 * p2.z * p2.z));
      c = (a * xy2 - 1) / (p2.z * p2.z);
    } while (isinf(a) || isnan(a));
  }
  BBox ObjectBound() const { return BBox(Point(-rmax, -rmax, zmin), Point(rmax, rmax, zmax)); }
#define PB_SQR(x) ((x) * (x))
#define PB_QUAD(x) ((PB_SQR(x)) * (PB_SQR(x)))
  float Area() const {
    return phiMax / 6.f *
           (2.f * PB_QUAD(p1.x) - 2.f * p1.x * p1.x * p1.x * p2.x + 2.f * PB_QUAD(p2.x) +
            2.f * (p1.y * p1.y + p1.y * p2.y + p2.y * p2.y) * (PB_SQR(p1.y - p2.y) + PB_SQR(p1.z - p2.z)) +
            p2.x * p2.x * (5.f * p1.y * p1.y + 2.f * p1.y * p2.y - 4.f * p2.y * p2.y + 2.f * PB_SQR(p1.z - p2.z)) +
            p1.x * p1.x * (-4.f * p1.y * p1.y + 2.f * p1.y * p2.y + 5.f * p2.y * p2.y + 2.f * PB_SQR(p1.z - p2.z)) -
            2.f * p1.x * p2.x * (p2.x * p2.x - p1.y * p1.y + 5.f * p1.y * p2.y - p2.y * p2.y - p1.z * p1.z +
                                 2.f * p1.z * p2.z - p2.z * p2.z));
  }
#undef PB_SQR
#undef PB_QUAD
  void Fill(PbrtB200Quadric *q) const {
    q->kind = PBRT_B200_QUADRIC_HYPERBOLOID;
    q->zmin = zmin; q->zmax = zmax; q->phi_max = phiMax; q->radius = rmax;
    q->p1[0] = p1.x; q->p1[1] = p1.y; q->p1[2] = p1.z;
    q->p2[0] = p2.x; q->p2[1] = p2.y; q->p2[2] = p2.z;
    q->a = a; q->c = c;
  }
 private:
  Point p1, p2;
  float zmin, zmax, phiMax, rmax, a, c;
};

// ------------------------------------------------------------ trianglemesh
class TriangleMesh : public Shape {
 public:
  // shapes/trianglemesh.cpp:141-169: vertices are moved to world space once.
  TriangleMesh(const Transform &o2w, bool ro, int nt, int nv, const int *vi, const Point *P,
               const Normal *N, const Vector *S, const float *uv)
      : Shape(o2w, ro), ntris(nt), nverts(nv) {
    vertexIndex.assign(vi, vi + 3 * ntris);
    if (uv) uvs.assign(uv, uv + 2 * nverts);
    if (N) n.assign(N, N + nverts);
    if (S) s.assign(S, S + nverts);
    p.resize(nverts);
    for (int i = 0; i < nverts; ++i) p[i] = ObjectToWorld(P[i]);
  }
  BBox ObjectBound() const {
    BBox bobj;
    for (int i = 0; i < nverts; i++) bobj = Union(bobj, WorldToObject(p[i]));
    return bobj;
  }
  BBox WorldBound() const {
    BBox worldBounds;
    for (int i = 0; i < nverts; i++) worldBounds = Union(worldBounds, p[i]);
    return worldBounds;
  }
  bool CanIntersect() const { return false; }
  void Refine(vector<Reference<Shape> > &refined) const;
  float Area() const { Severe("Unimplemented Shape::Area() method called"); return 0.f; }
  void Flatten(FlatBuilder &, PbrtB200Prim *, PbrtB200PrimAux *) const {
    Severe("TriangleMesh must be refined before export");
  }
  int ntris, nverts;
  vector<int> vertexIndex;
  vector<Point> p;
  vector<Normal> n;
  vector<Vector> s;
  vector<float> uvs;
};

class Triangle : public Shape {
 public:
  Triangle(const Transform &o2w, bool ro, const TriangleMesh *m, int num)
      : Shape(o2w, ro), mesh(const_cast<TriangleMesh *>(m)), v(&m->vertexIndex[3 * num]) {}
  BBox ObjectBound() const {
    const Point &p1 = mesh->p[v[0]], &p2 = mesh->p[v[1]], &p3 = mesh->p[v[2]];
    return Union(BBox(WorldToObject(p1), WorldToObject(p2)), WorldToObject(p3));
  }
  // shapes/trianglemesh.cpp:205-212
  BBox WorldBound() const {
    const Point &p1 = mesh->p[v[0]], &p2 = mesh->p[v[1]], &p3 = mesh->p[v[2]];
    return Union(BBox(p1, p2), p3);
  }
  // shapes/trianglemesh.cpp:329-335
  float Area() const {
    const Point &p1 = mesh->p[v[0]], &p2 = mesh->p[v[1]], &p3 = mesh->p[v[2]];
    return 0.5f * Cross(p2 - p1, p3 - p1).Length();
  }
  void Identity(const void **owner, int *index) const {
    *owner = mesh.get();
    *index = (int)((v - &mesh->vertexIndex[0]) / 3);
  }
  void Flatten(FlatBuilder &fb, PbrtB200Prim *prim, PbrtB200PrimAux *aux) const {
    const TriangleMesh *m = mesh.get();
    uint32_t meshId, vbase;
    map<const void *, uint32_t>::iterator it = fb.meshIds.find(m);
    if (it == fb.meshIds.end()) {
      // first triangle of this mesh: register the mesh and its per-vertex data
      meshId = (uint32_t)fb.meshes.size();
      fb.meshIds[m] = meshId;
      PbrtB200Mesh pm;
      CopyMatrix(m->ObjectToWorld.GetMatrix(), pm.o2w);
      CopyMatrix(m->WorldToObject.GetMatrix(), pm.w2o);
      fb.meshes.push_back(pm);
      vbase = fb.nVertices;
      fb.meshVertexBase[m] = vbase;
      for (int i = 0; i < m->nverts; ++i) {
        for (int k = 0; k < 3; ++k) fb.vertexN.push_back(m->n.empty() ? 0.f : m->n[i][k]);
        for (int k = 0; k < 3; ++k) fb.vertexS.push_back(m->s.empty() ? 0.f : m->s[i][k]);
        for (int k = 0; k < 2; ++k) fb.vertexUV.push_back(m->uvs.empty() ? 0.f : m->uvs[2 * i + k]);
      }
      fb.nVertices += (uint32_t)m->nverts;
      fb.anyN |= !m->n.empty();
      fb.anyS |= !m->s.empty();
      fb.anyUV |= !m->uvs.empty();
    } else {
      meshId = it->second;
      vbase = fb.meshVertexBase[m];
    }
    for (int i = 0; i < 3; ++i) {
      const Point &q = m->p[v[i]];
      prim->p[3 * i] = q.x; prim->p[3 * i + 1] = q.y; prim->p[3 * i + 2] = q.z;
      aux->v[i] = vbase + (uint32_t)v[i];
    }
    aux->owner = meshId;
    prim->flags = PBRT_B200_PRIM_TRIANGLE | OrientationFlags() |
                  (m->uvs.empty() ? 0u : PBRT_B200_PF_HAS_UV) |
                  (m->n.empty() ? 0u : PBRT_B200_PF_HAS_N) |
                  (m->s.empty() ? 0u : PBRT_B200_PF_HAS_S);
  }
 private:
  Reference<TriangleMesh> mesh;
  const int *v;
};

// -------------------------------------------------------------- heightfield
// shapes/heightfield.cpp:28-109: a shape that only exists to be refined into one triangle mesh over
// the unit square (two triangles per cell, uv = xy); everything downstream sees triangles.
class Heightfield : public Shape {
 public:
  Heightfield(const Transform &o2w, bool ro, int x, int y, const float *zs) : Shape(o2w, ro), nx(x), ny(y) {
    z.assign(zs, zs + (size_t)nx * ny);
  }
  BBox ObjectBound() const {
    float minz = z[0], maxz = z[0];
    for (int i = 1; i < nx * ny; ++i) {
      if (z[i] < minz) minz = z[i];
      if (z[i] > maxz) maxz = z[i];
    }
    return BBox(Point(0, 0, minz), Point(1, 1, maxz));
  }
  bool CanIntersect() const { return false; }
  float Area() const { Severe("Unimplemented Shape::Area() method called"); return 0.f; }
  void Flatten(FlatBuilder &, PbrtB200Prim *, PbrtB200PrimAux *) const {
    Severe("Heightfield must be refined before export");
  }
  void Refine(vector<Reference<Shape> > &refined) const {
    int ntris = 2 * (nx - 1) * (ny - 1);
    vector<int> verts(3 * (size_t)ntris);
    vector<Point> P((size_t)nx * ny);
    vector<float> uvs(2 * (size_t)nx * ny);
    int pos = 0;
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x) {
        P[pos].x = uvs[2 * pos] = (float)x / (float)(nx - 1);
        P[pos].y = uvs[2 * pos + 1] = (float)y / (float)(ny - 1);
        P[pos].z = z[pos];
        ++pos;
      }
    int *vp = &verts[0];
    for (int y = 0; y < ny - 1; ++y)
      for (int x = 0; x < nx - 1; ++x) {
        int v00 = x + y * nx, v10 = x + 1 + y * nx, v11 = x + 1 + (y + 1) * nx, v01 = x + (y + 1) * nx;
        *vp++ = v00; *vp++ = v10; *vp++ = v11;
        *vp++ = v00; *vp++ = v11; *vp++ = v01;
      }
    ParamSet paramSet;
    paramSet.AddInt("indices", &verts[0], 3 * ntris);
    paramSet.AddFloat("uv", &uvs[0], 2 * nx * ny);
    paramSet.AddPoint("P", &P[0], nx * ny);
    refined.push_back(CreateTriangleMeshShape(ObjectToWorld, reverseOrientation, paramSet));
  }
 private:
  int nx, ny;
  vector<float> z;
};

void TriangleMesh::Refine(vector<Reference<Shape> > &refined) const {
  for (int i = 0; i < ntris; ++i)
    refined.push_back(new Triangle(ObjectToWorld, reverseOrientation, this, i));
}

}  // namespace

// shapes/sphere.cpp:252-263
Shape *CreateSphereShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  float radius = params.FindOneFloat("radius", 1.f);
  float zmin = params.FindOneFloat("zmin", -radius);
  float zmax = params.FindOneFloat("zmax", radius);
  float phimax = params.FindOneFloat("phimax", 360.f);
  return new Sphere(o2w, reverseOrientation, radius, zmin, zmax, phimax);
}

// shapes/heightfield.cpp:110-119
Shape *CreateHeightfieldShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  int nu = params.FindOneInt("nu", -1);
  int nv = params.FindOneInt("nv", -1);
  int nitems = 0;
  const float *Pz = params.FindFloat("Pz", &nitems);
  if (nu < 2 || nv < 2 || Pz == NULL || nitems != nu * nv) {
    Error("heightfield: \"nu\", \"nv\" (>= 2) and nu*nv \"Pz\" values are required; shape ignored.");
    return NULL;
  }
  return new Heightfield(o2w, reverseOrientation, nu, nv, Pz);
}
// shapes/cylinder.cpp:221-228
Shape *CreateCylinderShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  float radius = params.FindOneFloat("radius", 1);
  float zmin = params.FindOneFloat("zmin", -1);
  float zmax = params.FindOneFloat("zmax", 1);
  float phimax = params.FindOneFloat("phimax", 360);
  return new Cylinder(o2w, reverseOrientation, radius, zmin, zmax, phimax);
}
// shapes/disk.cpp:151-158
Shape *CreateDiskShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  float height = params.FindOneFloat("height", 0.);
  float radius = params.FindOneFloat("radius", 1);
  float inner_radius = params.FindOneFloat("innerradius", 0);
  float phimax = params.FindOneFloat("phimax", 360);
  return new Disk(o2w, reverseOrientation, height, radius, inner_radius, phimax);
}
// shapes/cone.cpp:187-193
Shape *CreateConeShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  float radius = params.FindOneFloat("radius", 1);
  float height = params.FindOneFloat("height", 1);
  float phimax = params.FindOneFloat("phimax", 360);
  return new Cone(o2w, reverseOrientation, height, radius, phimax);
}
// shapes/paraboloid.cpp:191-198
Shape *CreateParaboloidShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  float radius = params.FindOneFloat("radius", 1);
  float zmin = params.FindOneFloat("zmin", 0);
  float zmax = params.FindOneFloat("zmax", 1);
  float phimax = params.FindOneFloat("phimax", 360);
  return new Paraboloid(o2w, reverseOrientation, radius, zmin, zmax, phimax);
}
// shapes/hyperboloid.cpp:238-244
Shape *CreateHyperboloidShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  Point p1 = params.FindOnePoint("p1", Point(0, 0, 0));
  Point p2 = params.FindOnePoint("p2", Point(1, 1, 1));
  float phimax = params.FindOneFloat("phimax", 360);
  return new Hyperboloid(o2w, reverseOrientation, p1, p2, phimax);
}
// Shapes whose Shape::Sample the reference implements (an area light needs it, core/shape.h:85-88)
bool ShapeCanSample(const Shape *s) {
  const Quadric *q = dynamic_cast<const Quadric *>(s);
  return q == NULL || q->CanSample();
}

// shapes/trianglemesh.cpp:350-405
Shape *CreateTriangleMeshShape(const Transform &o2w, bool reverseOrientation,
                               const ParamSet &params) {
  int nvi = 0, npi = 0, nuvi = 0, nsi = 0, nni = 0;
  const int *vi = params.FindInt("indices", &nvi);
  const Point *P = params.FindPoint("P", &npi);
  const float *uvs = params.FindFloat("uv", &nuvi);
  if (!uvs) uvs = params.FindFloat("st", &nuvi);
  if (uvs) {
    if (nuvi < 2 * npi) {
      Error("Not enough of \"uv\"s for triangle mesh.  Expencted %d, found %d.  Discarding.\n",
            2 * npi, nuvi);
      uvs = NULL;
    } else if (nuvi > 2 * npi)
      Warning("More \"uv\"s provided than will be used for triangle mesh.  (%d expcted, %d found)\n",
              2 * npi, nuvi);
  }
  if (!vi || !P) return NULL;
  const Vector *S = params.FindVector("S", &nsi);
  if (S && nsi != npi) {
    Error("Number of \"S\"s for triangle mesh must match \"P\"s");
    S = NULL;
  }
  const Normal *N = params.FindNormal("N", &nni);
  if (N && nni != npi) {
    Error("Number of \"N\"s for triangle mesh must match \"P\"s");
    N = NULL;
  }
  if (uvs && N) {
    // degenerate uv mappings are discarded when normals are present (:378-396)
    const int *vp = vi;
    for (int i = 0; i < nvi; i += 3, vp += 3) {
      float area = .5f * Cross(P[vp[0]] - P[vp[1]], P[vp[2]] - P[vp[1]]).Length();
      if (area < 1e-7) continue;
      if ((uvs[2 * vp[0]] == uvs[2 * vp[1]] && uvs[2 * vp[0] + 1] == uvs[2 * vp[1] + 1]) ||
          (uvs[2 * vp[1]] == uvs[2 * vp[2]] && uvs[2 * vp[1] + 1] == uvs[2 * vp[2] + 1]) ||
          (uvs[2 * vp[2]] == uvs[2 * vp[0]] && uvs[2 * vp[2] + 1] == uvs[2 * vp[0] + 1])) {
        Warning("Degenerate uv coordinates in triangle mesh.  Discarding all uvs.");
        uvs = NULL;
        break;
      }
    }
  }
  for (int i = 0; i < nvi; ++i)
    if (vi[i] >= npi || vi[i] < 0) {
      Error("trianglemesh has out of-bounds vertex index %d (%d \"P\" values were given",
            vi[i], npi);
      return NULL;
    }
  return new TriangleMesh(o2w, reverseOrientation, nvi / 3, npi, vi, P, N, S, uvs);
}
