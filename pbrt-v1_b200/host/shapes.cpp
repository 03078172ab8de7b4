// shapes.cpp -- "sphere" and "trianglemesh" shape plugins, host half.
// Parameter handling follows shapes/sphere.cpp:252-263 and
// shapes/trianglemesh.cpp:350-405; the intersection code of those plugins is
// what the CUDA kernels implement (pbrt-v1_b200/csrc).
#include "core.h"

namespace {

void CopyMatrix(const Matrix4x4 &m, float out[16]) { memcpy(out, m.m, 16 * sizeof(float)); }

// ------------------------------------------------------------------ sphere
class Sphere : public Shape {
 public:
  // shapes/sphere.cpp:80-92
  Sphere(const Transform &o2w, bool ro, float rad, float z0, float z1, float pm) : Shape(o2w, ro) {
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
  void Flatten(FlatBuilder &fb, PbrtB200Prim *prim, PbrtB200PrimAux *aux) const {
    PbrtB200Sphere s;
    memset(&s, 0, sizeof(s));
    CopyMatrix(ObjectToWorld.GetMatrix(), s.o2w);
    CopyMatrix(WorldToObject.GetMatrix(), s.w2o);
    s.radius = radius; s.zmin = zmin; s.zmax = zmax;
    s.theta_min = thetaMin; s.theta_max = thetaMax; s.phi_max = phiMax;
    s.flags = OrientationFlags();
    memset(prim->p, 0, sizeof(prim->p));
    prim->flags = PBRT_B200_PRIM_SPHERE | OrientationFlags();
    aux->v[0] = aux->v[1] = aux->v[2] = 0;
    aux->owner = (uint32_t)fb.spheres.size();
    fb.spheres.push_back(s);
  }
 private:
  float radius, phiMax, zmin, zmax, thetaMin, thetaMax;
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
