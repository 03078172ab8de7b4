// loopsubdiv.cpp -- "loopsubdiv" shape plugin, host half: Loop subdivision of a triangle control mesh,
// pushed to the limit surface, handed on as ONE triangle mesh with per-vertex normals
// (shapes/loopsubdiv.cpp:28-494).  Like the heightfield, the shape only exists to be refined; the
// device never sees it.
//
// The reference links vertices and faces with pointers; here a subdivision level is two arrays and
// every link is an index.  What has to be kept for the refined mesh to come out vertex for vertex
// and bit for bit: the order in which faces / edges are visited (it numbers the new odd vertices and
// picks each vertex's start face), the walk order around a vertex (one-ring order = summation order
// of the weights) and the association of every float expression.
#include "core.h"
#include <map>

Shape *CreateTriangleMeshShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params);

namespace {

inline int NextI(int i) { return (i + 1) % 3; }
inline int PrevI(int i) { return (i + 2) % 3; }

struct SubVertex {
  Point P;
  int startFace, child;
  bool regular, boundary;
  SubVertex() : startFace(-1), child(-1), regular(false), boundary(false) {}
};
struct SubFace {
  int v[3], f[3], children[4];
  SubFace() {
    for (int i = 0; i < 3; ++i) v[i] = f[i] = -1;
    for (int i = 0; i < 4; ++i) children[i] = -1;
  }
};
// One level of the mesh tree.
struct SubLevel {
  vector<SubVertex> verts;
  vector<SubFace> faces;
  int vnum(int face, int vert) const {
    for (int i = 0; i < 3; ++i)
      if (faces[face].v[i] == vert) return i;
    Severe("Basic logic error in SDFace::vnum()");
    return -1;
  }
  int nextFace(int face, int vert) const { return faces[face].f[vnum(face, vert)]; }
  int prevFace(int face, int vert) const { return faces[face].f[PrevI(vnum(face, vert))]; }
  int nextVert(int face, int vert) const { return faces[face].v[NextI(vnum(face, vert))]; }
  int prevVert(int face, int vert) const { return faces[face].v[PrevI(vnum(face, vert))]; }
  int otherVert(int face, int v0, int v1) const {
    for (int i = 0; i < 3; ++i)
      if (faces[face].v[i] != v0 && faces[face].v[i] != v1) return faces[face].v[i];
    Severe("Basic logic error in SDVertex::otherVert()");
    return -1;
  }
  // loopsubdiv.cpp:137-157
  int valence(int vert) const {
    int f = verts[vert].startFace;
    if (!verts[vert].boundary) {
      int nf = 1;
      while ((f = nextFace(f, vert)) != verts[vert].startFace) ++nf;
      return nf;
    }
    int nf = 1;
    while ((f = nextFace(f, vert)) != -1) ++nf;
    f = verts[vert].startFace;
    while ((f = prevFace(f, vert)) != -1) ++nf;
    return nf + 1;
  }
  // loopsubdiv.cpp:445-466: the ring in walking order
  void oneRing(int vert, vector<Point> *ring) const {
    ring->clear();
    if (!verts[vert].boundary) {
      int face = verts[vert].startFace;
      do {
        ring->push_back(verts[nextVert(face, vert)].P);
        face = nextFace(face, vert);
      } while (face != verts[vert].startFace);
    } else {
      int face = verts[vert].startFace, f2;
      while ((f2 = nextFace(face, vert)) != -1) face = f2;
      ring->push_back(verts[nextVert(face, vert)].P);
      do {
        ring->push_back(verts[prevVert(face, vert)].P);
        face = prevFace(face, vert);
      } while (face != -1);
    }
  }
};
inline Point Scaled(float f, const Point &p) { return Point(f * p.x, f * p.y, f * p.z); }   // float * Point
inline void Accumulate(Point *a, const Point &b) { a->x += b.x; a->y += b.y; a->z += b.z; }   // Point += Point
inline Point Sum(const Point &a, const Point &b) { return Point(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float Beta(int valence) {
  if (valence == 3) return 3.f / 16.f;
  else return 3.f / (8.f * valence);
}
inline float Gamma(int valence) { return 1.f / (valence + 3.f / (8.f * Beta(valence))); }
// loopsubdiv.cpp:434-444
Point WeightOneRing(const SubLevel &lv, int vert, float beta) {
  int valence = lv.valence(vert);
  vector<Point> ring;
  lv.oneRing(vert, &ring);
  Point P = Scaled(1 - valence * beta, lv.verts[vert].P);
  for (int i = 0; i < valence; ++i) Accumulate(&P, Scaled(beta, ring[i]));
  return P;
}
// loopsubdiv.cpp:467-477
Point WeightBoundary(const SubLevel &lv, int vert, float beta) {
  int valence = lv.valence(vert);
  vector<Point> ring;
  lv.oneRing(vert, &ring);
  Point P = Scaled(1 - 2 * beta, lv.verts[vert].P);
  Accumulate(&P, Scaled(beta, ring[0]));
  Accumulate(&P, Scaled(beta, ring[valence - 1]));
  return P;
}
typedef std::pair<int, int> EdgeKey;
inline EdgeKey MakeEdge(int a, int b) { return a < b ? EdgeKey(a, b) : EdgeKey(b, a); }

class LoopSubdiv : public Shape {
 public:
  // loopsubdiv.cpp:159-226: control mesh topology
  LoopSubdiv(const Transform &o2w, bool ro, int nfaces, int nvertices, const int *vertexIndices, const Point *P,
             int nl)
      : Shape(o2w, ro), nLevels(nl) {
    base.verts.resize(nvertices);
    base.faces.resize(nfaces);
    for (int i = 0; i < nvertices; ++i) base.verts[i].P = P[i];
    const int *vp = vertexIndices;
    for (int i = 0; i < nfaces; ++i) {
      for (int j = 0; j < 3; ++j) {
        base.faces[i].v[j] = vp[j];
        base.verts[vp[j]].startFace = i;
      }
      vp += 3;
    }
    // neighbour links: an edge seen for the second time joins its two faces
    std::map<EdgeKey, std::pair<int, int> > edges;   // edge -> (first face, its edge number)
    for (int i = 0; i < nfaces; ++i)
      for (int edgeNum = 0; edgeNum < 3; ++edgeNum) {
        EdgeKey e = MakeEdge(base.faces[i].v[edgeNum], base.faces[i].v[NextI(edgeNum)]);
        std::map<EdgeKey, std::pair<int, int> >::iterator it = edges.find(e);
        if (it == edges.end()) edges[e] = std::make_pair(i, edgeNum);
        else {
          base.faces[it->second.first].f[it->second.second] = i;
          base.faces[i].f[edgeNum] = it->second.first;
          edges.erase(it);
        }
      }
    for (int i = 0; i < nvertices; ++i) {
      SubVertex &v = base.verts[i];
      if (v.startFace < 0) Severe("loopsubdiv: vertex %d is not used by any face", i);
      int f = v.startFace;
      do {
        f = base.nextFace(f, i);
      } while (f != -1 && f != v.startFace);
      v.boundary = (f == -1);
      if (!v.boundary && base.valence(i) == 6) v.regular = true;
      else if (v.boundary && base.valence(i) == 4) v.regular = true;
      else v.regular = false;
    }
  }
  BBox ObjectBound() const {
    BBox b;
    for (size_t i = 0; i < base.verts.size(); i++) b = Union(b, base.verts[i].P);
    return b;
  }
  BBox WorldBound() const {
    BBox b;
    for (size_t i = 0; i < base.verts.size(); i++) b = Union(b, ObjectToWorld(base.verts[i].P));
    return b;
  }
  bool CanIntersect() const { return false; }
  float Area() const { Severe("Unimplemented Shape::Area() method called"); return 0.f; }
  void Flatten(FlatBuilder &, PbrtB200Prim *, PbrtB200PrimAux *) const {
    Severe("LoopSubdiv must be refined before export");
  }
  void Refine(vector<Reference<Shape> > &refined) const;

 private:
  int nLevels;
  SubLevel base;
};

// loopsubdiv.cpp:245-432
void LoopSubdiv::Refine(vector<Reference<Shape> > &refined) const {
  SubLevel cur = base;
  for (int level = 0; level < nLevels; ++level) {
    SubLevel next;
    const int nv = (int)cur.verts.size(), nf = (int)cur.faces.size();
    // children: one even vertex per vertex, four faces per face
    next.verts.resize(nv);
    for (int j = 0; j < nv; ++j) {
      cur.verts[j].child = j;
      next.verts[j].regular = cur.verts[j].regular;
      next.verts[j].boundary = cur.verts[j].boundary;
    }
    next.faces.resize(4 * (size_t)nf);
    for (int j = 0; j < nf; ++j)
      for (int k = 0; k < 4; ++k) cur.faces[j].children[k] = 4 * j + k;
    // even vertices
    for (int j = 0; j < nv; ++j) {
      if (!cur.verts[j].boundary) {
        if (cur.verts[j].regular) next.verts[j].P = WeightOneRing(cur, j, 1.f / 16.f);
        else next.verts[j].P = WeightOneRing(cur, j, Beta(cur.valence(j)));
      } else
        next.verts[j].P = WeightBoundary(cur, j, 1.f / 8.f);
    }
    // odd vertices, numbered in the order their edge is first met
    std::map<EdgeKey, int> edgeVerts;
    for (int j = 0; j < nf; ++j) {
      const SubFace &face = cur.faces[j];
      for (int k = 0; k < 3; ++k) {
        EdgeKey edge = MakeEdge(face.v[k], face.v[NextI(k)]);
        if (edgeVerts.find(edge) != edgeVerts.end()) continue;
        SubVertex vert;
        vert.regular = true;
        vert.boundary = (face.f[k] == -1);
        vert.startFace = face.children[3];
        const Point &P0 = cur.verts[edge.first].P, &P1 = cur.verts[edge.second].P;
        if (vert.boundary) {
          vert.P = Scaled(0.5f, P0);
          Accumulate(&vert.P, Scaled(0.5f, P1));
        } else {
          vert.P = Scaled(3.f / 8.f, P0);
          Accumulate(&vert.P, Scaled(3.f / 8.f, P1));
          Accumulate(&vert.P, Scaled(1.f / 8.f, cur.verts[cur.otherVert(j, edge.first, edge.second)].P));
          Accumulate(&vert.P, Scaled(1.f / 8.f, cur.verts[cur.otherVert(face.f[k], edge.first, edge.second)].P));
        }
        edgeVerts[edge] = (int)next.verts.size();
        next.verts.push_back(vert);
      }
    }
    // topology of the next level
    for (int j = 0; j < nv; ++j) {
      int sf = cur.verts[j].startFace;
      next.verts[j].startFace = cur.faces[sf].children[cur.vnum(sf, j)];
    }
    for (int j = 0; j < nf; ++j) {
      const SubFace &face = cur.faces[j];
      for (int k = 0; k < 3; ++k) {
        next.faces[face.children[3]].f[k] = face.children[NextI(k)];
        next.faces[face.children[k]].f[NextI(k)] = face.children[3];
        int f2 = face.f[k];
        next.faces[face.children[k]].f[k] = f2 != -1 ? cur.faces[f2].children[cur.vnum(f2, face.v[k])] : -1;
        f2 = face.f[PrevI(k)];
        next.faces[face.children[k]].f[PrevI(k)] = f2 != -1 ? cur.faces[f2].children[cur.vnum(f2, face.v[k])] : -1;
      }
    }
    for (int j = 0; j < nf; ++j) {
      const SubFace &face = cur.faces[j];
      for (int k = 0; k < 3; ++k) {
        next.faces[face.children[k]].v[k] = cur.verts[face.v[k]].child;
        int vert = edgeVerts[MakeEdge(face.v[k], face.v[NextI(k)])];
        next.faces[face.children[k]].v[NextI(k)] = vert;
        next.faces[face.children[NextI(k)]].v[k] = vert;
        next.faces[face.children[3]].v[k] = vert;
      }
    }
    cur.verts.swap(next.verts);
    cur.faces.swap(next.faces);
  }
  // limit surface
  const int nv = (int)cur.verts.size();
  vector<Point> Plimit(nv);
  for (int i = 0; i < nv; ++i) {
    if (cur.verts[i].boundary) Plimit[i] = WeightBoundary(cur, i, 1.f / 5.f);
    else Plimit[i] = WeightOneRing(cur, i, Gamma(cur.valence(i)));
  }
  for (int i = 0; i < nv; ++i) cur.verts[i].P = Plimit[i];
  // tangents -> normals
  vector<Normal> Ns;
  Ns.reserve(nv);
  vector<Point> Pring;
  for (int i = 0; i < nv; ++i) {
    Vector S(0, 0, 0), T(0, 0, 0);
    int valence = cur.valence(i);
    cur.oneRing(i, &Pring);
    const Point &vP = cur.verts[i].P;
    if (!cur.verts[i].boundary) {
      for (int k = 0; k < valence; ++k) {
        float c = cosf(2.f * M_PI * k / valence), s = sinf(2.f * M_PI * k / valence);
        S = S + Vector(c * Pring[k].x, c * Pring[k].y, c * Pring[k].z);
        T = T + Vector(s * Pring[k].x, s * Pring[k].y, s * Pring[k].z);
      }
    } else {
      S = Pring[valence - 1] - Pring[0];
      if (valence == 2) {
        Point a = Sum(Pring[0], Pring[1]);
        T = a - Scaled(2, vP);
      } else if (valence == 3)
        T = Pring[1] - vP;
      else if (valence == 4) {
        Point a = Sum(Sum(Sum(Sum(Scaled(-1, Pring[0]), Scaled(2, Pring[1])), Scaled(2, Pring[2])), Scaled(-1, Pring[3])),
                      Scaled(-2, vP));
        T = Vector(a);
      } else {
        float theta = M_PI / float(valence - 1);
        T = Vector(Scaled(sinf(theta), Sum(Pring[0], Pring[valence - 1])));
        for (int k = 1; k < valence - 1; ++k) {
          float wt = (2 * cosf(theta) - 2) * sinf((k) * theta);
          T = T + Vector(Scaled(wt, Pring[k]));
        }
        T = -T;
      }
    }
    Ns.push_back(Normal(Cross(S, T)));
  }
  const int ntris = (int)cur.faces.size();
  vector<int> verts(3 * (size_t)ntris);
  for (int i = 0; i < ntris; ++i)
    for (int j = 0; j < 3; ++j) verts[3 * i + j] = cur.faces[i].v[j];
  ParamSet paramSet;
  paramSet.AddInt("indices", &verts[0], 3 * ntris);
  paramSet.AddPoint("P", &Plimit[0], nv);
  paramSet.AddNormal("N", &Ns[0], (int)Ns.size());
  refined.push_back(CreateTriangleMeshShape(ObjectToWorld, reverseOrientation, paramSet));
}

}  // namespace

// loopsubdiv.cpp:478-494
Shape *CreateLoopSubdivShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  int nlevels = params.FindOneInt("nlevels", 3);
  int nps = 0, nIndices = 0;
  const int *vi = params.FindInt("indices", &nIndices);
  const Point *P = params.FindPoint("P", &nps);
  if (!vi || !P) return NULL;
  params.FindOneString("scheme", "loop");   // read (and ignored) by the reference as well
  for (int i = 0; i < nIndices; ++i)
    if (vi[i] < 0 || vi[i] >= nps) {
      Error("loopsubdiv has out of-bounds vertex index %d (%d \"P\" values were given)", vi[i], nps);
      return NULL;
    }
  return new LoopSubdiv(o2w, reverseOrientation, nIndices / 3, nps, vi, P, nlevels);
}
