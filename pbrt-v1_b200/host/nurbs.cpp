// nurbs.cpp -- "nurbs" shape plugin, host half (shapes/nurbs.cpp:28-332): a NURBS patch diced into a
// 30 x 30 grid by de Boor evaluation in homogeneous coordinates, handed on as ONE triangle mesh with
// per-vertex normals and uv.  Only exists to be refined; the device never sees it.
#include "core.h"

Shape *CreateTriangleMeshShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params);

namespace {

struct Hom { float x, y, z, w; };

// nurbs.cpp:51-60: the knot span that holds t
int KnotSpan(const vector<float> &knot, int order, float t) {
  int span = order - 1;
  while (t > knot[span + 1]) ++span;
  return span;
}
// nurbs.cpp:70-112: one curve point (and optionally its derivative) by repeated knot-interval
// interpolation of the `order` control points of the span; control point i is cp[base + i * stride]
Hom EvalCurve(int order, const vector<float> &knots, const Hom *cp, int stride, float t, Vector *deriv) {
  int span = KnotSpan(knots, order, t);
  const float *knot = &knots[span];
  int first = span - order + 1;
  vector<Hom> w(order);
  for (int i = 0; i < order; ++i) w[i] = cp[(first + i) * stride];
  float alpha;
  for (int i = 0; i < order - 2; ++i)
    for (int j = 0; j < order - 1 - i; ++j) {
      alpha = (knot[1 + j] - t) / (knot[1 + j] - knot[j + 2 - order + i]);
      w[j].x = w[j].x * alpha + w[j + 1].x * (1 - alpha);
      w[j].y = w[j].y * alpha + w[j + 1].y * (1 - alpha);
      w[j].z = w[j].z * alpha + w[j + 1].z * (1 - alpha);
      w[j].w = w[j].w * alpha + w[j + 1].w * (1 - alpha);
    }
  alpha = (knot[1] - t) / (knot[1] - knot[0]);
  Hom val = {w[0].x * alpha + w[1].x * (1 - alpha), w[0].y * alpha + w[1].y * (1 - alpha),
             w[0].z * alpha + w[1].z * (1 - alpha), w[0].w * alpha + w[1].w * (1 - alpha)};
  if (deriv) {
    float factor = (order - 1) / (knot[1] - knot[0]);
    Hom delta = {(w[1].x - w[0].x) * factor, (w[1].y - w[0].y) * factor, (w[1].z - w[0].z) * factor,
                 (w[1].w - w[0].w) * factor};
    deriv->x = delta.x / val.w - (val.x * delta.w / (val.w * val.w));
    deriv->y = delta.y / val.w - (val.y * delta.w / (val.w * val.w));
    deriv->z = delta.z / val.w - (val.z * delta.w / (val.w * val.w));
  }
  return val;
}

class NURBS : public Shape {
 public:
  NURBS(const Transform &o2w, bool ro, int numu, int uo, const float *uk, float u0, float u1, int numv, int vo,
        const float *vk, float v0, float v1, const float *p, bool homogeneous)
      : Shape(o2w, ro), nu(numu), uorder(uo), nv(numv), vorder(vo), umin(u0), umax(u1), vmin(v0), vmax(v1),
        isHomogeneous(homogeneous) {
    P.assign(p, p + (size_t)(homogeneous ? 4 : 3) * nu * nv);
    uknot.assign(uk, uk + nu + uorder);
    vknot.assign(vk, vk + nv + vorder);
  }
  Point ControlPoint(int i) const {
    if (!isHomogeneous) return Point(P[3 * i], P[3 * i + 1], P[3 * i + 2]);
    return Point(P[4 * i] / P[4 * i + 3], P[4 * i + 1] / P[4 * i + 3], P[4 * i + 2] / P[4 * i + 3]);
  }
  BBox ObjectBound() const {   // nurbs.cpp:160-177
    BBox bound = ControlPoint(0);
    for (int i = 0; i < nu * nv; ++i) bound = Union(bound, ControlPoint(i));
    return bound;
  }
  BBox WorldBound() const {    // nurbs.cpp:178-203
    BBox bound = ObjectToWorld(ControlPoint(0));
    for (int i = 0; i < nu * nv; ++i) bound = Union(bound, ObjectToWorld(ControlPoint(i)));
    return bound;
  }
  bool CanIntersect() const { return false; }
  float Area() const { Severe("Unimplemented Shape::Area() method called"); return 0.f; }
  void Flatten(FlatBuilder &, PbrtB200Prim *, PbrtB200PrimAux *) const { Severe("NURBS must be refined before export"); }
  void Refine(vector<Reference<Shape> > &refined) const;

 private:
  // nurbs.cpp:113-137: surface point with both partial derivatives: an isoparametric curve of
  // `uorder` points (each from a curve in v), evaluated in u -- and the other way round for dP/dv
  Point EvalSurface(const vector<Hom> &cp, float u, float v, Vector *dPdu, Vector *dPdv) const {
    int uFirst = KnotSpan(uknot, uorder, u) - uorder + 1;
    int vFirst = KnotSpan(vknot, vorder, v) - vorder + 1;
    // the curve evaluator addresses control point (first + i) of whatever it is given: pad the front
    vector<Hom> iso((size_t)std::max(uorder, vorder) + std::max(uFirst, vFirst));
    for (int i = 0; i < uorder; ++i) iso[uFirst + i] = EvalCurve(vorder, vknot, &cp[uFirst + i], nu, v, NULL);
    Hom Pt = EvalCurve(uorder, uknot, &iso[0], 1, u, dPdu);
    for (int i = 0; i < vorder; ++i) iso[vFirst + i] = EvalCurve(uorder, uknot, &cp[(size_t)(vFirst + i) * nu], 1, u, NULL);
    EvalCurve(vorder, vknot, &iso[0], 1, v, dPdv);
    return Point(Pt.x / Pt.w, Pt.y / Pt.w, Pt.z / Pt.w);
  }
  int nu, uorder, nv, vorder;
  float umin, umax, vmin, vmax;
  bool isHomogeneous;
  vector<float> P, uknot, vknot;
};

// The patch sampled on a regular parameter grid (what the reference hands to its triangle mesh,
// shapes/nurbs.cpp:204-281): GRID x GRID surface points with normals and (u, v), row = one v value.
struct SampledPatch {
  static const int GRID = 30;
  vector<Point> position;
  vector<Normal> normal;
  vector<float> uv;
  SampledPatch() : position((size_t)GRID * GRID), normal((size_t)GRID * GRID), uv(2 * (size_t)GRID * GRID) {}
  static int vertex(int iu, int iv) { return iv * GRID + iu; }
  // two triangles per grid cell, both fanned from the cell's (u, v) corner
  vector<int> triangulate() const {
    vector<int> tri;
    tri.reserve(6 * (size_t)(GRID - 1) * (GRID - 1));
    for (int iv = 0; iv + 1 < GRID; ++iv)
      for (int iu = 0; iu + 1 < GRID; ++iu) {
        const int corner = vertex(iu, iv), opposite = vertex(iu + 1, iv + 1);
        const int fan[2][3] = {{corner, vertex(iu + 1, iv), opposite}, {corner, opposite, vertex(iu, iv + 1)}};
        for (int t = 0; t < 2; ++t) tri.insert(tri.end(), fan[t], fan[t] + 3);
      }
    return tri;
  }
};

void NURBS::Refine(vector<Reference<Shape> > &refined) const {
  const int G = SampledPatch::GRID;
  // every control point in homogeneous form (weight 1 for a polynomial patch)
  const int stride = isHomogeneous ? 4 : 3;
  vector<Hom> control((size_t)nu * nv);
  for (size_t i = 0; i < control.size(); ++i) {
    const float *c = &P[stride * i];
    Hom h = {c[0], c[1], c[2], isHomogeneous ? c[3] : 1.f};
    control[i] = h;
  }
  // the grid's parameter values: G evenly spaced steps across [umin, umax] x [vmin, vmax]
  float uAt[SampledPatch::GRID], vAt[SampledPatch::GRID];
  for (int k = 0; k < G; ++k) {
    const float frac = (float)k / (float)(G - 1);
    uAt[k] = Lerp(frac, umin, umax);
    vAt[k] = Lerp(frac, vmin, vmax);
  }
  SampledPatch grid;
  for (int iv = 0; iv < G; ++iv)
    for (int iu = 0; iu < G; ++iu) {
      const int at = SampledPatch::vertex(iu, iv);
      Vector du, dv;
      grid.position[at] = EvalSurface(control, uAt[iu], vAt[iv], &du, &dv);
      grid.normal[at] = Normal(Normalize(Cross(du, dv)));
      grid.uv[2 * at] = uAt[iu];
      grid.uv[2 * at + 1] = vAt[iv];
    }
  const vector<int> indices = grid.triangulate();
  ParamSet mesh;
  mesh.AddInt("indices", &indices[0], (int)indices.size());
  mesh.AddPoint("P", &grid.position[0], G * G);
  mesh.AddFloat("uv", &grid.uv[0], 2 * G * G);
  mesh.AddNormal("N", &grid.normal[0], G * G);
  refined.push_back(CreateTriangleMeshShape(ObjectToWorld, reverseOrientation, mesh));
}

}  // namespace

// nurbs.cpp:282-332
Shape *CreateNURBSShape(const Transform &o2w, bool reverseOrientation, const ParamSet &params) {
  int nu = params.FindOneInt("nu", -1);
  int uorder = params.FindOneInt("uorder", -1);
  int nuknots = 0, nvknots = 0;
  const float *uknots = params.FindFloat("uknots", &nuknots);
  int nv = params.FindOneInt("nv", -1);
  int vorder = params.FindOneInt("vorder", -1);
  const float *vknots = params.FindFloat("vknots", &nvknots);
  if (nu < 2 || uorder < 2 || !uknots || nuknots != nu + uorder || nv < 2 || vorder < 2 || !vknots ||
      nvknots != nv + vorder || uorder > nu || vorder > nv) {
    Error("nurbs: \"nu\", \"uorder\", nu+uorder \"uknots\" (and the same in v) are required; shape ignored.");
    return NULL;
  }
  float u0 = params.FindOneFloat("u0", uknots[uorder - 1]);
  float u1 = params.FindOneFloat("u1", uknots[nu]);
  float v0 = params.FindOneFloat("v0", vknots[vorder - 1]);
  float v1 = params.FindOneFloat("v1", vknots[nv]);
  bool isHomogeneous = false;
  int npts = 0;
  const float *P = (const float *)params.FindPoint("P", &npts);
  if (!P) {
    P = params.FindFloat("Pw", &npts);
    npts /= 4;
    if (!P) return NULL;
    isHomogeneous = true;
  }
  if (npts != nu * nv) {
    Error("nurbs: %d control points given, nu * nv = %d expected; shape ignored.", npts, nu * nv);
    return NULL;
  }
  return new NURBS(o2w, reverseOrientation, nu, uorder, uknots, u0, u1, nv, vorder, vknots, v0, v1, P, isHomogeneous);
}
