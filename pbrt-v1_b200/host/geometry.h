// geometry.h -- Vector / Point / Normal / BBox for the host front end.
// Same operator semantics as the reference (core/geometry.h:30-334); in
// particular division multiplies by a rounded reciprocal (geometry.h:66-70).
#ifndef PBRT_B200_HOST_GEOMETRY_H
#define PBRT_B200_HOST_GEOMETRY_H
#include "pbrt.h"

class Point;
class Normal;
class Vector {
 public:
  Vector(float _x = 0, float _y = 0, float _z = 0) : x(_x), y(_y), z(_z) {}
  explicit Vector(const Point &p);
  explicit Vector(const Normal &n);
  Vector operator+(const Vector &v) const { return Vector(x + v.x, y + v.y, z + v.z); }
  Vector operator-(const Vector &v) const { return Vector(x - v.x, y - v.y, z - v.z); }
  Vector operator*(float f) const { return Vector(f * x, f * y, f * z); }
  Vector operator/(float f) const {
    float inv = 1.f / f;
    return Vector(x * inv, y * inv, z * inv);
  }
  Vector operator-() const { return Vector(-x, -y, -z); }
  float operator[](int i) const { return (&x)[i]; }
  float &operator[](int i) { return (&x)[i]; }
  float LengthSquared() const { return x * x + y * y + z * z; }
  float Length() const { return sqrtf(LengthSquared()); }
  float x, y, z;
};
class Point {
 public:
  Point(float _x = 0, float _y = 0, float _z = 0) : x(_x), y(_y), z(_z) {}
  Point operator+(const Vector &v) const { return Point(x + v.x, y + v.y, z + v.z); }
  Point &operator+=(const Vector &v) { x += v.x; y += v.y; z += v.z; return *this; }   // geometry.h:110-113
  Vector operator-(const Point &p) const { return Vector(x - p.x, y - p.y, z - p.z); }
  Point operator-(const Vector &v) const { return Point(x - v.x, y - v.y, z - v.z); }
  Point operator*(float f) const { return Point(f * x, f * y, f * z); }
  Point operator/(float f) const {
    float inv = 1.f / f;
    return Point(inv * x, inv * y, inv * z);
  }
  Point &operator/=(float f) {
    float inv = 1.f / f;
    x *= inv; y *= inv; z *= inv;
    return *this;
  }
  float operator[](int i) const { return (&x)[i]; }
  float &operator[](int i) { return (&x)[i]; }
  float x, y, z;
};
class Normal {
 public:
  Normal(float _x = 0, float _y = 0, float _z = 0) : x(_x), y(_y), z(_z) {}
  explicit Normal(const Vector &v) : x(v.x), y(v.y), z(v.z) {}
  Normal operator-() const { return Normal(-x, -y, -z); }
  Normal operator*(float f) const { return Normal(f * x, f * y, f * z); }
  float operator[](int i) const { return (&x)[i]; }
  float &operator[](int i) { return (&x)[i]; }
  float x, y, z;
};
inline Vector::Vector(const Point &p) : x(p.x), y(p.y), z(p.z) {}
inline Vector::Vector(const Normal &n) : x(n.x), y(n.y), z(n.z) {}
inline Vector operator*(float f, const Vector &v) { return v * f; }
inline float Dot(const Vector &v1, const Vector &v2) {
  return v1.x * v2.x + v1.y * v2.y + v1.z * v2.z;
}
inline Vector Cross(const Vector &v1, const Vector &v2) {
  return Vector((v1.y * v2.z) - (v1.z * v2.y), (v1.z * v2.x) - (v1.x * v2.z),
                (v1.x * v2.y) - (v1.y * v2.x));
}
inline Vector Normalize(const Vector &v) { return v / v.Length(); }
// core/geometry.h:320-331 (Cross is spelled out so this does not depend on declaration order)
inline void CoordinateSystem(const Vector &v1, Vector *v2, Vector *v3) {
  if (fabsf(v1.x) > fabsf(v1.y)) {
    float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
    *v2 = Vector(-v1.z * invLen, 0.f, v1.x * invLen);
  } else {
    float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
    *v2 = Vector(0.f, v1.z * invLen, -v1.y * invLen);
  }
  *v3 = Vector((v1.y * v2->z) - (v1.z * v2->y), (v1.z * v2->x) - (v1.x * v2->z), (v1.x * v2->y) - (v1.y * v2->x));
}

// core/geometry.h:235-287, core/geometry.cpp:27-48
class BBox {
 public:
  BBox() {
    pMin = Point(INFINITY, INFINITY, INFINITY);
    pMax = Point(-INFINITY, -INFINITY, -INFINITY);
  }
  BBox(const Point &p) : pMin(p), pMax(p) {}
  BBox(const Point &p1, const Point &p2) {
    pMin = Point(min(p1.x, p2.x), min(p1.y, p2.y), min(p1.z, p2.z));
    pMax = Point(max(p1.x, p2.x), max(p1.y, p2.y), max(p1.z, p2.z));
  }
  Point pMin, pMax;
};
inline BBox Union(const BBox &b, const Point &p) {
  BBox ret = b;
  ret.pMin.x = min(b.pMin.x, p.x); ret.pMin.y = min(b.pMin.y, p.y);
  ret.pMin.z = min(b.pMin.z, p.z);
  ret.pMax.x = max(b.pMax.x, p.x); ret.pMax.y = max(b.pMax.y, p.y);
  ret.pMax.z = max(b.pMax.z, p.z);
  return ret;
}
inline BBox Union(const BBox &b, const BBox &b2) {
  BBox ret;
  ret.pMin.x = min(b.pMin.x, b2.pMin.x); ret.pMin.y = min(b.pMin.y, b2.pMin.y);
  ret.pMin.z = min(b.pMin.z, b2.pMin.z);
  ret.pMax.x = max(b.pMax.x, b2.pMax.x); ret.pMax.y = max(b.pMax.y, b2.pMax.y);
  ret.pMax.z = max(b.pMax.z, b2.pMax.z);
  return ret;
}

// core/color.h:29-210 -- RGB "spectrum"; the host only stores and clamps it.
class Spectrum {
 public:
  Spectrum(float v = 0.f) { c[0] = c[1] = c[2] = v; }
  Spectrum(float cs[COLOR_SAMPLES]) { c[0] = cs[0]; c[1] = cs[1]; c[2] = cs[2]; }
  Spectrum Clamp(float low = 0.f, float high = INFINITY) const {
    Spectrum ret;
    for (int i = 0; i < COLOR_SAMPLES; ++i) ret.c[i] = ::Clamp(c[i], low, high);
    return ret;
  }
  bool Black() const {
    for (int i = 0; i < COLOR_SAMPLES; ++i)
      if (c[i] != 0.) return false;
    return true;
  }
  float c[COLOR_SAMPLES];
};
#endif
