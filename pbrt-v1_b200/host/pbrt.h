// pbrt.h -- common declarations of the B200 host front end.
//
// The host side keeps pbrt-v1's public surface (core/pbrt.h, core/api.h) so
// scene files and API callers work unmodified; everything below
// Scene::Render() is flattened to the POD arrays of include/pbrt_b200.h and
// executed on the GPU.  All setup arithmetic here is strict fp32 in the
// reference's operation order (compile with -ffp-contract=off) because the
// matrices/bounds it produces feed the device bit-for-bit.
#ifndef PBRT_B200_HOST_PBRT_H
#define PBRT_B200_HOST_PBRT_H

#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <iostream>
using std::string;
using std::vector;
using std::map;
using std::min;
using std::max;
using std::swap;
using std::ostream;

typedef unsigned int u_int;

#define PBRT_VERSION 1.05
// core/pbrt.h:201-212: M_PI is redefined as a *float* literal.
#ifdef M_PI
#undef M_PI
#endif
#define M_PI 3.14159265358979323846f
#define INV_PI 0.31830988618379067154f
#define INV_TWOPI 0.15915494309189533577f
#ifndef INFINITY
#define INFINITY FLT_MAX
#endif
#define RAY_EPSILON 1e-3f
#define COLOR_SAMPLES 3
#define COREDLL

class ParamSet;
class TextureParams;
class Transform;

// core/util.cpp:36-97 -- same prefixes, same "Line n, file f" suffix.
void Info(const char *, ...) __attribute__((__format__(__printf__, 1, 2)));
void Warning(const char *, ...) __attribute__((__format__(__printf__, 1, 2)));
void Error(const char *, ...) __attribute__((__format__(__printf__, 1, 2)));
void Severe(const char *, ...) __attribute__((__format__(__printf__, 1, 2)));

void pbrtInit();
void pbrtCleanup();
bool ParseFile(const char *filename);
std::string ResolveDataFile(const std::string &name);   // parser.cpp

#define Assert(expr) ((void)0)

// core/pbrt.h:549-613
inline float Lerp(float t, float v1, float v2) { return (1.f - t) * v1 + t * v2; }
inline float Clamp(float val, float low, float high) {
  if (val < low) return low;
  else if (val > high) return high;
  else return val;
}
inline int Clamp(int val, int low, int high) {
  if (val < low) return low;
  else if (val > high) return high;
  else return val;
}
inline float Radians(float deg) { return ((float)M_PI / 180.f) * deg; }
inline float Degrees(float rad) { return (180.f / (float)M_PI) * rad; }
inline float Log2(float x) {
  static float invLog2 = 1.f / logf(2.f);
  return logf(x) * invLog2;
}
#define PBRT_DOUBLEMAGICROUNDEPS (.5 - 1.4e-11)
inline int Log2Int(float v) { return int(Log2(v) + PBRT_DOUBLEMAGICROUNDEPS); }
inline int Round2Int(double val) { return int(val + PBRT_DOUBLEMAGICROUNDEPS); }
inline int Float2Int(double val) { return (int)val; }
inline int Floor2Int(double val) { return (int)floor(val); }
inline int Ceil2Int(double val) { return (int)ceil(val); }

// core/pbrt.h:322-365 -- intrusive reference counting.
class ReferenceCounted {
 public:
  ReferenceCounted() { nReferences = 0; }
  int nReferences;
 private:
  ReferenceCounted(const ReferenceCounted &);
  ReferenceCounted &operator=(const ReferenceCounted &);
};
template <class T> class Reference {
 public:
  Reference(T *p = NULL) { ptr = p; if (ptr) ++ptr->nReferences; }
  Reference(const Reference<T> &r) { ptr = r.ptr; if (ptr) ++ptr->nReferences; }
  Reference &operator=(const Reference<T> &r) {
    if (r.ptr) r.ptr->nReferences++;
    if (ptr && --ptr->nReferences == 0) delete ptr;
    ptr = r.ptr;
    return *this;
  }
  Reference &operator=(T *p) {
    if (p) p->nReferences++;
    if (ptr && --ptr->nReferences == 0) delete ptr;
    ptr = p;
    return *this;
  }
  ~Reference() { if (ptr && --ptr->nReferences == 0) delete ptr; }
  T *operator->() { return ptr; }
  const T *operator->() const { return ptr; }
  operator bool() const { return ptr != NULL; }
  bool operator<(const Reference<T> &t2) const { return ptr < t2.ptr; }
  T *get() const { return ptr; }
 private:
  T *ptr;
};

// core/pbrt.h:499-539, core/util.cpp:109-184
struct Matrix4x4 : public ReferenceCounted {
  Matrix4x4() {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.f : 0.f;
  }
  Matrix4x4(float mat[4][4]) { memcpy(m, mat, 16 * sizeof(float)); }
  Matrix4x4(float t00, float t01, float t02, float t03, float t10, float t11,
            float t12, float t13, float t20, float t21, float t22, float t23,
            float t30, float t31, float t32, float t33) {
    m[0][0] = t00; m[0][1] = t01; m[0][2] = t02; m[0][3] = t03;
    m[1][0] = t10; m[1][1] = t11; m[1][2] = t12; m[1][3] = t13;
    m[2][0] = t20; m[2][1] = t21; m[2][2] = t22; m[2][3] = t23;
    m[3][0] = t30; m[3][1] = t31; m[3][2] = t32; m[3][3] = t33;
  }
  Reference<Matrix4x4> Transpose() const;
  static Reference<Matrix4x4> Mul(const Reference<Matrix4x4> &m1,
                                  const Reference<Matrix4x4> &m2) {
    float r[4][4];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
        r[i][j] = m1->m[i][0] * m2->m[0][j] + m1->m[i][1] * m2->m[1][j] +
                  m1->m[i][2] * m2->m[2][j] + m1->m[i][3] * m2->m[3][j];
    return new Matrix4x4(r);
  }
  Reference<Matrix4x4> Inverse() const;
  float m[4][4];
};

#endif
