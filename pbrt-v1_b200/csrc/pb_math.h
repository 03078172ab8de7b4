// pb_math.h -- fp32 vector / spectrum / transform arithmetic of the render
// hot path, written so that every operation rounds exactly like the
// reference's strict-fp32, no-FMA build (Makefile:24-25: -msse2 -mfpmath=sse):
//
//   * the CUDA side is compiled with -fmad=false -prec-div=true
//     -prec-sqrt=true -ftz=false, so + - * / sqrtf are IEEE identical;
//   * expression association follows the reference line by line
//     (core/geometry.h:30-334, core/color.h:29-210, core/transform.h:73-135);
//   * "x / f" on vectors and spectra multiplies by the rounded reciprocal,
//     as the reference's operator/ does (geometry.h:66-70, color.h:103-106);
//   * transcendental functions go through double precision and round once to
//     float, which reproduces glibc's sinf/cosf/atan2f/acosf results except
//     for a vanishing fraction of arguments ("FP parity").
//
// The functions are PB_HD so the same text also builds for the CPU twin
// (oracle/twin.cpp, test infrastructure only).
#ifndef PB_MATH_H
#define PB_MATH_H
#include <math.h>
#include <stdint.h>
#include <float.h>

#ifdef __CUDACC__
#define PB_HD __host__ __device__ __forceinline__
#define PB_NOINLINE __host__ __device__ __noinline__
#else
#define PB_HD inline
#define PB_NOINLINE inline
#endif

// core/pbrt.h:201-211 -- M_PI is a *float* literal in the reference
#define PB_PI 3.14159265358979323846f
#define PB_INV_PI 0.31830988618379067154f
#define PB_INV_TWOPI 0.15915494309189533577f
#define PB_RAY_EPSILON 1e-3f
#ifdef __CUDA_ARCH__
#define PB_INFINITY __int_as_float(0x7f800000)
#else
#define PB_INFINITY ((float)INFINITY)
#endif

#if (defined(PB_TWIN_LIBM_FLOAT) || defined(PB_HOST_LIBM_FLOAT)) && !defined(__CUDACC__)
// CPU builds that must round like the reference: the oracle twin's bit-comparison variant
// (PB_TWIN_LIBM_FLOAT) and the host front end's scene-load code (PB_HOST_LIBM_FLOAT, host/texture.cpp)
// call glibc's float functions directly
PB_HD float pb_sinf(float x) { return sinf(x); }
PB_HD float pb_cosf(float x) { return cosf(x); }
PB_HD float pb_acosf(float x) { return acosf(x); }
PB_HD float pb_atan2f(float y, float x) { return atan2f(y, x); }
PB_HD float pb_tanf(float x) { return tanf(x); }
PB_HD float pb_atanf(float x) { return atanf(x); }
#else
PB_NOINLINE float pb_sinf(float x) { return (float)sin((double)x); }
PB_NOINLINE float pb_cosf(float x) { return (float)cos((double)x); }
PB_NOINLINE float pb_acosf(float x) { return (float)acos((double)x); }
PB_NOINLINE float pb_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
PB_NOINLINE float pb_tanf(float x) { return (float)tan((double)x); }
PB_NOINLINE float pb_atanf(float x) { return (float)atan((double)x); }
#endif
#if (defined(PB_TWIN_LIBM_FLOAT) || defined(PB_HOST_LIBM_FLOAT)) && !defined(__CUDACC__)
PB_HD float pb_powf(float x, float y) { return powf(x, y); }
#else
PB_NOINLINE float pb_powf(float x, float y) { return (float)pow((double)x, (double)y); }
#endif

PB_HD float pb_max(float a, float b) { return a < b ? b : a; }   // std::max(a,b)
PB_HD float pb_min(float a, float b) { return b < a ? b : a; }   // std::min(a,b)
PB_HD int pb_imin(int a, int b) { return b < a ? b : a; }
PB_HD int pb_imax(int a, int b) { return a < b ? b : a; }
// core/pbrt.h:549-560
PB_HD float pb_lerp(float t, float v1, float v2) { return (1.f - t) * v1 + t * v2; }
PB_HD float pb_clamp(float val, float low, float high) {
  if (val < low) return low;
  else if (val > high) return high;
  else return val;
}
// core/pbrt.h:622-635 (FAST_INT is off): the argument is promoted to double
// Floor2Int / Ceil2Int (core/pbrt.h:622-635) take a double; for a float argument floorf / ceilf give the same
// integer exactly (the floor of a float is a float), without the fp64 round trip
// (int)x as the reference's x86 build performs it: cvttss2si returns the "integer indefinite" 0x80000000 for a NaN
// and for every value outside [-2^31, 2^31); the device's cvt.rzi.s32.f32 returns 0 for a NaN and saturates.  Only
// degenerate inputs differ (a NaN texture coordinate, an infinite film coordinate), but they decide which texel or
// pixel such a sample lands on.
PB_HD int pb_f2i(float v) {
#ifdef __CUDA_ARCH__
  return (v >= -2147483648.f && v < 2147483648.f) ? (int)v : (int)0x80000000;
#else
  return (int)v;
#endif
}
PB_HD int pb_floor2int(float v) { return pb_f2i(floorf(v)); }
PB_HD int pb_ceil2int(float v) { return pb_f2i(ceilf(v)); }
// the same for arguments the caller knows to be finite and far inside the int range (film and sample coordinates):
// the plain conversion, without the range test (the film gather converts per sample and pixel)
PB_HD int pb_floor2int_fin(float v) { return (int)floorf(v); }
PB_HD int pb_ceil2int_fin(float v) { return (int)ceilf(v); }

struct V3 {
  float x, y, z;
  PB_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
PB_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
PB_HD V3 operator+(const V3 &a, const V3 &b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB_HD V3 operator-(const V3 &a, const V3 &b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB_HD V3 operator-(const V3 &a) { return v3(-a.x, -a.y, -a.z); }
PB_HD V3 operator*(const V3 &a, float f) { return v3(f * a.x, f * a.y, f * a.z); }
PB_HD V3 operator*(float f, const V3 &a) { return v3(f * a.x, f * a.y, f * a.z); }
PB_HD V3 operator/(const V3 &a, float f) {   // geometry.h:66-70
  float inv = 1.f / f;
  return v3(a.x * inv, a.y * inv, a.z * inv);
}
PB_HD float dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PB_HD float absdot(const V3 &a, const V3 &b) { return fabsf(dot(a, b)); }
PB_HD V3 cross(const V3 &a, const V3 &b) {
  return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
PB_HD float length_sq(const V3 &a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
PB_HD float length(const V3 &a) { return sqrtf(length_sq(a)); }
PB_HD V3 normalize(const V3 &a) { return a / length(a); }
PB_HD float distance_sq(const V3 &a, const V3 &b) { return length_sq(a - b); }
// core/geometry.h:320-331
PB_HD void coordinate_system(const V3 &v1, V3 *v2, V3 *v3o) {
  if (fabsf(v1.x) > fabsf(v1.y)) {
    float invLen = 1.f / sqrtf(v1.x * v1.x + v1.z * v1.z);
    *v2 = v3(-v1.z * invLen, 0.f, v1.x * invLen);
  } else {
    float invLen = 1.f / sqrtf(v1.y * v1.y + v1.z * v1.z);
    *v2 = v3(0.f, v1.z * invLen, -v1.y * invLen);
  }
  *v3o = cross(v1, *v2);
}

// Spectrum = 3 floats RGB (core/pbrt.h:212, core/color.h:29-210)
struct Spec {
  float c[3];
};
PB_HD Spec spec(float v) { Spec s; s.c[0] = s.c[1] = s.c[2] = v; return s; }
PB_HD Spec spec3(float r, float g, float b) { Spec s; s.c[0] = r; s.c[1] = g; s.c[2] = b; return s; }
PB_HD Spec operator+(const Spec &a, const Spec &b) {
  return spec3(a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]);
}
PB_HD Spec operator-(const Spec &a, const Spec &b) {
  return spec3(a.c[0] - b.c[0], a.c[1] - b.c[1], a.c[2] - b.c[2]);
}
PB_HD Spec operator*(const Spec &a, const Spec &b) {
  return spec3(a.c[0] * b.c[0], a.c[1] * b.c[1], a.c[2] * b.c[2]);
}
PB_HD Spec operator/(const Spec &a, const Spec &b) {
  return spec3(a.c[0] / b.c[0], a.c[1] / b.c[1], a.c[2] / b.c[2]);
}
PB_HD Spec operator*(const Spec &a, float f) { return spec3(a.c[0] * f, a.c[1] * f, a.c[2] * f); }
PB_HD Spec operator*(float f, const Spec &a) { return spec3(a.c[0] * f, a.c[1] * f, a.c[2] * f); }
PB_HD Spec operator/(const Spec &a, float f) { return a * (1.f / f); }   // color.h:103-106
PB_HD bool is_black(const Spec &a) { return a.c[0] == 0.f && a.c[1] == 0.f && a.c[2] == 0.f; }
PB_HD bool spec_isnan(const Spec &a) { return a.c[0] != a.c[0] || a.c[1] != a.c[1] || a.c[2] != a.c[2]; }
// core/color.h:184-189, weights core/color.cpp:39-41
PB_HD float spec_y(const Spec &a) {
  float v = 0.f;
  v += 0.212671f * a.c[0];
  v += 0.715160f * a.c[1];
  v += 0.072169f * a.c[2];
  return v;
}

// Row-major 4x4 as exported by the host (Matrix4x4::m, core/pbrt.h:499-539).
struct M44 {
  float m[4][4];
};
// core/transform.h:73-83 (Point, by value)
PB_HD V3 xf_point(const float *M, const V3 &pt) {
  float x = pt.x, y = pt.y, z = pt.z;
  float xp = M[0] * x + M[1] * y + M[2] * z + M[3];
  float yp = M[4] * x + M[5] * y + M[6] * z + M[7];
  float zp = M[8] * x + M[9] * y + M[10] * z + M[11];
  float wp = M[12] * x + M[13] * y + M[14] * z + M[15];
  if (wp == 1.f) return v3(xp, yp, zp);
  float inv = 1.f / wp;   // Point::operator/ (geometry.h:134-138)
  return v3(inv * xp, inv * yp, inv * zp);
}
// core/transform.h:95-101
PB_HD V3 xf_vector(const float *M, const V3 &v) {
  float x = v.x, y = v.y, z = v.z;
  return v3(M[0] * x + M[1] * y + M[2] * z, M[4] * x + M[5] * y + M[6] * z,
            M[8] * x + M[9] * y + M[10] * z);
}
// core/transform.h:109-115: normals use the transposed inverse; Minv is the
// inverse matrix of the transform being applied.
PB_HD V3 xf_normal(const float *Minv, const V3 &n) {
  float x = n.x, y = n.y, z = n.z;
  return v3(Minv[0] * x + Minv[4] * y + Minv[8] * z, Minv[1] * x + Minv[5] * y + Minv[9] * z,
            Minv[2] * x + Minv[6] * y + Minv[10] * z);
}

struct Ray {
  V3 o, d;
  float mint, maxt;
};
PB_HD V3 ray_at(const Ray &r, float t) { return r.o + r.d * t; }   // geometry.h:214
#endif
