// transform.h -- 4x4 transform with cached inverse (core/transform.h:31-135).
#ifndef PBRT_B200_HOST_TRANSFORM_H
#define PBRT_B200_HOST_TRANSFORM_H
#include "pbrt.h"
#include "geometry.h"

class Transform {
 public:
  Transform() { m = mInv = new Matrix4x4; }
  Transform(float mat[4][4]) {
    m = new Matrix4x4(mat[0][0], mat[0][1], mat[0][2], mat[0][3], mat[1][0],
                      mat[1][1], mat[1][2], mat[1][3], mat[2][0], mat[2][1],
                      mat[2][2], mat[2][3], mat[3][0], mat[3][1], mat[3][2],
                      mat[3][3]);
    mInv = m->Inverse();
  }
  Transform(const Reference<Matrix4x4> &mat) { m = mat; mInv = m->Inverse(); }
  Transform(const Reference<Matrix4x4> &mat, const Reference<Matrix4x4> &minv) {
    m = mat; mInv = minv;
  }
  Transform GetInverse() const { return Transform(mInv, m); }
  bool HasScale() const { return false; }   // transform.cpp:148-156 (disabled there too)
  // transform.h:73-83: row sums left to right, divide only when w != 1
  Point operator()(const Point &pt) const {
    float x = pt.x, y = pt.y, z = pt.z;
    float xp = m->m[0][0] * x + m->m[0][1] * y + m->m[0][2] * z + m->m[0][3];
    float yp = m->m[1][0] * x + m->m[1][1] * y + m->m[1][2] * z + m->m[1][3];
    float zp = m->m[2][0] * x + m->m[2][1] * y + m->m[2][2] * z + m->m[2][3];
    float wp = m->m[3][0] * x + m->m[3][1] * y + m->m[3][2] * z + m->m[3][3];
    if (wp == 1.) return Point(xp, yp, zp);
    else return Point(xp, yp, zp) / wp;
  }
  Vector operator()(const Vector &v) const {
    float x = v.x, y = v.y, z = v.z;
    return Vector(m->m[0][0] * x + m->m[0][1] * y + m->m[0][2] * z,
                  m->m[1][0] * x + m->m[1][1] * y + m->m[1][2] * z,
                  m->m[2][0] * x + m->m[2][1] * y + m->m[2][2] * z);
  }
  BBox operator()(const BBox &b) const;
  Transform operator*(const Transform &t2) const;
  bool SwapsHandedness() const;
  const Matrix4x4 &GetMatrix() const { return *m.get(); }
  const Matrix4x4 &GetInverseMatrix() const { return *mInv.get(); }
 private:
  Reference<Matrix4x4> m, mInv;
};

Transform Translate(const Vector &delta);
Transform Scale(float x, float y, float z);
Transform Rotate(float angle, const Vector &axis);
Transform LookAt(const Point &pos, const Point &look, const Vector &up);
Transform Perspective(float fov, float znear, float zfar);
#endif
