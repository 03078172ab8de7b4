// transform.cpp -- transform builders (core/transform.cpp:32-194) and the
// Gauss-Jordan inverse (core/util.cpp:127-184), restated in the reference's
// operation order so that CTMs match the reference bit-for-bit.
#include "transform.h"

Reference<Matrix4x4> Matrix4x4::Transpose() const {
  return new Matrix4x4(m[0][0], m[1][0], m[2][0], m[3][0], m[0][1], m[1][1],
                       m[2][1], m[3][1], m[0][2], m[1][2], m[2][2], m[3][2],
                       m[0][3], m[1][3], m[2][3], m[3][3]);
}

// core/util.cpp:127-184: full-pivot Gauss-Jordan, pivot chosen with >=.
Reference<Matrix4x4> Matrix4x4::Inverse() const {
  int indxc[4], indxr[4];
  int ipiv[4] = {0, 0, 0, 0};
  float minv[4][4];
  memcpy(minv, m, 4 * 4 * sizeof(float));
  for (int i = 0; i < 4; i++) {
    int irow = -1, icol = -1;
    float big = 0.;
    for (int j = 0; j < 4; j++) {
      if (ipiv[j] != 1) {
        for (int k = 0; k < 4; k++) {
          if (ipiv[k] == 0) {
            if (fabsf(minv[j][k]) >= big) {
              big = float(fabsf(minv[j][k]));
              irow = j;
              icol = k;
            }
          } else if (ipiv[k] > 1)
            Error("Singular matrix in MatrixInvert");
        }
      }
    }
    ++ipiv[icol];
    if (irow != icol) {
      for (int k = 0; k < 4; ++k) swap(minv[irow][k], minv[icol][k]);
    }
    indxr[i] = irow;
    indxc[i] = icol;
    if (minv[icol][icol] == 0.) Error("Singular matrix in MatrixInvert");
    float pivinv = 1.f / minv[icol][icol];
    minv[icol][icol] = 1.f;
    for (int j = 0; j < 4; j++) minv[icol][j] *= pivinv;
    for (int j = 0; j < 4; j++) {
      if (j != icol) {
        float save = minv[j][icol];
        minv[j][icol] = 0;
        for (int k = 0; k < 4; k++) minv[j][k] -= minv[icol][k] * save;
      }
    }
  }
  for (int j = 3; j >= 0; j--) {
    if (indxr[j] != indxc[j]) {
      for (int k = 0; k < 4; k++) swap(minv[k][indxr[j]], minv[k][indxc[j]]);
    }
  }
  return new Matrix4x4(minv);
}

Transform Translate(const Vector &delta) {
  Matrix4x4 *m = new Matrix4x4(1, 0, 0, delta.x, 0, 1, 0, delta.y, 0, 0, 1,
                               delta.z, 0, 0, 0, 1);
  Matrix4x4 *minv = new Matrix4x4(1, 0, 0, -delta.x, 0, 1, 0, -delta.y, 0, 0,
                                  1, -delta.z, 0, 0, 0, 1);
  return Transform(m, minv);
}
Transform Scale(float x, float y, float z) {
  Matrix4x4 *m = new Matrix4x4(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1);
  Matrix4x4 *minv = new Matrix4x4(1.f / x, 0, 0, 0, 0, 1.f / y, 0, 0, 0, 0,
                                  1.f / z, 0, 0, 0, 0, 1);
  return Transform(m, minv);
}
// transform.cpp:87-117 (Rodrigues form; inverse is the transpose)
Transform Rotate(float angle, const Vector &axis) {
  Vector a = Normalize(axis);
  float s = sinf(Radians(angle));
  float c = cosf(Radians(angle));
  float m[4][4];
  m[0][0] = a.x * a.x + (1.f - a.x * a.x) * c;
  m[0][1] = a.x * a.y * (1.f - c) - a.z * s;
  m[0][2] = a.x * a.z * (1.f - c) + a.y * s;
  m[0][3] = 0;
  m[1][0] = a.x * a.y * (1.f - c) + a.z * s;
  m[1][1] = a.y * a.y + (1.f - a.y * a.y) * c;
  m[1][2] = a.y * a.z * (1.f - c) - a.x * s;
  m[1][3] = 0;
  m[2][0] = a.x * a.z * (1.f - c) - a.y * s;
  m[2][1] = a.y * a.z * (1.f - c) + a.x * s;
  m[2][2] = a.z * a.z + (1.f - a.z * a.z) * c;
  m[2][3] = 0;
  m[3][0] = 0; m[3][1] = 0; m[3][2] = 0; m[3][3] = 1;
  Matrix4x4 *mat = new Matrix4x4(m);
  return Transform(mat, mat->Transpose());
}
// transform.cpp:118-146
Transform LookAt(const Point &pos, const Point &look, const Vector &up) {
  float m[4][4];
  m[0][3] = pos.x; m[1][3] = pos.y; m[2][3] = pos.z; m[3][3] = 1;
  Vector dir = Normalize(look - pos);
  Vector right = Normalize(Cross(dir, up));
  Vector newUp = Cross(right, dir);
  m[0][0] = right.x; m[1][0] = right.y; m[2][0] = right.z; m[3][0] = 0.;
  m[0][1] = newUp.x; m[1][1] = newUp.y; m[2][1] = newUp.z; m[3][1] = 0.;
  m[0][2] = dir.x; m[1][2] = dir.y; m[2][2] = dir.z; m[3][2] = 0.;
  Matrix4x4 *camToWorld = new Matrix4x4(m);
  return Transform(camToWorld->Inverse(), camToWorld);
}
BBox Transform::operator()(const BBox &b) const {
  const Transform &M = *this;
  BBox ret(M(Point(b.pMin.x, b.pMin.y, b.pMin.z)));
  ret = Union(ret, M(Point(b.pMax.x, b.pMin.y, b.pMin.z)));
  ret = Union(ret, M(Point(b.pMin.x, b.pMax.y, b.pMin.z)));
  ret = Union(ret, M(Point(b.pMin.x, b.pMin.y, b.pMax.z)));
  ret = Union(ret, M(Point(b.pMin.x, b.pMax.y, b.pMax.z)));
  ret = Union(ret, M(Point(b.pMax.x, b.pMax.y, b.pMin.z)));
  ret = Union(ret, M(Point(b.pMax.x, b.pMin.y, b.pMax.z)));
  ret = Union(ret, M(Point(b.pMax.x, b.pMax.y, b.pMax.z)));
  return ret;
}
Transform Transform::operator*(const Transform &t2) const {
  Reference<Matrix4x4> m1 = Matrix4x4::Mul(m, t2.m);
  Reference<Matrix4x4> m2 = Matrix4x4::Mul(t2.mInv, mInv);
  return Transform(m1, m2);
}
bool Transform::SwapsHandedness() const {
  float det = ((m->m[0][0] * (m->m[1][1] * m->m[2][2] - m->m[1][2] * m->m[2][1])) -
               (m->m[0][1] * (m->m[1][0] * m->m[2][2] - m->m[1][2] * m->m[2][0])) +
               (m->m[0][2] * (m->m[1][0] * m->m[2][1] - m->m[1][1] * m->m[2][0])));
  return det < 0.f;
}
// transform.cpp:180-194
Transform Perspective(float fov, float n, float f) {
  float inv_denom = 1.f / (f - n);
  Matrix4x4 *persp = new Matrix4x4(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, f * inv_denom,
                                   -f * n * inv_denom, 0, 0, 1, 0);
  float invTanAng = 1.f / tanf(Radians(fov) / 2.f);
  return Scale(invTanAng, invTanAng, 1) * Transform(persp);
}
