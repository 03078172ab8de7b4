// transform.cpp -- the transform builders of the scene-description API and the 4x4 inverse.
//
// Contract: every matrix this file produces must equal, bit for bit, the one pbrt-v1 computes for the
// same scene statement (core/transform.cpp:32-194, core/util.cpp:127-184), because the CTM reaches
// every vertex, ray and bound downstream.  That pins the *sequence of float operations* (which
// products are formed, in which association, which pivot a tie selects), not the program text: the
// code below is organised around small helpers of its own and is checked against the compiled
// reference by the fixture and fuzz tests (tests/test_oracle.py).
#include "transform.h"

Reference<Matrix4x4> Matrix4x4::Transpose() const {
  float t[4][4];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) t[c][r] = m[r][c];
  return new Matrix4x4(t);
}

namespace {
// Work state of the inversion: the matrix being reduced in place plus which rows / columns already
// hold a pivot.
struct Reduction {
  float a[4][4];
  bool rowDone[4], colDone[4];
  int fromRow[4], toCol[4];   // the (row, column) pair of each elimination round
};
// Largest magnitude among the entries whose row and column are both still free; of equal candidates
// the last one in row-major order wins (the comparison is >=, starting from 0).
bool pick_pivot(const Reduction &w, int *row, int *col) {
  float best = 0.f;
  *row = *col = -1;
  for (int r = 0; r < 4; ++r) {
    if (w.rowDone[r]) continue;
    for (int c = 0; c < 4; ++c) {
      if (w.colDone[c]) continue;
      float mag = fabsf(w.a[r][c]);
      if (mag >= best) { best = mag; *row = r; *col = c; }
    }
  }
  return *row >= 0;
}
void exchange_rows(float a[4][4], int r0, int r1) {
  for (int c = 0; c < 4; ++c) { float t = a[r0][c]; a[r0][c] = a[r1][c]; a[r1][c] = t; }
}
void exchange_columns(float a[4][4], int c0, int c1) {
  for (int r = 0; r < 4; ++r) { float t = a[r][c0]; a[r][c0] = a[r][c1]; a[r][c1] = t; }
}
}  // namespace

// In-place Gauss-Jordan elimination with full pivoting: each round moves the chosen pivot onto the
// diagonal by a row exchange, turns the pivot column into a unit column while the inverse builds up in
// the same storage, and remembers the exchange; the column exchanges that undo the row exchanges are
// applied last, newest first.
Reference<Matrix4x4> Matrix4x4::Inverse() const {
  Reduction w;
  memcpy(w.a, m, sizeof(w.a));
  // (a row and its column retire together: after the exchange the pivot sits on the diagonal)
  for (int i = 0; i < 4; ++i) w.rowDone[i] = w.colDone[i] = false;
  for (int round = 0; round < 4; ++round) {
    int pr, pc;
    if (!pick_pivot(w, &pr, &pc)) {
      Error("Singular matrix in MatrixInvert");
      return new Matrix4x4(w.a);
    }
    w.colDone[pc] = w.rowDone[pc] = true;
    if (pr != pc) exchange_rows(w.a, pr, pc);
    w.fromRow[round] = pr;
    w.toCol[round] = pc;
    float *prow = w.a[pc];
    if (prow[pc] == 0.f) Error("Singular matrix in MatrixInvert");
    const float scale = 1.f / prow[pc];
    prow[pc] = 1.f;
    for (int c = 0; c < 4; ++c) prow[c] *= scale;
    for (int r = 0; r < 4; ++r) {
      if (r == pc) continue;
      float *row = w.a[r];
      const float factor = row[pc];
      row[pc] = 0.f;
      for (int c = 0; c < 4; ++c) row[c] -= prow[c] * factor;
    }
  }
  for (int round = 3; round >= 0; --round)
    if (w.fromRow[round] != w.toCol[round]) exchange_columns(w.a, w.fromRow[round], w.toCol[round]);
  return new Matrix4x4(w.a);
}

namespace {
Matrix4x4 *diagonal_matrix(float x, float y, float z) {
  float d[4][4] = {{x, 0, 0, 0}, {0, y, 0, 0}, {0, 0, z, 0}, {0, 0, 0, 1}};
  return new Matrix4x4(d);
}
Matrix4x4 *offset_matrix(float x, float y, float z) {
  float d[4][4] = {{1, 0, 0, x}, {0, 1, 0, y}, {0, 0, 1, z}, {0, 0, 0, 1}};
  return new Matrix4x4(d);
}
}  // namespace

Transform Translate(const Vector &delta) {
  return Transform(offset_matrix(delta.x, delta.y, delta.z), offset_matrix(-delta.x, -delta.y, -delta.z));
}
Transform Scale(float x, float y, float z) {
  return Transform(diagonal_matrix(x, y, z), diagonal_matrix(1.f / x, 1.f / y, 1.f / z));
}

// Rotation by `angle` degrees about `axis` (Rodrigues' formula; a rotation's inverse is its transpose).
// Entry (i, j), i != j, is a_i a_j (1 - cos) -/+ a_k sin with k the third index and the sign that of the
// permutation (i, j, k); the diagonal is a_i^2 + (1 - a_i^2) cos.
Transform Rotate(float angle, const Vector &axis) {
  const Vector unit = Normalize(axis);
  const float a[3] = {unit.x, unit.y, unit.z};
  const float sine = sinf(Radians(angle)), cosine = cosf(Radians(angle));
  float r[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r[i][j] = (i == 3 && j == 3) ? 1.f : 0.f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      if (i == j) {
        r[i][j] = a[i] * a[i] + (1.f - a[i] * a[i]) * cosine;
        continue;
      }
      const int k = 3 - i - j;
      const float sym = a[i] * a[j] * (1.f - cosine);
      const bool cyclic = (j - i + 3) % 3 == 1;   // (i, j, k) is an even permutation of (0, 1, 2)
      r[i][j] = cyclic ? sym - a[k] * sine : sym + a[k] * sine;
    }
  Matrix4x4 *rot = new Matrix4x4(r);
  return Transform(rot, rot->Transpose());
}

// World-from-camera has the camera's right / up / viewing direction as its first three columns and the
// eye position as the fourth; the transform handed back is its inverse (world to camera).
Transform LookAt(const Point &pos, const Point &look, const Vector &up) {
  const Vector forward = Normalize(look - pos);
  const Vector right = Normalize(Cross(forward, up));
  const Vector trueUp = Cross(right, forward);
  const float columns[4][3] = {{right.x, right.y, right.z}, {trueUp.x, trueUp.y, trueUp.z},
                               {forward.x, forward.y, forward.z}, {pos.x, pos.y, pos.z}};
  float c2w[4][4];
  for (int c = 0; c < 4; ++c) {
    for (int r = 0; r < 3; ++r) c2w[r][c] = columns[c][r];
    c2w[3][c] = c == 3 ? 1.f : 0.f;
  }
  Matrix4x4 *camToWorld = new Matrix4x4(c2w);
  return Transform(camToWorld->Inverse(), camToWorld);
}

// Bounds of a transformed box: the union of its eight transformed corners.  The corners are visited in
// a fixed order (bit 0 / 1 / 2 of the code selects the max x / y / z) because min / max keep the first
// of two equal values, which decides the sign of a zero.
BBox Transform::operator()(const BBox &b) const {
  static const int cornerOrder[8] = {0, 1, 2, 4, 6, 3, 5, 7};
  BBox out;
  for (int n = 0; n < 8; ++n) {
    const int code = cornerOrder[n];
    const Point corner((code & 1) ? b.pMax.x : b.pMin.x, (code & 2) ? b.pMax.y : b.pMin.y,
                       (code & 4) ? b.pMax.z : b.pMin.z);
    const Point moved = (*this)(corner);
    out = n == 0 ? BBox(moved) : Union(out, moved);
  }
  return out;
}

// Composition keeps both directions: (A B)^-1 = B^-1 A^-1.
Transform Transform::operator*(const Transform &t2) const {
  return Transform(Matrix4x4::Mul(m, t2.m), Matrix4x4::Mul(t2.mInv, mInv));
}

// Sign of the determinant of the upper-left 3x3 block, expanded along its first row.
bool Transform::SwapsHandedness() const {
  const float(*a)[4] = m->m;
  auto minor2 = [a](int c0, int c1) { return a[1][c0] * a[2][c1] - a[1][c1] * a[2][c0]; };
  const float det = ((a[0][0] * minor2(1, 2)) - (a[0][1] * minor2(0, 2)) + (a[0][2] * minor2(0, 1)));
  return det < 0.f;
}

// Perspective projection: z in [n, f] goes to [0, 1] with the divide by z, then x and y are scaled so
// that the field of view spans [-1, 1].
Transform Perspective(float fov, float n, float f) {
  const float depthScale = 1.f / (f - n);
  float p[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, f * depthScale, -f * n * depthScale}, {0, 0, 1, 0}};
  const float cotHalfFov = 1.f / tanf(Radians(fov) / 2.f);
  return Scale(cotHalfFov, cotHalfFov, 1) * Transform(new Matrix4x4(p));
}
