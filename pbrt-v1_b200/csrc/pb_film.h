// pb_film.h -- ImageFilm arithmetic.
//   AddSample      film/image.cpp:103-141   (filter-table splat)
//   WriteImage     film/image.cpp:157-203   (XYZ round trip, normalise, clamp, premultiply)
//   sample extent  film/image.cpp:148-156
// The device accumulates by *gathering*: one thread per pixel walks the image
// samples that can reach it in the order the reference's serial loop adds
// them, so every pixel sees the same sequence of fp32 additions as on the CPU
// and a tile rendered alone is bit-identical to the same tile of a full render.
#ifndef PB_FILM_H
#define PB_FILM_H
#include "pb_scene.h"

#define PB_FILTER_TABLE_SIZE 16

// Does the sample at (imageX, imageY) touch pixel (x, y), and with what table weight?
PB_HD bool film_sample_weight(const SceneView &sv, const float *filterTable, float imageX,
                              float imageY, int x, int y, float *wt) {
  float dImageX = imageX - 0.5f;
  float dImageY = imageY - 0.5f;
  int x0 = pb_ceil2int_fin(dImageX - sv.filterXW);
  int x1 = pb_floor2int_fin(dImageX + sv.filterXW);
  int y0 = pb_ceil2int_fin(dImageY - sv.filterYW);
  int y1 = pb_floor2int_fin(dImageY + sv.filterYW);
  if (x < x0 || x > x1 || y < y0 || y > y1) return false;
  float fx = fabsf((x - dImageX) * sv.filterInvXW * PB_FILTER_TABLE_SIZE);
  int ifx = pb_imin(pb_floor2int_fin(fx), PB_FILTER_TABLE_SIZE - 1);
  float fy = fabsf((y - dImageY) * sv.filterInvYW * PB_FILTER_TABLE_SIZE);
  int ify = pb_imin(pb_floor2int_fin(fy), PB_FILTER_TABLE_SIZE - 1);
  *wt = filterTable[ify * PB_FILTER_TABLE_SIZE + ifx];
  return true;
}
// film/image.cpp:57-64
struct FilmPixel {
  float L[3], alpha, weightSum;
};
PB_HD void film_pixel_add(FilmPixel *px, float wt, const float L[3], float alpha) {
  px->L[0] += wt * L[0]; px->L[1] += wt * L[1]; px->L[2] += wt * L[2];   // AddWeighted
  px->alpha += alpha * wt;
  px->weightSum += wt;
}
// film/image.cpp:165-200
PB_HD void film_pixel_resolve(const FilmPixel &px, bool premultiply, float rgba[4]) {
  float xyz[3];
  xyz[0] = xyz[1] = xyz[2] = 0.f;
  const float XW[3] = {0.412453f, 0.357580f, 0.180423f};
  const float YW[3] = {0.212671f, 0.715160f, 0.072169f};
  const float ZW[3] = {0.019334f, 0.119193f, 0.950227f};
  for (int i = 0; i < 3; ++i) {
    xyz[0] += XW[i] * px.L[i];
    xyz[1] += YW[i] * px.L[i];
    xyz[2] += ZW[i] * px.L[i];
  }
  const float rWeight[3] = {3.240479f, -1.537150f, -0.498535f};
  const float gWeight[3] = {-0.969256f, 1.875991f, 0.041556f};
  const float bWeight[3] = {0.055648f, -0.204043f, 1.057311f};
  float r = rWeight[0] * xyz[0] + rWeight[1] * xyz[1] + rWeight[2] * xyz[2];
  float g = gWeight[0] * xyz[0] + gWeight[1] * xyz[1] + gWeight[2] * xyz[2];
  float b = bWeight[0] * xyz[0] + bWeight[1] * xyz[1] + bWeight[2] * xyz[2];
  float a = px.alpha;
  if (px.weightSum != 0.f) {
    float invWt = 1.f / px.weightSum;
    r = pb_clamp(r * invWt, 0.f, PB_INFINITY);
    g = pb_clamp(g * invWt, 0.f, PB_INFINITY);
    b = pb_clamp(b * invWt, 0.f, PB_INFINITY);
    a = pb_clamp(a * invWt, 0.f, 1.f);
  }
  if (premultiply) { r *= a; g *= a; b *= a; }
  rgba[0] = r; rgba[1] = g; rgba[2] = b; rgba[3] = a;
}
// film/image.cpp:148-156 for a pixel window used as the crop
PB_HD void film_sample_extent(const SceneView &sv, int px0, int py0, int pw, int ph, int *xstart,
                              int *xend, int *ystart, int *yend) {
  *xstart = pb_floor2int_fin(px0 + .5f - sv.filterXW);
  *xend = pb_floor2int_fin(px0 + .5f + pw + sv.filterXW);
  *ystart = pb_floor2int_fin(py0 + .5f - sv.filterYW);
  *yend = pb_floor2int_fin(py0 + .5f + ph + sv.filterYW);
}
#endif
