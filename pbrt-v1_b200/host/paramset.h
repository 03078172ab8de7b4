// paramset.h -- typed name->array parameter bag with the reference's lookup
// contract (core/paramset.h:54-132): Add* replaces an entry of the same name,
// FindOne* only matches single-item entries, every successful lookup marks the
// entry as used, ReportUnused() warns about the rest.
#ifndef PBRT_B200_HOST_PARAMSET_H
#define PBRT_B200_HOST_PARAMSET_H
#include "pbrt.h"
#include "geometry.h"

template <class T> struct ParamSetItem {
  ParamSetItem(const string &n, const T *v, int ni) : name(n), data(v, v + ni), lookedUp(false) {}
  string name;
  vector<T> data;
  mutable bool lookedUp;
};

template <class T> class ParamList {
 public:
  void Add(const string &name, const T *v, int n) {
    Erase(name);
    items.push_back(ParamSetItem<T>(name, v, n));
  }
  bool Erase(const string &name) {
    for (size_t i = 0; i < items.size(); ++i)
      if (items[i].name == name) { items.erase(items.begin() + i); return true; }
    return false;
  }
  const T *Find(const string &name, int *nItems) const {
    for (size_t i = 0; i < items.size(); ++i)
      if (items[i].name == name) {
        *nItems = (int)items[i].data.size();
        items[i].lookedUp = true;
        return items[i].data.empty() ? NULL : &items[i].data[0];
      }
    return NULL;
  }
  T FindOne(const string &name, const T &d) const {
    for (size_t i = 0; i < items.size(); ++i)
      if (items[i].name == name && items[i].data.size() == 1) {
        items[i].lookedUp = true;
        return items[i].data[0];
      }
    return d;
  }
  void ReportUnused() const {
    for (size_t i = 0; i < items.size(); ++i)
      if (!items[i].lookedUp) Warning("Parameter \"%s\" not used", items[i].name.c_str());
  }
  vector<ParamSetItem<T> > items;
};

// std::vector<bool> has no addressable storage, so bools are kept as chars.
class ParamSet {
 public:
  void AddFloat(const string &n, const float *d, int nItems = 1) { floats.Add(n, d, nItems); }
  void AddInt(const string &n, const int *d, int nItems = 1) { ints.Add(n, d, nItems); }
  void AddBool(const string &n, const bool *d, int nItems = 1) {
    vector<char> tmp(d, d + nItems);
    bools.Add(n, nItems ? &tmp[0] : NULL, nItems);
  }
  void AddPoint(const string &n, const Point *d, int nItems = 1) { points.Add(n, d, nItems); }
  void AddVector(const string &n, const Vector *d, int nItems = 1) { vectors.Add(n, d, nItems); }
  void AddNormal(const string &n, const Normal *d, int nItems = 1) { normals.Add(n, d, nItems); }
  void AddSpectrum(const string &n, const Spectrum *d, int nItems = 1) { spectra.Add(n, d, nItems); }
  void AddString(const string &n, const string *d, int nItems = 1) { strings.Add(n, d, nItems); }
  void AddTexture(const string &n, const string &value) { textures.Add(n, &value, 1); }
  bool EraseInt(const string &n) { return ints.Erase(n); }
  bool EraseBool(const string &n) { return bools.Erase(n); }
  bool EraseFloat(const string &n) { return floats.Erase(n); }
  bool ErasePoint(const string &n) { return points.Erase(n); }
  bool EraseVector(const string &n) { return vectors.Erase(n); }
  bool EraseNormal(const string &n) { return normals.Erase(n); }
  bool EraseSpectrum(const string &n) { return spectra.Erase(n); }
  bool EraseString(const string &n) { return strings.Erase(n); }
  bool EraseTexture(const string &n) { return textures.Erase(n); }
  float FindOneFloat(const string &n, float d) const { return floats.FindOne(n, d); }
  int FindOneInt(const string &n, int d) const { return ints.FindOne(n, d); }
  bool FindOneBool(const string &n, bool d) const { return bools.FindOne(n, (char)d) != 0; }
  Point FindOnePoint(const string &n, const Point &d) const { return points.FindOne(n, d); }
  Vector FindOneVector(const string &n, const Vector &d) const { return vectors.FindOne(n, d); }
  Normal FindOneNormal(const string &n, const Normal &d) const { return normals.FindOne(n, d); }
  Spectrum FindOneSpectrum(const string &n, const Spectrum &d) const { return spectra.FindOne(n, d); }
  string FindOneString(const string &n, const string &d) const { return strings.FindOne(n, d); }
  string FindTexture(const string &n) const { return textures.FindOne(n, string("")); }
  const float *FindFloat(const string &n, int *nItems) const { return floats.Find(n, nItems); }
  const int *FindInt(const string &n, int *nItems) const { return ints.Find(n, nItems); }
  const Point *FindPoint(const string &n, int *nItems) const { return points.Find(n, nItems); }
  const Vector *FindVector(const string &n, int *nItems) const { return vectors.Find(n, nItems); }
  const Normal *FindNormal(const string &n, int *nItems) const { return normals.Find(n, nItems); }
  const Spectrum *FindSpectrum(const string &n, int *nItems) const { return spectra.Find(n, nItems); }
  const string *FindString(const string &n, int *nItems) const { return strings.Find(n, nItems); }
  void ReportUnused() const;
  void Clear();
 private:
  ParamList<int> ints;
  ParamList<char> bools;
  ParamList<float> floats;
  ParamList<Point> points;
  ParamList<Vector> vectors;
  ParamList<Normal> normals;
  ParamList<Spectrum> spectra;
  ParamList<string> strings;
  ParamList<string> textures;
};

// core/paramset.h:165-215: geometry parameters shadow material parameters;
// a "texture" parameter names a previously declared texture, otherwise the
// literal becomes a ConstantTexture (core/paramset.cpp:435-465).  Only
// constant textures are inside the hot-path scope, so a texture is its value.
// `node` >= 0: the texture varies over the surface (image maps and expressions over them); it is entry
// `node` of the scene's texture table (texture.h) and `value` is unused.
struct FloatTexture : public ReferenceCounted { FloatTexture() : value(0.f), node(-1) {} float value; int node; };
struct SpectrumTexture : public ReferenceCounted { SpectrumTexture() : node(-1) {} Spectrum value; int node; };
class TextureParams {
 public:
  TextureParams(const ParamSet &geomp, const ParamSet &matp,
                map<string, Reference<FloatTexture> > &ft,
                map<string, Reference<SpectrumTexture> > &st)
      : geomParams(geomp), materialParams(matp), floatTextures(ft), spectrumTextures(st) {}
  // *node receives the texture-table entry when the operand is not constant (the returned value is then unused)
  Spectrum GetSpectrumTexture(const string &name, const Spectrum &def, int32_t *node = NULL) const;
  float GetFloatTexture(const string &name, float def, int32_t *node = NULL) const;
  string FindString(const string &n, const string &d = "") const {
    return geomParams.FindOneString(n, materialParams.FindOneString(n, d));
  }
  int FindInt(const string &n, int d) const { return geomParams.FindOneInt(n, materialParams.FindOneInt(n, d)); }
  bool FindBool(const string &n, bool d) const { return geomParams.FindOneBool(n, materialParams.FindOneBool(n, d)); }
  Vector FindVector(const string &n, const Vector &d) const {
    return geomParams.FindOneVector(n, materialParams.FindOneVector(n, d));
  }
  float FindFloat(const string &n, float d) const {
    return geomParams.FindOneFloat(n, materialParams.FindOneFloat(n, d));
  }
  Spectrum FindSpectrum(const string &n, const Spectrum &d) const {
    return geomParams.FindOneSpectrum(n, materialParams.FindOneSpectrum(n, d));
  }
  void ReportUnused() const { geomParams.ReportUnused(); materialParams.ReportUnused(); }
 private:
  const ParamSet &geomParams, &materialParams;
  map<string, Reference<FloatTexture> > &floatTextures;
  map<string, Reference<SpectrumTexture> > &spectrumTextures;
};
#endif
