#include "paramset.h"

void ParamSet::ReportUnused() const {
  // same category order as core/paramset.cpp:242-254
  ints.ReportUnused(); bools.ReportUnused(); floats.ReportUnused(); points.ReportUnused();
  vectors.ReportUnused(); normals.ReportUnused(); spectra.ReportUnused();
  strings.ReportUnused(); textures.ReportUnused();
}
void ParamSet::Clear() {
  ints.items.clear(); bools.items.clear(); floats.items.clear(); points.items.clear();
  vectors.items.clear(); normals.items.clear(); spectra.items.clear();
  strings.items.clear(); textures.items.clear();
}
Spectrum TextureParams::GetSpectrumTexture(const string &n, const Spectrum &def, int32_t *node) const {
  if (node) *node = -1;
  string name = geomParams.FindTexture(n);
  if (name == "") name = materialParams.FindTexture(n);
  if (name != "") {
    map<string, Reference<SpectrumTexture> >::const_iterator it = spectrumTextures.find(name);
    if (it != spectrumTextures.end()) {
      if (it->second->node >= 0) {
        if (node) *node = it->second->node;
        else Error("Texture \"%s\" varies over the surface; parameter \"%s\" only takes constant textures here", name.c_str(), n.c_str());
      }
      return it->second->value;
    }
    Error("Couldn't find spectrumtexture named \"%s\"", n.c_str());
  }
  return geomParams.FindOneSpectrum(n, materialParams.FindOneSpectrum(n, def));
}
float TextureParams::GetFloatTexture(const string &n, float def, int32_t *node) const {
  if (node) *node = -1;
  string name = geomParams.FindTexture(n);
  if (name == "") name = materialParams.FindTexture(n);
  if (name != "") {
    map<string, Reference<FloatTexture> >::const_iterator it = floatTextures.find(name);
    if (it != floatTextures.end()) {
      if (it->second->node >= 0) {
        if (node) *node = it->second->node;
        else Error("Texture \"%s\" varies over the surface; parameter \"%s\" only takes constant textures here", name.c_str(), n.c_str());
      }
      return it->second->value;
    }
    Error("Couldn't find float texture named \"%s\"", n.c_str());
  }
  return geomParams.FindOneFloat(n, materialParams.FindOneFloat(n, def));
}
