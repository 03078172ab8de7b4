// texture.h -- host half of the texture operands: the node table the materials refer to, image loading
// and MIPMap pyramid construction (core/mipmap.h:89-177, textures/imagemap.cpp:75-121, core/exrio.cpp:40-73).
#ifndef PBRT_B200_HOST_TEXTURE_H
#define PBRT_B200_HOST_TEXTURE_H
#include "paramset.h"
#include "transform.h"
#include "../../include/pbrt_b200.h"
#include <map>
#include <vector>

struct TextureStore {
  TextureStore();
  std::vector<PbrtB200Texture> nodes;
  std::vector<PbrtB200TexLevel> levels;
  std::vector<float> texels;
  float ewaLut[PBRT_B200_EWA_LUT_SIZE];
  // ImageTexture::GetTexture's cache (imagemap.cpp:75-89): one pyramid per (file, filter, anisotropy, wrap, type)
  std::map<string, std::pair<uint32_t, uint32_t> > pyramids;
  void Clear() { nodes.clear(); levels.clear(); texels.clear(); pyramids.clear(); imageSizes.clear(); lightDistributions.clear(); }
  int AddConstant(const float v[3], int channels);
  int AddImageMap(const Transform &tex2world, const TextureParams &tp, int channels);
  // InfiniteAreaLight's radiance map: `new MIPMap<Spectrum>(w, h, texels)` with the constructor's defaults
  // (lights/infinite.cpp:63-73); -1 when the file cannot be read (the light then keeps its constant Lbase)
  int AddRadianceMap(const string &filename);
  // MIPMap::Lookup(s, t, width) (trilinear) of an image-map node, on the host: InfiniteAreaLightIS's constructor
  // resamples its radiance map with it (lights/infinitesample.cpp:103-111)
  void LookupWidth(int node, float s, float t, float width, float rgb[3]) const;
  std::vector<float> lightDistributions;   // Distribution1D tables of the "infinitesample" lights
  // level-0 size of a pyramid's source image before resampling (ProjectionLight's aspect ratio)
  std::map<string, std::pair<int, int> > imageSizes;
  int AddScale(int a, int b, int channels);
  // textures/{uv,bilerp,checkerboard,dots,fbm,wrinkled,windy,marble}.cpp; -1 (with an error message) when the
  // parameters ask for something outside the scope (3D or supersampled checkerboards, float uv / marble)
  int AddProcedural(const string &texname, const Transform &tex2world, const TextureParams &tp, int channels,
                    int arg0, int arg1);
  int AddMix(int a, int b, int amount, int channels);
  int Depth(int node) const;
};
// Spectrum texels of an OpenEXR file (R, G, B read as HALF), NULL with an error message when unreadable
Spectrum *ReadImage(const string &name, int *width, int *height);
TextureStore &CurrentTextureStore();   // the texture table of the world block being described (api.cpp)
#endif
