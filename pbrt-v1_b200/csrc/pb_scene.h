// pb_scene.h -- device-side view of a flattened scene (the arrays of
// include/pbrt_b200.h after upload) plus the per-render constants the host
// library derives from them.
#ifndef PB_SCENE_H
#define PB_SCENE_H
#include "pb_math.h"
#include "pb_rng.h"
#include "../../include/pbrt_b200.h"
#ifndef __CUDACC__
struct float4 { float x, y, z, w; };   // (host builds of the shared headers: the twin never dereferences these tables)
#endif

#define PB_MAX_PATTERNS 64      // sample patterns addressable from constant memory
#define PB_KD_STACK 64          // MAX_TODO, accelerators/kdtree.cpp:323

// 16-byte view of the 48-byte primitive record: three aligned 128-bit loads
// fetch a triangle's nine vertex floats plus its flags/material/light words.
struct PrimQuad {
  float a, b, c, d;
};

struct SceneView {
  const PbrtB200KdNode *nodes;
  const uint32_t *primRefs;
  const PrimQuad *prims;            // 3 per primitive
  const PbrtB200PrimAux *aux;
  const PbrtB200Mesh *meshes;
  const PbrtB200Quadric *quadrics;
  const PbrtB200Instance *instances;
  const float *vertexN, *vertexS, *vertexUV;
  const PbrtB200Material *materials;
  const PbrtB200Light *lights;
  const uint32_t *lightShapes;
  const float *lightCdf;
  uint32_t nNodes, nPrims, nLights;   // nNodes: the top-level tree's
  uint32_t pairedNodes;               // device copy in treelet order (pb_treelet.h): word1 = below child, above = below + 1
  uint32_t nInstances;
  uint32_t nInfiniteLights;         // lights whose Le(ray) is not black (lights/infinite.cpp)
  float bmin[3], bmax[3];
  PbrtB200Camera camera;
  PbrtB200Sampler sampler;
  PbrtB200Integrator integrator;
  // film constants (film/image.cpp:69-102)
  int32_t xRes, yRes;
  float filterXW, filterYW, filterInvXW, filterInvYW;
  int32_t premultiply;
  // sample patterns (core/sampling.cpp:41-70 layout, see pb_integrator.h)
  int32_t nPatterns;
  SamplePattern patterns[PB_MAX_PATTERNS];
  // per-light pattern indices for directlighting "all" live in patterns[]:
  // light i uses patterns 2D {2i, 2i+1} and 1D {i}; see PatternIndex helpers.
  int32_t n1D, n2D;               // pattern counts; 2D patterns start at index n1D
  int32_t directRecords;          // EstimateDirect calls per shaded vertex
  // lowdiscrepancy sampler: the pixel's camera tables (image, lens, time) and the size of one
  // sample's value record (image 2 + lens 2 + every integrator pattern)
  SamplePattern ldCamera[3];
  int32_t ldFloatsPerSample;
  // bestcandidate sampler: the 4096 x 5 sample table and the width of one placement in pixels
  const float *bcTable;
  float bcTableWidth;
  // texture operands (pb_texture.h): node table, MIPMap levels, texels, EWA weight table
  const PbrtB200Texture *textures;
  const PbrtB200TexLevel *texLevels;
  const float *texels;
  const float *ewaLut;
  const float *lightDist;           // Distribution1D tables of the importance-sampled infinite lights
  const unsigned char *noisePerm;   // Perlin permutation, 512 entries (pb_noise_table.h), when a texture needs it
  uint32_t nTextures;
  // Device-only tables, filled once per pbrt_b200_scene_create by k_prim_frames / k_light_frames with the very
  // code a hit would run (null in the CPU twin, which evaluates every hit from scratch):
  //   primFrame   3 x float4 per primitive: the BSDF frame (ng | nn | sn | tn) of a flat triangle (no per-vertex
  //               N / S).  It depends on the triangle only -- dpdu / dpdv come from the vertices and uvs
  //               (trianglemesh.cpp:248-268), Bump with the constant displacement 0 and the BSDF constructor
  //               (material.cpp:29-71, reflection.cpp:471-479) add nothing hit dependent -- so a hit reads 48 bytes
  //               instead of re-deriving it.  Valid where the primitive's device flags carry PB_PF_DEV_FRAME.
  //   lightFrame  2 x float4 per entry of lightShapes[] that is a triangle: the DifferentialGeometry normal
  //               Shape::Pdf uses (shape.h:96-107) and the Ns Triangle::Sample returns (trianglemesh.cpp:336-349).
  const float4 *primFrame;
  const float4 *lightFrame;
};
#define PB_PF_DEV_FRAME (1u << 31)

struct Hit {
  int32_t prim;   // -1 = miss
  float t, b1, b2;
  int32_t inst;   // instance the primitive was reached through, or -1
};
// One kd-tree inside the scene's node / reference / primitive arrays (the top-level tree or an
// instanced object's).
struct KdTreeView {
  const PbrtB200KdNode *nodes;
  const uint32_t *primRefs;
  uint32_t nNodes, primBase;
  uint32_t paired;   // node order of pb_treelet.h (device copy of the top-level tree only)
  float bmin[3], bmax[3];
};
#endif
