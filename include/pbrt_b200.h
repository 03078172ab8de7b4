/*
 * pbrt_b200.h -- C ABI of the B200 render hot path (libpbrt_b200.so).
 *
 * This is the drop-in boundary for pbrt-v1's per-pixel render loop.  The one
 * reference seam it replaces is `void Scene::Render()` (core/scene.h:34,
 * core/scene.cpp:32-88), entered from pbrtWorldEnd (core/api.cpp:473-476):
 * the host flattens the Scene it built through the unchanged api.h surface
 * into the plain arrays below and calls pbrt_b200_render*().  Everything the
 * reference does per image sample below that seam -- Sampler::GetNextSample,
 * Camera::GenerateRay, Scene::Intersect/IntersectP (KdTreeAccel), the
 * SurfaceIntegrator::Li recursion, Film::AddSample and the resolve half of
 * ImageFilm::WriteImage -- runs on the GPU behind this interface.
 *
 * Plain C: pointers + sizes only, no C++/torch types.  All arrays are
 * borrowed for the duration of the call that receives them.
 *
 * Random numbers: the library implements the per-sample keyed counter RNG of
 * SURVEY.md App. H (oracle/keyed_rng.h is the normative statement) -- there
 * is no seed argument because the keys are (pixel x, pixel y, sample, stream).
 */
#ifndef PBRT_B200_H
#define PBRT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBRT_B200_ABI_VERSION 4

/* ---- error codes (Scene::Render has none: core/scene.cpp:32; the host maps
 *      non-zero codes to Severe(), core/util.cpp:90-97) ---- */
enum {
  PBRT_B200_OK = 0,
  PBRT_B200_ERR_INVALID = 1,      /* malformed scene / arguments            */
  PBRT_B200_ERR_UNSUPPORTED = 2,  /* feature outside the hot-path scope      */
  PBRT_B200_ERR_CUDA = 3,         /* CUDA runtime failure                    */
  PBRT_B200_ERR_NO_DEVICE = 4     /* no sm_100 device visible                */
};

/* ---- kd-tree: accelerators/kdtree.cpp:36-88 (KdAccelNode) ----
 * 8 bytes.  Interior: `split` float whose 2 low mantissa bits were replaced
 * by the axis (0..2) exactly as the reference's flags/split union does
 * (kdtree.cpp:66-72), `above` = index of the above child (below child is
 * node+1).  Leaf: low 2 bits == 3, nPrims = word0 >> 2; word1 = the single
 * primitive id when nPrims == 1, else offset of the list in prim_refs[]. */
typedef struct PbrtB200KdNode {
  uint32_t word0;
  uint32_t word1;
} PbrtB200KdNode;

/* ---- refined primitive: core/primitive.h:72-76 (GeometricPrimitive) with the
 *      shape data it points at pre-gathered (shapes/trianglemesh.cpp:213-222
 *      reads mesh->p[v[0..2]]) ---- */
enum { PBRT_B200_PRIM_TRIANGLE = 0, PBRT_B200_PRIM_QUADRIC = 1 /* sphere, cylinder, disk, cone, paraboloid, hyperboloid */,
       PBRT_B200_PRIM_INSTANCE = 2 /* InstancePrimitive, core/primitive.cpp:66-85 */ };
enum {
  PBRT_B200_PF_KIND_MASK = 0x3,
  PBRT_B200_PF_REVERSE   = 1u << 2,  /* Shape::reverseOrientation            */
  PBRT_B200_PF_SWAPS     = 1u << 3,  /* Shape::transformSwapsHandedness      */
  PBRT_B200_PF_HAS_UV    = 1u << 4,
  PBRT_B200_PF_HAS_N     = 1u << 5,
  PBRT_B200_PF_HAS_S     = 1u << 6
};
typedef struct PbrtB200Prim {     /* 48 bytes, 16-byte aligned records */
  float    p[9];                   /* triangle: world-space p1,p2,p3; quadric: unused */
  uint32_t flags;                  /* PBRT_B200_PF_* ; kind in the low 2 bits */
  uint32_t material;               /* index into materials[]                  */
  int32_t  light;                  /* area light index or -1                  */
} PbrtB200Prim;

/* per-primitive side record, read only by shading (never by traversal) */
typedef struct PbrtB200PrimAux {
  uint32_t v[3];                   /* triangle: global vertex ids into N/S/UV */
  uint32_t owner;                  /* triangle: mesh id; quadric: quadric id; instance: instance id */
} PbrtB200PrimAux;

/* shapes/trianglemesh.cpp:141-169: object<->world of the mesh (only needed to
 * move per-vertex N/S into world space, trianglemesh.cpp:98-103) */
typedef struct PbrtB200Mesh {
  float o2w[16];                   /* ObjectToWorld m, row major              */
  float w2o[16];                   /* its inverse                             */
} PbrtB200Mesh;

/* The quadrics: shapes/sphere.cpp:80-97, cylinder.cpp:51-59, disk.cpp:49-57, cone.cpp:41-48,
 * paraboloid.cpp:43-52, hyperboloid.cpp:46-71.  Intersected in object space (the reference's
 * WorldToObject(ray)), so each carries both matrices. */
enum { PBRT_B200_QUADRIC_SPHERE = 0, PBRT_B200_QUADRIC_CYLINDER = 1, PBRT_B200_QUADRIC_DISK = 2,
       PBRT_B200_QUADRIC_CONE = 3, PBRT_B200_QUADRIC_PARABOLOID = 4, PBRT_B200_QUADRIC_HYPERBOLOID = 5 };
typedef struct PbrtB200Quadric {
  float o2w[16];
  float w2o[16];
  float radius;                    /* all but hyperboloid                     */
  float zmin, zmax;                /* sphere, cylinder, paraboloid, hyperboloid */
  float theta_min, theta_max;      /* sphere                                  */
  float phi_max;                   /* radians                                 */
  uint32_t flags;                  /* PBRT_B200_PF_REVERSE | _SWAPS           */
  uint32_t kind;                   /* PBRT_B200_QUADRIC_*                     */
  float height;                    /* disk, cone                              */
  float inner_radius;              /* disk                                    */
  float p1[3], p2[3];              /* hyperboloid (after the constructor's swap) */
  float a, c;                      /* hyperboloid implicit-form coefficients  */
} PbrtB200Quadric;

/* InstancePrimitive (core/primitive.h:93-113, primitive.cpp:66-85; created by pbrtObjectInstance,
 * core/api.cpp:428-457): a ray is moved into instance space by WorldToInstance and intersected
 * with the instanced primitive -- the object's own kd-tree (KdTreeAccel over the ObjectBegin/End
 * block's primitives) or, when the block holds one intersectable shape, that primitive itself.
 * Nested trees live in the same nodes[] / prim_refs[] / prims[] arrays at the given offsets; their
 * node child indices, leaf offsets and primitive numbers are relative to those offsets. */
typedef struct PbrtB200Instance {
  float i2w[16];                   /* InstanceToWorld m, row major            */
  float w2i[16];                   /* its inverse (WorldToInstance m)         */
  float bounds[6];                 /* the nested KdTreeAccel::bounds (instance space) */
  uint32_t node_base, n_nodes;     /* n_nodes == 0: single primitive prims[prim_base] */
  uint32_t ref_base;
  uint32_t prim_base;
} PbrtB200Instance;

/* materials/{matte,glass,mirror,plastic,uber,shinymetal}.cpp with ConstantTexture
 * operands (core/texture.h:113-123); spectra are already Clamp()ed >= 0 */
enum { PBRT_B200_MAT_MATTE = 0, PBRT_B200_MAT_GLASS = 1, PBRT_B200_MAT_MIRROR = 2,
       PBRT_B200_MAT_PLASTIC = 3,    /* materials/plastic.cpp: Lambertian + Blinn microfacet */
       PBRT_B200_MAT_UBER = 4,       /* materials/uber.cpp: opacity pass-through + Lambertian + microfacet + specular */
       PBRT_B200_MAT_SHINYMETAL = 5, /* materials/shinymetal.cpp: conductor microfacet + conductor specular */
       PBRT_B200_MAT_TRANSLUCENT = 6,/* materials/translucent.cpp: Lambertian + microfacet, reflected (kr = reflect) and
                                        transmitted through BRDFToBTDF (kt = transmit) */
       PBRT_B200_MAT_SUBSTRATE = 7,  /* materials/substrate.cpp: FresnelBlend(Kd, Ks, Anisotropic(1/uroughness, 1/vroughness));
                                        roughness = uroughness, sigma = vroughness */
       PBRT_B200_MAT_MEASURED = 8    /* materials/{bluepaint,brushedmetal,clay,felt,primer,skin}.cpp: one three-lobe Lafortune
                                        BRDF with fixed fitted coefficients; index = which (PBRT_B200_MEASURED_*) */ };
enum { PBRT_B200_MEASURED_BLUEPAINT = 0, PBRT_B200_MEASURED_BRUSHEDMETAL = 1, PBRT_B200_MEASURED_CLAY = 2,
       PBRT_B200_MEASURED_FELT = 3, PBRT_B200_MEASURED_PRIMER = 4, PBRT_B200_MEASURED_SKIN = 5 };
/* Texture operands of the materials: a small expression graph whose leaves are constants and image maps
 * (textures/{constant,scale,mix,imagemap}.cpp).  A material slot whose tex[] entry is >= 0 is evaluated at
 * every hit (Material::GetBSDF's `Kd->Evaluate(dgs)`), the others keep the constant in the record.
 * Operand nodes always come before the node that uses them; graphs are at most 3 operators deep. */
enum { PBRT_B200_TEX_CONSTANT = 0, PBRT_B200_TEX_IMAGEMAP = 1, PBRT_B200_TEX_SCALE = 2, PBRT_B200_TEX_MIX = 3,
       /* procedural textures: textures/{uv,bilerp,checkerboard,dots,fbm,wrinkled,windy,marble}.cpp */
       PBRT_B200_TEX_UV = 4,            /* 2D mapping only */
       PBRT_B200_TEX_BILERP = 5,        /* param[0..11] = v00, v01, v10, v11 (3 floats each; a float texture repeats its value) */
       PBRT_B200_TEX_CHECKERBOARD = 6,  /* 2D; arg[0], arg[1] = tex1, tex2; iparam[0] = aamode (PBRT_B200_CHECKER_*) */
       PBRT_B200_TEX_DOTS = 7,          /* arg[0] = the "inside" parameter, arg[1] = "outside" (used swapped, like dots.cpp:38-42,89-90) */
       PBRT_B200_TEX_FBM = 8,           /* 3D: world_to_texture holds IdentityMapping3D's matrix; iparam[0] = octaves, param[0] = roughness */
       PBRT_B200_TEX_WRINKLED = 9,      /* as FBM, Turbulence() */
       PBRT_B200_TEX_WINDY = 10,
       PBRT_B200_TEX_MARBLE = 11        /* iparam[0] = octaves, param[0] = roughness, param[1] = scale, param[2] = variation */ };
enum { PBRT_B200_CHECKER_NONE = 0, PBRT_B200_CHECKER_CLOSEDFORM = 2 };   /* "supersample" draws random numbers: not supported */
enum { PBRT_B200_TEXMAP_UV = 0, PBRT_B200_TEXMAP_SPHERICAL = 1, PBRT_B200_TEXMAP_CYLINDRICAL = 2,
       PBRT_B200_TEXMAP_PLANAR = 3 };                       /* core/texture.cpp:66-149 */
enum { PBRT_B200_TEXWRAP_REPEAT = 0, PBRT_B200_TEXWRAP_BLACK = 1, PBRT_B200_TEXWRAP_CLAMP = 2 };   /* core/mipmap.h:30-34 */
#define PBRT_B200_EWA_LUT_SIZE 128                           /* WEIGHT_LUT_SIZE, core/mipmap.h:80 */
typedef struct PbrtB200TexLevel {    /* one level of a MIPMap pyramid (core/mipmap.h:136-152) */
  uint64_t offset;                   /* first float of the level in `texels`; row-major, `channels` floats per texel */
  uint32_t width, height;
} PbrtB200TexLevel;
typedef struct PbrtB200Texture {
  uint32_t kind;
  uint32_t channels;                 /* 3: Texture<Spectrum>, 1: Texture<float> */
  int32_t  arg[3];                   /* SCALE: tex1, tex2; MIX: tex1, tex2, amount */
  float    value[3];                 /* CONSTANT (a float constant repeats its value) */
  /* IMAGEMAP (textures/imagemap.cpp:130-195) */
  uint32_t mapping, wrap, trilinear;
  float    max_anisotropy;
  float    su, sv, du, dv;           /* uv: uscale, vscale, udelta, vdelta; planar: du, dv = udelta, vdelta */
  float    v1[3], v2[3];             /* planar */
  float    world_to_texture[16];     /* spherical / cylindrical */
  uint32_t first_level, n_levels;    /* tex_levels[first_level .. first_level + n_levels), finest first */
  float    param[12];                /* procedural textures, see PBRT_B200_TEX_* */
  int32_t  iparam[2];
} PbrtB200Texture;
enum { PBRT_B200_TEXSLOT_KD = 0, PBRT_B200_TEXSLOT_KS = 1, PBRT_B200_TEXSLOT_KR = 2, PBRT_B200_TEXSLOT_KT = 3,
       PBRT_B200_TEXSLOT_OPACITY = 4, PBRT_B200_TEXSLOT_SIGMA = 5, PBRT_B200_TEXSLOT_INDEX = 6,
       PBRT_B200_TEXSLOT_ROUGHNESS = 7,
       PBRT_B200_TEXSLOT_BUMP = 8,   /* "bumpmap": the displacement of Material::Bump (core/material.cpp:29-71); has no
                                        constant in the record -- absent (-1) means the reference's default constant 0 */
       PBRT_B200_TEXSLOTS = 9 };
typedef struct PbrtB200Material {    /* 112 bytes */
  int32_t kind;
  float   kd[3];                   /* matte / plastic / uber                  */
  float   ks[3];                   /* plastic / uber / shinymetal             */
  float   kr[3];                   /* glass / mirror / uber / shinymetal      */
  float   kt[3];                   /* glass                                   */
  float   opacity[3];              /* uber                                    */
  float   sigma;                   /* matte: clamped to [0,90] degrees        */
  float   index;                   /* glass: index of refraction              */
  float   roughness;               /* plastic / uber / shinymetal             */
  int32_t tex[PBRT_B200_TEXSLOTS]; /* texture node per operand (PBRT_B200_TEXSLOT_*), -1 = the constant above */
} PbrtB200Material;

/* lights/point.cpp:49-54, lights/area.cpp:28-57, core/shape.h:113-147 */
enum { PBRT_B200_LIGHT_POINT = 0, PBRT_B200_LIGHT_AREA = 1,
       PBRT_B200_LIGHT_DISTANT = 2 /* lights/distant.cpp */, PBRT_B200_LIGHT_SPOT = 3 /* lights/spot.cpp */,
       PBRT_B200_LIGHT_INFINITE = 4 /* lights/infinite.cpp: constant Lbase, or Lbase x radiance map */,
       PBRT_B200_LIGHT_GONIOMETRIC = 5 /* lights/goniometric.cpp: point light x map at the direction's (phi, theta), y/z swapped */,
       PBRT_B200_LIGHT_PROJECTION = 6  /* lights/projection.cpp: point light x map at the projected direction */ };
typedef struct PbrtB200Light {
  int32_t  kind;
  int32_t  n_samples;              /* Light::nSamples = max(1, ns)            */
  float    color[3];               /* point/spot: I ; area: Lemit ; distant: L ; infinite: Lbase */
  float    pos[3];                 /* point/spot: LightToWorld(0,0,0) ; distant: lightDir (normalised, world) */
  uint32_t first_shape;            /* area: range in light_shapes[]/light_cdf[] */
  uint32_t n_shapes;               /* 1 => the shape itself, >1 => ShapeSet   */
  float    area;                   /* shape->Area()                           */
  float    cos_total_width;        /* spot: cosf(Radians(coneangle))          */
  float    cos_falloff_start;      /* spot: cosf(Radians(coneangle - conedeltaangle)) */
  float    w2l[9];                 /* spot, infinite: upper 3x3 of WorldToLight (rows): Falloff(), radiance-map lookup */
  uint32_t radiance_map;           /* infinite, goniometric, projection: 1 + index in textures[] of the IMAGEMAP node
                                      holding the "mapname" pyramid (lights/infinite.cpp:63-86), 0 = no map (factor 1) */
  uint32_t dist_offset;            /* infinite with importance sampling (lights/infinitesample.cpp): first float of
                                      the light's tables in light_distributions, laid out
                                      uFunc[nu] uCdf[nu+1] uFuncInt vFuncInt[nu] { vFunc[nv] vCdf[nv+1] } x nu */
  uint32_t dist_nu, dist_nv;       /* 0 = the plain "infinite" light (cosine / uniform sampling) */
  float    projection[16];         /* projection: lightProjection = Perspective(fov, 1e-3, 1e30) (projection.cpp:80-82) */
  float    screen[4];              /* projection: screenX0, screenX1, screenY0, screenY1 (projection.cpp:71-78) */
} PbrtB200Light;

/* cameras/perspective.cpp:51-82, cameras/orthographic.cpp:49-78, cameras/environment.cpp:49-63,
 * core/camera.cpp:50-70 */
enum { PBRT_B200_CAMERA_PERSPECTIVE = 0, PBRT_B200_CAMERA_ORTHOGRAPHIC = 1, PBRT_B200_CAMERA_ENVIRONMENT = 2 };
typedef struct PbrtB200Camera {
  int32_t kind;                    /* PBRT_B200_CAMERA_*                      */
  float ray_origin[3];             /* environment: CameraToWorld(0,0,0)       */
  float raster_to_camera[16];      /* projective cameras                      */
  float camera_to_world[16];
  float hither, yon;
  float shutter_open, shutter_close;
  float lens_radius, focal_distance;
} PbrtB200Camera;

/* samplers/stratified.cpp:51-86, samplers/lowdiscrepancy.cpp:57-74 */
enum { PBRT_B200_SAMPLER_STRATIFIED = 0, PBRT_B200_SAMPLER_LOWDISCREPANCY = 1,
       PBRT_B200_SAMPLER_BESTCANDIDATE = 2 /* samplers/bestcandidate.cpp: placements of the 4096-entry sample table */,
       PBRT_B200_SAMPLER_RANDOM = 3 /* samplers/random.cpp: x_samples * y_samples uniform samples per pixel */ };
#define PBRT_B200_BC_TABLE_SIZE 4096   /* SAMPLE_TABLE_SIZE, samplers/bestcandidate.cpp:29-31 */
typedef struct PbrtB200Sampler {
  int32_t x_samples, y_samples;    /* stratified: strata per axis; lowdiscrepancy: pixelsamples (a power of 2), 1;
                                      bestcandidate: pixelsamples, 1 */
  int32_t jitter;                  /* stratified only */
  int32_t kind;                    /* PBRT_B200_SAMPLER_* */
} PbrtB200Sampler;

/* integrators/{whitted,directlighting,path}.cpp; volume integrator emission
 * contributes 2 one-dimensional sample slots (integrators/emission.cpp:42-46) */
enum { PBRT_B200_INT_WHITTED = 0, PBRT_B200_INT_DIRECT = 1, PBRT_B200_INT_PATH = 2 };
enum { PBRT_B200_STRATEGY_ALL = 0, PBRT_B200_STRATEGY_ONE = 1 };
typedef struct PbrtB200Integrator {
  int32_t kind;
  int32_t max_depth;
  int32_t strategy;                /* directlighting only                     */
  int32_t n_volume_1d;             /* 2 for "emission"                        */
} PbrtB200Integrator;

/* film/image.cpp:69-102 + the filter it owns (core/sampling.h:118-130) */
typedef struct PbrtB200Film {
  int32_t x_resolution, y_resolution;
  int32_t x_pixel_start, y_pixel_start, x_pixel_count, y_pixel_count; /* crop */
  float   filter_x_width, filter_y_width, filter_inv_x_width, filter_inv_y_width;
  float   filter_table[256];       /* 16x16, image.cpp:89-101                 */
  int32_t premultiply_alpha;
} PbrtB200Film;

typedef struct PbrtB200Scene {
  uint32_t abi_version;            /* PBRT_B200_ABI_VERSION                   */
  /* accelerator */
  const PbrtB200KdNode *nodes;      uint32_t n_nodes;      /* top-level tree first, then the instanced objects' trees */
  uint32_t n_top_nodes;             /* nodes of the top-level tree (== n_nodes without instancing) */
  const uint32_t       *prim_refs;  uint32_t n_prim_refs;
  float                 bounds[6];  /* KdTreeAccel::bounds pMin, pMax         */
  /* refined primitives, in the order KdTreeAccel sees them (kdtree.cpp:147-149); the top-level
   * tree's primitives first, then each instanced object's */
  const PbrtB200Prim    *prims;     uint32_t n_prims;
  const PbrtB200PrimAux *prim_aux;  /* n_prims                                */
  const PbrtB200Mesh    *meshes;    uint32_t n_meshes;
  const PbrtB200Quadric *quadrics;  uint32_t n_quadrics;
  const PbrtB200Instance *instances; uint32_t n_instances;
  /* optional per-vertex shading data, object space, 3/3/2 floats per vertex  */
  const float *vertex_n; const float *vertex_s; const float *vertex_uv;
  uint32_t n_vertices;
  /* shading */
  const PbrtB200Material *materials; uint32_t n_materials;
  const PbrtB200Light    *lights;    uint32_t n_lights;
  /* area lights attached to shapes inside an ObjectBegin block follow the n_lights sampled ones in `lights`:
   * pbrtShape keeps them on the primitive but leaves them out of the scene's light vector
   * (core/api.cpp:383-387), so they only emit when hit (Intersection::Le) and are never sampled */
  uint32_t n_unsampled_lights;
  /* texture graphs and MIPMap pyramids (path integrator only; NULL / 0 when every operand is constant) */
  const PbrtB200Texture  *textures;   uint32_t n_textures;
  const PbrtB200TexLevel *tex_levels; uint32_t n_tex_levels;
  const float *texels;                uint64_t n_texel_floats;
  const float *ewa_weight_lut;        /* PBRT_B200_EWA_LUT_SIZE floats, MIPMap::weightLut (core/mipmap.h:167-176) */
  const float *light_distributions;   uint64_t n_light_distribution_floats;   /* "infinitesample" lights' Distribution1Ds */
  const uint32_t *light_shapes;      /* primitive ids, ShapeSet order         */
  const float    *light_cdf;         /* ShapeSet::areaCDF, same indexing      */
  uint32_t n_light_shapes;
  /* bestcandidate sampler only: BestCandidateSampler::sampleTable, 4096 x {imageX, imageY, time, lensU, lensV} */
  const float *bestcandidate_table;
  PbrtB200Camera     camera;
  PbrtB200Sampler    sampler;
  PbrtB200Integrator integrator;
  PbrtB200Film       film;
} PbrtB200Scene;

/* A pixel window [x0,x0+w) x [y0,y0+h) of the film -- the reference's
 * cropwindow contract (film/image.cpp:79-85,148-156): the window's samples
 * include the filter apron, so every returned pixel is complete. */
typedef struct PbrtB200Window {
  int32_t x0, y0, width, height;
} PbrtB200Window;

/* Counters: the reference's post-render statistics (core/scene.cpp:80,
 * core/light.cpp:32, core/primitive.cpp:137, shapes/trianglemesh.cpp:217-247)
 * plus the kd-walk counters the roofline uses. */
typedef struct PbrtB200Stats {
  uint64_t camera_rays;            /* image samples traced                    */
  uint64_t intersect_rays;         /* Scene::Intersect calls (closest hit)    */
  uint64_t shadow_rays;            /* Scene::IntersectP calls (any hit)       */
  uint64_t points_shaded;          /* Intersection::GetBSDF calls             */
  uint64_t interior_visits, leaf_visits, prim_refs_read, prim_tests;
  uint64_t nan_samples, negative_samples, infinite_samples; /* scene.cpp:60-74 */
  double   render_ms;              /* device time of the sample loop          */
  double   upload_ms, download_ms; /* host<->device copies (e2e calls)        */
  uint64_t kernel_launches;
  uint64_t h2d_bytes, d2h_bytes;
} PbrtB200Stats;

typedef struct PbrtB200Handle PbrtB200Handle;   /* device-resident scene */

/* Library / device queries. */
int         pbrt_b200_abi_version(void);
int         pbrt_b200_device_count(void);
const char *pbrt_b200_last_error(void);

/* Upload a flattened scene to `device` (cudaSetDevice ordinal). */
int pbrt_b200_scene_create(const PbrtB200Scene *scene, int device,
                           PbrtB200Handle **out);
void pbrt_b200_scene_destroy(PbrtB200Handle *h);

/* Render `n_windows` pixel windows of an uploaded scene.  Results stay on the
 * device; fetch them with pbrt_b200_film_download().  `stats` may be NULL.
 * Replaces the `while (sampler->GetNextSample(sample))` loop of
 * core/scene.cpp:43-84 for those windows. */
int pbrt_b200_render(PbrtB200Handle *h, const PbrtB200Window *windows,
                     int n_windows, PbrtB200Stats *stats);

/* The same for TILES of one frame.  Every window must lie inside the scene's film window
 * (film/image.cpp:79-85).  A crop render (above) samples the reference's crop extent (image.cpp:148-156),
 * which misses the rare sample of the pixel row / column before the window whose coordinate rounds up to
 * the window's first integer coordinate (it splats into both pixels, image.cpp:108-111).  A tile instead
 * draws from every sample pixel of the FRAME's extent that can reach it and traces exactly the samples
 * whose filter footprint touches one of its pixels: its pixels are bit-identical to the same pixels of the
 * whole frame, and apron samples that contribute nothing are not traced.  This is the unit of the multi-GPU
 * partition below; the reference's counterpart is cropwindow renders pasted by tools/exrassemble.cpp:53-66.
 * Not available with the bestcandidate sampler (PBRT_B200_ERR_UNSUPPORTED). */
int pbrt_b200_render_tiles(PbrtB200Handle *h, const PbrtB200Window *tiles,
                           int n_tiles, PbrtB200Stats *stats);

/* ---- one frame on several GPUs (pbrt-v1_b200/csrc/multi.cuh) -------------------------------------------
 * A communicator names this rank among `world` ranks, one GPU each: processes of one node (torchrun: rank 0
 * calls pbrt_b200_comm_unique_id and broadcasts the 128 bytes) or threads of one process.  NCCL is loaded at
 * run time; without it these return PBRT_B200_ERR_UNSUPPORTED and single-GPU rendering is unaffected. */
typedef struct PbrtB200Comm PbrtB200Comm;
int  pbrt_b200_comm_unique_id(void *id128);
int  pbrt_b200_comm_create(const void *id128, int rank, int world, int device, PbrtB200Comm **out);
void pbrt_b200_comm_destroy(PbrtB200Comm *comm);
/* Render the scene's whole film window: the window is cut into square tiles (`tile_edge` pixels, 0 = chosen
 * from the frame size and the rank count), dealt statically to the ranks; each rank renders its share as
 * tiles (above) and ONE exchange -- an ncclSend of the rank's contiguous tile block, matched by grouped
 * ncclRecv's on rank 0 -- plus one paste kernel assemble the frame on rank 0's device, bit-identical to the
 * single-GPU frame.  rank 0: `rgba_host` (may be NULL) receives width*height*4 floats; other ranks ignore it.
 * comm == NULL renders on this device alone.  `stats` describes THIS rank's share. */
int pbrt_b200_render_frame(PbrtB200Handle *h, PbrtB200Comm *comm, int tile_edge, float *rgba_host,
                           PbrtB200Stats *stats);
/* The partition pbrt_b200_render_frame uses (pure host arithmetic): tiles of `frame` in raster order and the
 * rank that owns each; returns the number of tiles, writes at most max_tiles entries. */
int pbrt_b200_frame_partition(const PbrtB200Window *frame, int world, int tile_edge,
                              PbrtB200Window *tiles_out, int32_t *owner_out, int max_tiles);
/* rank 0: device pointer of the frame assembled by the last pbrt_b200_render_frame (float4 per pixel). */
int pbrt_b200_frame_device_ptr(PbrtB200Handle *h, void **dptr, uint64_t *n_floats);
/* Host-buffer call for `n_devices` GPUs of this process -- what Scene::Render() makes when PBRT_GPUS > 1
 * (SURVEY.md section 5): one thread per device uploads the scene, renders its tiles and joins the exchange;
 * `rgba_out` receives the film window; statistics are summed over the devices. */
int pbrt_b200_render_host_multi(const PbrtB200Scene *scene, const int *devices, int n_devices,
                                float *rgba_out, PbrtB200Stats *stats);

/* Copy the resolved pixels of window `w` of the last render to host memory:
 * `rgba` receives width*height*4 floats = exactly what ImageFilm::WriteImage
 * hands to WriteRGBAImage (film/image.cpp:157-209: XYZ round trip, /weightSum,
 * clamps, premultiply), row major.  `raw` (optional, may be NULL) receives the
 * unresolved film, 5 floats per pixel: L.r L.g L.b alpha weightSum
 * (film/image.cpp:57-64). */
int pbrt_b200_film_download(PbrtB200Handle *h, int w, float *rgba, float *raw);

/* Device pointer (float4 per pixel, row major) of window `w`'s resolved film,
 * for callers that exchange films GPU-to-GPU (NCCL) without a host bounce. */
int pbrt_b200_film_device_ptr(PbrtB200Handle *h, int w, void **dptr,
                              uint64_t *n_floats);

/* One-call form with HOST buffers on both sides: upload scene, render the
 * windows, download each window's rgba into rgba_out[i] (raw_out may be NULL
 * or hold NULL entries), free device memory.  This is what the host-side
 * Scene::Render() calls. */
int pbrt_b200_render_host(const PbrtB200Scene *scene, int device,
                          const PbrtB200Window *windows, int n_windows,
                          float *const *rgba_out, float *const *raw_out,
                          PbrtB200Stats *stats);

/* Tunables; returns the previous value, value < 0 only queries.
 * key 0: image samples per wave (default 128 Mi = up to ~43 GB of wave state for the path integrator, 16 Mi for whitted /
 *        directlighting; never more than half of the device's free memory; a frame whose sample count exceeds it
 *        is cut into equal waves)
 * key 1: collect the kd-walk counters of PbrtB200Stats (default off: they cost time)
 * key 2: resident CTAs per SM of the persistent kernels
 * key 3: bracket every kernel launch with CUDA events (pbrt_b200_kernel_stats)
 * key 5: streams of one wave: 0 = one; 1 = a path vertex's shadow + probe traversal on a second stream,
 *        beside the continuation kernel and the next closest-hit launch; 2 = only the probes on the second stream;
 *        3 (default) = like 1 with the probes on a third stream
 * key 6: print per-iteration queue sizes of every wave to stderr (profiling aid)
 * key 7: precompute the hit-independent BSDF frames of flat triangles and the light triangles' normals at
 *        scene create (default on; 0 evaluates them at every hit -- same bits, A/B switch for the profiles)
 * key 8: shade plain vertices (flat triangle, matte material) in the single-pass kernel k_path_fast (default on)
 * key 9: node order of the device copy of the top-level kd-tree, read at pbrt_b200_scene_create: 0 = the array as
 *        given (depth first, default), 1 = sibling pairs packed in 128-byte treelets (same tree, same visits)
 * key 10: waves of one frame in flight at once (path integrator): 1 (default; 2 measured no gain), or 2 = frames of 2 Mi samples and more are cut
 *        into an even number of waves that alternate between two lanes (arena + streams each), so that one wave's
 *        nearly empty last launches overlap the other's kernels; 3 = two lanes whatever the size (tests)
 * key 11: percent of the L2 set aside as a persisting access window over the kd-tree nodes, read at
 *        pbrt_b200_scene_create (default 0 = off; measured no gain) */
int64_t pbrt_b200_tunable(int key, int64_t value);

/* Per-kernel-class totals of the last pbrt_b200_render(): launches, summed
 * CUDA-event duration (key 3 on) and, for the three traversal kernels, the
 * kd-walk counters (key 1 on) that the roofline's algorithmic bytes are
 * computed from. */
typedef struct PbrtB200KernelStats {
  char     name[24];
  uint64_t launches;
  double   ms;
  uint64_t rays, interior_visits, leaf_visits, prim_refs_read, prim_tests;
} PbrtB200KernelStats;
int pbrt_b200_kernel_stats(PbrtB200Handle *h, PbrtB200KernelStats *out, int max_out, int *n_out);

/* ---- per-stage entry points used by the parity tests (tests/), each the
 *      device twin of one reference function ---- */

/* KdTreeAccel::Intersect (kdtree.cpp:313-403) for n rays given as 8 floats
 * (o.xyz, d.xyz, mint, maxt).  Out per ray: hit_prim (-1 = miss), t, b1, b2
 * (barycentrics, triangles only). */
int pbrt_b200_intersect(PbrtB200Handle *h, const float *rays, uint64_t n,
                        int32_t *hit_prim, float *hit_t, float *hit_b1,
                        float *hit_b2);
/* KdTreeAccel::IntersectP (kdtree.cpp:404-488): out 1 = occluded. */
int pbrt_b200_intersect_p(PbrtB200Handle *h, const float *rays, uint64_t n,
                          uint8_t *occluded);
/* Camera rays for the first `n` samples of window w, reference order
 * (pixel-major, sample-minor): 8 floats per ray + imageX, imageY. */
int pbrt_b200_camera_rays(PbrtB200Handle *h, const PbrtB200Window *w,
                          uint64_t n, float *rays, float *image_xy);
/* Per-sample radiance (after the NaN/neg/inf guards of scene.cpp:60-74) of
 * window w in reference order: 4 floats (L.rgb, alpha) per sample.  Keeping
 * it costs memory, so it is opt-in: call with ls == NULL and n != 0 BEFORE
 * pbrt_b200_render() to arm it (n == 0 disarms), then with a buffer after. */
int pbrt_b200_sample_radiance(PbrtB200Handle *h, int w, uint64_t n, float *ls);
/* Number of per-sample records window w of the last render holds.  Pixel samplers: the window's
 * image samples.  bestcandidate: 4096 per table placement, including the entries that fell outside
 * the sampler's extent and were skipped -- those carry alpha == -1. */
uint64_t pbrt_b200_sample_slots(PbrtB200Handle *h, int w);

#ifdef __cplusplus
}
#endif
#endif /* PBRT_B200_H */
