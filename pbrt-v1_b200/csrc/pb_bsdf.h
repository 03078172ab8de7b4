// pb_bsdf.h -- BSDF assembly and evaluation for the in-scope materials.
//   Material::Bump (constant displacement 0)   core/material.cpp:29-71
//   Matte / Glass / Mirror ::GetBSDF            materials/{matte,glass,mirror}.cpp
//   BSDF ctor, f, Sample_f, Pdf                 core/reflection.cpp:402-494
//   Lambertian, OrenNayar, SpecularReflection,
//   SpecularTransmission, FresnelDielectric     core/reflection.cpp:31-156, reflection.h:197-276
//   MC warps                                    core/mc.cpp:92-141, core/mc.h:38-59
// No arena, no virtual calls: a BSDF is at most two lobes described by the
// material record.
#ifndef PB_BSDF_H
#define PB_BSDF_H
#include "pb_shapes.h"

// BxDFType, core/reflection.h:54-71
enum {
  PB_REFLECTION = 1 << 0, PB_TRANSMISSION = 1 << 1, PB_DIFFUSE = 1 << 2, PB_GLOSSY = 1 << 3,
  PB_SPECULAR = 1 << 4,
  PB_ALL_TYPES = PB_DIFFUSE | PB_GLOSSY | PB_SPECULAR,
  PB_ALL = PB_REFLECTION | PB_TRANSMISSION | PB_ALL_TYPES
};
enum { PB_LOBE_LAMBERT = 0, PB_LOBE_ORENNAYAR = 1, PB_LOBE_SPEC_REFL_DIEL = 2,
       PB_LOBE_SPEC_REFL_NOOP = 3, PB_LOBE_SPEC_TRANS = 4,
       PB_LOBE_MICROFACET_BLINN = 5,      /* Microfacet(FresnelDielectric, Blinn): A = exponent */
       PB_LOBE_SPEC_REFL_COND = 6,        /* SpecularReflection(FresnelConductor(E, k = 0))      */
       PB_LOBE_MICROFACET_BLINN_COND = 7, /* Microfacet(FresnelConductor(E, k = 0), Blinn)       */
       PB_LOBE_FRESNEL_BLEND = 8,         /* FresnelBlend(R = Rd, E = Rs, Anisotropic(A, B))     */
       PB_LOBE_LAFORTUNE = 9              /* Lafortune(R = diffuse, three lobes): A = row of the coefficient table */ };
// Lobes a BSDF can hold: 2 covers matte / glass / mirror; the kernels compiled for scenes with
// glossy materials (template flag G) carry 4 (uber: transmission + Lambertian + microfacet + specular).
#define PB_MAX_LOBES 4
template <bool G> struct LobeLimit { static const int n = G ? PB_MAX_LOBES : 2; };

struct Lobe {
  int kind;
  int type;      // BxDFType bits
  Spec R;        // reflectance / transmittance
  float A, B;    // OrenNayar
  float etai, etat;
  Spec E;        // conductor lobes: FresnelConductor::eta (k = 0)
  int btdf;      // wrapped in BRDFToBTDF (core/reflection.h:143-167): evaluated with the other hemisphere's wi
};
struct Bsdf {
  V3 p;          // dgShading.p
  V3 nn, ng, sn, tn;
  float eta;
  int nLobes;
  Lobe lobes[PB_MAX_LOBES];
};

PB_HD bool lobe_matches(const Lobe &l, int flags) { return (l.type & flags) == l.type; }

// core/mc.cpp:92-135
PB_HD void concentric_sample_disk(float u1, float u2, float *dx, float *dy) {
  float r, theta;
  float sx = 2 * u1 - 1;
  float sy = 2 * u2 - 1;
  if (sx == 0.0f && sy == 0.0f) { *dx = 0.0f; *dy = 0.0f; return; }
  if (sx >= -sy) {
    if (sx > sy) {
      r = sx;
      if (sy > 0.0f) theta = sy / r;
      else theta = 8.0f + sy / r;
    } else {
      r = sy;
      theta = 2.0f - sx / r;
    }
  } else {
    if (sx <= sy) {
      r = -sx;
      theta = 4.0f - sy / r;
    } else {
      r = -sy;
      theta = 6.0f + sx / r;
    }
  }
  theta *= PB_PI / 4.f;
  *dx = r * pb_cosf(theta);
  *dy = r * pb_sinf(theta);
}
// core/mc.h:38-44
PB_HD V3 cosine_sample_hemisphere(float u1, float u2) {
  V3 ret;
  concentric_sample_disk(u1, u2, &ret.x, &ret.y);
  ret.z = sqrtf(pb_max(0.f, 1.f - ret.x * ret.x - ret.y * ret.y));
  return ret;
}
// core/mc.h:55-59
PB_HD float power_heuristic(int nf, float fPdf, int ng, float gPdf) {
  float f = nf * fPdf, g = ng * gPdf;
  return (f * f) / (f * f + g * g);
}

// core/reflection.cpp:31-39 with etai/etat the same in every channel
PB_HD float fr_diel(float cosi, float cost, float etai, float etat) {
  float Rparl = ((etat * cosi) - (etai * cost)) / ((etat * cosi) + (etai * cost));
  float Rperp = ((etai * cosi) - (etat * cost)) / ((etai * cosi) + (etat * cost));
  return (Rparl * Rparl + Rperp * Rperp) * (1.f / 2.f);   // Spectrum / float
}
// core/reflection.cpp:77-95
PB_HD float fresnel_dielectric(float cosi, float eta_i, float eta_t) {
  cosi = pb_clamp(cosi, -1.f, 1.f);
  bool entering = cosi > 0.f;
  float ei = eta_i, et = eta_t;
  if (!entering) { float tmp = ei; ei = et; et = tmp; }
  float sint = ei / et * sqrtf(pb_max(0.f, 1.f - cosi * cosi));
  if (sint >= 1.f) return 1.f;
  float cost = sqrtf(pb_max(0.f, 1.f - sint * sint));
  return fr_diel(fabsf(cosi), cost, ei, et);
}

// core/reflection.cpp:40-51 (FrCond) for FresnelConductor::Evaluate (:74-76) with k = 0 as
// shinymetal builds it; Spectrum arithmetic channel by channel in the reference's association
PB_HD Spec fr_cond(float cosi, const Spec &eta) {
  cosi = fabsf(cosi);
  Spec k = spec(0.f);
  Spec tmp = (eta * eta + k * k) * cosi * cosi;
  Spec Rparl2 = (tmp - (2.f * eta * cosi) + spec(1.f)) / (tmp + (2.f * eta * cosi) + spec(1.f));
  Spec tmp_f = eta * eta + k * k;
  Spec Rperp2 = (tmp_f - (2.f * eta * cosi) + spec(cosi * cosi)) / (tmp_f + (2.f * eta * cosi) + spec(cosi * cosi));
  return (Rparl2 + Rperp2) / 2.f;
}
// core/reflection.cpp:52-56
PB_HD Spec fresnel_approx_eta(const Spec &Fr) {
  Spec r = spec3(pb_clamp(Fr.c[0], 0.f, .999f), pb_clamp(Fr.c[1], 0.f, .999f), pb_clamp(Fr.c[2], 0.f, .999f));
  Spec sq = spec3(sqrtf(r.c[0]), sqrtf(r.c[1]), sqrtf(r.c[2]));
  return (spec(1.f) + sq) / (spec(1.f) - sq);
}

// reflection.h:31-52
PB_HD float sin_theta(const V3 &w) { return sqrtf(pb_max(0.f, 1.f - w.z * w.z)); }
PB_HD float sin_theta2(const V3 &w) { return pb_max(0.f, 1.f - w.z * w.z); }
PB_HD float cos_phi(const V3 &w) {
  float st = sin_theta(w);
  if (st == 0.f) return 1.f;
  return pb_clamp(w.x / st, -1.f, 1.f);
}
PB_HD float sin_phi(const V3 &w) {
  float st = sin_theta(w);
  if (st == 0.f) return 0.f;
  return pb_clamp(w.y / st, -1.f, 1.f);
}

// Blinn microfacet distribution, core/reflection.h:311-325, reflection.cpp:244-272
PB_HD float blinn_D(float exponent, const V3 &wh) {
  float costhetah = fabsf(wh.z);
  return (exponent + 2) * PB_INV_TWOPI * pb_powf(costhetah, exponent);
}
PB_NOINLINE float blinn_pdf(float exponent, const V3 &wo, const V3 &wi) {
  V3 H = normalize(wo + wi);
  float costheta = fabsf(H.z);
  float blinn = ((exponent + 1.f) * pb_powf(costheta, exponent)) / (2.f * PB_PI * 4.f * dot(wo, H));
  if (dot(wo, H) <= 0.f) blinn = 0.f;
  return blinn;
}
PB_HD void blinn_sample_f(float exponent, const V3 &wo, V3 *wi, float u1, float u2, float *pdf) {
  float costheta = pb_powf(u1, 1.f / (exponent + 1));
  float sintheta = sqrtf(pb_max(0.f, 1.f - costheta * costheta));
  float phi = u2 * 2.f * PB_PI;
  V3 H = v3(sintheta * pb_cosf(phi), sintheta * pb_sinf(phi), costheta);   // SphericalDirection
  if (!(wo.z * H.z > 0.f)) H = -H;
  *wi = -wo + 2.f * dot(wo, H) * H;
  float blinn = ((exponent + 1.f) * pb_powf(costheta, exponent)) / (2.f * PB_PI * 4.f * dot(wo, H));
  if (dot(wo, H) <= 0.f) blinn = 0.f;
  *pdf = blinn;
}
// Microfacet::G, core/reflection.h:293-301
PB_HD float microfacet_G(const V3 &wo, const V3 &wi, const V3 &wh) {
  float NdotWh = fabsf(wh.z);
  float NdotWo = fabsf(wo.z);
  float NdotWi = fabsf(wi.z);
  float WOdotWh = absdot(wo, wh);
  return pb_min(1.f, pb_min((2.f * NdotWh * NdotWo / WOdotWh), (2.f * NdotWh * NdotWi / WOdotWh)));
}
PB_HD float fresnel_dielectric(float cosi, float eta_i, float eta_t);
// Microfacet::f, core/reflection.cpp:163-175
PB_NOINLINE Spec microfacet_f(const Spec &R, float exponent, float etai, float etat, const V3 &wo, const V3 &wi,
                              const Spec *condEta = nullptr) {
  float cosThetaO = fabsf(wo.z);
  float cosThetaI = fabsf(wi.z);
  if (cosThetaI == 0.f || cosThetaO == 0.f) return spec(0.f);
  V3 wh = wi + wo;
  if (wh.x == 0.f && wh.y == 0.f && wh.z == 0.f) return spec(0.f);
  wh = normalize(wh);
  float cosThetaH = dot(wi, wh);
  Spec F = condEta ? fr_cond(cosThetaH, *condEta) : spec(fresnel_dielectric(cosThetaH, etai, etat));
  return R * blinn_D(exponent, wh) * microfacet_G(wo, wi, wh) * F / (4.f * cosThetaI * cosThetaO);
}

// Microfacet::Sample_f, reflection.cpp:235-239 (out of line: scenes without glossy materials only
// pay a branch for it)
PB_NOINLINE Spec microfacet_sample_f(const Spec &R, float exponent, float etai, float etat, const V3 &wo, V3 *wi,
                                     float u1, float u2, float *pdf, const Spec *condEta = nullptr) {
  blinn_sample_f(exponent, wo, wi, u1, u2, pdf);
  if (!(wo.z * wi->z > 0.f)) return spec(0.f);
  return microfacet_f(R, exponent, etai, etat, wo, *wi, condEta);
}

// Anisotropic microfacet distribution (core/reflection.h:326-345, reflection.cpp:273-332)
PB_HD float aniso_D(float ex, float ey, const V3 &wh) {
  float costhetah = fabsf(wh.z);
  float d = (1.f - costhetah * costhetah);
  if (d == 0.f) return 0.f;
  float e = (ex * wh.x * wh.x + ey * wh.y * wh.y) / d;
  return sqrtf((ex + 2.f) * (ey + 2.f)) * PB_INV_TWOPI * pb_powf(costhetah, e);
}
PB_HD float aniso_pdf_H(float ex, float ey, const V3 &wo, const V3 &H) {
  float costhetah = fabsf(H.z);
  float ds = 1.f - costhetah * costhetah;
  float anisotropic_pdf = 0.f;
  if (ds > 0.f && dot(wo, H) > 0.f) {
    float e = (ex * H.x * H.x + ey * H.y * H.y) / ds;
    float d = sqrtf((ex + 1.f) * (ey + 1.f)) * PB_INV_TWOPI * pb_powf(costhetah, e);
    anisotropic_pdf = d / (4.f * dot(wo, H));
  }
  return anisotropic_pdf;
}
PB_HD void aniso_first_quadrant(float ex, float ey, float u1, float u2, float *phi, float *costheta) {
  if (ex == ey) *phi = PB_PI * u1 * 0.5f;
  else *phi = pb_atanf(sqrtf((ex + 1) / (ey + 1)) * pb_tanf(PB_PI * u1 * 0.5f));
  float cosphi = pb_cosf(*phi), sinphi = pb_sinf(*phi);
  *costheta = pb_powf(u2, 1.f / (ex * cosphi * cosphi + ey * sinphi * sinphi + 1));
}
PB_NOINLINE void aniso_sample_f(float ex, float ey, const V3 &wo, V3 *wi, float u1, float u2, float *pdf) {
  float phi, costheta;
  if (u1 < .25f) {
    aniso_first_quadrant(ex, ey, 4.f * u1, u2, &phi, &costheta);
  } else if (u1 < .5f) {
    u1 = 4.f * (.5f - u1);
    aniso_first_quadrant(ex, ey, u1, u2, &phi, &costheta);
    phi = PB_PI - phi;
  } else if (u1 < .75f) {
    u1 = 4.f * (u1 - .5f);
    aniso_first_quadrant(ex, ey, u1, u2, &phi, &costheta);
    phi += PB_PI;
  } else {
    u1 = 4.f * (1.f - u1);
    aniso_first_quadrant(ex, ey, u1, u2, &phi, &costheta);
    phi = 2.f * PB_PI - phi;
  }
  float sintheta = sqrtf(pb_max(0.f, 1.f - costheta * costheta));
  V3 H = v3(sintheta * pb_cosf(phi), sintheta * pb_sinf(phi), costheta);   // SphericalDirection
  if (!(wo.z * H.z > 0.f)) H = -H;
  *wi = -wo + 2.f * dot(wo, H) * H;
  *pdf = aniso_pdf_H(ex, ey, wo, H);
}
// FresnelBlend (core/reflection.cpp:199-218,333-355, reflection.h:359-376): Rd = l.R, Rs = l.E
PB_NOINLINE Spec fresnel_blend_f(const Lobe &l, const V3 &wo, const V3 &wi) {
  const Spec Rd = l.R, Rs = l.E;
  Spec diffuse = (28.f / (23.f * PB_PI)) * Rd * (spec(1.f) - Rs) * (1 - pb_powf(1 - .5f * fabsf(wi.z), 5)) *
                 (1 - pb_powf(1 - .5f * fabsf(wo.z), 5));
  V3 H = normalize(wi + wo);
  float costheta = dot(wi, H);
  Spec schlick = Rs + pb_powf(1 - costheta, 5.f) * (spec(1.f) - Rs);
  Spec specular = aniso_D(l.A, l.B, H) / (4.f * absdot(wi, H) * pb_max(fabsf(wi.z), fabsf(wo.z))) * schlick;
  return diffuse + specular;
}
PB_HD float fresnel_blend_pdf(const Lobe &l, const V3 &wo, const V3 &wi) {
  if (!(wo.z * wi.z > 0.f)) return 0.f;
  return .5f * (fabsf(wi.z) * PB_INV_PI + aniso_pdf_H(l.A, l.B, wo, normalize(wo + wi)));
}

// The measured materials (materials/{bluepaint,brushedmetal,clay,felt,primer,skin}.cpp): Lafortune et al.'s fits to
// the Cornell reflectance measurements, three cosine lobes each.  Row = PBRT_B200_MEASURED_*; per row: diffuse rgb,
// then for each lobe its xy, z and exponent rgb triples.  These numbers are the materials' definition (published
// measurement data, as quoted by the reference's plugins); the code around them is not.
#define PB_LAFORTUNE_ROWS { \
  {0.3094f, 0.39667f, 0.70837f, 0.870567f, 0.857255f, 0.670982f, 0.803624f, 0.774290f, 0.586674f, 21.820103f, 18.597755f, 7.472717f, \
   -0.451218f, -0.406681f, -0.477976f, 0.023123f, 0.017625f, 0.227295f, 2.774499f, 2.581499f, 3.677653f, \
   -1.031545f, -1.029426f, -1.026588f, 0.706734f, 0.696530f, 0.687715f, 66.899060f, 63.767912f, 57.489181f}, \
  {0.f, 0.f, 0.f, -1.11854f, -1.11845f, -1.11999f, 1.01272f, 1.01469f, 1.01942f, 15.8708f, 15.6489f, 15.4571f, \
   -1.05334f, -1.06409f, -1.08378f, 0.69541f, 0.662178f, 0.626672f, 111.267f, 88.9222f, 65.2179f, \
   -1.01684f, -1.01635f, -1.01529f, 1.00132f, 1.00112f, 1.00108f, 180.181f, 184.152f, 195.773f}, \
  {0.383626f, 0.260749f, 0.274207f, -1.089701f, -1.102701f, -1.107603f, -1.354682f, -2.714801f, -1.569866f, 17.968505f, 11.024489f, 21.270282f, \
   -0.733381f, -0.793320f, -0.848206f, 0.676108f, 0.679314f, 0.726031f, 8.219745f, 9.055139f, 11.261951f, \
   -1.010548f, -1.012378f, -1.011263f, 0.910783f, 0.885239f, 0.892451f, 152.912795f, 141.937171f, 201.046802f}, \
  {0.025865f, 0.025865f, 0.025865f, -0.304075f, -0.304075f, -0.304075f, -0.065992f, -0.065992f, -0.065992f, 3.047892f, 3.047892f, 3.047892f, \
   -0.749561f, -0.749561f, -0.749561f, -1.167929f, -1.167929f, -1.167929f, 6.931827f, 6.931827f, 6.931827f, \
   1.004921f, 1.004921f, 1.004921f, -0.205529f, -0.205529f, -0.205529f, 94.117332f, 94.117332f, 94.117332f}, \
  {0.118230f, 0.121218f, 0.133209f, -0.399286f, -1.033473f, -1.058104f, 0.167504f, 0.009545f, -0.068002f, 2.466633f, 7.637253f, 8.117645f, \
   -1.041861f, -1.100108f, -1.087779f, 0.014375f, -0.198147f, -0.053605f, 7.993722f, 29.446268f, 41.988990f, \
   -1.098605f, -0.379883f, -0.449038f, -0.145110f, 0.159127f, 0.173224f, 31.899719f, 2.372852f, 2.636161f}, \
  {0.428425f, 0.301341f, 0.331054f, -1.131747f, -1.016939f, -0.966018f, -1.209182f, -1.462488f, -1.222419f, 6.421658f, 3.699932f, 3.524889f, \
   -0.546570f, -0.643533f, -0.638934f, 0.380123f, 0.410559f, 0.437367f, 3.685044f, 4.266495f, 4.539742f, \
   -0.998888f, -1.020153f, -1.027479f, 0.857998f, 0.703913f, 0.573625f, 64.208486f, 63.919687f, 43.809866f}}
#ifdef __CUDACC__
static __device__ const float pb_lafortune_dev[6][30] = PB_LAFORTUNE_ROWS;
#endif
PB_HD const float *lafortune_row(int which) {
#ifdef __CUDA_ARCH__
  return pb_lafortune_dev[which];
#else
  static const float rows[6][30] = PB_LAFORTUNE_ROWS;
  return rows[which];
#endif
}
// Lafortune::f, core/reflection.cpp:188-198 (Spectrum::Pow: c > 0 ? powf(c, e) : 0, color.h:141-147)
PB_NOINLINE Spec lafortune_f(const Lobe &l, const V3 &wo, const V3 &wi) {
  const float *t = lafortune_row((int)l.A);
  Spec ret = l.R * PB_INV_PI;
  for (int i = 0; i < 3; ++i) {
    const float *xy = t + 3 + 9 * i, *z = xy + 3, *e = xy + 6;
    Spec v;
    for (int c = 0; c < 3; ++c) {
      float vc = xy[c] * wo.x * wi.x + xy[c] * wo.y * wi.y + z[c] * wo.z * wi.z;
      v.c[c] = vc > 0 ? pb_powf(vc, e[c]) : 0.f;
    }
    ret = ret + v;
  }
  return ret;
}

// BxDF::f for the non-specular lobes (reflection.cpp:128-156); specular lobes return 0.
// G = false compiles the glossy (microfacet) lobe out: kernels for scenes without such materials
template <bool G = true>
PB_HD Spec lobe_f_brdf(const Lobe &l, const V3 &wo, const V3 &wi) {
  if (l.kind == PB_LOBE_LAMBERT) return l.R;   // R already holds R * INV_PI (reflection.h:249)
  if (G && l.kind == PB_LOBE_MICROFACET_BLINN) return microfacet_f(l.R, l.A, l.etai, l.etat, wo, wi);
  if (G && l.kind == PB_LOBE_MICROFACET_BLINN_COND) return microfacet_f(l.R, l.A, l.etai, l.etat, wo, wi, &l.E);
  if (G && l.kind == PB_LOBE_FRESNEL_BLEND) return fresnel_blend_f(l, wo, wi);
  if (G && l.kind == PB_LOBE_LAFORTUNE) return lafortune_f(l, wo, wi);
  if (l.kind == PB_LOBE_ORENNAYAR) {
    float sinthetai = sin_theta(wi);
    float sinthetao = sin_theta(wo);
    float maxcos = 0.f;
    if (sinthetai > 1e-4 && sinthetao > 1e-4) {
      float sinphii = sin_phi(wi), cosphii = cos_phi(wi);
      float sinphio = sin_phi(wo), cosphio = cos_phi(wo);
      float dcos = cosphii * cosphio + sinphii * sinphio;
      maxcos = pb_max(0.f, dcos);
    }
    float sinalpha, tanbeta;
    if (fabsf(wi.z) > fabsf(wo.z)) {
      sinalpha = sinthetao;
      tanbeta = sinthetai / fabsf(wi.z);
    } else {
      sinalpha = sinthetai;
      tanbeta = sinthetao / fabsf(wo.z);
    }
    return l.R * PB_INV_PI * (l.A + l.B * maxcos * sinalpha * tanbeta);
  }
  return spec(0.f);
}
// BxDF::Pdf (reflection.cpp:226-230); specular lobes return 0 (reflection.h:213,235)
// BRDFToBTDF::f (reflection.cpp:63-66): the wrapped BRDF with wi mirrored into wo's hemisphere
template <bool G = true>
PB_HD Spec lobe_f(const Lobe &l, const V3 &wo, const V3 &wi) {
  if (G && l.btdf) return lobe_f_brdf<G>(l, wo, v3(wi.x, wi.y, -wi.z));
  return lobe_f_brdf<G>(l, wo, wi);
}
template <bool G = true>
PB_HD float lobe_pdf_brdf(const Lobe &l, const V3 &wo, const V3 &wi) {
  if (l.type & PB_SPECULAR) return 0.f;
  if (G && (l.kind == PB_LOBE_MICROFACET_BLINN || l.kind == PB_LOBE_MICROFACET_BLINN_COND))   // Microfacet::Pdf, reflection.cpp:240-243
    return (wo.z * wi.z > 0.f) ? blinn_pdf(l.A, wo, wi) : 0.f;
  if (G && l.kind == PB_LOBE_FRESNEL_BLEND) return fresnel_blend_pdf(l, wo, wi);
  return (wo.z * wi.z > 0.f) ? fabsf(wi.z) * PB_INV_PI : 0.f;
}
// BRDFToBTDF::Pdf (reflection.cpp:231-234): brdf->Pdf(wo, -wi)
template <bool G = true>
PB_HD float lobe_pdf(const Lobe &l, const V3 &wo, const V3 &wi) {
  if (G && l.btdf) return lobe_pdf_brdf<G>(l, wo, -wi);
  return lobe_pdf_brdf<G>(l, wo, wi);
}
// BxDF::Sample_f for each lobe kind (reflection.cpp:96-127, 219-225)
template <bool G = true>
PB_HD Spec lobe_sample_f_brdf(const Lobe &l, const V3 &wo, V3 *wi, float u1, float u2, float *pdf) {
  if (l.kind == PB_LOBE_SPEC_REFL_DIEL || l.kind == PB_LOBE_SPEC_REFL_NOOP) {
    *wi = v3(-wo.x, -wo.y, wo.z);
    *pdf = 1.f;
    float F = (l.kind == PB_LOBE_SPEC_REFL_NOOP) ? 1.f : fresnel_dielectric(wo.z, l.etai, l.etat);
    // fresnel->Evaluate(..) * R / fabsf(CosTheta(*wi))
    return (spec(F) * l.R) / fabsf(wi->z);
  }
  if (G && l.kind == PB_LOBE_SPEC_REFL_COND) {
    *wi = v3(-wo.x, -wo.y, wo.z);
    *pdf = 1.f;
    return (fr_cond(wo.z, l.E) * l.R) / fabsf(wi->z);
  }
  if (l.kind == PB_LOBE_SPEC_TRANS) {
    bool entering = wo.z > 0.f;
    float ei = l.etai, et = l.etat;
    if (!entering) { float tmp = ei; ei = et; et = tmp; }
    float sini2 = sin_theta2(wo);
    float eta = ei / et;
    float sint2 = eta * eta * sini2;
    if (sint2 >= 1.f) return spec(0.f);   // *pdf stays 0 (set by the caller)
    float cost = sqrtf(pb_max(0.f, 1.f - sint2));
    if (entering) cost = -cost;
    float sintOverSini = eta;
    *wi = v3(sintOverSini * -wo.x, sintOverSini * -wo.y, cost);
    *pdf = 1.f;
    float F = fresnel_dielectric(wo.z, l.etai, l.etat);
    return ((et * et) / (ei * ei) * (spec(1.f) - spec(F)) * l.R) / fabsf(wi->z);
  }
  if (G && l.kind == PB_LOBE_MICROFACET_BLINN) return microfacet_sample_f(l.R, l.A, l.etai, l.etat, wo, wi, u1, u2, pdf);
  if (G && l.kind == PB_LOBE_MICROFACET_BLINN_COND)
    return microfacet_sample_f(l.R, l.A, l.etai, l.etat, wo, wi, u1, u2, pdf, &l.E);
  if (G && l.kind == PB_LOBE_FRESNEL_BLEND) {   // FresnelBlend::Sample_f, reflection.cpp:333-348
    if (u1 < .5) {
      u1 = 2.f * u1;
      *wi = cosine_sample_hemisphere(u1, u2);
      if (wo.z < 0.f) wi->z *= -1.f;
    } else {
      u1 = 2.f * (u1 - .5f);
      aniso_sample_f(l.A, l.B, wo, wi, u1, u2, pdf);
      if (!(wo.z * wi->z > 0.f)) return spec(0.f);   // (*pdf keeps the distribution's value)
    }
    *pdf = fresnel_blend_pdf(l, wo, *wi);
    return fresnel_blend_f(l, wo, *wi);
  }
  *wi = cosine_sample_hemisphere(u1, u2);
  if (wo.z < 0.f) wi->z *= -1.f;
  *pdf = lobe_pdf_brdf<G>(l, wo, *wi);
  return lobe_f_brdf<G>(l, wo, *wi);
}
// BRDFToBTDF::Sample_f (reflection.cpp:67-72): sample the BRDF, then mirror wi
template <bool G = true>
PB_HD Spec lobe_sample_f(const Lobe &l, const V3 &wo, V3 *wi, float u1, float u2, float *pdf) {
  Spec f = lobe_sample_f_brdf<G>(l, wo, wi, u1, u2, pdf);
  if (G && l.btdf) *wi = v3(wi->x, wi->y, -wi->z);
  return f;
}

PB_HD V3 bsdf_world_to_local(const Bsdf &b, const V3 &v) {
  return v3(dot(v, b.sn), dot(v, b.tn), dot(v, b.nn));
}
PB_HD V3 bsdf_local_to_world(const Bsdf &b, const V3 &v) {
  return v3(b.sn.x * v.x + b.tn.x * v.y + b.nn.x * v.z, b.sn.y * v.x + b.tn.y * v.y + b.nn.y * v.z,
            b.sn.z * v.x + b.tn.z * v.y + b.nn.z * v.z);
}
template <bool G = true>
PB_HD int bsdf_num_components(const Bsdf &b, int flags) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < LobeLimit<G>::n; ++i)
    if (i < b.nLobes && lobe_matches(b.lobes[i], flags)) ++n;
  return n;
}
// BSDF::f, reflection.cpp:480-494
template <bool G = true>
PB_HD Spec bsdf_f(const Bsdf &b, const V3 &woW, const V3 &wiW, int flags = PB_ALL) {
  V3 wi = bsdf_world_to_local(b, wiW), wo = bsdf_world_to_local(b, woW);
  if (dot(wiW, b.ng) * dot(woW, b.ng) > 0) flags &= ~PB_TRANSMISSION;
  else flags &= ~PB_REFLECTION;
  Spec f = spec(0.f);
#pragma unroll
  for (int i = 0; i < LobeLimit<G>::n; ++i)
    if (i < b.nLobes && lobe_matches(b.lobes[i], flags)) f = f + lobe_f<G>(b.lobes[i], wo, wi);
  return f;
}
// BSDF::Pdf, reflection.cpp:458-470
template <bool G = true>
PB_HD float bsdf_pdf(const Bsdf &b, const V3 &woW, const V3 &wiW, int flags = PB_ALL) {
  if (b.nLobes == 0) return 0.f;
  V3 wo = bsdf_world_to_local(b, woW), wi = bsdf_world_to_local(b, wiW);
  float pdf = 0.f;
  int matchingComps = 0;
#pragma unroll
  for (int i = 0; i < LobeLimit<G>::n; ++i)
    if (i < b.nLobes && lobe_matches(b.lobes[i], flags)) {
      ++matchingComps;
      pdf += lobe_pdf<G>(b.lobes[i], wo, wi);
    }
  return matchingComps > 0 ? pdf / matchingComps : 0.f;
}
// BSDF::Sample_f with explicit sample values, reflection.cpp:402-457
template <bool G = true>
PB_NOINLINE Spec bsdf_sample_f(const Bsdf &b, const V3 &woW, V3 *wiW, float u1, float u2, float u3,
                         float *pdf, int flags, int *sampledType) {
  int matchingComps = bsdf_num_components<G>(b, flags);
  if (matchingComps == 0) {
    *pdf = 0.f;
    *sampledType = 0;
    return spec(0.f);
  }
  int which = pb_imin(pb_floor2int(u3 * matchingComps), matchingComps - 1);
  int chosen = -1;
  int count = which;
  Lobe bx = b.lobes[0];
#pragma unroll
  for (int i = 0; i < LobeLimit<G>::n; ++i)
    if (chosen < 0 && i < b.nLobes && lobe_matches(b.lobes[i], flags))
      if (count-- == 0) { chosen = i; bx = b.lobes[i]; }
  V3 wi = v3(0, 0, 0);
  V3 wo = bsdf_world_to_local(b, woW);
  *pdf = 0.f;
  Spec f = lobe_sample_f<G>(bx, wo, &wi, u1, u2, pdf);
  if (*pdf == 0.f) {
    *sampledType = 0;
    return spec(0.f);
  }
  *sampledType = bx.type;
  *wiW = bsdf_local_to_world(b, wi);
  if (!(bx.type & PB_SPECULAR) && matchingComps > 1) {
#pragma unroll
    for (int i = 0; i < LobeLimit<G>::n; ++i)
      if (i < b.nLobes && i != chosen && lobe_matches(b.lobes[i], flags)) *pdf += lobe_pdf<G>(b.lobes[i], wo, wi);
  }
  if (matchingComps > 1) *pdf /= matchingComps;
  if (!(bx.type & PB_SPECULAR)) {
    f = spec(0.f);
    if (dot(*wiW, b.ng) * dot(woW, b.ng) > 0) flags &= ~PB_TRANSMISSION;
    else flags &= ~PB_REFLECTION;
#pragma unroll
    for (int i = 0; i < LobeLimit<G>::n; ++i)
      if (i < b.nLobes && lobe_matches(b.lobes[i], flags)) f = f + lobe_f<G>(b.lobes[i], wo, wi);
  }
  return f;
}

// ---- A BSDF holding exactly one Lambertian lobe: matte with sigma == 0 (materials/matte.cpp:46-64), the
// material of most path vertices.  bsdf_f / bsdf_pdf / bsdf_sample_f above specialised by hand for nLobes == 1,
// lobes[0] = {PB_LOBE_LAMBERT, PB_REFLECTION | PB_DIFFUSE, R = Kd * INV_PI}: the same operations on the values
// that reach a result (a Lambertian lobe reads only the z components of the local directions), without the
// lobe table.  The kernels use it where the material allows; the CPU twin always runs the general code, so the
// GPU == twin tests pin the two forms to each other.
struct LambertBsdf {
  V3 p;
  V3 nn, ng, sn, tn;
  Spec R;   // Lambertian::RoverPI
};
PB_HD V3 bsdf_local_to_world(const LambertBsdf &b, const V3 &v) {
  return v3(b.sn.x * v.x + b.tn.x * v.y + b.nn.x * v.z, b.sn.y * v.x + b.tn.y * v.y + b.nn.y * v.z,
            b.sn.z * v.x + b.tn.z * v.y + b.nn.z * v.z);
}
template <bool G = true>
PB_HD Spec bsdf_f(const LambertBsdf &b, const V3 &woW, const V3 &wiW, int = PB_ALL) {
  Spec f = spec(0.f);
  if (dot(wiW, b.ng) * dot(woW, b.ng) > 0) f = f + b.R;   // the reflection lobe counts on the same side only
  return f;
}
template <bool G = true>
PB_HD float bsdf_pdf(const LambertBsdf &b, const V3 &woW, const V3 &wiW, int = PB_ALL) {
  float woz = dot(woW, b.nn), wiz = dot(wiW, b.nn);
  float pdf = 0.f;
  pdf += (woz * wiz > 0.f) ? fabsf(wiz) * PB_INV_PI : 0.f;
  return pdf;   // (/ matchingComps == 1)
}
// (the component choice min(Floor2Int(u3 * 1), 0) is 0 for every u3 in [0, 1); the single diffuse lobe matches
// both flag sets the path integrator passes: BSDF_ALL and BSDF_ALL & ~BSDF_SPECULAR)
template <bool G = true>
PB_HD Spec bsdf_sample_f(const LambertBsdf &b, const V3 &woW, V3 *wiW, float u1, float u2, float, float *pdf,
                         int, int *sampledType) {
  float woz = dot(woW, b.nn);
  V3 wi = cosine_sample_hemisphere(u1, u2);
  if (woz < 0.f) wi.z *= -1.f;
  *pdf = (woz * wi.z > 0.f) ? fabsf(wi.z) * PB_INV_PI : 0.f;
  if (*pdf == 0.f) {
    *sampledType = 0;
    return spec(0.f);
  }
  *sampledType = PB_REFLECTION | PB_DIFFUSE;
  *wiW = bsdf_local_to_world(b, wi);
  Spec f = spec(0.f);
  if (dot(*wiW, b.ng) * dot(woW, b.ng) > 0) f = f + b.R;
  return f;
}

struct Bsdf;
template <bool G = true>
PB_HD void bsdf_set_lobes(const PbrtB200Material &m, Bsdf *b);
// Material::Bump with a constant-zero displacement (core/material.cpp:29-71):
// dpdu/dpdv are unchanged (x + 0*nn + 0*dndu), the normal is re-derived from
// them, re-oriented by the shape flags and then towards the geometric normal.
// Followed by the BSDF constructor (reflection.cpp:471-479) and the lobes the
// material adds.
// LOBES = false builds only the shading frame (the path integrator's surface kernel stores the frame
// and a later kernel rebuilds the lobes from the material record).
// The displacement samples of Material::Bump (core/material.cpp:33-56).  Every material calls Bump(): an absent
// "bumpmap" is the constant texture 0 (materials/*.cpp CreateMaterial), for which all three samples are 0.
struct BumpSamples {
  float du, dv;               // shift distances (.01 when the screen-space derivatives are zero)
  float uDisplace, vDisplace, displace;
};
template <bool LOBES = true>
PB_HD void make_bsdf(const PbrtB200Material &m, const DiffGeom &dgGeom, const DiffGeom &dgs,
                     uint32_t shapeFlags, Bsdf *b, const BumpSamples *bump = (const BumpSamples *)0) {
  // dgBump->dpdu = dgs.dpdu + (uDisplace - displace) / du * nn + displace * dndu (material.cpp:58-63),
  // evaluated literally also for the constant-0 displacement so that signed zeros / non-finite normals
  // behave as in the reference (0 / du is the same 0 for every du)
  float du = .01f, dv = .01f, uDisplace = 0.f, vDisplace = 0.f, displace = 0.f;
  if (bump) { du = bump->du; dv = bump->dv; uDisplace = bump->uDisplace; vDisplace = bump->vDisplace; displace = bump->displace; }
  V3 dpdu = dgs.dpdu + (uDisplace - displace) / du * dgs.nn + displace * dgs.dndu;
  V3 dpdv = dgs.dpdv + (vDisplace - displace) / dv * dgs.nn + displace * dgs.dndv;
  V3 nn = normalize(cross(dpdu, dpdv));
  bool ro = (shapeFlags & PBRT_B200_PF_REVERSE) != 0, sh = (shapeFlags & PBRT_B200_PF_SWAPS) != 0;
  if (ro ^ sh) nn = nn * -1.f;
  if (dot(dgGeom.nn, nn) < 0.f) nn = nn * -1.f;
  b->p = dgs.p;
  b->ng = dgGeom.nn;
  b->nn = nn;
  b->sn = normalize(dpdu);
  b->tn = cross(b->nn, b->sn);
  if (LOBES) bsdf_set_lobes<true>(m, b);
  else { b->eta = 1.f; b->nLobes = 0; }
}
// The BxDFs a material adds (materials/{matte,glass,mirror}.cpp); separate from the frame so a
// kernel can rebuild them from the material record instead of carrying them through memory.
template <bool G>
PB_HD void bsdf_set_lobes(const PbrtB200Material &m, Bsdf *b) {
  b->eta = 1.f;
  b->nLobes = 0;
#pragma unroll
  for (int i = 0; i < LobeLimit<G>::n; ++i) b->lobes[i].btdf = 0;
  const bool reflects = m.kind == PBRT_B200_MAT_GLASS || m.kind == PBRT_B200_MAT_MIRROR;
  Spec R = reflects ? spec3(m.kr[0], m.kr[1], m.kr[2]) : spec3(m.kd[0], m.kd[1], m.kd[2]);
  if (m.kind == PBRT_B200_MAT_MATTE) {
    Lobe &l = b->lobes[b->nLobes++];
    l.type = PB_REFLECTION | PB_DIFFUSE;
    l.etai = l.etat = 1.f;
    if (m.sigma == 0.f) {
      l.kind = PB_LOBE_LAMBERT;
      l.R = R * PB_INV_PI;   // Lambertian::RoverPI
      l.A = l.B = 0.f;
    } else {
      l.kind = PB_LOBE_ORENNAYAR;
      l.R = R;
      float sigma = (PB_PI / 180.f) * m.sigma;   // Radians()
      float sigma2 = sigma * sigma;
      l.A = 1.f - (sigma2 / (2.f * (sigma2 + 0.33f)));
      l.B = 0.45f * sigma2 / (sigma2 + 0.09f);
    }
  } else if (m.kind == PBRT_B200_MAT_GLASS) {
    b->eta = m.index;
    Spec T = spec3(m.kt[0], m.kt[1], m.kt[2]);
    if (!is_black(R)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_SPEC_REFL_DIEL; l.type = PB_REFLECTION | PB_SPECULAR;
      l.R = R; l.A = l.B = 0.f; l.etai = 1.f; l.etat = m.index;
    }
    if (!is_black(T)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_SPEC_TRANS; l.type = PB_TRANSMISSION | PB_SPECULAR;
      l.R = T; l.A = l.B = 0.f; l.etai = 1.f; l.etat = m.index;
    }
  } else if (G && m.kind == PBRT_B200_MAT_PLASTIC) {
    // materials/plastic.cpp:57-67: Lambertian(kd) then Microfacet(ks, FresnelDielectric(1.5, 1), Blinn(1/rough))
    Lobe &d = b->lobes[b->nLobes++];
    d.kind = PB_LOBE_LAMBERT; d.type = PB_REFLECTION | PB_DIFFUSE;
    d.R = R * PB_INV_PI; d.A = d.B = 0.f; d.etai = d.etat = 1.f;
    Lobe &g = b->lobes[b->nLobes++];
    g.kind = PB_LOBE_MICROFACET_BLINN; g.type = PB_REFLECTION | PB_GLOSSY;
    g.R = spec3(m.ks[0], m.ks[1], m.ks[2]);
    float e = 1.f / m.roughness;
    if (e > 1000.f || e != e) e = 1000.f;   // Blinn::Blinn, reflection.h:313
    g.A = e; g.B = 0.f; g.etai = 1.5f; g.etat = 1.f;
  } else if (G && m.kind == PBRT_B200_MAT_UBER) {
    // materials/uber.cpp:60-83: SpecularTransmission(1 - op, 1, 1) when op != 1, Lambertian(op * Kd),
    // Microfacet(op * Ks, FresnelDielectric(1.5, 1), Blinn(1/rough)), SpecularReflection(op * Kr, same Fresnel)
    Spec op = spec3(m.opacity[0], m.opacity[1], m.opacity[2]);
    if (!(op.c[0] == 1.f && op.c[1] == 1.f && op.c[2] == 1.f)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_SPEC_TRANS; l.type = PB_TRANSMISSION | PB_SPECULAR;
      l.R = spec3(-op.c[0], -op.c[1], -op.c[2]) + spec(1.f);
      l.A = l.B = 0.f; l.etai = 1.f; l.etat = 1.f;
    }
    Spec kd = op * spec3(m.kd[0], m.kd[1], m.kd[2]);
    if (!is_black(kd)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_LAMBERT; l.type = PB_REFLECTION | PB_DIFFUSE;
      l.R = kd * PB_INV_PI; l.A = l.B = 0.f; l.etai = l.etat = 1.f;
    }
    Spec ks = op * spec3(m.ks[0], m.ks[1], m.ks[2]);
    if (!is_black(ks)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_MICROFACET_BLINN; l.type = PB_REFLECTION | PB_GLOSSY;
      l.R = ks;
      float e = 1.f / m.roughness;
      if (e > 1000.f || e != e) e = 1000.f;
      l.A = e; l.B = 0.f; l.etai = 1.5f; l.etat = 1.f;
    }
    Spec kr = op * spec3(m.kr[0], m.kr[1], m.kr[2]);
    if (!is_black(kr)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_SPEC_REFL_DIEL; l.type = PB_REFLECTION | PB_SPECULAR;
      l.R = kr; l.A = l.B = 0.f; l.etai = 1.5f; l.etat = 1.f;
    }
  } else if (G && m.kind == PBRT_B200_MAT_TRANSLUCENT) {
    // materials/translucent.cpp:52-82: Lambertian(r * kd), BRDFToBTDF(Lambertian(t * kd)),
    // Microfacet(r * ks, FresnelDielectric(1.5, 1), Blinn(1/rough)), BRDFToBTDF(Microfacet(t * ks, ..))
    Spec r = spec3(m.kr[0], m.kr[1], m.kr[2]), t = spec3(m.kt[0], m.kt[1], m.kt[2]);
    if (!(is_black(r) && is_black(t))) {
      Spec kd = spec3(m.kd[0], m.kd[1], m.kd[2]);
      if (!is_black(kd)) {
        if (!is_black(r)) {
          Lobe &l = b->lobes[b->nLobes++];
          l.kind = PB_LOBE_LAMBERT; l.type = PB_REFLECTION | PB_DIFFUSE;
          l.R = (r * kd) * PB_INV_PI; l.A = l.B = 0.f; l.etai = l.etat = 1.f;
        }
        if (!is_black(t)) {
          Lobe &l = b->lobes[b->nLobes++];
          l.kind = PB_LOBE_LAMBERT; l.type = PB_TRANSMISSION | PB_DIFFUSE; l.btdf = 1;
          l.R = (t * kd) * PB_INV_PI; l.A = l.B = 0.f; l.etai = l.etat = 1.f;
        }
      }
      Spec ks = spec3(m.ks[0], m.ks[1], m.ks[2]);
      if (!is_black(ks)) {
        float e = 1.f / m.roughness;
        if (e > 1000.f || e != e) e = 1000.f;
        if (!is_black(r)) {
          Lobe &l = b->lobes[b->nLobes++];
          l.kind = PB_LOBE_MICROFACET_BLINN; l.type = PB_REFLECTION | PB_GLOSSY;
          l.R = r * ks; l.A = e; l.B = 0.f; l.etai = 1.5f; l.etat = 1.f;
        }
        if (!is_black(t)) {
          Lobe &l = b->lobes[b->nLobes++];
          l.kind = PB_LOBE_MICROFACET_BLINN; l.type = PB_TRANSMISSION | PB_GLOSSY; l.btdf = 1;
          l.R = t * ks; l.A = e; l.B = 0.f; l.etai = 1.5f; l.etat = 1.f;
        }
      }
    }
  } else if (G && m.kind == PBRT_B200_MAT_SUBSTRATE) {
    // materials/substrate.cpp:58-62: FresnelBlend(d, s, Anisotropic(1/u, 1/v)); exponents capped at
    // 1000 (reflection.h:329-332)
    Lobe &l = b->lobes[b->nLobes++];
    l.kind = PB_LOBE_FRESNEL_BLEND; l.type = PB_REFLECTION | PB_GLOSSY;
    l.R = spec3(m.kd[0], m.kd[1], m.kd[2]);
    l.E = spec3(m.ks[0], m.ks[1], m.ks[2]);
    float ex = 1.f / m.roughness, ey = 1.f / m.sigma;
    if (ex > 1000.f || ex != ex) ex = 1000.f;
    if (ey > 1000.f || ey != ey) ey = 1000.f;
    l.A = ex; l.B = ey; l.etai = l.etat = 1.f;
  } else if (G && m.kind == PBRT_B200_MAT_SHINYMETAL) {
    // materials/shinymetal.cpp:54-63: Microfacet(1, FresnelConductor(FresnelApproxEta(Ks), 0), Blinn(1/rough)),
    // SpecularReflection(1, FresnelConductor(FresnelApproxEta(Kr), 0))
    Lobe &g = b->lobes[b->nLobes++];
    g.kind = PB_LOBE_MICROFACET_BLINN_COND; g.type = PB_REFLECTION | PB_GLOSSY;
    g.R = spec(1.f);
    float e = 1.f / m.roughness;
    if (e > 1000.f || e != e) e = 1000.f;
    g.A = e; g.B = 0.f; g.etai = g.etat = 1.f;
    g.E = fresnel_approx_eta(spec3(m.ks[0], m.ks[1], m.ks[2]));
    Lobe &l = b->lobes[b->nLobes++];
    l.kind = PB_LOBE_SPEC_REFL_COND; l.type = PB_REFLECTION | PB_SPECULAR;
    l.R = spec(1.f); l.A = l.B = 0.f; l.etai = l.etat = 1.f;
    l.E = fresnel_approx_eta(spec3(m.kr[0], m.kr[1], m.kr[2]));
  } else if (G && m.kind == PBRT_B200_MAT_MEASURED) {
    // e.g. materials/bluepaint.cpp:54-56: Lafortune(diffuse, 3, xy, xy, z, e, REFLECTION | DIFFUSE); brushedmetal's is GLOSSY
    int which = (int)m.index;
    const float *t = lafortune_row(which);
    Lobe &l = b->lobes[b->nLobes++];
    l.kind = PB_LOBE_LAFORTUNE;
    l.type = PB_REFLECTION | (which == PBRT_B200_MEASURED_BRUSHEDMETAL ? PB_GLOSSY : PB_DIFFUSE);
    l.R = spec3(t[0], t[1], t[2]);
    l.A = (float)which; l.B = 0.f; l.etai = l.etat = 1.f;
  } else {   // mirror
    if (!is_black(R)) {
      Lobe &l = b->lobes[b->nLobes++];
      l.kind = PB_LOBE_SPEC_REFL_NOOP; l.type = PB_REFLECTION | PB_SPECULAR;
      l.R = R; l.A = l.B = 0.f; l.etai = l.etat = 1.f;
    }
  }
}
#endif
