// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// CPU restatement of the material closures the integrator calls through the "inner boundary"
// (SURVEY.md 8b): mdl_bsdf_scattering_{init,sample,evaluate}, mdl_edf_emission_{evaluate,intensity},
// mdl_volume_absorption_coefficient, mdl_thin_walled (call sites: rp_main.chit:157,170,306-329,
// 352-357,376-381,425).
//
// THIRD-PARTY ARITHMETIC, PARITY UNPINNED: in the reference these bodies are generated at run time
// by the NVIDIA MDL SDK 2024.1.4 (libbsdf) from MDL emitted by MaterialX's MdlShaderGenerator;
// neither is under /root/reference.  What follows restates the *published*
// definitions: MDL 1.8 spec section 21.1 (df::diffuse_reflection_bsdf, df::microfacet_ggx_smith_bsdf,
// df::custom_curve_layer, df::fresnel_layer, df::weighted_layer, df::diffuse_edf), Heitz 2018
// "Sampling the GGX Distribution of Visible Normals" (JCGT 7(4)), Shirley-Chiu concentric map,
// MaterialX 1.38/1.39 `ND_UsdPreviewSurface_surfaceshader` and the in-tree OpenPBR graph
// /root/reference/src/gi/mtlx/open_pbr_surface.mtlx:99-635.  Every choice that the SDK could make
// differently is listed in ORACLE_DEVIATIONS.md.
#pragma once
#include "orc_math.h"
#include "../include/gtlc.h"

namespace orc {

// mdl_types.glsl:123-137
enum {
  BSDF_EVENT_ABSORB = 0, BSDF_EVENT_DIFFUSE = 1, BSDF_EVENT_GLOSSY = 2, BSDF_EVENT_SPECULAR = 4,
  BSDF_EVENT_REFLECTION = 8, BSDF_EVENT_TRANSMISSION = 16
};

// mdl_types.glsl:32-113 (fields the closures read)
struct State {
  V3 normal, geom_normal, position, tangent_u, tangent_v;
  V3 text_coords;
  V3 local_position;   // object space (mdl_shading_state.glsl:23): the coordinate of 3-D texture bindings
};

// mdl_types.glsl:158-170
struct BsdfSampleData { V3 ior1, ior2, k1; V3 k2; V4 xi; float pdf; V3 bsdf_over_pdf; int event_type; };
// mdl_types.glsl:173-184
struct BsdfEvaluateData { V3 ior1, ior2, k1, k2; V3 bsdf_diffuse, bsdf_glossy; float pdf; };
// mdl_types.glsl:218-225
struct EdfEvaluateData { V3 k1; float cos; V3 edf; float pdf; };

static const float MIN_ALPHA = 1.0e-4f;   // roughness clamp (alpha), see ORACLE_DEVIATIONS.md
static const float INV_PI = 0.31830988618379067f;

struct Frame { V3 t, b, n; };
static inline Frame make_frame(const State& s)
{
  // libbsdf builds the local frame from the shading normal and tangent_u; tangent_v's sign
  // (bitangentSign) does not enter isotropic lobes.
  Frame f; f.n = s.normal; f.t = s.tangent_u; f.b = cross(s.normal, s.tangent_u);
  return f;
}
static inline V3 to_local(const Frame& f, V3 v) { return v3(dot(v, f.t), dot(v, f.b), dot(v, f.n)); }
static inline V3 to_world(const Frame& f, V3 v) { return (f.t * v.x + f.b * v.y) + f.n * v.z; }

// Shirley-Chiu concentric map, z-up cosine-weighted hemisphere.
static inline V3 cosine_hemisphere(float u, float v)
{
  float a = 2.0f * u - 1.0f;
  float b = 2.0f * v - 1.0f;
  if (a == 0.0f && b == 0.0f) return v3(0.0f, 0.0f, 1.0f);
  float r, phi;
  if ((a * a) > (b * b)) { r = a; phi = (PI / 4.0f) * (b / a); }
  else { r = b; phi = (PI / 2.0f) - (PI / 4.0f) * (a / b); }
  float x = r * cosf(phi), y = r * sinf(phi);
  float z = sqrtf(fmaxf(0.0f, 1.0f - r * r));
  return v3(x, y, z);
}

// ---- GGX (Trowbridge-Reitz), isotropic, separable Smith masking -----------------------------
static inline float ggx_D(V3 h, float a)
{
  float a2 = a * a;
  float d = ((h.x * h.x + h.y * h.y) / a2) + h.z * h.z;
  return 1.0f / ((PI * a2) * (d * d));
}
static inline float ggx_lambda(V3 v, float a)
{
  float z2 = v.z * v.z;
  if (z2 <= 0.0f) return FLOAT_MAX;
  float t2 = (a * a) * ((v.x * v.x + v.y * v.y) / z2);
  return (sqrtf(1.0f + t2) - 1.0f) * 0.5f;
}
static inline float ggx_G1(V3 v, float a) { return 1.0f / (1.0f + ggx_lambda(v, a)); }

// Heitz 2018, listing 1 (isotropic alpha).
static inline V3 ggx_sample_vndf(V3 Ve, float a, float u1, float u2)
{
  V3 Vh = normalize(v3(a * Ve.x, a * Ve.y, Ve.z));
  float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
  V3 T1 = (lensq > 0.0f) ? (v3(-Vh.y, Vh.x, 0.0f) * (1.0f / sqrtf(lensq))) : v3(1.0f, 0.0f, 0.0f);
  V3 T2 = cross(Vh, T1);
  float r = sqrtf(u1);
  float phi = (2.0f * PI) * u2;
  float t1 = r * cosf(phi);
  float t2 = r * sinf(phi);
  float s = 0.5f * (1.0f + Vh.z);
  t2 = (1.0f - s) * sqrtf(fmaxf(0.0f, 1.0f - t1 * t1)) + s * t2;
  V3 Nh = (T1 * t1 + T2 * t2) + Vh * sqrtf(fmaxf(0.0f, (1.0f - t1 * t1) - t2 * t2));
  return normalize(v3(a * Nh.x, a * Nh.y, fmaxf(0.0f, Nh.z)));
}

static inline float schlick_w(float c) { float m = clampf(1.0f - c, 0.0f, 1.0f); float m2 = m * m; return (m2 * m2) * m; }
static inline V3 schlick(V3 f0, V3 f90, float c) { float w = schlick_w(c); return f0 * (1.0f - w) + f90 * w; }
static inline float avg3(V3 c) { return ((c.x + c.y) + c.z) * (1.0f / 3.0f); }

// Unpolarised dielectric Fresnel reflectance, eta = n_transmitted / n_incident, cos_i >= 0.
static inline float fresnel_dielectric(float cos_i, float eta)
{
  cos_i = clampf(cos_i, 0.0f, 1.0f);
  float sin2_t = (1.0f - cos_i * cos_i) / (eta * eta);
  if (sin2_t >= 1.0f) return 1.0f;
  float cos_t = sqrtf(1.0f - sin2_t);
  float rs = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
  float rp = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
  return 0.5f * (rs * rs + rp * rp);
}

// One glossy reflection lobe sampled in the local frame.  Returns false on absorb.
static inline bool ggx_reflect_sample(V3 k1l, float a, float u1, float u2, V3& k2l, V3& h, float& g1_k2)
{
  h = ggx_sample_vndf(k1l, a, u1, u2);
  float kh = dot(k1l, h);
  k2l = h * (2.0f * kh) - k1l;
  if (k2l.z <= 0.0f || kh <= 0.0f) return false;
  g1_k2 = ggx_G1(k2l, a);
  return true;
}
// bsdf * cos(k2) and the VNDF pdf of a GGX reflection lobe (Fresnel excluded).
static inline void ggx_reflect_eval(V3 k1l, V3 k2l, float a, float& f_cos, float& pdf, V3& h)
{
  h = normalize(k1l + k2l);
  float D = ggx_D(h, a);
  float g1a = ggx_G1(k1l, a), g1b = ggx_G1(k2l, a);
  float denom = 4.0f * k1l.z;
  f_cos = ((D * g1a) * g1b) / denom;
  pdf = (D * g1a) / denom;
}

// --------------------------------------------------------------------------------------------
// Material view: closure id + parameter block (GtlcMaterialDesc.params)
// --------------------------------------------------------------------------------------------
struct Material { int closure; float p[GTLC_MATERIAL_PARAM_COUNT]; };

static inline V3 p3(const Material& m, int i) { return v3(m.p[i], m.p[i + 1], m.p[i + 2]); }

// mc/Material.h:28-45 flags the reference passes as #defines
static inline bool mat_is_emissive(const Material& m)
{
  if (m.closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) { V3 e = p3(m, GTLC_UPS_EMISSIVE_R); return e.x > 0.0f || e.y > 0.0f || e.z > 0.0f; }
  if (m.closure == GTLC_CLOSURE_OPENPBR) {
    V3 e = p3(m, GTLC_OPBR_EMISSION_COLOR_R);
    return m.p[GTLC_OPBR_EMISSION_LUMINANCE] > 0.0f && (e.x > 0.0f || e.y > 0.0f || e.z > 0.0f);
  }
  return false;
}
// mdl_cutout_opacity (rp_main.ahit:42): UsdPreviewSurface `opacity`, OpenPBR `geometry_opacity`
static inline float mat_cutout_opacity(const Material& m)
{
  float o = 1.0f;
  if (m.closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) o = m.p[GTLC_UPS_OPACITY];
  else if (m.closure == GTLC_CLOSURE_OPENPBR) o = m.p[GTLC_OPBR_GEOMETRY_OPACITY];
  return o < 0.0f ? 0.0f : (o > 1.0f ? 1.0f : o);
}
static inline bool mat_is_thin_walled(const Material& m)
{
  return m.closure == GTLC_CLOSURE_OPENPBR && m.p[GTLC_OPBR_GEOMETRY_THIN_WALLED] != 0.0f;
}
// mdl_edf_emission_intensity: MaterialX uniform_edf -> df::diffuse_edf() with intensity color*PI.
static inline V3 mat_emission_intensity(const Material& m)
{
  if (m.closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) return p3(m, GTLC_UPS_EMISSIVE_R) * PI;
  if (m.closure == GTLC_CLOSURE_OPENPBR)   // open_pbr_surface.mtlx:582-623: uniform_edf tinted by the coat colour
    return ((p3(m, GTLC_OPBR_EMISSION_COLOR_R) * m.p[GTLC_OPBR_EMISSION_LUMINANCE]) * PI) * mix(v3(1.0f), p3(m, GTLC_OPBR_COAT_COLOR_R), clampf(m.p[GTLC_OPBR_COAT_WEIGHT], 0.0f, 1.0f));
  return v3(0.0f);
}
// mdl_edf_emission_evaluate for df::diffuse_edf()
static inline void edf_evaluate(EdfEvaluateData& d, const State& s)
{
  d.cos = fmaxf(dot(d.k1, s.normal), 0.0f);
  d.edf = v3(INV_PI);
  d.pdf = d.cos * INV_PI;
}
// mdl_volume_scattering_coefficient (OpenPBR: open_pbr_surface.mtlx:238-245,286-292: transmission_scatter / transmission_depth)
static inline V3 mat_volume_scattering(const Material& m)
{
  if (m.closure != GTLC_CLOSURE_OPENPBR) return v3(0.0f);
  float depth = m.p[GTLC_OPBR_TRANSMISSION_DEPTH];
  if (!(depth > 0.0f)) return v3(0.0f);
  V3 sc = p3(m, GTLC_OPBR_TRANSMISSION_SCATTER_R);
  return v3(sc.x / depth, sc.y / depth, sc.z / depth);
}
// mdl_volume_absorption_coefficient (OpenPBR: open_pbr_surface.mtlx:221-285): extinction = -ln(color) / depth,
// absorption = extinction - scattering, shifted to be non-negative
static inline V3 mat_volume_absorption(const Material& m)
{
  if (m.closure != GTLC_CLOSURE_OPENPBR) return v3(0.0f);
  float depth = m.p[GTLC_OPBR_TRANSMISSION_DEPTH];
  if (!(depth > 0.0f)) return v3(0.0f);
  V3 c = p3(m, GTLC_OPBR_TRANSMISSION_COLOR_R);
  // colour clamped away from 0 (ln(0) would give an infinite coefficient)
  V3 ext = v3(-logf(fmaxf(c.x, 1.0e-6f)) / depth, -logf(fmaxf(c.y, 1.0e-6f)) / depth, -logf(fmaxf(c.z, 1.0e-6f)) / depth);
  V3 a = ext - mat_volume_scattering(m);
  float mn = fminf(fminf(a.x, a.y), a.z);
  if (0.0f > mn) a = a - v3(mn);                                     // mtlx:263-279
  return a;
}
// mdl_ior: the MaterialX surface constructor (open_pbr_surface.mtlx:626-630) leaves material.ior at the MDL default
static inline V3 mat_ior(const Material&) { return v3(1.0f); }
// MEDIUM_DIRECTIONAL_BIAS (GlslShaderGen.cpp:212, Frontend.cpp:137-157): volume.scattering.directional_bias when constant
static inline float mat_directional_bias(const Material& m)
{
  return m.closure == GTLC_CLOSURE_OPENPBR ? m.p[GTLC_OPBR_TRANSMISSION_SCATTER_ANISOTROPY] : 0.0f;
}

// ---- closure 0: df::diffuse_reflection_bsdf(tint) -------------------------------------------
static inline void diffuse_sample(const Material& m, BsdfSampleData& d, const State& s)
{
  Frame f = make_frame(s);
  V3 l = cosine_hemisphere(d.xi.x, d.xi.y);
  d.k2 = normalize(to_world(f, l));
  d.pdf = l.z * INV_PI;
  if (l.z <= 0.0f || dot(d.k2, s.geom_normal) <= 0.0f) { d.event_type = BSDF_EVENT_ABSORB; d.bsdf_over_pdf = v3(0.0f); d.pdf = 0.0f; return; }
  d.bsdf_over_pdf = p3(m, GTLC_DIFFUSE_TINT_R);
  d.event_type = BSDF_EVENT_DIFFUSE | BSDF_EVENT_REFLECTION;
}
static inline void diffuse_evaluate(const Material& m, BsdfEvaluateData& d, const State& s)
{
  float nk2 = dot(d.k2, s.normal);
  d.bsdf_glossy = v3(0.0f);
  if (nk2 <= 0.0f) { d.bsdf_diffuse = v3(0.0f); d.pdf = 0.0f; return; }
  d.pdf = nk2 * INV_PI;
  d.bsdf_diffuse = p3(m, GTLC_DIFFUSE_TINT_R) * d.pdf;
}

// ---- closure 1: UsdPreviewSurface ------------------------------------------------------------
// coat  = df::fresnel_layer(ior 1.5, weight clearcoat, layer GGX(clearcoatRoughness^2), base inner)
// inner = df::custom_curve_layer(F0, F90 = 1, exponent 5, weight 1, layer GGX(roughness^2), base diffuse(Kd))
struct UpsLobes { V3 F0, F90, Kd; float a, ac, coat; };
static inline UpsLobes ups_lobes(const Material& m)
{
  UpsLobes u;
  V3 dc = p3(m, GTLC_UPS_DIFFUSE_R);
  float metallic = clampf(m.p[GTLC_UPS_METALLIC], 0.0f, 1.0f);
  float ior = m.p[GTLC_UPS_IOR];
  float r0 = (1.0f - ior) / (1.0f + ior);
  float f0 = r0 * r0;
  if (m.p[GTLC_UPS_USE_SPECULAR_WORKFLOW] != 0.0f) { u.F0 = p3(m, GTLC_UPS_SPECULAR_R); u.Kd = dc; }
  else { u.F0 = mix(v3(f0), dc, metallic); u.Kd = dc * (1.0f - metallic); }
  u.F90 = v3(1.0f);
  float r = m.p[GTLC_UPS_ROUGHNESS], cr = m.p[GTLC_UPS_CLEARCOAT_ROUGHNESS];
  u.a = fmaxf(r * r, MIN_ALPHA);
  u.ac = fmaxf(cr * cr, MIN_ALPHA);
  u.coat = clampf(m.p[GTLC_UPS_CLEARCOAT], 0.0f, 1.0f);
  return u;
}

static inline void ups_sample(const Material& m, BsdfSampleData& d, const State& s)
{
  UpsLobes u = ups_lobes(m);
  Frame f = make_frame(s);
  V3 k1 = to_local(f, d.k1);
  d.bsdf_over_pdf = v3(0.0f); d.pdf = 0.0f; d.event_type = BSDF_EVENT_ABSORB; d.k2 = v3(0.0f);
  if (k1.z <= 0.0f) return;
  float nk1 = k1.z;
  float z = d.xi.z;
  V3 weight = v3(1.0f);
  V3 k2l, h; float g1;

  float p_coat = (u.coat > 0.0f) ? (u.coat * fresnel_dielectric(nk1, 1.5f)) : 0.0f;
  bool pending_coat_base = false;
  if (u.coat > 0.0f) {
    if (z < p_coat) {
      if (!ggx_reflect_sample(k1, u.ac, d.xi.x, d.xi.y, k2l, h, g1)) return;
      float fc = u.coat * fresnel_dielectric(dot(k1, h), 1.5f);
      d.bsdf_over_pdf = v3(g1 * (fc / p_coat));
      d.event_type = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
      d.k2 = normalize(to_world(f, k2l));
      if (dot(d.k2, s.geom_normal) <= 0.0f) { d.event_type = BSDF_EVENT_ABSORB; d.bsdf_over_pdf = v3(0.0f); return; }
      d.pdf = 1.0f;
      return;
    }
    z = (z - p_coat) / (1.0f - p_coat);
    pending_coat_base = true;
  }

  float p_spec = clampf(avg3(schlick(u.F0, u.F90, nk1)), 0.0f, 1.0f);
  int ev;
  if (z < p_spec) {
    if (!ggx_reflect_sample(k1, u.a, d.xi.x, d.xi.y, k2l, h, g1)) return;
    V3 F = schlick(u.F0, u.F90, dot(k1, h));
    weight = F * (g1 / p_spec);
    ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
  } else {
    k2l = cosine_hemisphere(d.xi.x, d.xi.y);
    if (k2l.z <= 0.0f) return;
    V3 c1 = schlick(u.F0, u.F90, nk1), c2 = schlick(u.F0, u.F90, k2l.z);
    V3 att = v3(1.0f - fmaxf(c1.x, c2.x), 1.0f - fmaxf(c1.y, c2.y), 1.0f - fmaxf(c1.z, c2.z));
    weight = (u.Kd * att) * (1.0f / (1.0f - p_spec));
    ev = BSDF_EVENT_DIFFUSE | BSDF_EVENT_REFLECTION;
  }
  if (pending_coat_base) {
    float fb = 1.0f - u.coat * fmaxf(fresnel_dielectric(nk1, 1.5f), fresnel_dielectric(k2l.z, 1.5f));
    weight = weight * (fb / (1.0f - p_coat));
  }
  d.k2 = normalize(to_world(f, k2l));
  if (dot(d.k2, s.geom_normal) <= 0.0f) return;
  d.bsdf_over_pdf = weight;
  d.event_type = ev;
  d.pdf = 1.0f;
}

static inline void ups_evaluate(const Material& m, BsdfEvaluateData& d, const State& s)
{
  UpsLobes u = ups_lobes(m);
  Frame f = make_frame(s);
  V3 k1 = to_local(f, d.k1), k2 = to_local(f, d.k2);
  d.bsdf_diffuse = v3(0.0f); d.bsdf_glossy = v3(0.0f); d.pdf = 0.0f;
  if (k1.z <= 0.0f || k2.z <= 0.0f) return;
  float nk1 = k1.z, nk2 = k2.z;
  V3 h; float fcos, pdfg;
  float p_coat = (u.coat > 0.0f) ? (u.coat * fresnel_dielectric(nk1, 1.5f)) : 0.0f;
  float fb = 1.0f;
  float pdf = 0.0f;
  if (u.coat > 0.0f) {
    ggx_reflect_eval(k1, k2, u.ac, fcos, pdfg, h);
    float fc = u.coat * fresnel_dielectric(dot(k1, h), 1.5f);
    d.bsdf_glossy += v3(fc * fcos);
    pdf += p_coat * pdfg;
    fb = 1.0f - u.coat * fmaxf(fresnel_dielectric(nk1, 1.5f), fresnel_dielectric(nk2, 1.5f));
  }
  float p_spec = clampf(avg3(schlick(u.F0, u.F90, nk1)), 0.0f, 1.0f);
  ggx_reflect_eval(k1, k2, u.a, fcos, pdfg, h);
  V3 F = schlick(u.F0, u.F90, dot(k1, h));
  d.bsdf_glossy += F * (fb * fcos);
  V3 c1 = schlick(u.F0, u.F90, nk1), c2 = schlick(u.F0, u.F90, nk2);
  V3 att = v3(1.0f - fmaxf(c1.x, c2.x), 1.0f - fmaxf(c1.y, c2.y), 1.0f - fmaxf(c1.z, c2.z));
  float pd = nk2 * INV_PI;
  d.bsdf_diffuse = (u.Kd * att) * (fb * pd);
  pdf += (1.0f - p_coat) * (p_spec * pdfg + (1.0f - p_spec) * pd);
  d.pdf = pdf;
}

static inline V3 ups_albedo(const Material& m, V3 k1, const State& s)
{
  UpsLobes u = ups_lobes(m);
  const float nk1 = clampf(dot(k1, s.normal), 0.0f, 1.0f);
  const V3 F = schlick(u.F0, u.F90, nk1);
  const V3 inner = F + (v3(1.0f) - F) * u.Kd;
  if (!(u.coat > 0.0f)) return inner;
  const float fc = u.coat * fresnel_dielectric(nk1, 1.5f);
  return v3(fc) + inner * (1.0f - fc);
}

// ---- closure 2: OpenPBR (subset, see orc_openpbr.h) -----------------------------------------
void openpbr_sample(const Material& m, BsdfSampleData& d, const State& s);
void openpbr_evaluate(const Material& m, BsdfEvaluateData& d, const State& s);
V3 openpbr_albedo(const Material& m, V3 k1, const State& s);

// ---- dispatch = the generated mdl_bsdf_scattering_* entry points ---------------------------
static inline void bsdf_sample(const Material& m, BsdfSampleData& d, const State& s)
{
  switch (m.closure) {
    case GTLC_CLOSURE_DIFFUSE: diffuse_sample(m, d, s); break;
    case GTLC_CLOSURE_USD_PREVIEW_SURFACE: ups_sample(m, d, s); break;
    default: openpbr_sample(m, d, s); break;
  }
}
static inline void bsdf_evaluate(const Material& m, BsdfEvaluateData& d, const State& s)
{
  switch (m.closure) {
    case GTLC_CLOSURE_DIFFUSE: diffuse_evaluate(m, d, s); break;
    case GTLC_CLOSURE_USD_PREVIEW_SURFACE: ups_evaluate(m, d, s); break;
    default: openpbr_evaluate(m, d, s); break;
  }
}

// mdl_bsdf_scattering_auxiliary (rp_main.chit:264-280): albedo_diffuse + albedo_glossy
static inline V3 bsdf_albedo(const Material& m, V3 k1, const State& s)
{
  switch (m.closure) {
    case GTLC_CLOSURE_DIFFUSE: return p3(m, GTLC_DIFFUSE_TINT_R);
    case GTLC_CLOSURE_USD_PREVIEW_SURFACE: return ups_albedo(m, k1, s);
    default: return openpbr_albedo(m, k1, s);
  }
}

} // namespace orc
