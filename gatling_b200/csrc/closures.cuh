// Ahead-of-time CUDA closures standing in for the per-material GLSL the reference generates at run
// time through the MDL SDK (SURVEY.md 8b "inner boundary"; entry-point names fixed at
// /root/reference/src/gi/impl/GlslShaderGen.cpp:181-193, data structs
// /root/reference/src/gi/shaders/mdl_types.glsl:32-225).  A material is a closure id plus a
// parameter block (class-compiled), selected per hit by the shade kernel.
//
// Definitions follow the published specs (MDL 1.8 section 21.1 df::*; Heitz 2018 VNDF sampling;
// MaterialX UsdPreviewSurface graph; in-tree /root/reference/src/gi/mtlx/open_pbr_surface.mtlx) --
// the SDK's generated code itself is not available (see ORACLE_DEVIATIONS.md).
#pragma once
#include "device_math.cuh"
#include "../../include/gtlc.h"

namespace gtlc {

// mdl_types.glsl:123-137
enum : int {
  BSDF_EVENT_ABSORB = 0, BSDF_EVENT_DIFFUSE = 1, BSDF_EVENT_GLOSSY = 2, BSDF_EVENT_SPECULAR = 4,
  BSDF_EVENT_REFLECTION = 8, BSDF_EVENT_TRANSMISSION = 16
};

struct ShadingState { V3 normal, geom_normal, position, tangent_u, tangent_v; float u, v; uint32_t hitIndices[3]; V3 local_position; };  // mdl_types.glsl:32-113 (+ renderer_state.hitIndices)
struct BsdfSample { V3 ior1, ior2, k1, k2; V4 xi; float pdf; V3 bsdf_over_pdf; int event_type; float memo[2]; };  // :158-170; memo: see BsdfEval
// memo: view-dependent Fresnel terms (coat, base) that *_sample() already evaluated for the same state, k1 and IORs; a negative
// entry means "not available" (the evaluation then computes it -- the same expression, hence the same bits)
struct BsdfEval { V3 ior1, ior2, k1, k2; V3 bsdf_diffuse, bsdf_glossy; float pdf; float memo[2]; };              // :173-184

constexpr float kMinAlpha = 1.0e-4f;

struct Frame { V3 t, b, n; };
GTLC_D Frame make_frame(const ShadingState& s) { Frame f; f.n = s.normal; f.t = s.tangent_u; f.b = cross(s.normal, s.tangent_u); return f; }
GTLC_D V3 to_local(const Frame& f, V3 v) { return v3(dot(v, f.t), dot(v, f.b), dot(v, f.n)); }
GTLC_D V3 to_world(const Frame& f, V3 v) { return (f.t * v.x + f.b * v.y) + f.n * v.z; }

// Shirley-Chiu concentric map, z-up cosine-weighted hemisphere
GTLC_D V3 cosine_hemisphere(float u, float v)
{
  float a = 2.0f * u - 1.0f;
  float b = 2.0f * v - 1.0f;
  if (a == 0.0f && b == 0.0f) return v3(0.0f, 0.0f, 1.0f);
  float r, phi;
  if ((a * a) > (b * b)) { r = a; phi = (kPi / 4.0f) * (b / a); }
  else { r = b; phi = (kPi / 2.0f) - (kPi / 4.0f) * (a / b); }
  const SinCos sc = sincos_of(phi);
  float x = r * sc.c, y = r * sc.s;
  float z = sqrtf(fmaxf(0.0f, 1.0f - r * r));
  return v3(x, y, z);
}

GTLC_D float ggx_D(V3 h, float a)
{
  float a2 = a * a;
  float d = ((h.x * h.x + h.y * h.y) / a2) + h.z * h.z;
  return 1.0f / ((kPi * a2) * (d * d));
}
GTLC_D float ggx_lambda(V3 v, float a)
{
  float z2 = v.z * v.z;
  if (z2 <= 0.0f) return kFloatMax;
  float t2 = (a * a) * ((v.x * v.x + v.y * v.y) / z2);
  return (sqrtf(1.0f + t2) - 1.0f) * 0.5f;
}
GTLC_D float ggx_G1(V3 v, float a) { return 1.0f / (1.0f + ggx_lambda(v, a)); }

// Heitz 2018, listing 1 (isotropic alpha)
GTLC_D V3 ggx_sample_vndf(V3 Ve, float a, float u1, float u2)
{
  V3 Vh = normalize(v3(a * Ve.x, a * Ve.y, Ve.z));
  float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
  V3 T1 = (lensq > 0.0f) ? (v3(-Vh.y, Vh.x, 0.0f) * (1.0f / sqrtf(lensq))) : v3(1.0f, 0.0f, 0.0f);
  V3 T2 = cross(Vh, T1);
  float r = sqrtf(u1);
  float phi = (2.0f * kPi) * u2;
  const SinCos sc = sincos_of(phi);
  float t1 = r * sc.c;
  float t2 = r * sc.s;
  float s = 0.5f * (1.0f + Vh.z);
  t2 = (1.0f - s) * sqrtf(fmaxf(0.0f, 1.0f - t1 * t1)) + s * t2;
  V3 Nh = (T1 * t1 + T2 * t2) + Vh * sqrtf(fmaxf(0.0f, (1.0f - t1 * t1) - t2 * t2));
  return normalize(v3(a * Nh.x, a * Nh.y, fmaxf(0.0f, Nh.z)));
}

GTLC_D float schlick_w(float c) { float m = clampf(1.0f - c, 0.0f, 1.0f); float m2 = m * m; return (m2 * m2) * m; }
GTLC_D V3 schlick(V3 f0, V3 f90, float c) { float w = schlick_w(c); return f0 * (1.0f - w) + f90 * w; }
GTLC_D float avg3(V3 c) { return ((c.x + c.y) + c.z) * (1.0f / 3.0f); }

GTLC_D float fresnel_dielectric(float cos_i, float eta)
{
  cos_i = clampf(cos_i, 0.0f, 1.0f);
  float sin2_t = (1.0f - cos_i * cos_i) / (eta * eta);
  if (sin2_t >= 1.0f) return 1.0f;
  float cos_t = sqrtf(1.0f - sin2_t);
  float rs = (cos_i - eta * cos_t) / (cos_i + eta * cos_t);
  float rp = (eta * cos_i - cos_t) / (eta * cos_i + cos_t);
  return 0.5f * (rs * rs + rp * rp);
}

GTLC_D bool ggx_reflect_sample(V3 k1l, float a, float u1, float u2, V3& k2l, V3& h, float& g1_k2)
{
  h = ggx_sample_vndf(k1l, a, u1, u2);
  float kh = dot(k1l, h);
  k2l = h * (2.0f * kh) - k1l;
  if (k2l.z <= 0.0f || kh <= 0.0f) return false;
  g1_k2 = ggx_G1(k2l, a);
  return true;
}
GTLC_D void ggx_reflect_eval(V3 k1l, V3 k2l, float a, float& f_cos, float& pdf, V3& h)
{
  h = normalize(k1l + k2l);
  float D = ggx_D(h, a);
  float g1a = ggx_G1(k1l, a), g1b = ggx_G1(k2l, a);
  float denom = 4.0f * k1l.z;
  f_cos = ((D * g1a) * g1b) / denom;
  pdf = (D * g1a) / denom;
}

GTLC_D V3 p3(const float* p, int i) { return v3(p[i], p[i + 1], p[i + 2]); }

// mc/gtl/mc/Material.h:28-45 flags, evaluated per material at scene build (host) -- see gtlc_api.cu
GTLC_HD bool mat_is_emissive(int closure, const float* p)
{
  if (closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) return p[GTLC_UPS_EMISSIVE_R] > 0.0f || p[GTLC_UPS_EMISSIVE_G] > 0.0f || p[GTLC_UPS_EMISSIVE_B] > 0.0f;
  if (closure == GTLC_CLOSURE_OPENPBR)
    return p[GTLC_OPBR_EMISSION_LUMINANCE] > 0.0f && (p[GTLC_OPBR_EMISSION_COLOR_R] > 0.0f || p[GTLC_OPBR_EMISSION_COLOR_G] > 0.0f || p[GTLC_OPBR_EMISSION_COLOR_B] > 0.0f);
  return false;
}
// mdl_cutout_opacity (rp_main.ahit:42): UsdPreviewSurface `opacity`, OpenPBR `geometry_opacity`
GTLC_HD float mat_cutout_opacity(int closure, const float* p)
{
  float o = 1.0f;
  if (closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) o = p[GTLC_UPS_OPACITY];
  else if (closure == GTLC_CLOSURE_OPENPBR) o = p[GTLC_OPBR_GEOMETRY_OPACITY];
  return o < 0.0f ? 0.0f : (o > 1.0f ? 1.0f : o);
}
GTLC_HD bool mat_is_thin_walled(int closure, const float* p) { return closure == GTLC_CLOSURE_OPENPBR && p[GTLC_OPBR_GEOMETRY_THIN_WALLED] != 0.0f; }

// mdl_edf_emission_intensity (MaterialX uniform_edf -> df::diffuse_edf, intensity = color * PI)
GTLC_D V3 mat_emission_intensity(int closure, const float* p)
{
  if (closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) return p3(p, GTLC_UPS_EMISSIVE_R) * kPi;
  if (closure == GTLC_CLOSURE_OPENPBR)   // open_pbr_surface.mtlx:582-623: uniform_edf tinted by the coat colour
    return ((p3(p, GTLC_OPBR_EMISSION_COLOR_R) * p[GTLC_OPBR_EMISSION_LUMINANCE]) * kPi) * mix(v3(1.0f), p3(p, GTLC_OPBR_COAT_COLOR_R), clampf(p[GTLC_OPBR_COAT_WEIGHT], 0.0f, 1.0f));
  return v3(0.0f);
}
// mdl_volume_scattering_coefficient (open_pbr_surface.mtlx:238-245,286-292: transmission_scatter / transmission_depth)
GTLC_D V3 mat_volume_scattering(int closure, const float* p)
{
  if (closure != GTLC_CLOSURE_OPENPBR) return v3(0.0f);
  float depth = p[GTLC_OPBR_TRANSMISSION_DEPTH];
  if (!(depth > 0.0f)) return v3(0.0f);
  V3 sc = p3(p, GTLC_OPBR_TRANSMISSION_SCATTER_R);
  return v3(sc.x / depth, sc.y / depth, sc.z / depth);
}
// mdl_volume_absorption_coefficient (open_pbr_surface.mtlx:221-285): extinction - scattering, shifted to be non-negative
GTLC_D V3 mat_volume_absorption(int closure, const float* p)
{
  if (closure != GTLC_CLOSURE_OPENPBR) return v3(0.0f);
  float depth = p[GTLC_OPBR_TRANSMISSION_DEPTH];
  if (!(depth > 0.0f)) return v3(0.0f);
  V3 c = p3(p, GTLC_OPBR_TRANSMISSION_COLOR_R);
  V3 ext = v3(-logf(fmaxf(c.x, 1.0e-6f)) / depth, -logf(fmaxf(c.y, 1.0e-6f)) / depth, -logf(fmaxf(c.z, 1.0e-6f)) / depth);
  V3 a = ext - mat_volume_scattering(closure, p);
  float mn = fminf(fminf(a.x, a.y), a.z);
  if (0.0f > mn) a = a - v3(mn);
  return a;
}
// mdl_ior: material.ior stays at the MDL default for MaterialX surfaces (open_pbr_surface.mtlx:626-630)
GTLC_D V3 mat_ior(int, const float*) { return v3(1.0f); }
// MEDIUM_DIRECTIONAL_BIAS (GlslShaderGen.cpp:212, Frontend.cpp:137-157)
GTLC_D float mat_directional_bias(int closure, const float* p) { return closure == GTLC_CLOSURE_OPENPBR ? p[GTLC_OPBR_TRANSMISSION_SCATTER_ANISOTROPY] : 0.0f; }

// ---- closure 0: df::diffuse_reflection_bsdf(tint) --------------------------------------------
GTLC_D void diffuse_sample(const float* p, BsdfSample& d, const ShadingState& s)
{
  d.memo[0] = -1.0f; d.memo[1] = -1.0f;
  Frame f = make_frame(s);
  V3 l = cosine_hemisphere(d.xi.x, d.xi.y);
  d.k2 = normalize(to_world(f, l));
  d.pdf = l.z * kInvPi;
  if (l.z <= 0.0f || dot(d.k2, s.geom_normal) <= 0.0f) { d.event_type = BSDF_EVENT_ABSORB; d.bsdf_over_pdf = v3(0.0f); d.pdf = 0.0f; return; }
  d.bsdf_over_pdf = p3(p, GTLC_DIFFUSE_TINT_R);
  d.event_type = BSDF_EVENT_DIFFUSE | BSDF_EVENT_REFLECTION;
}
GTLC_D void diffuse_evaluate(const float* p, BsdfEval& d, const ShadingState& s)
{
  float nk2 = dot(d.k2, s.normal);
  d.bsdf_glossy = v3(0.0f);
  if (nk2 <= 0.0f) { d.bsdf_diffuse = v3(0.0f); d.pdf = 0.0f; return; }
  d.pdf = nk2 * kInvPi;
  d.bsdf_diffuse = p3(p, GTLC_DIFFUSE_TINT_R) * d.pdf;
}

// ---- closure 1: UsdPreviewSurface -------------------------------------------------------------
struct UpsLobes { V3 F0, F90, Kd; float a, ac, coat; };
GTLC_D UpsLobes ups_lobes(const float* p)
{
  UpsLobes u;
  V3 dc = p3(p, GTLC_UPS_DIFFUSE_R);
  float metallic = clampf(p[GTLC_UPS_METALLIC], 0.0f, 1.0f);
  float ior = p[GTLC_UPS_IOR];
  float r0 = (1.0f - ior) / (1.0f + ior);
  float f0 = r0 * r0;
  if (p[GTLC_UPS_USE_SPECULAR_WORKFLOW] != 0.0f) { u.F0 = p3(p, GTLC_UPS_SPECULAR_R); u.Kd = dc; }
  else { u.F0 = mix(v3(f0), dc, metallic); u.Kd = dc * (1.0f - metallic); }
  u.F90 = v3(1.0f);
  float r = p[GTLC_UPS_ROUGHNESS], cr = p[GTLC_UPS_CLEARCOAT_ROUGHNESS];
  u.a = fmaxf(r * r, kMinAlpha);
  u.ac = fmaxf(cr * cr, kMinAlpha);
  u.coat = clampf(p[GTLC_UPS_CLEARCOAT], 0.0f, 1.0f);
  return u;
}

GTLC_D void ups_sample(const float* p, BsdfSample& d, const ShadingState& s)
{
  d.memo[0] = -1.0f; d.memo[1] = -1.0f;
  UpsLobes u = ups_lobes(p);
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

GTLC_D void ups_evaluate(const float* p, BsdfEval& d, const ShadingState& s)
{
  UpsLobes u = ups_lobes(p);
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
  float pd = nk2 * kInvPi;
  d.bsdf_diffuse = (u.Kd * att) * (fb * pd);
  pdf += (1.0f - p_coat) * (p_spec * pdfg + (1.0f - p_spec) * pd);
  d.pdf = pdf;
}

} // namespace gtlc

#include "closure_openpbr.cuh"

namespace gtlc {

// The generated mdl_bsdf_scattering_sample / _evaluate entry points, selected by closure id.  MAXC = highest
// closure id present in the scene: kernels are instantiated per closure set so that a diffuse-only scene does
// not pay the registers of the layered closures (the counterpart of the reference's per-material shader
// permutations, GlslShaderGen.cpp:211-262).
// `lowered`: the material's row of the lowered-material table (closure_openpbr.cuh: OpbrLobes), or NULL when the
// parameter block was patched per hit (textured / primvar-driven inputs).
template <int MAXC>
GTLC_D void bsdf_sample(int closure, const float* p, const float* lowered, BsdfSample& d, const ShadingState& s)
{
  if (MAXC == GTLC_CLOSURE_DIFFUSE || closure == GTLC_CLOSURE_DIFFUSE) { diffuse_sample(p, d, s); return; }
  if (MAXC == GTLC_CLOSURE_USD_PREVIEW_SURFACE || closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) { ups_sample(p, d, s); return; }
  openpbr_sample(p, lowered, d, s);
}
GTLC_D V3 ups_albedo(const float* p, V3 k1, const ShadingState& s)
{
  UpsLobes u = ups_lobes(p);
  const float nk1 = clampf(dot(k1, s.normal), 0.0f, 1.0f);
  const V3 F = schlick(u.F0, u.F90, nk1);
  const V3 inner = F + (v3(1.0f) - F) * u.Kd;
  if (!(u.coat > 0.0f)) return inner;
  const float fc = u.coat * fresnel_dielectric(nk1, 1.5f);
  return v3(fc) + inner * (1.0f - fc);
}
// mdl_bsdf_scattering_auxiliary (rp_main.chit:264-280): albedo_diffuse + albedo_glossy
template <int MAXC>
GTLC_D V3 bsdf_albedo(int closure, const float* p, const float* lowered, V3 k1, const ShadingState& s)
{
  if (MAXC == GTLC_CLOSURE_DIFFUSE || closure == GTLC_CLOSURE_DIFFUSE) return p3(p, GTLC_DIFFUSE_TINT_R);
  if (MAXC == GTLC_CLOSURE_USD_PREVIEW_SURFACE || closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) return ups_albedo(p, k1, s);
  return openpbr_albedo(p, lowered, k1, s);
}
template <int MAXC>
GTLC_D void bsdf_evaluate(int closure, const float* p, const float* lowered, BsdfEval& d, const ShadingState& s)
{
  if (MAXC == GTLC_CLOSURE_DIFFUSE || closure == GTLC_CLOSURE_DIFFUSE) { diffuse_evaluate(p, d, s); return; }
  if (MAXC == GTLC_CLOSURE_USD_PREVIEW_SURFACE || closure == GTLC_CLOSURE_USD_PREVIEW_SURFACE) { ups_evaluate(p, d, s); return; }
  openpbr_evaluate(p, lowered, d, s);
}

// One thread per material: evaluates the hit-independent part of the OpenPBR closure once (see OpbrLobes).
__global__ void lower_materials_kernel(const float* __restrict__ params, const int32_t* __restrict__ closure, uint32_t count, float* __restrict__ lowered)
{
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= count) return;
  float* dst = lowered + (size_t) m * kLobeFloats;
  if (closure[m] == GTLC_CLOSURE_OPENPBR) opbr_lobes_store(opbr_lobes(params + (size_t) m * GTLC_MATERIAL_PARAM_COUNT), dst);
  else for (int k = 0; k < kLobeFloats; k++) dst[k] = 0.0f;
}

} // namespace gtlc
