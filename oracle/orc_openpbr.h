// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).
//
// OpenPBR closure: restatement of the node graph /root/reference/src/gi/mtlx/open_pbr_surface.mtlx:99-635
// (the only in-tree specification of the layering; the MDL the reference generates from it is compiled
// by the closed MDL SDK).  PARITY: the hit-independent values (opbr_lobes, the volume coefficients, the emission weight) are pinned to that
// node graph -- tests/golden/make_openpbr_graph_vectors.py evaluates the reference's file node by node, tests/test_openpbr_graph_cpu.py
// compares (a few ulp: pow() vs products); the lobes themselves (what each BSDF node does with its inputs) are UNPINNED, they live in the
// MDL SDK.  Implemented subset, top to bottom of the layer stack:
//   coat            dielectric_bsdf R (coat_weight, coat_ior, coat_roughness^2)          mtlx:542-568
//   coat tint       base * mix(1, coat_color, coat_weight)                                mtlx:542-550
//   coat darkening  base * mix(1, (1-Kcoat)/(1-Ebase*Kcoat), coat_weight*coat_darkening)  mtlx:472-540
//   metal           generalized_schlick(color0 = base_color*base_weight, color82 = specular_color*specular_weight,
//                   weight specular_weight), mixed by base_metalness                      mtlx:434-470
//   dielectric      dielectric_bsdf R (tint specular_color, ior modulated_eta_s) layered over the substrate  mtlx:306-432
//   roughness       coat roughening of the specular lobe, alpha = roughness^2             mtlx:102-137
//   substrate       mix(opaque base, dielectric_bsdf T (same ior/roughness), transmission_weight)           mtlx:374-396
//   opaque base     mix(diffuse(base_weight*base_color), diffuse(subsurface_color), subsurface_weight)      mtlx:195-218
//   volume          Beer absorption from transmission_color/depth (orc_bsdf.h mat_volume_absorption)        mtlx:221-262
//   emission        uniform_edf(emission_color*emission_luminance), tinted by the coat colour              mtlx:582-623
// Not modelled (treated as weight 0 / default): fuzz (Zeltner sheen, mtlx:570-580), thin film (mtlx:301-304),
// anisotropy, Oren-Nayar roughness of the diffuse base, subsurface random walk / thin-walled translucency,
// transmission scatter & dispersion, the Fresnel factor of the coated EDF, geometry_opacity.
#pragma once
#include "orc_bsdf.h"

namespace orc {

struct OpbrLobes {
  float coat_w, a_c, coat_ior; V3 coat_att;
  float metal; V3 m_f0, m_f82; float spec_w;
  float eta; V3 spec_color; float a;
  float trans_w; V3 trans_tint; V3 Kd;
};

static inline OpbrLobes opbr_lobes(const Material& m)
{
  const float* p = m.p;
  OpbrLobes o;
  const float coat_w = clampf(p[GTLC_OPBR_COAT_WEIGHT], 0.0f, 1.0f);
  const float coat_r = p[GTLC_OPBR_COAT_ROUGHNESS], spec_r = p[GTLC_OPBR_SPECULAR_ROUGHNESS];
  const float coat_ior = p[GTLC_OPBR_COAT_IOR], spec_ior = p[GTLC_OPBR_SPECULAR_IOR], spec_w = p[GTLC_OPBR_SPECULAR_WEIGHT];
  // mtlx:102-131 roughening due to coat
  const float c4 = (coat_r * coat_r) * (coat_r * coat_r), s4 = (spec_r * spec_r) * (spec_r * spec_r);
  const float affected = sqrtf(sqrtf(fminf(1.0f, 2.0f * c4 + s4)));
  const float eff_r = mixf(spec_r, affected, coat_w);
  o.a = fmaxf(eff_r * eff_r, MIN_ALPHA);
  o.a_c = fmaxf(coat_r * coat_r, MIN_ALPHA);
  o.coat_w = coat_w; o.coat_ior = coat_ior;
  // mtlx:306-373 specular F0 modulation (IOR ratio inversion under the coat, specular_weight folded into the IOR)
  const float ratio = spec_ior / coat_ior;
  const float tirfix = (ratio > 1.0f) ? ratio : (coat_ior / spec_ior);
  const float eta_s = mixf(spec_ior, tirfix, coat_w);
  const float f0s = (eta_s - 1.0f) / (eta_s + 1.0f);
  const float scaled = clampf(spec_w * (f0s * f0s), 0.0f, 0.99999f);
  const float eps = ((eta_s - 1.0f) > 0.0f ? 1.0f : ((eta_s - 1.0f) < 0.0f ? -1.0f : 0.0f)) * sqrtf(scaled);
  o.eta = (1.0f + eps) / (1.0f - eps);
  o.spec_color = p3(m, GTLC_OPBR_SPECULAR_COLOR_R);
  o.spec_w = spec_w;
  // mtlx:434-470 metal
  o.metal = clampf(p[GTLC_OPBR_BASE_METALNESS], 0.0f, 1.0f);
  const V3 base_color = vmax(p3(m, GTLC_OPBR_BASE_COLOR_R), 0.0f);
  o.m_f0 = base_color * p[GTLC_OPBR_BASE_WEIGHT];
  o.m_f82 = o.spec_color * spec_w;
  // mtlx:195-218 opaque base, mtlx:374-396 transmission
  const float ssw = clampf(p[GTLC_OPBR_SUBSURFACE_WEIGHT], 0.0f, 1.0f);
  const V3 sss = vmax(p3(m, GTLC_OPBR_SUBSURFACE_COLOR_R), 0.0f);
  o.Kd = mix(base_color * p[GTLC_OPBR_BASE_WEIGHT], sss, ssw);
  o.trans_w = clampf(p[GTLC_OPBR_TRANSMISSION_WEIGHT], 0.0f, 1.0f);
  o.trans_tint = (p[GTLC_OPBR_TRANSMISSION_DEPTH] > 0.0f) ? v3(1.0f) : p3(m, GTLC_OPBR_TRANSMISSION_COLOR_R);
  // mtlx:472-550 coat darkening and tint
  const float f0c_s = (coat_ior - 1.0f) / (coat_ior + 1.0f);
  const float Kcoat = 1.0f - (1.0f - f0c_s * f0c_s) / (coat_ior * coat_ior);
  const V3 Edielectric = mix(base_color, p3(m, GTLC_OPBR_SUBSURFACE_COLOR_R), ssw);
  const V3 Ebase = mix(Edielectric, base_color * spec_w, o.metal);
  const V3 darkening = v3((1.0f - Kcoat) / (1.0f - Ebase.x * Kcoat), (1.0f - Kcoat) / (1.0f - Ebase.y * Kcoat), (1.0f - Kcoat) / (1.0f - Ebase.z * Kcoat));
  const V3 modulated = mix(v3(1.0f), darkening, coat_w * p[GTLC_OPBR_COAT_DARKENING]);
  o.coat_att = modulated * mix(v3(1.0f), p3(m, GTLC_OPBR_COAT_COLOR_R), coat_w);
  return o;
}

// generalized_schlick with the OpenPBR "F82-tint" edge colour (OpenPBR spec 1.1, metal base)
static inline V3 opbr_metal_fresnel(const OpbrLobes& o, float c)
{
  const float mu = 1.0f / 7.0f;
  const float k = mu * ((1.0f - mu) * (1.0f - mu)) * ((1.0f - mu) * (1.0f - mu)) * ((1.0f - mu) * (1.0f - mu));   // mu (1-mu)^6
  const V3 fbar = schlick(o.m_f0, v3(1.0f), mu);
  const V3 b = (fbar * (v3(1.0f) - o.m_f82)) * (1.0f / k);
  c = clampf(c, 0.0f, 1.0f);
  const float om = 1.0f - c;
  const float w6 = c * ((om * om) * (om * om) * (om * om));
  const V3 f = schlick(o.m_f0, v3(1.0f), c) - b * w6;
  return v3(clampf(f.x, 0.0f, 1.0f), clampf(f.y, 0.0f, 1.0f), clampf(f.z, 0.0f, 1.0f)) * o.spec_w;
}

static inline float opbr_rel_eta(const V3& ior1, const V3& ior2, float material_ior)
{
  const float n1 = (ior1.x == -1.0f) ? material_ior : ior1.x;   // BSDF_USE_MATERIAL_IOR, mdl_types.glsl:155
  const float n2 = (ior2.x == -1.0f) ? material_ior : ior2.x;
  return n2 / n1;
}

inline void openpbr_sample(const Material& m, BsdfSampleData& d, const State& s)
{
  const OpbrLobes o = opbr_lobes(m);
  const Frame f = make_frame(s);
  const V3 k1 = to_local(f, d.k1);
  d.bsdf_over_pdf = v3(0.0f); d.pdf = 0.0f; d.event_type = BSDF_EVENT_ABSORB; d.k2 = v3(0.0f);
  if (k1.z <= 0.0f) return;
  const float nk1 = k1.z;
  const float eta = opbr_rel_eta(d.ior1, d.ior2, o.eta);
  float z = d.xi.z;
  V3 k2l, h; float g1;
  V3 weight; int ev;

  // ---- coat ----
  const float fc1 = (o.coat_w > 0.0f) ? fresnel_dielectric(nk1, o.coat_ior) : 0.0f;
  const float p_coat = o.coat_w * fc1;
  if (o.coat_w > 0.0f && z < p_coat) {
    if (!ggx_reflect_sample(k1, o.a_c, d.xi.x, d.xi.y, k2l, h, g1)) return;
    const float fc = o.coat_w * fresnel_dielectric(dot(k1, h), o.coat_ior);
    weight = v3(g1 * (fc / p_coat));
    ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
  } else {
    if (o.coat_w > 0.0f) z = (z - p_coat) / (1.0f - p_coat);
    if (z < o.metal) {
      // ---- metal ----
      if (!ggx_reflect_sample(k1, o.a, d.xi.x, d.xi.y, k2l, h, g1)) return;
      weight = opbr_metal_fresnel(o, dot(k1, h)) * g1;
      ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
    } else {
      if (o.metal > 0.0f) z = (z - o.metal) / (1.0f - o.metal);
      // ---- dielectric base ----
      const float fs1 = fresnel_dielectric(nk1, eta);
      if (z < fs1) {
        if (!ggx_reflect_sample(k1, o.a, d.xi.x, d.xi.y, k2l, h, g1)) return;
        weight = o.spec_color * (g1 * (fresnel_dielectric(dot(k1, h), eta) / fs1));
        ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
      } else {
        z = (z - fs1) / (1.0f - fs1);
        if (z < o.trans_w) {
          // rough transmission through the microfacet normal (dielectric_bsdf, scatter_mode T)
          h = ggx_sample_vndf(k1, o.a, d.xi.x, d.xi.y);
          const float kh = dot(k1, h);
          if (kh <= 0.0f) return;
          const float sin2_t = (1.0f - kh * kh) / (eta * eta);
          if (sin2_t >= 1.0f) {                                   // total internal reflection
            k2l = h * (2.0f * kh) - k1;
            if (k2l.z <= 0.0f) return;
            weight = v3(ggx_G1(k2l, o.a));
            ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
          } else {
            const float cos_t = sqrtf(1.0f - sin2_t);
            k2l = h * (kh / eta - cos_t) - k1 * (1.0f / eta);
            if (k2l.z >= 0.0f) return;
            const float fs2 = fresnel_dielectric(fabsf(k2l.z), 1.0f / eta);
            weight = o.trans_tint * (ggx_G1(v3(k2l.x, k2l.y, -k2l.z), o.a) * ((1.0f - fmaxf(fs1, fs2)) / (1.0f - fs1)));
            ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_TRANSMISSION;
          }
        } else {
          k2l = cosine_hemisphere(d.xi.x, d.xi.y);
          if (k2l.z <= 0.0f) return;
          const float fs2 = fresnel_dielectric(k2l.z, eta);
          weight = o.Kd * ((1.0f - fmaxf(fs1, fs2)) / (1.0f - fs1));
          ev = BSDF_EVENT_DIFFUSE | BSDF_EVENT_REFLECTION;
        }
      }
    }
    if (o.coat_w > 0.0f) {
      const float fc2 = fresnel_dielectric(fabsf(k2l.z), o.coat_ior);
      weight = (weight * o.coat_att) * ((1.0f - o.coat_w * fmaxf(fc1, fc2)) / (1.0f - p_coat));
    }
  }
  d.k2 = normalize(to_world(f, k2l));
  const float side = dot(d.k2, s.geom_normal);
  if ((ev & BSDF_EVENT_TRANSMISSION) ? (side >= 0.0f) : (side <= 0.0f)) return;
  d.bsdf_over_pdf = weight;
  d.event_type = ev;
  d.pdf = 1.0f;
}

inline void openpbr_evaluate(const Material& m, BsdfEvaluateData& d, const State& s)
{
  const OpbrLobes o = opbr_lobes(m);
  const Frame f = make_frame(s);
  const V3 k1 = to_local(f, d.k1), k2 = to_local(f, d.k2);
  d.bsdf_diffuse = v3(0.0f); d.bsdf_glossy = v3(0.0f); d.pdf = 0.0f;
  if (k1.z <= 0.0f || k2.z <= 0.0f) return;          // NEE only asks for the reflection side (rp_main.chit:415)
  const float nk1 = k1.z, nk2 = k2.z;
  const float eta = opbr_rel_eta(d.ior1, d.ior2, o.eta);
  V3 h; float fcos, pdfg;
  float pdf = 0.0f;
  V3 base_scale = v3(1.0f);
  float p_coat = 0.0f;
  if (o.coat_w > 0.0f) {
    const float fc1 = fresnel_dielectric(nk1, o.coat_ior), fc2 = fresnel_dielectric(nk2, o.coat_ior);
    p_coat = o.coat_w * fc1;
    ggx_reflect_eval(k1, k2, o.a_c, fcos, pdfg, h);
    d.bsdf_glossy += v3((o.coat_w * fresnel_dielectric(dot(k1, h), o.coat_ior)) * fcos);
    pdf += p_coat * pdfg;
    base_scale = o.coat_att * (1.0f - o.coat_w * fmaxf(fc1, fc2));
  }
  ggx_reflect_eval(k1, k2, o.a, fcos, pdfg, h);
  const float kh = dot(k1, h);
  const float fs1 = fresnel_dielectric(nk1, eta), fs2 = fresnel_dielectric(nk2, eta);
  const V3 metal = opbr_metal_fresnel(o, kh) * fcos;
  const V3 spec = o.spec_color * (fresnel_dielectric(kh, eta) * fcos);
  const float pd = nk2 * INV_PI;
  const V3 diff = o.Kd * ((1.0f - fmaxf(fs1, fs2)) * (1.0f - o.trans_w) * pd);
  d.bsdf_glossy += base_scale * (metal * o.metal + spec * (1.0f - o.metal));
  d.bsdf_diffuse = base_scale * (diff * (1.0f - o.metal));
  pdf += (1.0f - p_coat) * (o.metal * pdfg + (1.0f - o.metal) * (fs1 * pdfg + (1.0f - fs1) * (1.0f - o.trans_w) * pd));
  d.pdf = pdf;
}

// mdl_bsdf_scattering_auxiliary: albedo_diffuse + albedo_glossy of the layer stack at the viewing angle
// (directional albedo of every specular lobe taken as its Fresnel weight at n.k1).
inline V3 openpbr_albedo(const Material& m, V3 k1, const State& s)
{
  const OpbrLobes o = opbr_lobes(m);
  const float nk1 = clampf(dot(k1, s.normal), 0.0f, 1.0f);
  const float fs = fresnel_dielectric(nk1, o.eta);
  const V3 substrate = o.Kd * (1.0f - o.trans_w) + o.trans_tint * o.trans_w;
  const V3 dielectric = o.spec_color * fs + substrate * (1.0f - fs);
  const V3 base = opbr_metal_fresnel(o, nk1) * o.metal + dielectric * (1.0f - o.metal);
  if (!(o.coat_w > 0.0f)) return base;
  const float fc = o.coat_w * fresnel_dielectric(nk1, o.coat_ior);
  return v3(fc) + (base * o.coat_att) * (1.0f - fc);
}

} // namespace orc
