// OpenPBR closure (CUDA).  Subset of the in-tree node graph
// /root/reference/src/gi/mtlx/open_pbr_surface.mtlx:99-635 -- coat (+tint, darkening, roughening of the base),
// metal (generalized Schlick with F82 tint), dielectric specular with the coat-modulated IOR, diffuse / rough
// transmission substrate, Beer absorption, coat-tinted emission.  Not modelled: fuzz, thin film, anisotropy,
// Oren-Nayar roughness, subsurface walk, dispersion, opacity (see ORACLE_DEVIATIONS.md #5).
#pragma once
namespace gtlc {

GTLC_D float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

// What the closure needs of a material: constants of the parameter block, NOT of the hit.  For materials without
// textured / primvar-driven inputs lower_materials_kernel evaluates opbr_lobes() once per material at upload (the same
// device code, hence the same bits) and the shade kernel loads the 32-float record instead of spending ~10 IEEE divisions
// and 3 square roots per call -- the instance-compiled MDL of the reference folds these at material compile time too
// (MdlMaterialCompiler.cpp:262).  Textured materials (TEX instantiation) keep evaluating it per hit.
struct OpbrLobes {
  float coat_w, a_c, coat_ior; V3 coat_att;
  float metal; V3 m_f0, m_f82; float spec_w;
  float eta; V3 spec_color; float a;
  float trans_w; V3 trans_tint; V3 Kd;
  V3 m_b;     // F82-tint edge term of the metal Fresnel (opbr_metal_fresnel)
};
constexpr int kLobeFloats = 32;    // record size of the lowered-material table (>= sizeof(OpbrLobes) / 4, 128-byte rows)
static_assert(sizeof(OpbrLobes) <= kLobeFloats * 4, "lobe record too small");

GTLC_D OpbrLobes opbr_lobes(const float* p)
{
  OpbrLobes o;
  const float coat_w = clampf(p[GTLC_OPBR_COAT_WEIGHT], 0.0f, 1.0f);
  const float coat_r = p[GTLC_OPBR_COAT_ROUGHNESS], spec_r = p[GTLC_OPBR_SPECULAR_ROUGHNESS];
  const float coat_ior = p[GTLC_OPBR_COAT_IOR], spec_ior = p[GTLC_OPBR_SPECULAR_IOR], spec_w = p[GTLC_OPBR_SPECULAR_WEIGHT];
  // mtlx:102-131 roughening due to coat
  const float c4 = (coat_r * coat_r) * (coat_r * coat_r), s4 = (spec_r * spec_r) * (spec_r * spec_r);
  const float affected = sqrtf(sqrtf(fminf(1.0f, 2.0f * c4 + s4)));
  const float eff_r = mixf(spec_r, affected, coat_w);
  o.a = fmaxf(eff_r * eff_r, kMinAlpha);
  o.a_c = fmaxf(coat_r * coat_r, kMinAlpha);
  o.coat_w = coat_w; o.coat_ior = coat_ior;
  // mtlx:306-373 specular F0 modulation (IOR ratio inversion under the coat, specular_weight folded into the IOR)
  const float ratio = spec_ior / coat_ior;
  const float tirfix = (ratio > 1.0f) ? ratio : (coat_ior / spec_ior);
  const float eta_s = mixf(spec_ior, tirfix, coat_w);
  const float f0s = (eta_s - 1.0f) / (eta_s + 1.0f);
  const float scaled = clampf(spec_w * (f0s * f0s), 0.0f, 0.99999f);
  const float eps = ((eta_s - 1.0f) > 0.0f ? 1.0f : ((eta_s - 1.0f) < 0.0f ? -1.0f : 0.0f)) * sqrtf(scaled);
  o.eta = (1.0f + eps) / (1.0f - eps);
  o.spec_color = p3(p, GTLC_OPBR_SPECULAR_COLOR_R);
  o.spec_w = spec_w;
  // mtlx:434-470 metal
  o.metal = clampf(p[GTLC_OPBR_BASE_METALNESS], 0.0f, 1.0f);
  const V3 base_color = vmax(p3(p, GTLC_OPBR_BASE_COLOR_R), 0.0f);
  o.m_f0 = base_color * p[GTLC_OPBR_BASE_WEIGHT];
  o.m_f82 = o.spec_color * spec_w;
  // mtlx:195-218 opaque base, mtlx:374-396 transmission
  const float ssw = clampf(p[GTLC_OPBR_SUBSURFACE_WEIGHT], 0.0f, 1.0f);
  const V3 sss = vmax(p3(p, GTLC_OPBR_SUBSURFACE_COLOR_R), 0.0f);
  o.Kd = mix(base_color * p[GTLC_OPBR_BASE_WEIGHT], sss, ssw);
  o.trans_w = clampf(p[GTLC_OPBR_TRANSMISSION_WEIGHT], 0.0f, 1.0f);
  o.trans_tint = (p[GTLC_OPBR_TRANSMISSION_DEPTH] > 0.0f) ? v3(1.0f) : p3(p, GTLC_OPBR_TRANSMISSION_COLOR_R);
  // mtlx:472-550 coat darkening and tint
  const float f0c_s = (coat_ior - 1.0f) / (coat_ior + 1.0f);
  const float Kcoat = 1.0f - (1.0f - f0c_s * f0c_s) / (coat_ior * coat_ior);
  const V3 Edielectric = mix(base_color, p3(p, GTLC_OPBR_SUBSURFACE_COLOR_R), ssw);
  const V3 Ebase = mix(Edielectric, base_color * spec_w, o.metal);
  const V3 darkening = v3((1.0f - Kcoat) / (1.0f - Ebase.x * Kcoat), (1.0f - Kcoat) / (1.0f - Ebase.y * Kcoat), (1.0f - Kcoat) / (1.0f - Ebase.z * Kcoat));
  const V3 modulated = mix(v3(1.0f), darkening, coat_w * p[GTLC_OPBR_COAT_DARKENING]);
  o.coat_att = modulated * mix(v3(1.0f), p3(p, GTLC_OPBR_COAT_COLOR_R), coat_w);
  {  // generalized_schlick with the OpenPBR "F82-tint" edge colour (OpenPBR spec 1.1, metal base): the hit-independent part
    const float mu = 1.0f / 7.0f;
    const float k = mu * ((1.0f - mu) * (1.0f - mu)) * ((1.0f - mu) * (1.0f - mu)) * ((1.0f - mu) * (1.0f - mu));   // mu (1-mu)^6
    const V3 fbar = schlick(o.m_f0, v3(1.0f), mu);
    o.m_b = (fbar * (v3(1.0f) - o.m_f82)) * (1.0f / k);
  }
  return o;
}
GTLC_D OpbrLobes opbr_lobes_of(const float* p, const float* lowered)
{
  if (lowered) {
    OpbrLobes o;
    const float4* q = reinterpret_cast<const float4*>(lowered);
    const float4 a = __ldg(q + 0), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3), e = __ldg(q + 4), f = __ldg(q + 5), g = __ldg(q + 6);
    o.coat_w = a.x; o.a_c = a.y; o.coat_ior = a.z; o.coat_att = v3(a.w, b.x, b.y);
    o.metal = b.z; o.m_f0 = v3(b.w, c.x, c.y); o.m_f82 = v3(c.z, c.w, d.x); o.spec_w = d.y;
    o.eta = d.z; o.spec_color = v3(d.w, e.x, e.y); o.a = e.z;
    o.trans_w = e.w; o.trans_tint = v3(f.x, f.y, f.z); o.Kd = v3(f.w, g.x, g.y);
    o.m_b = v3(g.z, g.w, __ldg(lowered + 28));
    return o;
  }
  return opbr_lobes(p);
}
GTLC_D void opbr_lobes_store(const OpbrLobes& o, float* dst)
{
  const float v[kLobeFloats] = { o.coat_w, o.a_c, o.coat_ior, o.coat_att.x, o.coat_att.y, o.coat_att.z, o.metal, o.m_f0.x, o.m_f0.y, o.m_f0.z,
                                 o.m_f82.x, o.m_f82.y, o.m_f82.z, o.spec_w, o.eta, o.spec_color.x, o.spec_color.y, o.spec_color.z, o.a,
                                 o.trans_w, o.trans_tint.x, o.trans_tint.y, o.trans_tint.z, o.Kd.x, o.Kd.y, o.Kd.z, o.m_b.x, o.m_b.y, o.m_b.z, 0.0f, 0.0f, 0.0f };
  for (int k = 0; k < kLobeFloats; k++) dst[k] = v[k];
}

// generalized_schlick with the OpenPBR "F82-tint" edge colour (OpenPBR spec 1.1, metal base); b = o.m_b (opbr_lobes)
GTLC_D V3 opbr_metal_fresnel(const OpbrLobes& o, float c)
{
  const V3 b = o.m_b;
  c = clampf(c, 0.0f, 1.0f);
  const float om = 1.0f - c;
  const float w6 = c * ((om * om) * (om * om) * (om * om));
  const V3 f = schlick(o.m_f0, v3(1.0f), c) - b * w6;
  return v3(clampf(f.x, 0.0f, 1.0f), clampf(f.y, 0.0f, 1.0f), clampf(f.z, 0.0f, 1.0f)) * o.spec_w;
}

GTLC_D float opbr_rel_eta(const V3& ior1, const V3& ior2, float material_ior)
{
  const float n1 = (ior1.x == -1.0f) ? material_ior : ior1.x;   // BSDF_USE_MATERIAL_IOR, mdl_types.glsl:155
  const float n2 = (ior2.x == -1.0f) ? material_ior : ior2.x;
  return n2 / n1;
}

// Importance sampling.  The lobe is drawn first (xi.z, stochastic: the lanes of a warp disagree even when they share a
// material); everything expensive that more than one lobe needs -- the VNDF half vector, the masking term G1, the second
// Fresnel evaluation -- is then executed ONCE, by all lanes together, with per-lane arguments, instead of once per lobe
// branch with a few lanes each (ncu, C3 stand-in: the closure code ran at 11 of 32 lanes, ggx_sample_vndf at 3).  Per lane
// the sequence of floating-point operations is exactly the one of the per-lobe formulation (and of the oracle).
GTLC_D void openpbr_sample(const float* p, const float* lowered, BsdfSample& d, const ShadingState& s)
{
  const OpbrLobes o = opbr_lobes_of(p, lowered);
  const Frame f = make_frame(s);
  const V3 k1 = to_local(f, d.k1);
  d.bsdf_over_pdf = v3(0.0f); d.pdf = 0.0f; d.event_type = BSDF_EVENT_ABSORB; d.k2 = v3(0.0f);
  d.memo[0] = -1.0f; d.memo[1] = -1.0f;
  if (k1.z <= 0.0f) return;
  const float nk1 = k1.z;
  const float eta = opbr_rel_eta(d.ior1, d.ior2, o.eta);
  float z = d.xi.z;
  enum { LOBE_COAT, LOBE_METAL, LOBE_SPEC, LOBE_TRANS, LOBE_DIFF };

  // ---- lobe selection ----
  const float fc1 = (o.coat_w > 0.0f) ? fresnel_dielectric(nk1, o.coat_ior) : 0.0f;
  const float p_coat = o.coat_w * fc1;
  if (o.coat_w > 0.0f) d.memo[0] = fc1;
  int lobe; float fs1 = 0.0f;
  if (o.coat_w > 0.0f && z < p_coat) lobe = LOBE_COAT;
  else {
    if (o.coat_w > 0.0f) z = (z - p_coat) / (1.0f - p_coat);
    if (z < o.metal) lobe = LOBE_METAL;
    else {
      if (o.metal > 0.0f) z = (z - o.metal) / (1.0f - o.metal);
      fs1 = fresnel_dielectric(nk1, eta);
      d.memo[1] = fs1;
      if (z < fs1) lobe = LOBE_SPEC;
      else {
        z = (z - fs1) / (1.0f - fs1);
        lobe = (z < o.trans_w) ? LOBE_TRANS : LOBE_DIFF;
      }
    }
  }

  // ---- direction ----
  const float alpha = (lobe == LOBE_COAT) ? o.a_c : o.a;
  V3 k2l, h = v3(0.0f, 0.0f, 1.0f); float kh = 0.0f;
  bool refracted = false;
  if (lobe != LOBE_DIFF) {
    h = ggx_sample_vndf(k1, alpha, d.xi.x, d.xi.y);
    kh = dot(k1, h);
  }
  if (lobe == LOBE_DIFF) {
    k2l = cosine_hemisphere(d.xi.x, d.xi.y);
    if (k2l.z <= 0.0f) return;
  } else if (lobe == LOBE_TRANS) {
    // rough transmission through the microfacet normal (dielectric_bsdf, scatter_mode T)
    if (kh <= 0.0f) return;
    const float sin2_t = (1.0f - kh * kh) / (eta * eta);
    if (sin2_t >= 1.0f) {                                     // total internal reflection
      k2l = h * (2.0f * kh) - k1;
      if (k2l.z <= 0.0f) return;
    } else {
      const float cos_t = sqrtf(1.0f - sin2_t);
      k2l = h * (kh / eta - cos_t) - k1 * (1.0f / eta);
      if (k2l.z >= 0.0f) return;
      refracted = true;
    }
  } else {                                                    // ggx_reflect_sample
    k2l = h * (2.0f * kh) - k1;
    if (k2l.z <= 0.0f || kh <= 0.0f) return;
  }

  // ---- masking term of the sampled direction (every microfacet lobe), second Fresnel evaluation (every lobe but metal / TIR) ----
  float g1 = 0.0f, f2 = 0.0f;
  if (lobe != LOBE_DIFF) g1 = ggx_G1(refracted ? v3(k2l.x, k2l.y, -k2l.z) : k2l, alpha);
  const bool tir = (lobe == LOBE_TRANS) && !refracted;
  if (lobe != LOBE_METAL && !tir) {
    const float c2 = (lobe == LOBE_COAT || lobe == LOBE_SPEC) ? kh : ((lobe == LOBE_TRANS) ? fabsf(k2l.z) : k2l.z);
    const float e2 = (lobe == LOBE_COAT) ? o.coat_ior : ((lobe == LOBE_TRANS) ? (1.0f / eta) : eta);
    f2 = fresnel_dielectric(c2, e2);
  }

  // ---- weight ----
  V3 weight; int ev;
  if (lobe == LOBE_COAT) {
    const float fc = o.coat_w * f2;
    weight = v3(g1 * (fc / p_coat));
    ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
  } else {
    if (lobe == LOBE_METAL) {
      weight = opbr_metal_fresnel(o, kh) * g1;
      ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
    } else if (lobe == LOBE_SPEC) {
      weight = o.spec_color * (g1 * (f2 / fs1));
      ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION;
    } else if (lobe == LOBE_TRANS) {
      if (tir) { weight = v3(g1); ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_REFLECTION; }
      else { weight = o.trans_tint * (g1 * ((1.0f - fmaxf(fs1, f2)) / (1.0f - fs1))); ev = BSDF_EVENT_GLOSSY | BSDF_EVENT_TRANSMISSION; }
    } else {
      weight = o.Kd * ((1.0f - fmaxf(fs1, f2)) / (1.0f - fs1));
      ev = BSDF_EVENT_DIFFUSE | BSDF_EVENT_REFLECTION;
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

GTLC_D void openpbr_evaluate(const float* p, const float* lowered, BsdfEval& d, const ShadingState& s)
{
  const OpbrLobes o = opbr_lobes_of(p, lowered);
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
    const float fc1 = (d.memo[0] >= 0.0f) ? d.memo[0] : fresnel_dielectric(nk1, o.coat_ior), fc2 = fresnel_dielectric(nk2, o.coat_ior);
    p_coat = o.coat_w * fc1;
    ggx_reflect_eval(k1, k2, o.a_c, fcos, pdfg, h);
    d.bsdf_glossy += v3((o.coat_w * fresnel_dielectric(dot(k1, h), o.coat_ior)) * fcos);
    pdf += p_coat * pdfg;
    base_scale = o.coat_att * (1.0f - o.coat_w * fmaxf(fc1, fc2));
  }
  ggx_reflect_eval(k1, k2, o.a, fcos, pdfg, h);
  const float kh = dot(k1, h);
  const float fs1 = (d.memo[1] >= 0.0f) ? d.memo[1] : fresnel_dielectric(nk1, eta), fs2 = fresnel_dielectric(nk2, eta);
  const V3 metal = opbr_metal_fresnel(o, kh) * fcos;
  const V3 spec = o.spec_color * (fresnel_dielectric(kh, eta) * fcos);
  const float pd = nk2 * kInvPi;
  const V3 diff = o.Kd * ((1.0f - fmaxf(fs1, fs2)) * (1.0f - o.trans_w) * pd);
  d.bsdf_glossy += base_scale * (metal * o.metal + spec * (1.0f - o.metal));
  d.bsdf_diffuse = base_scale * (diff * (1.0f - o.metal));
  pdf += (1.0f - p_coat) * (o.metal * pdfg + (1.0f - o.metal) * (fs1 * pdfg + (1.0f - fs1) * (1.0f - o.trans_w) * pd));
  d.pdf = pdf;
}


// mdl_bsdf_scattering_auxiliary: albedo_diffuse + albedo_glossy at the viewing angle
GTLC_D V3 openpbr_albedo(const float* p, const float* lowered, V3 k1, const ShadingState& s)
{
  const OpbrLobes o = opbr_lobes_of(p, lowered);
  const float nk1 = clampf(dot(k1, s.normal), 0.0f, 1.0f);
  const float fs = fresnel_dielectric(nk1, o.eta);
  const V3 substrate = o.Kd * (1.0f - o.trans_w) + o.trans_tint * o.trans_w;
  const V3 dielectric = o.spec_color * fs + substrate * (1.0f - fs);
  const V3 base = opbr_metal_fresnel(o, nk1) * o.metal + dielectric * (1.0f - o.metal);
  if (!(o.coat_w > 0.0f)) return base;
  const float fc = o.coat_w * fresnel_dielectric(nk1, o.coat_ior);
  return v3(fc) + (base * o.coat_att) * (1.0f - fc);
}

} // namespace gtlc
