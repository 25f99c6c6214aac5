// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's 2D texture lookup and of the way a textured
// MaterialX input reaches the surface shader.
//   tex_lookup_float4_2d / apply_wrap_and_crop : /root/reference/src/gi/shaders/mdl_interface.glsl:8-38,127-145
//   tex_lookup_float4_3d, tex_texel_float4_2d/_3d : mdl_interface.glsl:45-64, 91-110, 167-185
//   sampler (linear, Repeat, one mip)         : /root/reference/src/gi/impl/Gi.cpp:388-392, CgpuVk.cpp:1861,1985-1990
//   image formats                             : /root/reference/src/gi/impl/TextureManager.cpp:135,196
//   mdl_adapt_normal                          : /root/reference/src/gi/shaders/mdl_interface.glsl:238-256
// The node semantics (UsdUVTexture scale/bias/fallback, UsdTransform2d, srgb_texture, normalmap) live in MaterialX,
// which is not under /root/reference: restated from the MaterialX 1.38/1.39 node definitions, see
// ORACLE_DEVIATIONS.md.  Parity unpinned.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "orc_math.h"
#include "../include/gtlc.h"

namespace orc {

struct Texture { uint32_t width = 0, height = 0, depth = 1; int format = 0; std::vector<uint8_t> data; };

static inline float wrap_coord(float coord, int wrap, int res)   // apply_wrap_and_crop with crop = (0, 1)
{
  if (wrap == GTLC_WRAP_REPEAT) coord -= floorf(coord);
  else {
    if (wrap == GTLC_WRAP_MIRRORED_REPEAT) {
      float tmp = floorf(coord);
      if ((((int) tmp) & 1) != 0) coord = 1.0f - (coord - tmp);
      else coord -= tmp;
    }
    float inv_hdim = 0.5f / (float) res;
    coord = fminf(fmaxf(coord, inv_hdim), 1.0f - inv_hdim);
  }
  return coord * (1.0f - 0.0f) + 0.0f;
}

static inline V4 texel(const Texture& t, int x, int y, int z = 0)
{
  const size_t i = ((size_t) z * t.height + (size_t) y) * t.width + (size_t) x;
  V4 r;
  switch (t.format) {
    case GTLC_TEX_RGBA8_UNORM: { const uint8_t* c = &t.data[i * 4]; r = v4((float) c[0] / 255.0f, (float) c[1] / 255.0f, (float) c[2] / 255.0f, (float) c[3] / 255.0f); break; }
    case GTLC_TEX_RGBA16_FLOAT: { uint16_t c[4]; std::memcpy(c, &t.data[i * 8], 8); r = v4(half_to_float(c[0]), half_to_float(c[1]), half_to_float(c[2]), half_to_float(c[3])); break; }
    case GTLC_TEX_R32_FLOAT: { float c; std::memcpy(&c, &t.data[i * 4], 4); r = v4(c, 0.0f, 0.0f, 1.0f); break; }
    default: { float c[4]; std::memcpy(c, &t.data[i * 16], 16); r = v4(c[0], c[1], c[2], c[3]); break; }
  }
  return r;
}

static inline int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

static inline V4 sample_linear_repeat(const Texture& t, float u, float v)
{
  const int W = (int) t.width, H = (int) t.height;
  float x = u * (float) W - 0.5f, y = v * (float) H - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  int x0 = (int) fx, y0 = (int) fy;
  V4 t00 = texel(t, imod(x0, W), imod(y0, H)), t10 = texel(t, imod(x0 + 1, W), imod(y0, H));
  V4 t01 = texel(t, imod(x0, W), imod(y0 + 1, H)), t11 = texel(t, imod(x0 + 1, W), imod(y0 + 1, H));
  float bx = 1.0f - ax, by = 1.0f - ay;
  return v4((t00.x * bx + t10.x * ax) * by + (t01.x * bx + t11.x * ax) * ay,
            (t00.y * bx + t10.y * ax) * by + (t01.y * bx + t11.y * ax) * ay,
            (t00.z * bx + t10.z * ax) * by + (t01.z * bx + t11.z * ax) * ay,
            (t00.w * bx + t10.w * ax) * by + (t01.w * bx + t11.w * ax) * ay);
}

static inline V4 tex_lookup_float4_2d(const Texture& t, float u, float v, int wrap_u, int wrap_v, bool& clipped)
{
  clipped = (wrap_u == GTLC_WRAP_CLIP && (u < 0.0f || u > 1.0f)) || (wrap_v == GTLC_WRAP_CLIP && (v < 0.0f || v > 1.0f));
  if (clipped) return v4(0.0f, 0.0f, 0.0f, 0.0f);
  return sample_linear_repeat(t, wrap_coord(u, wrap_u, (int) t.width), wrap_coord(v, wrap_v, (int) t.height));
}

// texture(sampler3D): the same sampler on a 3-D image -- two bilinear slices blended along w
static inline V4 sample_linear_repeat_3d(const Texture& t, float u, float v, float w)
{
  const int W = (int) t.width, H = (int) t.height, D = (int) t.depth;
  float x = u * (float) W - 0.5f, y = v * (float) H - 0.5f, z = w * (float) D - 0.5f;
  float fx = floorf(x), fy = floorf(y), fz = floorf(z);
  float ax = x - fx, ay = y - fy, az = z - fz;
  int x0 = (int) fx, y0 = (int) fy, z0 = (int) fz;
  float bx = 1.0f - ax, by = 1.0f - ay, bz = 1.0f - az;
  V4 r = v4(0.0f, 0.0f, 0.0f, 0.0f);
  for (int k = 0; k < 2; k++) {
    const int zz = imod(z0 + k, D); const float wz = k ? az : bz;
    V4 t00 = texel(t, imod(x0, W), imod(y0, H), zz), t10 = texel(t, imod(x0 + 1, W), imod(y0, H), zz);
    V4 t01 = texel(t, imod(x0, W), imod(y0 + 1, H), zz), t11 = texel(t, imod(x0 + 1, W), imod(y0 + 1, H), zz);
    V4 sl = v4((t00.x * bx + t10.x * ax) * by + (t01.x * bx + t11.x * ax) * ay,
               (t00.y * bx + t10.y * ax) * by + (t01.y * bx + t11.y * ax) * ay,
               (t00.z * bx + t10.z * ax) * by + (t01.z * bx + t11.z * ax) * ay,
               (t00.w * bx + t10.w * ax) * by + (t01.w * bx + t11.w * ax) * ay);
    if (k == 0) r = v4(sl.x * wz, sl.y * wz, sl.z * wz, sl.w * wz);
    else r = v4(r.x + sl.x * wz, r.y + sl.y * wz, r.z + sl.z * wz, r.w + sl.w * wz);
  }
  return r;
}

static inline V4 tex_lookup_float4_3d(const Texture& t, float u, float v, float w, int wrap_u, int wrap_v, int wrap_w, bool& clipped)
{
  clipped = (wrap_u == GTLC_WRAP_CLIP && (u < 0.0f || u > 1.0f)) || (wrap_v == GTLC_WRAP_CLIP && (v < 0.0f || v > 1.0f)) ||
            (wrap_w == GTLC_WRAP_CLIP && (w < 0.0f || w > 1.0f));
  if (clipped) return v4(0.0f, 0.0f, 0.0f, 0.0f);
  return sample_linear_repeat_3d(t, wrap_coord(u, wrap_u, (int) t.width), wrap_coord(v, wrap_v, (int) t.height), wrap_coord(w, wrap_w, (int) t.depth));
}

static inline V4 tex_texel_float4(const Texture& t, int x, int y, int z)
{
  if (x < 0 || x >= (int) t.width || y < 0 || y >= (int) t.height || z < 0 || z >= (int) t.depth) return v4(0.0f, 0.0f, 0.0f, 0.0f);
  return texel(t, x, y, z);
}

static inline float srgb_texture_to_linear(float c)
{
  const float breakPnt = 0.03928571566939354f, slope = 0.07738015800714493f;
  const float gamma = 2.4f, scale = 0.9478672742843628f, offset = 0.05213269963860512f;
  return c > breakPnt ? powf(fmaxf(scale * c + offset, 0.0f), gamma) : c * slope;
}

static inline V3 mdl_adapt_normal(V3 worldRayDir, V3 geomNormal, V3 normal)
{
  V3 r = normalize(worldRayDir - normal * (2.0f * dot(normal, worldRayDir)));
  float a = fmaxf(0.0f, dot(r, -geomNormal));
  float b = dot(normal, geomNormal);
  V3 tangent = normalize(r + normal * (a / b));
  return normalize(-worldRayDir + tangent);
}

// One binding at a hit: coordinate (UsdTransform2d on uv, or object-space position * scale + offset for the 3-D lookups), the
// lookup its mode names, MaterialX `default` when clipped, colour space, UsdUVTexture scale / bias.  ch = four source channels.
static inline void eval_binding(const std::vector<Texture>& textures, const GtlcTextureBinding& b, float u0, float v0, V3 localPos, float ch[4])
{
  const double PI_D = 3.14159265358979323846;
  if (b.combine & GTLC_COMBINE_CONSTANT) { for (int k = 0; k < 4; k++) ch[k] = b.bias[k]; return; }   // a constant operand: no lookup
  const Texture& tex = textures[(size_t) b.texture - 1];
  bool clipped = false;
  V4 t;
  if (b.lookup == GTLC_LOOKUP_FILTERED_2D || b.lookup == GTLC_LOOKUP_TEXEL_2D) {
    const float c = (float) cos((double) b.uvRotation * PI_D / 180.0), s = (float) sin((double) b.uvRotation * PI_D / 180.0);
    const float su = u0 * b.uvScale[0], sv = v0 * b.uvScale[1];
    const float u = (su * c - sv * s) + b.uvTranslation[0];
    const float v = (su * s + sv * c) + b.uvTranslation[1];
    if (b.lookup == GTLC_LOOKUP_FILTERED_2D) t = tex_lookup_float4_2d(tex, u, v, b.wrapU, b.wrapV, clipped);
    else t = tex_texel_float4(tex, (int) floorf(u * (float) tex.width), (int) floorf(v * (float) tex.height), 0);
  } else {
    const float x = localPos.x * b.xyzScale[0] + b.xyzOffset[0];
    const float y = localPos.y * b.xyzScale[1] + b.xyzOffset[1];
    const float z = localPos.z * b.xyzScale[2] + b.xyzOffset[2];
    if (b.lookup == GTLC_LOOKUP_FILTERED_3D) t = tex_lookup_float4_3d(tex, x, y, z, b.wrapU, b.wrapV, b.wrapW, clipped);
    else t = tex_texel_float4(tex, (int) floorf(x * (float) tex.width), (int) floorf(y * (float) tex.height), (int) floorf(z * (float) tex.depth));
  }
  ch[0] = t.x; ch[1] = t.y; ch[2] = t.z; ch[3] = t.w;
  if (clipped) for (int k = 0; k < 4; k++) ch[k] = b.fallback[k];
  if (b.srgb) for (int k = 0; k < 3; k++) ch[k] = srgb_texture_to_linear(ch[k]);
  for (int k = 0; k < 4; k++) ch[k] = ch[k] * b.scale[k] + b.bias[k];
}

// Evaluates the bindings of one material at (u, v) / the object-space hit position: overwrites floats of `params`; returns
// true and the mapped shading normal when a GTLC_BIND_NORMAL binding is present.  GtlcTextureBinding.texture = 1-based index
// into `textures`.
static inline bool apply_texture_bindings(const std::vector<Texture>& textures, const std::vector<GtlcTextureBinding>& bindings,
                                          float u0, float v0, V3 localPos, V3 worldRayDir, V3 geomNormal, V3 normal, V3 tangentU, V3 tangentV,
                                          float* params, V3& mappedNormal)
{
  bool hasNormal = false;
  float scratch[GTLC_BIND_SCRATCH_COUNT] = { 0.0f, 0.0f, 0.0f, 0.0f };
  auto slot = [&](int i) -> float& { return i < GTLC_MATERIAL_PARAM_COUNT ? params[i] : scratch[i - GTLC_MATERIAL_PARAM_COUNT]; };
  for (const GtlcTextureBinding& b : bindings) {
    float ch[4];
    eval_binding(textures, b, u0, v0, localPos, ch);
    float val[3] = { 0.0f, 0.0f, 0.0f };
    for (int k = 0; k < b.componentCount && k < 3; k++) val[k] = ch[b.channels[k]];
    if (b.paramIndex == GTLC_BIND_NORMAL) {
      V3 n = normalize((tangentU * val[0] + tangentV * val[1]) + normal * val[2]);
      mappedNormal = mdl_adapt_normal(worldRayDir, geomNormal, n);
      hasNormal = true;
    } else {                                                              // gtlc.h GtlcCombineOp: bindings applied in order build an input from several lookups
      const int op = b.combine & 15;
      const float f = (op == GTLC_COMBINE_MIX) ? slot((b.combine >> 8) & 255) : 0.0f;
      for (int k = 0; k < b.componentCount && k < 3; k++) {
        float& x = slot(b.paramIndex + k);
        if (op == GTLC_COMBINE_SET) x = val[k];
        else if (op == GTLC_COMBINE_MULTIPLY) x = x * val[k];
        else if (op == GTLC_COMBINE_ADD) x = x + val[k];
        else x = x * (1.0f - f) + val[k] * f;                             // MDL math::lerp
      }
    }
  }
  return hasNormal;
}

} // namespace orc
