// TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's 2D texture lookup and of the way a textured
// MaterialX input reaches the surface shader.
//   tex_lookup_float4_2d / apply_wrap_and_crop : /root/reference/src/gi/shaders/mdl_interface.glsl:8-38,127-145
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

struct Texture { uint32_t width = 0, height = 0; int format = 0; std::vector<uint8_t> data; };

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

static inline V4 texel(const Texture& t, int x, int y)
{
  const size_t i = (size_t) y * t.width + (size_t) x;
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

// Evaluates the bindings of one material at (u, v): overwrites floats of `params`; returns true and the mapped
// shading normal when a GTLC_BIND_NORMAL binding is present.  GtlcTextureBinding.texture = 1-based index into `textures`.
static inline bool apply_texture_bindings(const std::vector<Texture>& textures, const std::vector<GtlcTextureBinding>& bindings,
                                          float u0, float v0, V3 worldRayDir, V3 geomNormal, V3 normal, V3 tangentU, V3 tangentV,
                                          float* params, V3& mappedNormal)
{
  bool hasNormal = false;
  const double PI_D = 3.14159265358979323846;
  for (const GtlcTextureBinding& b : bindings) {
    const float c = (float) cos((double) b.uvRotation * PI_D / 180.0), s = (float) sin((double) b.uvRotation * PI_D / 180.0);
    const float su = u0 * b.uvScale[0], sv = v0 * b.uvScale[1];
    const float u = (su * c - sv * s) + b.uvTranslation[0];
    const float v = (su * s + sv * c) + b.uvTranslation[1];
    bool clipped;
    V4 t = tex_lookup_float4_2d(textures[(size_t) b.texture - 1], u, v, b.wrapU, b.wrapV, clipped);
    float ch[4] = { t.x, t.y, t.z, t.w };
    if (clipped) for (int k = 0; k < 4; k++) ch[k] = b.fallback[k];
    if (b.srgb) for (int k = 0; k < 3; k++) ch[k] = srgb_texture_to_linear(ch[k]);
    for (int k = 0; k < 4; k++) ch[k] = ch[k] * b.scale[k] + b.bias[k];
    float val[3] = { 0.0f, 0.0f, 0.0f };
    for (int k = 0; k < b.componentCount && k < 3; k++) val[k] = ch[b.channels[k]];
    if (b.paramIndex == GTLC_BIND_NORMAL) {
      V3 n = normalize((tangentU * val[0] + tangentV * val[1]) + normal * val[2]);
      mappedNormal = mdl_adapt_normal(worldRayDir, geomNormal, n);
      hasNormal = true;
    } else {
      for (int k = 0; k < b.componentCount && k < 3; k++) params[b.paramIndex + k] = val[k];
    }
  }
  return hasNormal;
}

} // namespace orc
