// Texture lookups and per-hit material inputs.
//
// The reference samples images inside MDL-generated code through tex_lookup_float4_2d
// (/root/reference/src/gi/shaders/mdl_interface.glsl:127-145: clip test, apply_wrap_and_crop :8-38, then one
// `texture()` with a linear/linear sampler in Repeat mode on a single mip, Gi.cpp:388-392).  Here the same lookup is
// four explicit texel gathers and a float32 bilinear blend (fixed operator order, no texture unit: its 8-bit blend
// weights would not match the CPU oracle), driven by per-material bindings that overwrite floats of the closure's
// parameter block before the BSDF runs -- the ahead-of-time counterpart of MaterialX image / UsdUVTexture /
// UsdTransform2d nodes feeding a surface-shader input.
#pragma once
#include <cuda_fp16.h>
#include "device_math.cuh"
#include "../../include/gtlc.h"

namespace gtlc {

struct DeviceTexture { const void* texels; uint32_t width, height, format, depth; };   // depth 1 = 2-D image

struct DeviceTexBinding {
  int32_t paramIndex;
  uint32_t packed;          // componentCount | ch0 << 4 | ch1 << 6 | ch2 << 8 | wrapU << 10 | wrapV << 12 | srgb << 14 | lookup << 15 | wrapW << 17
  uint32_t texture;         // index into DeviceScene.textures
  uint32_t combine;         // GtlcTextureBinding.combine: op | constant << 4 | factorParam << 8
  float fallback[4], scale[4], bias[4];
  float uvCos, uvSin, uvScale[2], uvTranslation[2], pad2[2];
  float xyzScale[3], xyzOffset[3], pad3[2];   // 3-D lookups: coordinate = object-space position * scale + offset
};
static_assert(sizeof(DeviceTexBinding) == 128, "DeviceTexBinding is 128 bytes");

// mdl_interface.glsl:8-38 with crop = (0, 1) (the only crop MaterialX-generated MDL passes)
GTLC_D float apply_wrap_and_crop(float coord, int wrap, int res)
{
  if (wrap == GTLC_WRAP_REPEAT) {
    coord -= floorf(coord);
  } else {
    if (wrap == GTLC_WRAP_MIRRORED_REPEAT) {
      const float tmp = floorf(coord);
      if ((((int) tmp) & 1) != 0) coord = 1.0f - (coord - tmp);
      else coord -= tmp;
    }
    const float inv_hdim = 0.5f / (float) res;
    coord = fminf(fmaxf(coord, inv_hdim), 1.0f - inv_hdim);
  }
  return coord * (1.0f - 0.0f) + 0.0f;
}

GTLC_D V4 tex_fetch(const DeviceTexture& t, int x, int y, int z = 0)
{
  const size_t i = ((size_t) z * t.height + (size_t) y) * t.width + (size_t) x;
  V4 r;
  if (t.format == GTLC_TEX_RGBA8_UNORM) {
    const uchar4 c = __ldg((const uchar4*) t.texels + i);
    r.x = (float) c.x / 255.0f; r.y = (float) c.y / 255.0f; r.z = (float) c.z / 255.0f; r.w = (float) c.w / 255.0f;
  } else if (t.format == GTLC_TEX_RGBA16_FLOAT) {
    const uint2 c = __ldg((const uint2*) t.texels + i);
    const __half2 lo = *reinterpret_cast<const __half2*>(&c.x), hi = *reinterpret_cast<const __half2*>(&c.y);
    r.x = __low2float(lo); r.y = __high2float(lo); r.z = __low2float(hi); r.w = __high2float(hi);
  } else if (t.format == GTLC_TEX_R32_FLOAT) {
    const float c = __ldg((const float*) t.texels + i);
    r.x = c; r.y = 0.0f; r.z = 0.0f; r.w = 1.0f;       // Vulkan: missing components read (0, 0, 1)
  } else {
    const float4 c = __ldg((const float4*) t.texels + i);
    r.x = c.x; r.y = c.y; r.z = c.z; r.w = c.w;
  }
  return r;
}

// `texture(sampler2D(...), coord)`: VK_FILTER_LINEAR, addressing Repeat, unnormalised coordinate u*W - 0.5
GTLC_D V4 tex_bilinear(const DeviceTexture& t, float u, float v)
{
  const int W = (int) t.width, H = (int) t.height;
  const float x = u * (float) W - 0.5f, y = v * (float) H - 0.5f;
  const float fx = floorf(x), fy = floorf(y);
  const float ax = x - fx, ay = y - fy;
  const int x0 = (int) fx, y0 = (int) fy;
  int xa = x0 % W; if (xa < 0) xa += W;
  int xb = (x0 + 1) % W; if (xb < 0) xb += W;
  int ya = y0 % H; if (ya < 0) ya += H;
  int yb = (y0 + 1) % H; if (yb < 0) yb += H;
  const V4 t00 = tex_fetch(t, xa, ya), t10 = tex_fetch(t, xb, ya), t01 = tex_fetch(t, xa, yb), t11 = tex_fetch(t, xb, yb);
  const float bx = 1.0f - ax, by = 1.0f - ay;
  V4 r;
  r.x = (t00.x * bx + t10.x * ax) * by + (t01.x * bx + t11.x * ax) * ay;
  r.y = (t00.y * bx + t10.y * ax) * by + (t01.y * bx + t11.y * ax) * ay;
  r.z = (t00.z * bx + t10.z * ax) * by + (t01.z * bx + t11.z * ax) * ay;
  r.w = (t00.w * bx + t10.w * ax) * by + (t01.w * bx + t11.w * ax) * ay;
  return r;
}

// mdl_interface.glsl:127-145.  `clipped` tells the caller to substitute the MaterialX `default` (the MDL lookup itself
// returns 0 there; mx_image_* checks the same condition first and returns `default`).
GTLC_D V4 tex_lookup_float4_2d(const DeviceTexture& t, float u, float v, int wrapU, int wrapV, bool& clipped)
{
  clipped = (wrapU == GTLC_WRAP_CLIP && (u < 0.0f || u > 1.0f)) || (wrapV == GTLC_WRAP_CLIP && (v < 0.0f || v > 1.0f));
  V4 zero; zero.x = zero.y = zero.z = zero.w = 0.0f;
  if (clipped) return zero;
  u = apply_wrap_and_crop(u, wrapU, (int) t.width);
  v = apply_wrap_and_crop(v, wrapV, (int) t.height);
  return tex_bilinear(t, u, v);
}

// `texture(sampler3D(...), coord)`: the same sampler (linear, Repeat) on a 3-D image: two bilinear slices blended along w
GTLC_D V4 tex_trilinear(const DeviceTexture& t, float u, float v, float w)
{
  const int W = (int) t.width, H = (int) t.height, D = (int) t.depth;
  const float x = u * (float) W - 0.5f, y = v * (float) H - 0.5f, z = w * (float) D - 0.5f;
  const float fx = floorf(x), fy = floorf(y), fz = floorf(z);
  const float ax = x - fx, ay = y - fy, az = z - fz;
  const int x0 = (int) fx, y0 = (int) fy, z0 = (int) fz;
  int xa = x0 % W; if (xa < 0) xa += W;
  int xb = (x0 + 1) % W; if (xb < 0) xb += W;
  int ya = y0 % H; if (ya < 0) ya += H;
  int yb = (y0 + 1) % H; if (yb < 0) yb += H;
  int za = z0 % D; if (za < 0) za += D;
  int zb = (z0 + 1) % D; if (zb < 0) zb += D;
  const float bx = 1.0f - ax, by = 1.0f - ay, bz = 1.0f - az;
  V4 r; r.x = r.y = r.z = r.w = 0.0f;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int zz = k ? zb : za; const float wz = k ? az : bz;
    const V4 t00 = tex_fetch(t, xa, ya, zz), t10 = tex_fetch(t, xb, ya, zz), t01 = tex_fetch(t, xa, yb, zz), t11 = tex_fetch(t, xb, yb, zz);
    const float sx = (t00.x * bx + t10.x * ax) * by + (t01.x * bx + t11.x * ax) * ay;
    const float sy = (t00.y * bx + t10.y * ax) * by + (t01.y * bx + t11.y * ax) * ay;
    const float sz = (t00.z * bx + t10.z * ax) * by + (t01.z * bx + t11.z * ax) * ay;
    const float sw = (t00.w * bx + t10.w * ax) * by + (t01.w * bx + t11.w * ax) * ay;
    if (k == 0) { r.x = sx * wz; r.y = sy * wz; r.z = sz * wz; r.w = sw * wz; }
    else { r.x = r.x + sx * wz; r.y = r.y + sy * wz; r.z = r.z + sz * wz; r.w = r.w + sw * wz; }
  }
  return r;
}

// mdl_interface.glsl:45-64
GTLC_D V4 tex_lookup_float4_3d(const DeviceTexture& t, float u, float v, float w, int wrapU, int wrapV, int wrapW, bool& clipped)
{
  clipped = (wrapU == GTLC_WRAP_CLIP && (u < 0.0f || u > 1.0f)) || (wrapV == GTLC_WRAP_CLIP && (v < 0.0f || v > 1.0f)) ||
            (wrapW == GTLC_WRAP_CLIP && (w < 0.0f || w > 1.0f));
  V4 zero; zero.x = zero.y = zero.z = zero.w = 0.0f;
  if (clipped) return zero;
  u = apply_wrap_and_crop(u, wrapU, (int) t.width);
  v = apply_wrap_and_crop(v, wrapV, (int) t.height);
  w = apply_wrap_and_crop(w, wrapW, (int) t.depth);
  return tex_trilinear(t, u, v, w);
}

// mdl_interface.glsl:167-185 (2-D, z = 0) and :91-110 (3-D): texelFetch, (0,0,0,0) outside the image
GTLC_D V4 tex_texel_float4(const DeviceTexture& t, int x, int y, int z)
{
  V4 zero; zero.x = zero.y = zero.z = zero.w = 0.0f;
  if (x < 0 || x >= (int) t.width || y < 0 || y >= (int) t.height || z < 0 || z >= (int) t.depth) return zero;
  return tex_fetch(t, x, y, z);
}

// MaterialX srgb_texture -> lin_rec709 (stdlib/genglsl/lib/mx_transform_color.glsl, mx_srgb_texture_to_lin_rec709)
GTLC_D float srgb_texture_to_linear(float c)
{
  const float breakPnt = 0.03928571566939354f, slope = 0.07738015800714493f;
  const float gamma = 2.4f, scale = 0.9478672742843628f, offset = 0.05213269963860512f;
  const float powSeg = powf(fmaxf(scale * c + offset, 0.0f), gamma);
  const float linSeg = c * slope;
  return c > breakPnt ? powSeg : linSeg;
}

// mdl_interface.glsl:238-256
GTLC_D V3 mdl_adapt_normal(V3 worldRayDir, V3 geomNormal, V3 normal)
{
  const V3 r = normalize(worldRayDir - normal * (2.0f * dot(normal, worldRayDir)));   // reflect(I, N) = I - 2 dot(N, I) N
  const float a = fmaxf(0.0f, dot(r, -geomNormal));
  const float b = dot(normal, geomNormal);
  const V3 tangent = normalize(r + normal * (a / b));
  return normalize(-worldRayDir + tangent);
}

// One binding evaluated at a hit: texture coordinate (UsdTransform2d on the hit's uv, or the object-space position for
// the 3-D lookups), the lookup its mode names, MaterialX `default` when clipped, colour space, UsdUVTexture scale / bias.
// ch[0..3] = the four source channels; the caller picks the ones the binding routes to its input.
GTLC_D void eval_tex_binding(const DeviceTexBinding& tb, const DeviceTexture* __restrict__ textures, float stu, float stv, V3 localPos, float (&ch)[4])
{
  if (tb.combine & GTLC_COMBINE_CONSTANT) { ch[0] = tb.bias[0]; ch[1] = tb.bias[1]; ch[2] = tb.bias[2]; ch[3] = tb.bias[3]; return; }
  const DeviceTexture& t = textures[tb.texture];
  const int wrapU = (int) ((tb.packed >> 10) & 3u), wrapV = (int) ((tb.packed >> 12) & 3u), wrapW = (int) ((tb.packed >> 17) & 3u);
  const int lookup = (int) ((tb.packed >> 15) & 3u);
  bool clipped = false;
  V4 c;
  if (lookup == GTLC_LOOKUP_FILTERED_2D || lookup == GTLC_LOOKUP_TEXEL_2D) {
    const float su = stu * tb.uvScale[0], sv = stv * tb.uvScale[1];                    // UsdTransform2d: scale, rotate, translate
    const float u = (su * tb.uvCos - sv * tb.uvSin) + tb.uvTranslation[0];
    const float v = (su * tb.uvSin + sv * tb.uvCos) + tb.uvTranslation[1];
    if (lookup == GTLC_LOOKUP_FILTERED_2D) c = tex_lookup_float4_2d(t, u, v, wrapU, wrapV, clipped);
    else c = tex_texel_float4(t, (int) floorf(u * (float) t.width), (int) floorf(v * (float) t.height), 0);
  } else {
    const float x = localPos.x * tb.xyzScale[0] + tb.xyzOffset[0];
    const float y = localPos.y * tb.xyzScale[1] + tb.xyzOffset[1];
    const float z = localPos.z * tb.xyzScale[2] + tb.xyzOffset[2];
    if (lookup == GTLC_LOOKUP_FILTERED_3D) c = tex_lookup_float4_3d(t, x, y, z, wrapU, wrapV, wrapW, clipped);
    else c = tex_texel_float4(t, (int) floorf(x * (float) t.width), (int) floorf(y * (float) t.height), (int) floorf(z * (float) t.depth));
  }
  ch[0] = c.x; ch[1] = c.y; ch[2] = c.z; ch[3] = c.w;
  if (clipped) { ch[0] = tb.fallback[0]; ch[1] = tb.fallback[1]; ch[2] = tb.fallback[2]; ch[3] = tb.fallback[3]; }
  if ((tb.packed >> 14) & 1u) { ch[0] = srgb_texture_to_linear(ch[0]); ch[1] = srgb_texture_to_linear(ch[1]); ch[2] = srgb_texture_to_linear(ch[2]); }
#pragma unroll
  for (int k = 0; k < 4; k++) ch[k] = ch[k] * tb.scale[k] + tb.bias[k];               // UsdUVTexture scale / bias
}

struct ShadingState;   // wavefront.cuh

} // namespace gtlc
