// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under gatling_b200/ may
// include, link or call this.  CPU restatement of the pure functions of the reference's shader
// prelude, /root/reference/src/gi/shaders/common.glsl, and of the host-side encoders in
// /root/reference/src/gi/impl/Gi.cpp:287-300.  Parity vs the real reference: UNPINNED (no golden
// vectors exist in the reference for these functions) -- the known-answer vectors
// in tests/golden/ are frozen outputs of THIS file.
//
// Floating-point contract shared with the CUDA kernels: every + - * / sqrt is a single IEEE-754
// binary32 operation in the written order (compile with -ffp-contract=off / nvcc -fmad=false);
// fused multiply-adds appear only where `fmaf` is written explicitly.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace orc {

static const float FLOAT_MAX = 3.402823466e38f;   // common.glsl:7
static const float FLOAT_MIN = 1.175494351e-38f;  // common.glsl:8
static const float PI = 3.1415926535897932384626433832795f; // common.glsl:9

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
static inline V4 v4(float x, float y, float z, float w) { V4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

static inline V3 v3(float x, float y, float z) { V3 r = { x, y, z }; return r; }
static inline V3 v3(float s) { V3 r = { s, s, s }; return r; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
static inline V3& operator*=(V3& a, float s) { a = a * s; return a; }
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length(V3 a) { return sqrtf(dot(a, a)); }
// GLSL normalize(): v * (1 / length) -- one division, three multiplies (contract with the kernels).
static inline V3 normalize(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
static inline V3 vmax(V3 a, float s) { return v3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }
static inline float max3(V3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
static inline V3 mix(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

static inline uint32_t floatBitsToUint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float uintBitsToFloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline int32_t floatBitsToInt(float f) { int32_t u; memcpy(&u, &f, 4); return u; }
static inline float intBitsToFloat(int32_t u) { float f; memcpy(&f, &u, 4); return f; }

// common.glsl:18-36
static inline float safe_div(float f1, float f2) { return (f2 == 0.0f) ? 0.0f : (f1 / f2); }
static inline V3 safe_div(V3 v, float f) { return (f == 0.0f) ? v3(0.0f) : (v / f); }
static inline V3 safe_div(V3 a, V3 b)
{
  return v3((b.x == 0.0f) ? 0.0f : (a.x / b.x), (b.y == 0.0f) ? 0.0f : (a.y / b.y), (b.z == 0.0f) ? 0.0f : (a.z / b.z));
}

// common.glsl:44-47
static inline float uintAsFloat(uint32_t v) { return uintBitsToFloat(0x3f800000u | (v >> 9)) - 1.0f; }

// common.glsl:82-90
static inline uint32_t hash_theironborn(uint32_t x)
{
  x ^= x >> 16u; x *= 0x21f0aaadu; x ^= x >> 15u; x *= 0xd35a2d97u; x ^= x >> 15u;
  return x;
}

// common.glsl:93-98 -- note the reference passes `state` inout AND assigns the return value to it
// (rng1d_next1f, common.glsl:100-104): the stored state is the *output word*, not the LCG state.
static inline uint32_t hash_pcg32(uint32_t& state)
{
  state = state * 747796405u + 2891336453u;
  uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}

// common.glsl:100-104
static inline float rng1d_next1f(uint32_t& rng_state)
{
  rng_state = hash_pcg32(rng_state);
  return uintAsFloat(rng_state);
}
// common.glsl:106-109 (GLSL evaluates constructor arguments left to right)
static inline V2 rng1d_next2f(uint32_t& s) { V2 r; r.x = rng1d_next1f(s); r.y = rng1d_next1f(s); return r; }
// common.glsl:111-119
static inline V4 rng1d_next4f(uint32_t& s) { V4 r; r.x = rng1d_next1f(s); r.y = rng1d_next1f(s); r.z = rng1d_next1f(s); r.w = rng1d_next1f(s); return r; }
// common.glsl:121-124
static inline uint32_t rng1d_init(uint32_t pixel_index, uint32_t sampleIndex) { return hash_theironborn(pixel_index * (sampleIndex + 1u)); }

// common.glsl:129-137 (Duff et al. 2017)
static inline void orthonormal_basis(V3 n, V3& b1, V3& b2)
{
  float nsign = (n.z >= 0.0f ? 1.0f : -1.0f);
  float a = -1.0f / (nsign + n.z);
  float b = n.x * n.y * a;
  b1 = v3(1.0f + nsign * n.x * n.x * a, nsign * b, -nsign * n.x);
  b2 = v3(b, nsign + n.y * n.y * a, -n.y);
}

// common.glsl:143-162 (Waechter & Binder; intScale 64, origin 1/32, floatScale 1/65536)
static inline V3 offset_ray_origin(V3 p, V3 n)
{
  const float origin = 1.0f / 32.0f;
  const float floatScale = 1.0f / 65536.0f;
  const float intScale = 64.0f;
  int32_t ox = (int32_t)(n.x * intScale), oy = (int32_t)(n.y * intScale), oz = (int32_t)(n.z * intScale);
  V3 intPos = v3(intBitsToFloat(floatBitsToInt(p.x) + ((p.x >= 0.0f) ? ox : -ox)),
                 intBitsToFloat(floatBitsToInt(p.y) + ((p.y >= 0.0f) ? oy : -oy)),
                 intBitsToFloat(floatBitsToInt(p.z) + ((p.z >= 0.0f) ? oz : -oz)));
  V3 fo = n * floatScale;
  return v3((fabsf(p.x) >= origin) ? intPos.x : (p.x + fo.x),
            (fabsf(p.y) >= origin) ? intPos.y : (p.y + fo.y),
            (fabsf(p.z) >= origin) ? intPos.z : (p.z + fo.z));
}

// GLSL packUnorm2x16 / unpackUnorm2x16 (x in the low 16 bits)
static inline uint32_t packUnorm2x16(V2 v)
{
  uint32_t x = (uint32_t) roundf(clampf(v.x, 0.0f, 1.0f) * 65535.0f);
  uint32_t y = (uint32_t) roundf(clampf(v.y, 0.0f, 1.0f) * 65535.0f);
  return x | (y << 16);
}
static inline V2 unpackUnorm2x16(uint32_t e) { V2 r; r.x = (float)(e & 0xFFFFu) / 65535.0f; r.y = (float)(e >> 16) / 65535.0f; return r; }

// common.glsl:173-184 (shader-side encoder) -- kept for the KATs
static inline V2 encode_octahedral(V3 v)
{
  float s = (fabsf(v.x) + fabsf(v.y)) + fabsf(v.z);
  v = v / s;
  V2 e;
  if (v.z < 0.0f) {
    e.x = (1.0f - fabsf(v.y)) * ((v.x >= 0.0f) ? 1.0f : -1.0f);
    e.y = (1.0f - fabsf(v.x)) * ((v.y >= 0.0f) ? 1.0f : -1.0f);
  } else { e.x = v.x; e.y = v.y; }
  e.x = e.x * 0.5f + 0.5f; e.y = e.y * 0.5f + 0.5f;
  return e;
}
// common.glsl:186-196
static inline V3 decode_octahedral(V2 e)
{
  e.x = e.x * 2.0f - 1.0f; e.y = e.y * 2.0f - 1.0f;
  V3 v = v3(e.x, e.y, (1.0f - fabsf(e.x)) - fabsf(e.y));
  float t = fmaxf(-v.z, 0.0f);
  v.x += (v.x >= 0.0f) ? -t : t;
  v.y += (v.y >= 0.0f) ? -t : t;
  return normalize(v);
}
// common.glsl:198-207
static inline uint32_t encode_direction(V3 d) { return packUnorm2x16(encode_octahedral(d)); }
static inline V3 decode_direction(uint32_t e) { return decode_octahedral(unpackUnorm2x16(e)); }

// Host encoder, Gi.cpp:287-300 (_EncodeOctahedral/_EncodeDirection): normalises first.
static inline uint32_t host_encode_direction(V3 v)
{
  v = normalize(v);
  return encode_direction(v);
}

// common.glsl:210-220
static inline V3 sample_hemisphere(V2 xi)
{
  float a = sqrtf(xi.x);
  float b = PI * 2.0f * xi.y;
  return v3(a * cosf(b), a * sinf(b), sqrtf(1.0f - xi.x));
}
// common.glsl:223-230
static inline V3 sample_sphere(V2 xi, V3 radius)
{
  float a = 1.0f - 2.0f * xi.x;
  float b = sqrtf(1.0f - a * a);
  float phi = 2.0f * PI * xi.y;
  return v3(b * cosf(phi), b * sinf(phi), a) * radius;
}
// common.glsl:233-252
static inline V2 sample_disk(V2 xi, V2 radius)
{
  float a = 2.0f * xi.x - 1.0f;
  float b = 2.0f * xi.y - 1.0f;
  V2 r; float phi;
  if ((a * a) > (b * b)) { r.x = radius.x * a; r.y = radius.y * a; phi = (PI / 4.0f) * (b / a); }
  else { r.x = radius.x * b; r.y = radius.y * b; phi = (PI / 2.0f) - (PI / 4.0f) * safe_div(a, b); }
  V2 o; o.x = r.x * cosf(phi); o.y = r.y * sinf(phi);
  return o;
}
// common.glsl:254-257
static inline float luminance(V3 c) { return (c.x * 0.2126f + c.y * 0.7152f) + c.z * 0.0722f; }

// IEEE binary16 conversion, round-to-nearest-even (GLSL packHalf2x16 / glm::packHalf2x16).
static inline uint16_t float_to_half(float f)
{
  uint32_t x = floatBitsToUint(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t mant = x & 0x007FFFFFu;
  int32_t exp = (int32_t)((x >> 23) & 0xFF);
  if (exp == 255) return (uint16_t)(sign | 0x7C00u | (mant ? 0x200u : 0u));
  int32_t e = exp - 127 + 15;
  if (e >= 31) return (uint16_t)(sign | 0x7C00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t) sign;
    mant |= 0x00800000u;
    uint32_t shift = (uint32_t)(14 - e);
    uint32_t h = mant >> shift;
    uint32_t rem = mant & ((1u << shift) - 1u);
    uint32_t half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) h++;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t) e << 10) | (mant >> 13);
  uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
  return (uint16_t)(sign | h);
}
static inline float half_to_float(uint16_t h)
{
  uint32_t sign = ((uint32_t) h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1Fu;
  uint32_t mant = h & 0x3FFu;
  if (exp == 0) {
    if (mant == 0) return uintBitsToFloat(sign);
    float m = (float) mant * (1.0f / 16777216.0f); // mant * 2^-24
    return (sign ? -m : m);
  }
  if (exp == 31) return uintBitsToFloat(sign | 0x7F800000u | (mant << 13));
  return uintBitsToFloat(sign | ((exp + 112u) << 23) | (mant << 13));
}
static inline uint32_t packHalf2x16(float a, float b) { return (uint32_t) float_to_half(a) | ((uint32_t) float_to_half(b) << 16); }
static inline V2 unpackHalf2x16(uint32_t p) { V2 r; r.x = half_to_float((uint16_t)(p & 0xFFFFu)); r.y = half_to_float((uint16_t)(p >> 16)); return r; }

} // namespace orc
