// Device-side math prelude of the wavefront kernels: the CUDA counterpart of the reference's shader
// prelude /root/reference/src/gi/shaders/common.glsl (RNG :80-124, uintAsFloat :44-47,
// orthonormal_basis :129-137, offset_ray_origin :143-162, octahedral codec :173-207,
// sample_hemisphere/sphere/disk :210-252, luminance :254-257).
//
// Floating-point contract: this translation unit is compiled with -fmad=false, so every + - * /
// sqrt below is one IEEE-754 binary32 operation in the written order; fused multiply-adds appear
// only where fmaf() is written (BVH8 slab tests).  Same contract as the CPU oracle, which makes the
// integer/intersection parts of the path bit-reproducible on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace gtlc {

#define GTLC_HD __host__ __device__ __forceinline__
#define GTLC_D __device__ __forceinline__

constexpr float kFloatMax = 3.402823466e38f;
constexpr float kFloatMin = 1.175494351e-38f;
constexpr float kPi = 3.1415926535897932384626433832795f;
constexpr float kInvPi = 0.31830988618379067f;

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

GTLC_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GTLC_HD V3 v3(float s) { V3 r; r.x = s; r.y = s; r.z = s; return r; }
GTLC_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
GTLC_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
GTLC_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
GTLC_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
GTLC_HD V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
GTLC_HD V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
GTLC_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
GTLC_HD V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
GTLC_HD V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
GTLC_HD V3& operator*=(V3& a, float s) { a = a * s; return a; }
GTLC_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
GTLC_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GTLC_HD float length(V3 a) { return sqrtf(dot(a, a)); }
GTLC_HD V3 normalize(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
GTLC_HD V3 vmax(V3 a, float s) { return v3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }
GTLC_HD V3 mix(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }
GTLC_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
GTLC_HD float comp(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// sin and cos of one angle with ONE argument reduction (CUDA sincosf: same 2-ulp bound as sinf / cosf)
struct SinCos { float s, c; };
GTLC_D SinCos sincos_of(float a) { SinCos r; sincosf(a, &r.s, &r.c); return r; }

GTLC_D uint32_t f2u(float f) { return __float_as_uint(f); }
GTLC_D float u2f(uint32_t u) { return __uint_as_float(u); }

GTLC_D float safe_div(float f1, float f2) { return (f2 == 0.0f) ? 0.0f : (f1 / f2); }
GTLC_D V3 safe_div(V3 v, float f) { return (f == 0.0f) ? v3(0.0f) : (v / f); }

// common.glsl:44-47
GTLC_D float uintAsFloat(uint32_t v) { return u2f(0x3f800000u | (v >> 9)) - 1.0f; }
// common.glsl:82-90
GTLC_D uint32_t hash_theironborn(uint32_t x)
{
  x ^= x >> 16u; x *= 0x21f0aaadu; x ^= x >> 15u; x *= 0xd35a2d97u; x ^= x >> 15u;
  return x;
}
// common.glsl:93-98
GTLC_D uint32_t hash_pcg32(uint32_t& state)
{
  state = state * 747796405u + 2891336453u;
  uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}
// common.glsl:100-119 -- the stored state is the output word (the reference assigns the hash result
// back to the inout state it just advanced).
GTLC_D float rng1d_next1f(uint32_t& s) { s = hash_pcg32(s); return uintAsFloat(s); }
GTLC_D V2 rng1d_next2f(uint32_t& s) { V2 r; r.x = rng1d_next1f(s); r.y = rng1d_next1f(s); return r; }
GTLC_D V4 rng1d_next4f(uint32_t& s) { V4 r; r.x = rng1d_next1f(s); r.y = rng1d_next1f(s); r.z = rng1d_next1f(s); r.w = rng1d_next1f(s); return r; }
// common.glsl:121-124
GTLC_D uint32_t rng1d_init(uint32_t pixel_index, uint32_t sampleIndex) { return hash_theironborn(pixel_index * (sampleIndex + 1u)); }

// common.glsl:129-137
GTLC_D void orthonormal_basis(V3 n, V3& b1, V3& b2)
{
  float nsign = (n.z >= 0.0f ? 1.0f : -1.0f);
  float a = -1.0f / (nsign + n.z);
  float b = n.x * n.y * a;
  b1 = v3(1.0f + nsign * n.x * n.x * a, nsign * b, -nsign * n.x);
  b2 = v3(b, nsign + n.y * n.y * a, -n.y);
}

// common.glsl:143-162
GTLC_D V3 offset_ray_origin(V3 p, V3 n)
{
  const float origin = 1.0f / 32.0f;
  const float floatScale = 1.0f / 65536.0f;
  const float intScale = 64.0f;
  int ox = (int) (n.x * intScale), oy = (int) (n.y * intScale), oz = (int) (n.z * intScale);
  V3 ip = v3(__int_as_float(__float_as_int(p.x) + ((p.x >= 0.0f) ? ox : -ox)),
             __int_as_float(__float_as_int(p.y) + ((p.y >= 0.0f) ? oy : -oy)),
             __int_as_float(__float_as_int(p.z) + ((p.z >= 0.0f) ? oz : -oz)));
  V3 fo = n * floatScale;
  return v3((fabsf(p.x) >= origin) ? ip.x : (p.x + fo.x),
            (fabsf(p.y) >= origin) ? ip.y : (p.y + fo.y),
            (fabsf(p.z) >= origin) ? ip.z : (p.z + fo.z));
}

// GLSL unpackUnorm2x16 + common.glsl:186-196,203-207
GTLC_D V3 decode_direction(uint32_t e)
{
  float ex = (float) (e & 0xFFFFu) / 65535.0f;
  float ey = (float) (e >> 16) / 65535.0f;
  ex = ex * 2.0f - 1.0f; ey = ey * 2.0f - 1.0f;
  V3 v = v3(ex, ey, (1.0f - fabsf(ex)) - fabsf(ey));
  float t = fmaxf(-v.z, 0.0f);
  v.x += (v.x >= 0.0f) ? -t : t;
  v.y += (v.y >= 0.0f) ? -t : t;
  return normalize(v);
}

// common.glsl:210-220
GTLC_D V3 sample_hemisphere(V2 xi)
{
  float a = sqrtf(xi.x);
  float b = kPi * 2.0f * xi.y;
  const SinCos sc = sincos_of(b);
  return v3(a * sc.c, a * sc.s, sqrtf(1.0f - xi.x));
}
// common.glsl:223-230
GTLC_D V3 sample_sphere(V2 xi, V3 radius)
{
  float a = 1.0f - 2.0f * xi.x;
  float b = sqrtf(1.0f - a * a);
  float phi = 2.0f * kPi * xi.y;
  const SinCos sc = sincos_of(phi);
  return v3(b * sc.c, b * sc.s, a) * radius;
}
// common.glsl:233-252
GTLC_D V2 sample_disk(V2 xi, V2 radius)
{
  float a = 2.0f * xi.x - 1.0f;
  float b = 2.0f * xi.y - 1.0f;
  V2 r; float phi;
  if ((a * a) > (b * b)) { r.x = radius.x * a; r.y = radius.y * a; phi = (kPi / 4.0f) * (b / a); }
  else { r.x = radius.x * b; r.y = radius.y * b; phi = (kPi / 2.0f) - (kPi / 4.0f) * safe_div(a, b); }
  const SinCos sc = sincos_of(phi);
  V2 o; o.x = r.x * sc.c; o.y = r.y * sc.s;
  return o;
}
// common.glsl:254-257
GTLC_D float luminance(V3 c) { return (c.x * 0.2126f + c.y * 0.7152f) + c.z * 0.0722f; }

// GLSL unpackHalf2x16
GTLC_D V2 unpackHalf2x16(uint32_t p)
{
  V2 r;
  r.x = __half2float(__ushort_as_half((unsigned short) (p & 0xFFFFu)));
  r.y = __half2float(__ushort_as_half((unsigned short) (p >> 16)));
  return r;
}

// Row-major 3x4 affine helpers (VkTransformMatrixKHR layout, Gi.cpp:1191)
GTLC_D V3 xform_point(const float4 r0, const float4 r1, const float4 r2, V3 p)
{
  return v3(((r0.x * p.x + r0.y * p.y) + r0.z * p.z) + r0.w,
            ((r1.x * p.x + r1.y * p.y) + r1.z * p.z) + r1.w,
            ((r2.x * p.x + r2.y * p.y) + r2.z * p.z) + r2.w);
}
GTLC_D V3 xform_dir(const float4 r0, const float4 r1, const float4 r2, V3 p)
{
  return v3((r0.x * p.x + r0.y * p.y) + r0.z * p.z,
            (r1.x * p.x + r1.y * p.y) + r1.z * p.z,
            (r2.x * p.x + r2.y * p.y) + r2.z * p.z);
}
// v * mat3(M)  (GLSL row-vector product) = transpose(M3x3) * v
GTLC_D V3 xform_dir_transposed(const float4 r0, const float4 r1, const float4 r2, V3 p)
{
  return v3((r0.x * p.x + r1.x * p.y) + r2.x * p.z,
            (r0.y * p.x + r1.y * p.y) + r2.y * p.z,
            (r0.z * p.x + r1.z * p.y) + r2.z * p.z);
}

} // namespace gtlc
