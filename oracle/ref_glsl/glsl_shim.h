// TEST INFRASTRUCTURE.  A minimal GLSL-in-C++ layer: just enough types and built-ins for the PURE functions of the reference's
// shader sources (src/gi/shaders/common.glsl, colormap.glsl, rp_main_payload.glsl, fragments of rp_main.rgen / .chit / .miss and
// mdl_interface.glsl) to compile unmodified -- apart from the mechanical rewrites of oracle/ref_glsl/build_ref.py (parameter
// qualifiers, swizzles) -- with g++ -fsingle-precision-constant -ffp-contract=off, from where they lie under /root/reference, into
// oracle/_ref/libglslref.so.  tests/test_reference_glsl_cpu.py then demands that the oracle's restatements return the same bits.
// The built-ins are defined exactly as the oracle defines them (orc_math.h: dot = (x x + y y) + z z, normalize = v * (1 / sqrt(dot)),
// packUnorm2x16 = round(clamp * 65535), libm transcendentals): what the comparison pins is the reference's own code --
// constants, operation order, branches --, not the GPU driver's arithmetic.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef uint32_t uint;
struct vec2; struct vec3; struct vec4;
struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };
struct vec2 {
  union { float x, r; }; union { float y, g; };
  vec2() : x(0), y(0) {}
  explicit vec2(float s) : x(s), y(s) {}
  vec2(float a, float b) : x(a), y(b) {}
  vec2 xy() const { return *this; }
  vec2 yx() const { return vec2(y, x); }
  float& operator[](int i) { return (&x)[i]; }
};
struct vec3 {
  union { float x, r; }; union { float y, g; }; union { float z, b; };
  vec3() : x(0), y(0), z(0) {}
  explicit vec3(float s) : x(s), y(s), z(s) {}
  vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
  vec3(vec2 a, float c) : x(a.x), y(a.y), z(c) {}
  explicit vec3(bvec3 s) : x(s.x ? 1.0f : 0.0f), y(s.y ? 1.0f : 0.0f), z(s.z ? 1.0f : 0.0f) {}
  vec2 xy() const { return vec2(x, y); }
  vec2 yx() const { return vec2(y, x); }
  vec3 xyz() const { return *this; }
  vec3 rgb() const { return *this; }
  float& operator[](int i) { return (&x)[i]; }                 // (glm-style component access of the host code)
};
struct vec4 {
  union { float x, r; }; union { float y, g; }; union { float z, b; }; union { float w, a; };
  vec4() : x(0), y(0), z(0), w(0) {}
  explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  vec4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
  vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec2 xy() const { return vec2(x, y); }
  vec2 zw() const { return vec2(z, w); }
  vec3 xyz() const { return vec3(x, y, z); }
  vec3 rgb() const { return vec3(x, y, z); }
};
struct ivec2 { int x, y; ivec2() : x(0), y(0) {} ivec2(int a, int b) : x(a), y(b) {} };
struct ivec3 {
  int x, y, z;
  ivec3() : x(0), y(0), z(0) {}
  ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
  explicit ivec3(vec3 v) : x((int) v.x), y((int) v.y), z((int) v.z) {}       // GLSL float -> int conversion truncates
};
struct ivec4 { int x, y, z, w; };
struct uvec2 { uint x, y; uvec2() : x(0), y(0) {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct uvec3 {
  uint x, y, z;
  uvec3() : x(0), y(0), z(0) {}
  uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
  explicit uvec3(uint s) : x(s), y(s), z(s) {}
  explicit uvec3(int s) : x((uint) s), y((uint) s), z((uint) s) {}
  uvec2 xy() const { return uvec2(x, y); }
  uint operator[](int i) const { return (&x)[i]; }
};
static inline uvec3 operator*(uvec3 a, uint s) { return uvec3(a.x * s, a.y * s, a.z * s); }
static inline uvec3 operator*(uvec3 a, int s) { return uvec3(a.x * (uint) s, a.y * (uint) s, a.z * (uint) s); }
static inline uvec3 operator+(uvec3 a, int s) { return uvec3(a.x + (uint) s, a.y + (uint) s, a.z + (uint) s); }
struct uvec4 {
  uint x, y, z, w;
  uvec4() : x(0), y(0), z(0), w(0) {}
  uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
  uvec4& operator>>=(int s) { x >>= s; y >>= s; z >>= s; w >>= s; return *this; }
  uvec4& operator|=(uint m) { x |= m; y |= m; z |= m; w |= m; return *this; }
};
struct mat4 { vec4 c[4]; };
struct mat3x4 { vec4 c[3]; };
// gl_ObjectToWorldEXT / gl_WorldToObjectEXT: kept as the row-major 3x4 the host uploads (VkTransformMatrixKHR).  The products the shaders
// form with them are built-ins, evaluated as the oracle evaluates them (orc_scene.h xform_point / xform_dir / xform_dir_transposed).
struct mat4x3 { float m[12]; };
struct mat3 { float m[12]; mat3() {} explicit mat3(const mat4x3& a) { for (int i = 0; i < 12; i++) m[i] = a.m[i]; } };

// ---- arithmetic (component-wise, one rounding per operation) ----
#define GLSL_VEC_OPS(V, N, ...)                                                                                   \
  static inline V operator+(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] + (&b.x)[i]; return r; } \
  static inline V operator-(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] - (&b.x)[i]; return r; } \
  static inline V operator*(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] * (&b.x)[i]; return r; } \
  static inline V operator/(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] / (&b.x)[i]; return r; } \
  static inline V operator+(V a, float s) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] + s; return r; }   \
  static inline V operator-(V a, float s) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] - s; return r; }   \
  static inline V operator*(V a, float s) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] * s; return r; }   \
  static inline V operator/(V a, float s) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = (&a.x)[i] / s; return r; }   \
  static inline V operator+(float s, V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = s + (&a.x)[i]; return r; }   \
  static inline V operator-(float s, V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = s - (&a.x)[i]; return r; }   \
  static inline V operator*(float s, V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = s * (&a.x)[i]; return r; }   \
  static inline V operator-(V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = -(&a.x)[i]; return r; }                \
  static inline V& operator+=(V& a, V b) { a = a + b; return a; }                                                  \
  static inline V& operator-=(V& a, V b) { a = a - b; return a; }                                                  \
  static inline V& operator*=(V& a, V b) { a = a * b; return a; }                                                  \
  static inline V& operator*=(V& a, float s) { a = a * s; return a; }                                              \
  static inline V& operator/=(V& a, float s) { a = a / s; return a; }                                              \
  static inline V abs(V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = fabsf((&a.x)[i]); return r; }            \
  static inline V max(V a, float s) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = fmaxf((&a.x)[i], s); return r; }            \
  static inline V max(V a, V b) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = fmaxf((&a.x)[i], (&b.x)[i]); return r; }        \
  static inline V exp(V a) { V r; for (int i = 0; i < N; i++) (&r.x)[i] = expf((&a.x)[i]); return r; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
static inline ivec3 operator+(ivec3 a, ivec3 b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline ivec3 operator-(ivec3 a) { return ivec3(-a.x, -a.y, -a.z); }

// ---- built-ins, as orc_math.h defines them (abs / sqrt / cos / sin / log / exp2 / acos / floor of a float: <math.h>'s float overloads = libm's *f) ----
static inline float atan(float y, float x) { return atan2f(y, x); }
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline uint max(uint a, uint b) { return a > b ? a : b; }
static inline uint min(uint a, uint b) { return a < b ? a : b; }
static inline uint min(uint a, int b) { return min(a, (uint) b); }
static inline float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length(vec3 a) { return sqrtf(dot(a, a)); }
static inline vec3 normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
static inline vec3 reflect(vec3 i, vec3 n) { return i - n * (2.0f * dot(n, i)); }                  // I - 2 dot(N, I) N
static inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
static inline vec3 mix(vec3 a, vec3 b, bvec3 s) { return vec3(s.x ? b.x : a.x, s.y ? b.y : a.y, s.z ? b.z : a.z); }
static inline ivec3 mix(ivec3 a, ivec3 b, bvec3 s) { return ivec3(s.x ? b.x : a.x, s.y ? b.y : a.y, s.z ? b.z : a.z); }
static inline bvec3 greaterThanEqual(vec3 a, vec3 b) { bvec3 r = { a.x >= b.x, a.y >= b.y, a.z >= b.z }; return r; }
static inline bvec3 greaterThan(vec3 a, vec3 b) { bvec3 r = { a.x > b.x, a.y > b.y, a.z > b.z }; return r; }
static inline bool any(bvec3 a) { return a.x || a.y || a.z; }
static inline bvec3 equal(vec3 a, vec3 b) { bvec3 r = { a.x == b.x, a.y == b.y, a.z == b.z }; return r; }
static inline bvec2 notEqual(vec2 a, vec2 b) { bvec2 r = { a.x != b.x, a.y != b.y }; return r; }
static inline bool all(bvec2 a) { return a.x && a.y; }
static inline uint floatBitsToUint(float f) { uint u; memcpy(&u, &f, 4); return u; }
static inline float uintBitsToFloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
static inline int floatBitsToInt(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float intBitsToFloat(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline ivec3 floatBitsToInt(vec3 v) { return ivec3(floatBitsToInt(v.x), floatBitsToInt(v.y), floatBitsToInt(v.z)); }
static inline vec3 intBitsToFloat(ivec3 v) { return vec3(intBitsToFloat(v.x), intBitsToFloat(v.y), intBitsToFloat(v.z)); }
static inline vec4 uintBitsToFloat(uvec4 u) { return vec4(uintBitsToFloat(u.x), uintBitsToFloat(u.y), uintBitsToFloat(u.z), uintBitsToFloat(u.w)); }
static inline uint packUnorm2x16(vec2 v)
{
  const uint a = (uint) roundf(clamp(v.x, 0.0f, 1.0f) * 65535.0f), b = (uint) roundf(clamp(v.y, 0.0f, 1.0f) * 65535.0f);
  return a | (b << 16);
}
static inline vec2 unpackUnorm2x16(uint e) { return vec2((float) (e & 0xFFFFu) / 65535.0f, (float) (e >> 16) / 65535.0f); }
// M * vec4(p, 1) (a point), M * vec4(d, 0) (a direction), v * mat3(M) (= transpose(M3x3) * v: normals through the inverse matrix)
static inline vec3 operator*(const mat4x3& a, vec4 p)
{
  const float* m = a.m;
  if (p.w == 0.0f) return vec3((m[0] * p.x + m[1] * p.y) + m[2] * p.z, (m[4] * p.x + m[5] * p.y) + m[6] * p.z, (m[8] * p.x + m[9] * p.y) + m[10] * p.z);
  return vec3(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7], ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
}
static inline vec3 operator*(vec3 p, const mat3& a)
{
  const float* m = a.m;
  return vec3((m[0] * p.x + m[4] * p.y) + m[8] * p.z, (m[1] * p.x + m[5] * p.y) + m[9] * p.z, (m[2] * p.x + m[6] * p.y) + m[10] * p.z);
}
// binary16 -> float is exact; written independently of the oracle's decoder
static inline float glsl_half_to_float(uint h)
{
  const uint e = (h >> 10) & 31u, mnt = h & 1023u;
  float v;
  if (e == 0u) v = ldexpf((float) mnt, -24);
  else if (e == 31u) v = mnt ? NAN : INFINITY;
  else v = ldexpf((float) (1024u + mnt), (int) e - 25);
  return (h & 0x8000u) ? -v : v;
}
static inline vec2 unpackHalf2x16(uint p) { return vec2(glsl_half_to_float(p & 0xFFFFu), glsl_half_to_float(p >> 16)); }
