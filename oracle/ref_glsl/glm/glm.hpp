// TEST INFRASTRUCTURE.  Stand-in for <glm/glm.hpp> (not in this image) so that the reference's shared C++/GLSL header
// src/gi/shaders/interface/rp_main.h compiles in its C++ branch: the glm types it names are the shim's types (same size and layout).
#pragma once
#include "../glsl_shim.h"
namespace glm {
  using vec2 = ::vec2; using vec3 = ::vec3; using vec4 = ::vec4;
  using ivec2 = ::ivec2; using ivec3 = ::ivec3; using ivec4 = ::ivec4;
  using uvec2 = ::uvec2; using uvec3 = ::uvec3; using uvec4 = ::uvec4;
  using mat3 = ::mat3; using mat3x4 = ::mat3x4;
}
