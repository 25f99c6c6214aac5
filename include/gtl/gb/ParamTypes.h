// Plain parameter value types used by the gi API's material-parameter map
// (counterpart of the reference's src/gb/gtl/gb/ParamTypes.h so that hdGatling compiles unchanged).
#pragma once
#include <string>

namespace gtl
{
  struct GbVec2f { float x, y; };
  struct GbVec3f { float x, y, z; };
  struct GbVec4f { float x, y, z, w; };
  struct GbColor { float r, g, b; };
  struct GbTextureAsset { std::string absPath; bool isSrgb; };
}
