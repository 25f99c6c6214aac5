// Drives the C++ facade (include/gtl/gi/Gi.h) exactly the way hdGatling does: MaterialX source in,
// GiMeshDesc with std::vector references, giRender, giGetRenderBufferMem.
//   gi_facade_main lower            -> CPU only: MaterialX lowering checks, prints "lower ok"
//   gi_facade_main render out.bin   -> GPU: renders a 48x32 spp 4 two-material scene, writes float32 RGBA
#include <gtl/gi/Gi.h>
#include "../../include/gtlc.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace gtl { bool giLowerMtlxToClosure(const char* mtlxSrc, const char* name, GtlcMaterialDesc* out); }
using namespace gtl;

static const char* kUps =
  "<?xml version=\"1.0\"?>\n<materialx version=\"1.38\">\n"
  "  <UsdPreviewSurface name=\"SR_red\" type=\"surfaceshader\">\n"
  "    <input name=\"diffuseColor\" type=\"color3\" value=\"0.9, 0.1, 0.05\" />\n"
  "    <input name=\"roughness\" type=\"float\" value=\"0.35\" />\n"
  "    <input name=\"metallic\" type=\"float\" value=\"0.25\" />\n"
  "    <input name=\"normal\" type=\"vector3\" nodename=\"normalmap1\" />\n"
  "  </UsdPreviewSurface>\n"
  "  <surfacematerial name=\"Red\" type=\"material\"><input name=\"surfaceshader\" type=\"surfaceshader\" nodename=\"SR_red\" /></surfacematerial>\n"
  "</materialx>\n";
static const char* kOpbr =
  "<materialx version=\"1.39\"><open_pbr_surface name=\"S\" type=\"surfaceshader\">"
  "<input name=\"base_color\" type=\"color3\" value=\"0.2, 0.4, 0.8\"/><input name=\"specular_roughness\" type=\"float\" value=\"0.1\"/>"
  "<input name=\"coat_weight\" type=\"float\" value=\"1\"/><input name=\"geometry_thin_walled\" type=\"boolean\" value=\"true\"/>"
  "</open_pbr_surface></materialx>";
static const char* kEmissive =
  "<materialx version=\"1.38\"><UsdPreviewSurface name=\"L\" type=\"surfaceshader\"><input name=\"diffuseColor\" type=\"color3\" value=\"0,0,0\"/>"
  "<input name=\"emissiveColor\" type=\"color3\" value=\"6, 5, 4\"/></UsdPreviewSurface></materialx>";

static int lowerTests()
{
  GtlcMaterialDesc d;
  if (!giLowerMtlxToClosure(kUps, "Red", &d) || d.closure != GTLC_CLOSURE_USD_PREVIEW_SURFACE) return 1;
  if (d.params[GTLC_UPS_DIFFUSE_R] != 0.9f || d.params[GTLC_UPS_DIFFUSE_B] != 0.05f || d.params[GTLC_UPS_ROUGHNESS] != 0.35f || d.params[GTLC_UPS_METALLIC] != 0.25f) return 2;
  if (d.params[GTLC_UPS_IOR] != 1.5f || d.params[GTLC_UPS_OPACITY] != 1.0f) return 3;   // untouched inputs keep the spec defaults
  if (!giLowerMtlxToClosure(kOpbr, "S", &d) || d.closure != GTLC_CLOSURE_OPENPBR) return 4;
  if (d.params[GTLC_OPBR_BASE_COLOR_B] != 0.8f || d.params[GTLC_OPBR_SPECULAR_ROUGHNESS] != 0.1f || d.params[GTLC_OPBR_COAT_WEIGHT] != 1.0f || d.params[GTLC_OPBR_GEOMETRY_THIN_WALLED] != 1.0f) return 5;
  if (d.params[GTLC_OPBR_COAT_IOR] != 1.6f) return 6;
  if (giLowerMtlxToClosure("<materialx><standard_surface name=\"x\"/></materialx>", "x", &d)) return 7;   // unsupported model -> false
  printf("lower ok\n");
  return 0;
}

static GiVertex vert(float x, float y, float z, float nx, float ny, float nz)
{
  GiVertex v; memset(&v, 0, sizeof(v));
  v.pos[0] = x; v.pos[1] = y; v.pos[2] = z; v.norm[0] = nx; v.norm[1] = ny; v.norm[2] = nz;
  // Duff basis tangent (hdGatling fallback, mesh.cpp:225-233)
  float sign = nz >= 0.0f ? 1.0f : -1.0f, a = -1.0f / (sign + nz), b = nx * ny * a;
  v.tangent[0] = 1.0f + sign * nx * nx * a; v.tangent[1] = sign * b; v.tangent[2] = -sign * nx; v.bitangentSign = 1.0f;
  return v;
}

static int renderScene(const char* outPath)
{
  std::vector<std::string> noPaths;
  GiInitParams ip{ "", "", noPaths, nullptr, "" };
  if (giInitialize(ip) != GiStatus::Ok) { fprintf(stderr, "giInitialize failed: %s\n", giGetLastError()); return 10; }
  GiScene* scene = giCreateScene();
  GiMaterial* red = giCreateMaterialFromMtlxStr(scene, "Red", kUps);
  GiMaterial* light = giCreateMaterialFromMtlxStr(scene, "Light", kEmissive);
  if (!red || !light) return 11;
  // floor quad (two triangles) + emissive quad above it
  std::vector<GiVertex> fv = { vert(-1, -1, 0, 0, 0, 1), vert(1, -1, 0, 0, 0, 1), vert(1, 1, 0, 0, 0, 1), vert(-1, 1, 0, 0, 0, 1) };
  std::vector<GiFace> ff = { { { 0, 1, 2 } }, { { 0, 2, 3 } } };
  std::vector<int> faceIds = { 0, 0 };
  std::vector<GiPrimvarData> noPrimvars;
  GiMeshDesc fd{ 2, ff, faceIds, 1, true, false, "floor", 0, noPrimvars, 4, fv };
  GiMesh* floorMesh = giCreateMesh(scene, fd);
  std::vector<GiVertex> lv = { vert(-0.3f, -0.3f, 1.5f, 0, 0, -1), vert(-0.3f, 0.3f, 1.5f, 0, 0, -1), vert(0.3f, 0.3f, 1.5f, 0, 0, -1), vert(0.3f, -0.3f, 1.5f, 0, 0, -1) };
  GiMeshDesc ld{ 2, ff, faceIds, 2, true, false, "light", 0, noPrimvars, 4, lv };
  GiMesh* lightMesh = giCreateMesh(scene, ld);
  if (!floorMesh || !lightMesh) return 12;
  giSetMeshMaterial(floorMesh, red); giSetMeshMaterial(lightMesh, light);
  const float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
  const float shifted[2][4][4] = { { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } }, { { 0.5f, 0, 0, 0 }, { 0, 0.5f, 0, 0 }, { 0, 0, 0.5f, 0 }, { 1.2f, 0.4f, 0.6f, 1 } } };
  giSetMeshTransform(floorMesh, ident);
  giSetMeshInstanceTransforms(lightMesh, 2, shifted);
  GiSphereLight* sl = giCreateSphereLight(scene);
  float pos[3] = { -0.8f, 0.2f, 0.7f }, em[3] = { 3, 3, 5 };
  giSetSphereLightPosition(sl, pos); giSetSphereLightBaseEmission(sl, em); giSetSphereLightRadius(sl, 0.1f, 0.1f, 0.1f);

  const uint32_t W = 48, H = 32;
  GiRenderBuffer* rb = giCreateRenderBuffer(W, H, GiRenderBufferFormat::Float32Vec4);
  if (!rb) return 13;
  GiRenderParams rp;
  GiAovBinding bind; bind.aovId = GiAovId::Color; bind.renderBuffer = rb;
  const float clear[4] = { 0.25f, 0.5f, 1.0f, 1.0f };
  memcpy(bind.clearValue, clear, 16);
  rp.aovBindings.push_back(bind);
  rp.camera = GiCameraDesc{ { 0, -4, 1.5f }, { 0, 0.9701425f, -0.2425356f }, { 0, 0.2425356f, 0.9701425f }, 0.7f, 0, 0, 0.05f, 0.1f, 100.0f, 0 };
  rp.domeLight = nullptr;
  rp.renderSettings = GiRenderSettings{ false, false, true, true, 0, true, 1.0f, 5, 10.0f, 7, 0, 1.0f, true, true, 3, 0.95f, 4, 0 };
  rp.scene = scene;
  if (giRender(rp) != GiStatus::Ok) { fprintf(stderr, "giRender failed: %s\n", giGetLastError()); return 14; }
  const float* px = (const float*) giGetRenderBufferMem(rb);
  FILE* f = fopen(outPath, "wb");
  if (!f) return 15;
  fwrite(px, sizeof(float), (size_t) W * H * 4, f);
  fclose(f);
  // an unsupported setting must fail loudly, not render something else
  rp.renderSettings.mediumStackSize = 1;
  if (giRender(rp) != GiStatus::Error) return 16;
  giDestroyRenderBuffer(rb); giDestroySphereLight(scene, sl); giDestroyMesh(floorMesh); giDestroyMesh(lightMesh);
  giDestroyMaterial(red); giDestroyMaterial(light); giDestroyScene(scene); giTerminate();
  printf("render ok\n");
  return 0;
}

int main(int argc, char** argv)
{
  if (argc >= 2 && !strcmp(argv[1], "lower")) return lowerTests();
  if (argc >= 3 && !strcmp(argv[1], "render")) return renderScene(argv[2]);
  fprintf(stderr, "usage: %s lower | render <out.bin>\n", argv[0]);
  return 2;
}
