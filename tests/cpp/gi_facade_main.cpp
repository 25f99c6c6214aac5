// Drives the C++ facade (include/gtl/gi/Gi.h) exactly the way hdGatling does: MaterialX source in,
// GiMeshDesc with std::vector references, giRender, giGetRenderBufferMem.
//   gi_facade_main lower            -> CPU only: MaterialX lowering checks, prints "lower ok"
//   gi_facade_main objects          -> CPU only: object lifecycle through the facade (meshes with primvars, instancer primvars,
//                                      materials with primvar readers, lights), host BVH through the C-ABI; prints "objects ok"
//   gi_facade_main render out.bin   -> GPU: renders a 48x32 spp 4 two-material scene, writes float32 RGBA
//   gi_facade_main render out.bin checker.ppm -> same, the floor's diffuseColor read from the image (UsdUVTexture + UsdTransform2d)
#include <gtl/gi/Gi.h>
#include "../../include/gtlc.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace gtl {
  bool giLowerMtlxToClosure(const char* mtlxSrc, const char* name, GtlcMaterialDesc* out);
  bool giLowerMtlxToClosureWithTextures(const char* mtlxSrc, const char* name, GtlcMaterialDesc* out, std::vector<GtlcTextureBinding>* bindings, std::vector<std::string>* files);
}
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

// what hdMtlx emits for a UsdPreviewSurface network with UsdUVTexture / UsdTransform2d / UsdPrimvarReader_float2 prims
static std::string texturedUps(const char* file)
{
  return std::string("<materialx version=\"1.38\">\n"
    "  <UsdPrimvarReader name=\"stReader\" type=\"vector2\"><input name=\"varname\" type=\"string\" value=\"st\" /></UsdPrimvarReader>\n"
    "  <UsdTransform2d name=\"xf\" type=\"vector2\"><input name=\"in\" type=\"vector2\" nodename=\"stReader\" />"
    "<input name=\"rotation\" type=\"float\" value=\"30\" /><input name=\"scale\" type=\"vector2\" value=\"3, 2\" /><input name=\"translation\" type=\"vector2\" value=\"0.13, -0.4\" /></UsdTransform2d>\n"
    "  <UsdUVTexture name=\"diffuseTex\" type=\"multioutput\"><input name=\"file\" type=\"filename\" value=\"") + file + "\" />"
    "<input name=\"st\" type=\"vector2\" nodename=\"xf\" /><input name=\"wrapS\" type=\"string\" value=\"mirror\" /><input name=\"wrapT\" type=\"string\" value=\"black\" />"
    "<input name=\"fallback\" type=\"color4\" value=\"0.05, 0.6, 0.1, 1\" /></UsdUVTexture>\n"
    "  <UsdUVTexture name=\"dataTex\" type=\"multioutput\"><input name=\"file\" type=\"filename\" value=\"data.pfm\" /><input name=\"sourceColorSpace\" type=\"string\" value=\"raw\" />"
    "<input name=\"scale\" type=\"color4\" value=\"1, 0.5, 1, 1\" /><input name=\"bias\" type=\"color4\" value=\"0, 0.1, 0, 0\" /></UsdUVTexture>\n"
    "  <UsdUVTexture name=\"normalTex\" type=\"multioutput\"><input name=\"file\" type=\"filename\" value=\"n.ppm\" /><input name=\"scale\" type=\"color4\" value=\"2, 2, 2, 1\" />"
    "<input name=\"bias\" type=\"color4\" value=\"-1, -1, -1, 0\" /></UsdUVTexture>\n"
    "  <UsdPreviewSurface name=\"SR\" type=\"surfaceshader\">\n"
    "    <input name=\"diffuseColor\" type=\"color3\" nodename=\"diffuseTex\" output=\"rgb\" />\n"
    "    <input name=\"roughness\" type=\"float\" nodename=\"dataTex\" output=\"g\" />\n"
    "    <input name=\"normal\" type=\"vector3\" nodename=\"normalTex\" output=\"rgb\" />\n"
    "    <input name=\"metallic\" type=\"float\" value=\"0.25\" />\n"
    "    <input name=\"occlusion\" type=\"float\" nodename=\"dataTex\" output=\"r\" />\n"
    "  </UsdPreviewSurface>\n</materialx>\n";
}
static const char* kOpbrNormalMap =
  "<materialx version=\"1.39\"><image name=\"img\" type=\"vector3\"><input name=\"file\" type=\"filename\" value=\"nrm.ppm\" colorspace=\"lin_rec709\"/>"
  "<input name=\"uaddressmode\" type=\"string\" value=\"clamp\"/><input name=\"vaddressmode\" type=\"string\" value=\"constant\"/><input name=\"default\" type=\"vector3\" value=\"0.5, 0.5, 1\"/></image>"
  "<normalmap name=\"nm\" type=\"vector3\"><input name=\"in\" type=\"vector3\" nodename=\"img\"/></normalmap>"
  "<image name=\"col\" type=\"color3\"><input name=\"file\" type=\"filename\" value=\"c.ppm\" colorspace=\"srgb_texture\"/></image>"
  "<open_pbr_surface name=\"S\" type=\"surfaceshader\"><input name=\"geometry_normal\" type=\"vector3\" nodename=\"nm\"/>"
  "<input name=\"base_color\" type=\"color3\" nodename=\"col\"/><input name=\"specular_roughness\" type=\"float\" nodename=\"noise_graph\"/></open_pbr_surface></materialx>";

// arithmetic with constants between the lookup and the input: roughness = (tex.g * 0.5 + 0.1) * 2, diffuseColor = (1 - tex.rgb) * (0.5, 1, 2),
// metallic = 0.9 - tex.b / 4; clearcoat = tex.r * tex.g is not affine in one lookup: two bindings, the second one multiplies
static const char* kArithUps =
  "<materialx version=\"1.38\">"
  "<image name=\"img\" type=\"color3\"><input name=\"file\" type=\"filename\" value=\"a.ppm\" colorspace=\"lin_rec709\"/></image>"
  "<multiply name=\"m1\" type=\"float\"><input name=\"in1\" type=\"float\" nodename=\"img\" output=\"g\"/><input name=\"in2\" type=\"float\" value=\"0.5\"/></multiply>"
  "<add name=\"a1\" type=\"float\"><input name=\"in1\" type=\"float\" value=\"0.1\"/><input name=\"in2\" type=\"float\" nodename=\"m1\"/></add>"
  "<multiply name=\"m2\" type=\"float\"><input name=\"in1\" type=\"float\" nodename=\"a1\"/><input name=\"in2\" type=\"float\" value=\"2\"/></multiply>"
  "<invert name=\"inv\" type=\"color3\"><input name=\"in\" type=\"color3\" nodename=\"img\"/></invert>"
  "<multiply name=\"tint\" type=\"color3\"><input name=\"in1\" type=\"color3\" nodename=\"inv\"/><input name=\"in2\" type=\"color3\" value=\"0.5, 1, 2\"/></multiply>"
  "<divide name=\"d1\" type=\"float\"><input name=\"in1\" type=\"float\" nodename=\"img\" output=\"b\"/><input name=\"in2\" type=\"float\" value=\"4\"/></divide>"
  "<subtract name=\"s1\" type=\"float\"><input name=\"in1\" type=\"float\" value=\"0.9\"/><input name=\"in2\" type=\"float\" nodename=\"d1\"/></subtract>"
  "<multiply name=\"tt\" type=\"float\"><input name=\"in1\" type=\"float\" nodename=\"img\" output=\"r\"/><input name=\"in2\" type=\"float\" nodename=\"img\" output=\"g\"/></multiply>"
  "<UsdPreviewSurface name=\"S\" type=\"surfaceshader\"><input name=\"roughness\" type=\"float\" nodename=\"m2\"/><input name=\"diffuseColor\" type=\"color3\" nodename=\"tint\"/>"
  "<input name=\"metallic\" type=\"float\" nodename=\"s1\"/><input name=\"clearcoat\" type=\"float\" nodename=\"tt\"/></UsdPreviewSurface></materialx>";

// inputs computed from two or three lookups: mix(image, image, image mask), mix with a constant mask (folds into scale / bias),
// (image + image) * 0.5, image - image, mix(image, constant, image mask); a mix nested in a multiply is not lowered
static const char* kLayeredOpbr =
  "<materialx version=\"1.39\">"
  "<image name=\"A\" type=\"color3\"><input name=\"file\" type=\"filename\" value=\"a.ppm\"/></image>"
  "<image name=\"B\" type=\"color3\"><input name=\"file\" type=\"filename\" value=\"b.ppm\"/></image>"
  "<image name=\"M\" type=\"color3\"><input name=\"file\" type=\"filename\" value=\"m.pfm\"/></image>"
  "<mix name=\"mx\" type=\"color3\"><input name=\"bg\" type=\"color3\" nodename=\"A\"/><input name=\"fg\" type=\"color3\" nodename=\"B\"/><input name=\"mix\" type=\"float\" nodename=\"M\" output=\"r\"/></mix>"
  "<mix name=\"mr\" type=\"float\"><input name=\"bg\" type=\"float\" value=\"0.2\"/><input name=\"fg\" type=\"float\" nodename=\"A\" output=\"g\"/><input name=\"mix\" type=\"float\" value=\"0.25\"/></mix>"
  "<add name=\"sum\" type=\"color3\"><input name=\"in1\" type=\"color3\" nodename=\"A\"/><input name=\"in2\" type=\"color3\" nodename=\"B\"/></add>"
  "<multiply name=\"half\" type=\"color3\"><input name=\"in1\" type=\"color3\" nodename=\"sum\"/><input name=\"in2\" type=\"float\" value=\"0.5\"/></multiply>"
  "<subtract name=\"dif\" type=\"color3\"><input name=\"in1\" type=\"color3\" nodename=\"A\"/><input name=\"in2\" type=\"color3\" nodename=\"B\"/></subtract>"
  "<mix name=\"mc\" type=\"color3\"><input name=\"bg\" type=\"color3\" nodename=\"A\"/><input name=\"fg\" type=\"color3\" value=\"1, 0, 0.5\"/><input name=\"mix\" type=\"float\" nodename=\"M\" output=\"g\"/></mix>"
  "<multiply name=\"nest\" type=\"float\"><input name=\"in1\" type=\"float\" nodename=\"mr\"/><input name=\"in2\" type=\"float\" nodename=\"B\" output=\"r\"/></multiply>"
  "<open_pbr_surface name=\"S\" type=\"surfaceshader\"><input name=\"base_color\" type=\"color3\" nodename=\"mx\"/><input name=\"specular_roughness\" type=\"float\" nodename=\"mr\"/>"
  "<input name=\"coat_color\" type=\"color3\" nodename=\"half\"/><input name=\"emission_color\" type=\"color3\" nodename=\"dif\"/><input name=\"fuzz_color\" type=\"color3\" nodename=\"mc\"/>"
  "<input name=\"base_metalness\" type=\"float\" nodename=\"nest\"/></open_pbr_surface></materialx>";

static const char* kPrimvarUps =
  "<materialx version=\"1.38\"><UsdPrimvarReader name=\"dc\" type=\"color3\"><input name=\"varname\" type=\"string\" value=\"displayColor\"/>"
  "<input name=\"fallback\" type=\"color3\" value=\"0.1, 0.2, 0.9\"/></UsdPrimvarReader>"
  "<geompropvalue name=\"r\" type=\"float\"><input name=\"geomprop\" type=\"string\" value=\"rough\"/><input name=\"default\" type=\"float\" value=\"0.9\"/></geompropvalue>"
  "<UsdPreviewSurface name=\"S\" type=\"surfaceshader\"><input name=\"diffuseColor\" type=\"color3\" nodename=\"dc\"/>"
  "<input name=\"roughness\" type=\"float\" nodename=\"r\"/></UsdPreviewSurface></materialx>";

static int lowerTextureTests()
{
  {
    GtlcMaterialDesc pd; std::vector<GtlcTextureBinding> pb; std::vector<std::string> pf;
    if (!giLowerMtlxToClosureWithTextures(kPrimvarUps, "P", &pd, &pb, &pf) || !pb.empty() || pd.primvarCount != 2) return 40;
    if (pf.size() != 2 || pf[0] != "primvar:displayColor:" + std::to_string((int) GTLC_UPS_DIFFUSE_R) + ":3" || pf[1] != "primvar:rough:" + std::to_string((int) GTLC_UPS_ROUGHNESS) + ":1") return 41;
  }
  {
    GtlcMaterialDesc ad; std::vector<GtlcTextureBinding> ab; std::vector<std::string> af;
    if (!giLowerMtlxToClosureWithTextures(kArithUps, "A", &ad, &ab, &af) || ab.size() != 5) return 50;      // three plain inputs + clearcoat (texture x texture) as SET, MULTIPLY
    const GtlcTextureBinding *ro = nullptr, *dc = nullptr, *me = nullptr;
    const GtlcTextureBinding *cc0 = nullptr, *cc1 = nullptr;
    for (const GtlcTextureBinding& x : ab) {
      if (x.paramIndex == GTLC_UPS_ROUGHNESS) ro = &x; if (x.paramIndex == GTLC_UPS_DIFFUSE_R) dc = &x; if (x.paramIndex == GTLC_UPS_METALLIC) me = &x;
      if (x.paramIndex == GTLC_UPS_CLEARCOAT) { if (!cc0) cc0 = &x; else cc1 = &x; }
    }
    if (!ro || !dc || !me || !cc0 || !cc1) return 51;
    if (cc0->combine != 0 || cc0->channels[0] != 0 || cc1->combine != GTLC_COMBINE(GTLC_COMBINE_MULTIPLY, 0) || cc1->channels[0] != 1 || cc1->scale[1] != 1.0f || cc1->bias[1] != 0.0f) return 56;
    if (ro->channels[0] != 1 || ro->scale[1] != 1.0f || ro->bias[1] != 0.2f) return 52;                     // (g * 0.5 + 0.1) * 2 = g * 1 + 0.2
    if (dc->scale[0] != -0.5f || dc->bias[0] != 0.5f || dc->scale[1] != -1.0f || dc->bias[1] != 1.0f || dc->scale[2] != -2.0f || dc->bias[2] != 2.0f) return 53;   // (1 - c) * k
    if (me->channels[0] != 2 || me->scale[2] != -0.25f || me->bias[2] != 0.9f) return 54;                    // 0.9 - b / 4
    if (ad.params[GTLC_UPS_CLEARCOAT] != 0.0f) return 55;
  }
  {
    GtlcMaterialDesc ld; std::vector<GtlcTextureBinding> lb; std::vector<std::string> lf;
    if (!giLowerMtlxToClosureWithTextures(kLayeredOpbr, "L", &ld, &lb, &lf) || ld.closure != GTLC_CLOSURE_OPENPBR) return 60;
    // base_color: mask -> scratch0, A, B mixed in; roughness: one binding; coat: A/2 then + B/2; emission: A then - B; fuzz: mask -> scratch1, A, constant mixed in
    if (lb.size() != 3 + 1 + 2 + 2 + 3 || lf.size() != lb.size()) return 61;
    const int S0 = GTLC_BIND_SCRATCH, S1 = GTLC_BIND_SCRATCH + 1;
    if (lb[0].paramIndex != S0 || lb[0].componentCount != 1 || lb[0].channels[0] != 0 || lb[0].combine != 0 || lf[0] != "m.pfm") return 62;
    if (lb[1].paramIndex != GTLC_OPBR_BASE_COLOR_R || lb[1].combine != 0 || lf[1] != "a.ppm" || lb[2].paramIndex != GTLC_OPBR_BASE_COLOR_R || lb[2].combine != GTLC_COMBINE(GTLC_COMBINE_MIX, S0) || lf[2] != "b.ppm") return 63;
    if (lb[3].paramIndex != GTLC_OPBR_SPECULAR_ROUGHNESS || lb[3].combine != 0 || lb[3].channels[0] != 1 || lb[3].scale[1] != 0.25f || lb[3].bias[1] != 0.2f * 0.75f) return 64;   // 0.2 (1 - 0.25) + g 0.25
    if (lb[4].paramIndex != GTLC_OPBR_COAT_COLOR_R || lb[4].combine != 0 || lb[4].scale[0] != 0.5f || lb[5].combine != GTLC_COMBINE(GTLC_COMBINE_ADD, 0) || lb[5].scale[2] != 0.5f || lf[5] != "b.ppm") return 65;
    if (lb[6].paramIndex != GTLC_OPBR_EMISSION_COLOR_R || lb[6].scale[1] != 1.0f || lb[7].combine != GTLC_COMBINE(GTLC_COMBINE_ADD, 0) || lb[7].scale[1] != -1.0f) return 66;
    if (lb[8].paramIndex != S1 || lb[8].channels[0] != 1 || lb[9].paramIndex != GTLC_OPBR_FUZZ_COLOR_R || lb[9].combine != 0) return 67;
    if (lb[10].combine != (GTLC_COMBINE(GTLC_COMBINE_MIX, S1) | GTLC_COMBINE_CONSTANT) || lb[10].bias[0] != 1.0f || lb[10].bias[1] != 0.0f || lb[10].bias[2] != 0.5f || !lf[10].empty() || lb[10].texture != 0) return 68;
    if (ld.params[GTLC_OPBR_BASE_METALNESS] != 0.0f) return 69;                            // the nested graph is not lowered: default constant (with a warning)
  }
  GtlcMaterialDesc d; std::vector<GtlcTextureBinding> b; std::vector<std::string> files;
  if (!giLowerMtlxToClosureWithTextures(texturedUps("checker.ppm").c_str(), "T", &d, &b, &files)) return 20;
  if (b.size() != 3 || files.size() != 3 || d.params[GTLC_UPS_METALLIC] != 0.25f || d.params[GTLC_UPS_ROUGHNESS] != 0.5f) return 21;   // connected inputs keep the default constant
  const GtlcTextureBinding& dc = b[0];
  if (files[0] != "checker.ppm" || dc.paramIndex != GTLC_UPS_DIFFUSE_R || dc.componentCount != 3 || dc.channels[0] != 0 || dc.channels[1] != 1 || dc.channels[2] != 2) return 22;
  if (dc.wrapU != GTLC_WRAP_MIRRORED_REPEAT || dc.wrapV != GTLC_WRAP_CLIP || !dc.srgb || dc.fallback[1] != 0.6f) return 23;     // "auto" + diffuseColor -> srgb_texture
  if (dc.uvRotation != 30.0f || dc.uvScale[0] != 3.0f || dc.uvScale[1] != 2.0f || dc.uvTranslation[0] != 0.13f || dc.uvTranslation[1] != -0.4f) return 24;
  const GtlcTextureBinding& ro = b[1];
  if (files[1] != "data.pfm" || ro.paramIndex != GTLC_UPS_ROUGHNESS || ro.componentCount != 1 || ro.channels[0] != 1 || ro.srgb || ro.scale[1] != 0.5f || ro.bias[1] != 0.1f) return 25;
  if (ro.wrapU != GTLC_WRAP_REPEAT || ro.uvScale[0] != 1.0f || ro.uvRotation != 0.0f) return 26;
  const GtlcTextureBinding& nm = b[2];
  if (nm.paramIndex != GTLC_BIND_NORMAL || nm.componentCount != 3 || nm.srgb || nm.scale[0] != 2.0f || nm.bias[2] != -1.0f) return 27;
  if (!giLowerMtlxToClosureWithTextures(kOpbrNormalMap, "N", &d, &b, &files) || d.closure != GTLC_CLOSURE_OPENPBR) return 28;
  if (b.size() != 2 || files[0] != "nrm.ppm" || files[1] != "c.ppm") return 29;                 // the procedural roughness graph is not lowered (warning)
  if (b[0].paramIndex != GTLC_BIND_NORMAL || b[0].scale[0] != 2.0f || b[0].bias[0] != -1.0f || b[0].wrapU != GTLC_WRAP_CLAMP || b[0].wrapV != GTLC_WRAP_CLIP || b[0].fallback[2] != 1.0f || b[0].srgb) return 30;
  if (b[1].paramIndex != GTLC_OPBR_BASE_COLOR_R || !b[1].srgb || b[1].componentCount != 3) return 31;
  if (d.params[GTLC_OPBR_SPECULAR_ROUGHNESS] != 0.3f) return 32;
  return 0;
}

static int lowerTests()
{
  if (int rc = lowerTextureTests()) return rc;
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

static int renderScene(const char* outPath, const char* texturePath)
{
  std::vector<std::string> noPaths;
  GiInitParams ip{ "", "", noPaths, nullptr, "" };
  if (giInitialize(ip) != GiStatus::Ok) { fprintf(stderr, "giInitialize failed: %s\n", giGetLastError()); return 10; }
  GiScene* scene = giCreateScene();
  GiMaterial* red = texturePath ? giCreateMaterialFromMtlxStr(scene, "Red", texturedUps(texturePath).c_str()) : giCreateMaterialFromMtlxStr(scene, "Red", kUps);
  GiMaterial* light = giCreateMaterialFromMtlxStr(scene, "Light", kEmissive);
  if (!red || !light) return 11;
  // floor quad (two triangles) + emissive quad above it
  std::vector<GiVertex> fv = { vert(-1, -1, 0, 0, 0, 1), vert(1, -1, 0, 0, 0, 1), vert(1, 1, 0, 0, 0, 1), vert(-1, 1, 0, 0, 0, 1) };
  if (texturePath) for (GiVertex& v : fv) { v.u = (v.pos[0] + 1.0f) * 0.5f; v.v = (v.pos[1] + 1.0f) * 0.5f; }
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
  rp.renderSettings.mediumStackSize = 9;
  if (giRender(rp) != GiStatus::Error) return 16;
  giDestroyRenderBuffer(rb); giDestroySphereLight(scene, sl); giDestroyMesh(floorMesh); giDestroyMesh(lightMesh);
  giDestroyMaterial(red); giDestroyMaterial(light); giDestroyScene(scene); giTerminate();
  printf("render ok\n");
  return 0;
}

// No GPU: everything up to (not including) giRender works on the host, like Hydra's parallel prim sync.
static int objectTests()
{
  GiScene* scene = giCreateScene();
  if (!scene) return 50;
  GiMaterial* mat = giCreateMaterialFromMtlxStr(scene, "P", kPrimvarUps);
  if (!mat) return 51;
  std::vector<GiVertex> v = { vert(-1, -1, 0, 0, 0, 1), vert(1, -1, 0, 0, 0, 1), vert(1, 1, 0, 0, 0, 1), vert(-1, 1, 0, 0, 0, 1) };
  std::vector<GiFace> f = { { { 0, 1, 2 } }, { { 0, 2, 3 } } };
  std::vector<int> faceIds = { 7, 9 };
  const float colours[12] = { 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 0 };
  const float rough = 0.4f;
  std::vector<GiPrimvarData> primvars(2);
  primvars[0].name = "displayColor"; primvars[0].type = GiPrimvarType::Vec3; primvars[0].interpolation = GiPrimvarInterpolation::Vertex;
  primvars[0].data.assign((const uint8_t*) colours, (const uint8_t*) colours + sizeof(colours));
  primvars[1].name = "rough"; primvars[1].type = GiPrimvarType::Float; primvars[1].interpolation = GiPrimvarInterpolation::Constant;
  primvars[1].data.assign((const uint8_t*) &rough, (const uint8_t*) &rough + 4);
  GiMeshDesc md{ 2, f, faceIds, 3, true, false, "quad", 9, primvars, 4, v };
  GiMesh* mesh = giCreateMesh(scene, md);
  if (!mesh) { fprintf(stderr, "%s\n", giGetLastError()); return 52; }
  giSetMeshMaterial(mesh, mat);
  const float two[2][4][4] = { { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } }, { { 1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, 0, 1, 0 }, { 3, 0, 0, 1 } } };
  giSetMeshInstanceTransforms(mesh, 2, two);
  int ids[2] = { 1, 0 };
  giSetMeshInstanceIds(mesh, 2, ids);
  std::vector<GiPrimvarData> inst(1);
  const float instRough[2] = { 0.2f, 0.8f };
  inst[0].name = "rough"; inst[0].type = GiPrimvarType::Float; inst[0].interpolation = GiPrimvarInterpolation::Instance;
  inst[0].data.assign((const uint8_t*) instRough, (const uint8_t*) instRough + 8);
  giSetMeshInstancerPrimvars(mesh, inst);
  GiRectLight* rl = giCreateRectLight(scene);
  float o[3] = { 0, 0, 2 }, t0[3] = { 1, 0, 0 }, t1[3] = { 0, 1, 0 }, em[3] = { 5, 5, 5 };
  giSetRectLightOrigin(rl, o); giSetRectLightTangents(rl, t0, t1); giSetRectLightBaseEmission(rl, em); giSetRectLightDimensions(rl, 1, 1);
  // the flattened scene through the C-ABI: 2 instances x 2 triangles, second instance shifted by +3 in x
  struct Peek { GtlcHandle h; };
  GtlcBvhView view;
  if (gtlc_scene_get_bvh(reinterpret_cast<Peek*>(scene)->h, &view) != GTLC_OK || view.triCount != 4 || view.nodeCount < 1) return 53;
  float maxX = -1e30f;
  for (uint32_t i = 0; i < view.triCount * 3; i++) maxX = std::fmax(maxX, view.triangles[i * 4]);
  if (maxX != 4.0f) return 54;
  // a mesh whose primvar description is broken is refused
  std::vector<GiPrimvarData> bad(1);
  bad[0].name = "x"; bad[0].type = (GiPrimvarType) 42; bad[0].interpolation = GiPrimvarInterpolation::Vertex;
  GiMeshDesc bd{ 2, f, faceIds, 4, true, false, "bad", 9, bad, 4, v };
  if (giCreateMesh(scene, bd) != nullptr) return 55;
  giDestroyRectLight(scene, rl); giDestroyMesh(mesh); giDestroyMaterial(mat); giDestroyScene(scene);
  printf("objects ok\n");
  return 0;
}

int main(int argc, char** argv)
{
  if (argc >= 2 && !strcmp(argv[1], "objects")) return objectTests();
  if (argc >= 2 && !strcmp(argv[1], "lower")) return lowerTests();
  if (argc >= 3 && !strcmp(argv[1], "render")) return renderScene(argv[2], argc >= 4 ? argv[3] : nullptr);
  fprintf(stderr, "usage: %s lower | render <out.bin>\n", argv[0]);
  return 2;
}
