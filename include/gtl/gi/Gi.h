//
// The declarations in this header reproduce the public interface of gatling's gi library
// (src/gi/gtl/gi/Gi.h), Copyright (C) 2019-2022 Pablo Delgado Krämer, so that its callers build
// unchanged.  That interface is licensed as follows, and this header is distributed under the same terms:
//
// This program is free software: you can redistribute it and/or modify
// it under the terms of the GNU General Public License as published by
// the Free Software Foundation, either version 3 of the License, or
// (at your option) any later version.
//
// This program is distributed in the hope that it will be useful,
// but WITHOUT ANY WARRANTY; without even the implied warranty of
// MERCHANTABILITY or FITNESS FOR A PARTICULAR PURPOSE. See the
// GNU General Public License for more details.
//
// You should have received a copy of the GNU General Public License
// along with this program. If not, see <https://www.gnu.org/licenses/>.
//
// gtl/gi/Gi.h -- source-compatible `gi` renderer API of the B200-native wavefront core.
//
// hdGatling (the Hydra render delegate) and the gatling CLI include <gtl/gi/Gi.h> and link the
// functions declared here; this header keeps every type and signature of the reference's
// src/gi/gtl/gi/Gi.h:30-261 so those callers build unchanged, while the implementation
// (gatling_b200/gi_cpp/Gi.cpp) is a thin adapter onto the C-ABI in include/gtlc.h.
// What differs behind the signatures is documented in INTEGRATION.md (material lowering,
// dome-light image formats, AOV coverage).
#pragma once

#include <stddef.h>
#include <stdint.h>

#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <variant>
#include <vector>

#include <gtl/gb/ParamTypes.h>

namespace gtl
{
  // ---- opaque object types: the caller owns them and destroys them explicitly -----------------
  struct GiAsset;
  struct GiMaterial;
  struct GiMesh;
  struct GiShaderCache;
  struct GiScene;
  struct GiSphereLight;
  struct GiDistantLight;
  struct GiRectLight;
  struct GiDiskLight;
  struct GiDomeLight;
  struct GiRenderBuffer;

  // ---- enums and constants --------------------------------------------------------------------
  constexpr static const uint32_t GI_MAX_AOV_COMP_SIZE = 16; // one vec4

  enum class GiStatus { Ok, Error };

  enum class GiAovId
  {
    Color = 0, Normal, NEE, Barycentrics, Texcoords, Bounces, ClockCycles, Opacity, Tangents, Bitangents,
    ThinWalled, ObjectId, Depth, FaceId, InstanceId, DoubleSided, Albedo,
    COUNT
  };

  enum class GiRenderBufferFormat { Int32, Float32, Float32Vec4 };

  enum class GiPrimvarType { Float, Vec2, Vec3, Vec4, Int, Int2, Int3, Int4 };

  enum class GiPrimvarInterpolation { Constant, Instance, Uniform, Vertex, COUNT };

  // ---- geometry ---------------------------------------------------------------------------------
  struct GiVertex
  {
    float pos[3];
    float u;
    float norm[3];
    float v;
    float tangent[3];
    float bitangentSign;
  };

  struct GiFace { uint32_t v_i[3]; };

  struct GiPrimvarData
  {
    std::string name;
    GiPrimvarType type;
    GiPrimvarInterpolation interpolation;
    std::vector<uint8_t> data;
  };

  struct GiMeshDesc
  {
    uint32_t                          faceCount;
    const std::vector<GiFace>&        faces;
    const std::vector<int>&           faceIds;
    int                               id;
    bool                              isDoubleSided;
    bool                              isLeftHanded;
    const char*                       name;
    uint32_t                          maxFaceId;
    const std::vector<GiPrimvarData>& primvars;
    uint32_t                          vertexCount;
    const std::vector<GiVertex>&      vertices;
  };

  // ---- camera, settings, render parameters ------------------------------------------------------
  struct GiCameraDesc
  {
    float position[3];
    float forward[3];
    float up[3];
    float vfov;
    float fStop;
    float focusDistance;
    float focalLength;
    float clipStart;
    float clipEnd;
    float exposure;
  };

  struct GiRenderSettings
  {
    bool     clippingPlanes;
    bool     depthOfField;
    bool     domeLightCameraVisible;
    bool     filterImportanceSampling;
    float    frame;
    bool     jitteredSampling;
    float    lightIntensityMultiplier;
    uint32_t maxBounces;
    float    maxSampleValue;
    uint32_t maxVolumeWalkLength;
    uint32_t mediumStackSize;
    float    metersPerSceneUnit;
    bool     nextEventEstimation;
    bool     progressiveAccumulation;
    uint32_t rrBounceOffset;
    float    rrInvMinTermProb;
    uint32_t spp;
    float    time;
  };

  struct GiAovBinding
  {
    GiAovId         aovId;
    uint8_t         clearValue[GI_MAX_AOV_COMP_SIZE];
    GiRenderBuffer* renderBuffer;
  };

  struct GiRenderParams
  {
    std::vector<GiAovBinding> aovBindings;
    GiCameraDesc              camera;
    GiDomeLight*              domeLight;
    GiRenderSettings          renderSettings;
    GiScene*                  scene;
  };

  struct GiInitParams
  {
    std::string_view shaderPath;
    std::string_view mdlRuntimePath;
    const std::vector<std::string>& mdlSearchPaths;
    const std::shared_ptr<void/*MaterialX::Document*/> mtlxStdLib;
    std::string mtlxCustomNodesPath;
  };

  // ---- pluggable asset IO -------------------------------------------------------------------------
  class GiAssetReader
  {
  public:
    virtual GiAsset* open(const char* path) = 0;
    virtual size_t size(const GiAsset* asset) const = 0;
    virtual void* data(const GiAsset* asset) const = 0;
    virtual void close(GiAsset* asset) = 0;
    virtual ~GiAssetReader() = default;
  };

  using GiMaterialParameterValue = std::variant<bool, int, float, GbVec2f, GbVec3f, GbVec4f, GbColor, GbTextureAsset>;
  using GiMaterialParameters = std::unordered_map<std::string, GiMaterialParameterValue>;

  // ---- lifecycle ------------------------------------------------------------------------------------
  GiStatus giInitialize(const GiInitParams& params);
  void giTerminate();
  void giRegisterAssetReader(GiAssetReader* reader);

  // ---- scene ----------------------------------------------------------------------------------------
  GiScene* giCreateScene();
  void giDestroyScene(GiScene* scene);

  // ---- materials ------------------------------------------------------------------------------------
  GiMaterial* giCreateMaterialFromMtlxStr(GiScene* scene, const char* name, const char* mtlxSrc);
  GiMaterial* giCreateMaterialFromMtlxDoc(GiScene* scene, const char* name, const std::shared_ptr<void/*MaterialX::Document*/> doc);
  GiMaterial* giCreateMaterialFromMdlFile(GiScene* scene, const char* name, const char* filePath, const char* subIdentifier, const GiMaterialParameters& params = {});
  void giDestroyMaterial(GiMaterial* mat);

  // ---- meshes ---------------------------------------------------------------------------------------
  GiMesh* giCreateMesh(GiScene* scene, const GiMeshDesc& desc);
  void giSetMeshTransform(GiMesh* mesh, const float* mat4x4);
  void giSetMeshInstanceTransforms(GiMesh* mesh, uint32_t count, const float(*transforms)[4][4]);
  void giSetMeshInstancerPrimvars(GiMesh* mesh, const std::vector<GiPrimvarData>& instancerPrimvars);
  void giSetMeshInstanceIds(GiMesh* mesh, uint32_t count, int* ids);
  void giSetMeshMaterial(GiMesh* mesh, GiMaterial* mat);
  void giSetMeshVisibility(GiMesh* mesh, bool visible);
  void giDestroyMesh(GiMesh* mesh);

  // ---- rendering (synchronous; results are in host memory on return) ----------------------------------
  GiStatus giRender(const GiRenderParams& params);

  // ---- analytic lights ------------------------------------------------------------------------------
  GiSphereLight* giCreateSphereLight(GiScene* scene);
  void giDestroySphereLight(GiScene* scene, GiSphereLight* light);
  void giSetSphereLightPosition(GiSphereLight* light, float* position);
  void giSetSphereLightBaseEmission(GiSphereLight* light, float* rgb);
  void giSetSphereLightRadius(GiSphereLight* light, float radiusX, float radiusY, float radiusZ);
  void giSetSphereLightDiffuseSpecular(GiSphereLight* light, float diffuse, float specular);

  GiDistantLight* giCreateDistantLight(GiScene* scene);
  void giDestroyDistantLight(GiScene* scene, GiDistantLight* light);
  void giSetDistantLightDirection(GiDistantLight* light, float* direction);
  void giSetDistantLightBaseEmission(GiDistantLight* light, float* rgb);
  void giSetDistantLightAngle(GiDistantLight* light, float angle);
  void giSetDistantLightDiffuseSpecular(GiDistantLight* light, float diffuse, float specular);

  GiRectLight* giCreateRectLight(GiScene* scene);
  void giDestroyRectLight(GiScene* scene, GiRectLight* light);
  void giSetRectLightOrigin(GiRectLight* light, float* origin);
  void giSetRectLightTangents(GiRectLight* light, float* t0, float* t1);
  void giSetRectLightBaseEmission(GiRectLight* light, float* rgb);
  void giSetRectLightDimensions(GiRectLight* light, float width, float height);
  void giSetRectLightDiffuseSpecular(GiRectLight* light, float diffuse, float specular);

  GiDiskLight* giCreateDiskLight(GiScene* scene);
  void giDestroyDiskLight(GiScene* scene, GiDiskLight* light);
  void giSetDiskLightOrigin(GiDiskLight* light, float* origin);
  void giSetDiskLightTangents(GiDiskLight* light, float* t0, float* t1);
  void giSetDiskLightBaseEmission(GiDiskLight* light, float* rgb);
  void giSetDiskLightRadius(GiDiskLight* light, float radiusX, float radiusY);
  void giSetDiskLightDiffuseSpecular(GiDiskLight* light, float diffuse, float specular);

  // ---- dome light -----------------------------------------------------------------------------------
  GiDomeLight* giCreateDomeLight(GiScene* scene, const char* filePath);
  void giDestroyDomeLight(GiDomeLight* light);
  void giSetDomeLightRotation(GiDomeLight* light, float* quat);
  void giSetDomeLightBaseEmission(GiDomeLight* light, float* rgb);
  void giSetDomeLightDiffuseSpecular(GiDomeLight* light, float diffuse, float specular);

  // ---- render buffers -------------------------------------------------------------------------------
  GiRenderBuffer* giCreateRenderBuffer(uint32_t width, uint32_t height, GiRenderBufferFormat format);
  void giDestroyRenderBuffer(GiRenderBuffer* renderBuffer);
  void* giGetRenderBufferMem(GiRenderBuffer* renderBuffer);

  // ---- additions of this implementation (not in the reference header) ---------------------------------
  // Serialiser hook for giCreateMaterialFromMtlxDoc: the facade cannot look inside a MaterialX::Document
  // without linking MaterialX, so the caller (hdGatling, which links MaterialX anyway) registers
  // `MaterialX::writeToXmlString` once; the XML is then lowered like giCreateMaterialFromMtlxStr does.
  using GiMtlxDocToXmlFn = std::string (*)(const std::shared_ptr<void>& doc);
  void giRegisterMtlxDocSerializer(GiMtlxDocToXmlFn fn);
  // Last error line (the reference logs through GB_ERROR).
  const char* giGetLastError();
}
