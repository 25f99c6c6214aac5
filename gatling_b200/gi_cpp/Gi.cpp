// C++ facade: the reference's gi* API (include/gtl/gi/Gi.h == src/gi/gtl/gi/Gi.h:199-261) implemented
// as a thin adapter over the C-ABI (include/gtlc.h).  This file plus libgtlc.so is what replaces
// src/gi/impl/Gi.cpp in a gatling build; it contains no rendering code.
//
// Material lowering: the reference hands a MaterialX document to `mc`, which patches it, generates MDL
// and compiles that with the MDL SDK into per-material GLSL (src/mc, out of scope -- closed SDK).
// Here MaterialX XML is lowered directly to a closure id + parameter block for the two surface shaders
// the configs name (UsdPreviewSurface, open_pbr_surface).  Inputs connected to UsdUVTexture / image nodes
// (optionally through UsdTransform2d and normalmap) become texture bindings of the C-ABI material; what the
// reference's patcher decides about colour spaces (MtlxDocumentPatcher.cpp:201-253) is applied here.  Other
// connected inputs (procedural graphs, primvar readers other than `st`) keep the node's default with a warning.
#include <gtl/gi/Gi.h>

#include "../../include/gtlc.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <mutex>
#include <string>
#include <vector>

namespace gtl
{
  // Opaque reference types wrap a C-ABI handle.
  struct GiScene { GtlcHandle h; };
  struct GiMaterial { GtlcHandle h; };
  struct GiMesh { GtlcHandle h; };
  struct GiSphereLight { GtlcHandle h; };
  struct GiDistantLight { GtlcHandle h; };
  struct GiRectLight { GtlcHandle h; };
  struct GiDiskLight { GtlcHandle h; };
  struct GiDomeLight { GtlcHandle h; };
  struct GiRenderBuffer { GtlcHandle h; uint32_t width = 0, height = 0; };

  namespace
  {
    GiMtlxDocToXmlFn s_docSerializer = nullptr;
    GiAssetReader* s_assetReader = nullptr;
    std::string s_lastError;
    std::mutex s_errMutex;

    void logError(const std::string& msg)
    {
      std::lock_guard<std::mutex> g(s_errMutex);
      s_lastError = msg;
      fprintf(stderr, "[gi] error: %s\n", msg.c_str());
    }

    template <typename T> T* wrap(GtlcHandle h) { if (!h) { logError(gtlc_last_error()); return nullptr; } T* o = new T; o->h = h; return o; }

    // ---- minimal MaterialX XML scanning -------------------------------------------------------
    struct XmlInput { std::string name, value, nodename, output, colorspace; bool connected; };

    bool attr(const std::string& tag, const char* key, std::string& out)
    {
      std::string k = std::string(key) + "=\"";
      size_t p = tag.find(" " + k);
      if (p == std::string::npos) return false;
      p += k.size() + 1;
      size_t e = tag.find('"', p);
      if (e == std::string::npos) return false;
      out = tag.substr(p, e - p);
      return true;
    }

    // Finds the first <nodeName ...> element (surface shader instance, not the nodedef) and collects its <input> children.
    bool findShaderNode(const std::string& xml, const char* nodeName, std::vector<XmlInput>& inputs)
    {
      std::string open = std::string("<") + nodeName;
      size_t p = 0;
      while ((p = xml.find(open, p)) != std::string::npos) {
        char c = xml[p + open.size()];
        if (c != ' ' && c != '\t' && c != '\n' && c != '\r' && c != '>') { p += open.size(); continue; }
        size_t tagEnd = xml.find('>', p);
        if (tagEnd == std::string::npos) return false;
        bool selfClosing = xml[tagEnd - 1] == '/';
        inputs.clear();
        if (selfClosing) return true;
        std::string close = std::string("</") + nodeName + ">";
        size_t end = xml.find(close, tagEnd);
        if (end == std::string::npos) return false;
        size_t q = tagEnd;
        while ((q = xml.find("<input", q)) != std::string::npos && q < end) {
          size_t e = xml.find('>', q);
          std::string tag = xml.substr(q, e - q);
          XmlInput in; in.connected = false;
          std::string dummy;
          if (attr(tag, "name", in.name)) {
            attr(tag, "nodename", in.nodename); attr(tag, "output", in.output); attr(tag, "colorspace", in.colorspace);
            if (!attr(tag, "value", in.value)) in.connected = !in.nodename.empty() || attr(tag, "nodegraph", dummy) || !in.output.empty() || attr(tag, "interfacename", dummy);
            inputs.push_back(in);
          }
          q = e;
        }
        return true;
      }
      return false;
    }

    int parseFloats(const std::string& v, float* out, int maxCount)
    {
      int n = 0; const char* s = v.c_str();
      while (*s && n < maxCount) {
        char* end = nullptr;
        float f = strtof(s, &end);
        if (end == s) { if (!strncmp(s, "true", 4)) { f = 1.0f; end = (char*) s + 4; } else if (!strncmp(s, "false", 5)) { f = 0.0f; end = (char*) s + 5; } else { s++; continue; } }
        out[n++] = f; s = end;
      }
      return n;
    }


    // Finds the element <category name="nodeName" ...> (any category except input/output) and collects its inputs.
    bool findNodeByName(const std::string& xml, const std::string& nodeName, std::string& category, std::vector<XmlInput>& inputs)
    {
      const std::string key = " name=\"" + nodeName + "\"";
      size_t p = 0;
      while ((p = xml.find(key, p)) != std::string::npos) {
        size_t lt = xml.rfind('<', p);
        if (lt == std::string::npos) return false;
        size_t ce = lt + 1;
        while (ce < xml.size() && xml[ce] != ' ' && xml[ce] != '\t' && xml[ce] != '\n' && xml[ce] != '\r' && xml[ce] != '>') ce++;
        category = xml.substr(lt + 1, ce - lt - 1);
        if (category == "input" || category == "output" || xml.find('>', lt) < p) { p += key.size(); continue; }
        std::string rest = xml.substr(lt);     // re-use the category scanner on the tail that starts at this element
        return findShaderNode(rest, category.c_str(), inputs);
      }
      return false;
    }

    const XmlInput* findInput(const std::vector<XmlInput>& inputs, const char* name)
    {
      for (const XmlInput& in : inputs) if (in.name == name) return &in;
      return nullptr;
    }

    int wrapModeOf(const std::string& v)
    {
      if (v == "black" || v == "constant") return GTLC_WRAP_CLIP;
      if (v == "clamp") return GTLC_WRAP_CLAMP;
      if (v == "mirror") return GTLC_WRAP_MIRRORED_REPEAT;
      return GTLC_WRAP_REPEAT;     // "repeat" / "periodic" / "useMetadata" (ND_UsdUVTexture default: periodic)
    }

    // ---- lowering of the graph in front of a surface-shader input ----
    // Supported: (arithmetic with constants ->) (normalmap ->) UsdUVTexture | image (-> UsdTransform2d -> st reader), and ONE
    // multiply / add / subtract of two such chains or ONE mix of them (constant or image operands, constant or image mask) on the
    // way.  "Arithmetic with constants": multiply / add / subtract / divide / invert nodes whose other operand is a value -- what
    // authoring tools put between an image and a shader input to tint, scale or invert it.  They compose to one affine map per
    // component, which folds into a binding's per-channel scale / bias (UsdUVTexture semantics: value = texel * scale + bias,
    // after the colour-space transform, as in the graph).  Two lookups meet through the combine word of the bindings
    // (gtlc.h GtlcCombineOp), which the library applies in order; a mask image waits in a scratch float.  Anything else
    // (noise, nested mixes, ...) is not lowered: the input keeps its constant and the caller warns.  (The reference compiles
    // the whole graph to MDL: mc/impl/MtlxMdlCodeGen.cpp:189-278.)
    int parseConstant3(const XmlInput* in, float dflt, float k[3])
    {
      k[0] = k[1] = k[2] = dflt;
      if (!in || in->value.empty()) return 0;
      float v[4]; const int n = parseFloats(in->value, v, 4);
      if (n == 1) k[0] = k[1] = k[2] = v[0]; else for (int q = 0; q < n && q < 3; q++) k[q] = v[q];
      return n;
    }
    bool isConnected(const XmlInput* in) { return in && !in->nodename.empty(); }

    // Walks the constant-arithmetic nodes starting at (category, node), composing input = upstream * mulK + addK.  Stops at the
    // first node of another kind, or at a multiply / add / subtract whose operands are BOTH connected (`binary`).
    bool walkConstantArithmetic(const std::string& xml, std::string& category, std::vector<XmlInput>& node, std::string& output, float mulK[3], float addK[3], bool& binary)
    {
      binary = false;
      for (int hops = 0; hops < 16; hops++) {
        const bool isMul = category == "multiply", isAdd = category == "add", isSub = category == "subtract", isDiv = category == "divide", isInv = category == "invert";
        if (!isMul && !isAdd && !isSub && !isDiv && !isInv) return true;
        const XmlInput* a = findInput(node, isInv ? "in" : "in1");
        const XmlInput* c = findInput(node, isInv ? "amount" : "in2");
        const bool aConn = isConnected(a), cConn = isConnected(c);
        if (aConn && cConn) { binary = isMul || isAdd || isSub; return binary; }
        if (!aConn && !cConn) return false;
        const XmlInput* conn = aConn ? a : c;
        float k[3]; parseConstant3(aConn ? c : a, (isMul || isDiv || isInv) ? 1.0f : 0.0f, k);
        for (int q = 0; q < 3; q++) {
          if (isMul) mulK[q] *= k[q];
          else if (isAdd) addK[q] += k[q] * mulK[q];
          else if (isSub || isInv) {
            if (aConn && isSub) addK[q] -= k[q] * mulK[q];                           // upstream - k
            else { addK[q] += k[q] * mulK[q]; mulK[q] = -mulK[q]; }                  // k - upstream (subtract with in2 connected, invert: amount - in)
          } else {                                                                  // divide
            if (!aConn || k[q] == 0.0f) return false;                               // k / upstream is not affine
            mulK[q] /= k[q];
          }
        }
        output = conn->output;
        const std::string up = conn->nodename;
        if (!findNodeByName(xml, up, category, node)) return false;
      }
      return false;
    }

    void initBinding(GtlcTextureBinding& b, int paramIndex, int count, int combine)
    {
      memset(&b, 0, sizeof(b));
      for (int k = 0; k < 4; k++) { b.scale[k] = 1.0f; b.bias[k] = 0.0f; }
      b.fallback[3] = 1.0f; b.uvScale[0] = b.uvScale[1] = 1.0f;
      b.paramIndex = paramIndex; b.componentCount = count; b.combine = combine;
    }

    // One lookup chain: the node `nodename` (output `output`) must lead through constant arithmetic (and a normalmap) to an image.
    // value written = chain value * mulK0 + addK0, combined with the input's current value as `combine` says.
    bool lowerLookupChain(const std::string& xml, const std::string& nodename, std::string output, int paramIndex, int count, bool upsSrgbInput,
                          const float mulK0[3], const float addK0[3], int combine, GtlcTextureBinding& b, std::string& file)
    {
      std::string category; std::vector<XmlInput> node;
      if (nodename.empty() || !findNodeByName(xml, nodename, category, node)) return false;
      initBinding(b, paramIndex, count, combine);
      float mulK[3] = { mulK0[0], mulK0[1], mulK0[2] }, addK[3] = { addK0[0], addK0[1], addK0[2] };
      bool binary = false;
      if (!walkConstantArithmetic(xml, category, node, output, mulK, addK, binary) || binary) return false;
      bool decodeNormal = false;
      if (category == "normalmap") {                       // MaterialX normalmap decodes value * 2 - 1 itself
        const XmlInput* src = findInput(node, "in");
        if (!src || src->nodename.empty()) return false;
        output = src->output; decodeNormal = true;
        std::string up = src->nodename;
        if (!findNodeByName(xml, up, category, node)) return false;
      }
      const bool usdUv = category == "UsdUVTexture";
      if (!usdUv && category != "image") return false;
      const XmlInput* fileIn = findInput(node, "file");
      if (!fileIn || fileIn->value.empty()) return false;
      file = fileIn->value;
      float v[4];
      if (usdUv) {
        if (const XmlInput* i = findInput(node, "wrapS")) b.wrapU = wrapModeOf(i->value); else b.wrapU = GTLC_WRAP_REPEAT;
        if (const XmlInput* i = findInput(node, "wrapT")) b.wrapV = wrapModeOf(i->value); else b.wrapV = GTLC_WRAP_REPEAT;
        if (const XmlInput* i = findInput(node, "fallback")) if (parseFloats(i->value, v, 4) == 4) memcpy(b.fallback, v, 16);
        if (const XmlInput* i = findInput(node, "scale")) if (parseFloats(i->value, v, 4) == 4) memcpy(b.scale, v, 16);
        if (const XmlInput* i = findInput(node, "bias")) if (parseFloats(i->value, v, 4) == 4) memcpy(b.bias, v, 16);
        // _PatchUsdUVTextureColorSpaces (MtlxDocumentPatcher.cpp:201-253): an explicit colorspace on the file input wins,
        // then sourceColorSpace, then "auto" = sRGB for the three UsdPreviewSurface colour inputs
        std::string cs = "auto";
        if (const XmlInput* i = findInput(node, "sourceColorSpace")) cs = i->value;
        if (!fileIn->colorspace.empty()) b.srgb = fileIn->colorspace == "srgb_texture";
        else b.srgb = (cs == "sRGB") || (cs == "auto" && upsSrgbInput);
      } else {
        if (const XmlInput* i = findInput(node, "uaddressmode")) b.wrapU = wrapModeOf(i->value); else b.wrapU = GTLC_WRAP_REPEAT;
        if (const XmlInput* i = findInput(node, "vaddressmode")) b.wrapV = wrapModeOf(i->value); else b.wrapV = GTLC_WRAP_REPEAT;
        if (const XmlInput* i = findInput(node, "default")) { int n = parseFloats(i->value, v, 4); for (int k = 0; k < n; k++) b.fallback[k] = v[k]; if (n == 1) b.fallback[1] = b.fallback[2] = v[0]; }
        b.srgb = fileIn->colorspace == "srgb_texture";
      }
      if (decodeNormal) for (int k = 0; k < 3; k++) { b.bias[k] = b.bias[k] * 2.0f - 1.0f; b.scale[k] *= 2.0f; }
      // which channels feed the input
      if (output == "r" || output == "g" || output == "b" || output == "a") { const int c = output == "r" ? 0 : output == "g" ? 1 : output == "b" ? 2 : 3; b.channels[0] = b.channels[1] = b.channels[2] = c; }
      else { b.channels[0] = 0; b.channels[1] = count > 1 ? 1 : 0; b.channels[2] = count > 2 ? 2 : 0; }
      // fold the arithmetic between the lookup and the input into scale / bias of the source channels it reads
      {
        bool seen[4] = { false, false, false, false }; float ns[4], nb[4];
        for (int k = 0; k < count && k < 3; k++) {
          const int ch = b.channels[k];
          const float sc = b.scale[ch] * mulK[k], bi = b.bias[ch] * mulK[k] + addK[k];
          if (seen[ch] && (ns[ch] != sc || nb[ch] != bi)) return false;             // one source channel feeding two components with different maps
          seen[ch] = true; ns[ch] = sc; nb[ch] = bi;
        }
        for (int ch = 0; ch < 4; ch++) if (seen[ch]) { b.scale[ch] = ns[ch]; b.bias[ch] = nb[ch]; }
      }
      // texture coordinates: UsdTransform2d in front of the st reader
      const XmlInput* st = findInput(node, usdUv ? "st" : "texcoord");
      if (st && !st->nodename.empty()) {
        std::string c2; std::vector<XmlInput> xf;
        if (findNodeByName(xml, st->nodename, c2, xf) && c2 == "UsdTransform2d") {
          if (const XmlInput* i = findInput(xf, "rotation")) if (parseFloats(i->value, v, 1) == 1) b.uvRotation = v[0];
          if (const XmlInput* i = findInput(xf, "scale")) if (parseFloats(i->value, v, 2) == 2) { b.uvScale[0] = v[0]; b.uvScale[1] = v[1]; }
          if (const XmlInput* i = findInput(xf, "translation")) if (parseFloats(i->value, v, 2) == 2) { b.uvTranslation[0] = v[0]; b.uvTranslation[1] = v[1]; }
        }
      }
      return true;
    }

    // The bindings (and their image files; "" for a constant operand) that compute one shader input.  `scratchUsed` counts the
    // scratch floats this material's masks occupy so far.
    bool lowerTextureInput(const std::string& xml, const XmlInput& in, int paramIndex, int count, bool upsSrgbInput,
                           std::vector<GtlcTextureBinding>& bs, std::vector<std::string>& files, int& scratchUsed)
    {
      const float one[3] = { 1.0f, 1.0f, 1.0f }, zero[3] = { 0.0f, 0.0f, 0.0f };
      bs.clear(); files.clear();
      GtlcTextureBinding b; std::string file;
      if (lowerLookupChain(xml, in.nodename, in.output, paramIndex, count, upsSrgbInput, one, zero, GTLC_COMBINE_SET, b, file)) { bs.push_back(b); files.push_back(file); return true; }
      if (paramIndex == GTLC_BIND_NORMAL) return false;                              // the normal takes one plain lookup
      std::string category, output = in.output; std::vector<XmlInput> node;
      if (in.nodename.empty() || !findNodeByName(xml, in.nodename, category, node)) return false;
      float mulK[3] = { 1.0f, 1.0f, 1.0f }, addK[3] = { 0.0f, 0.0f, 0.0f };
      bool binary = false;
      if (!walkConstantArithmetic(xml, category, node, output, mulK, addK, binary)) return false;
      auto chain = [&](const XmlInput* operand, int param, int n, const float* mk, const float* ak, int combine) {
        GtlcTextureBinding cb; std::string cf;
        if (!lowerLookupChain(xml, operand->nodename, operand->output, param, n, upsSrgbInput, mk, ak, combine, cb, cf)) return false;
        bs.push_back(cb); files.push_back(cf); return true;
      };
      auto constant = [&](const float k[3], int combine) {
        GtlcTextureBinding cb; initBinding(cb, paramIndex, count, combine | GTLC_COMBINE_CONSTANT);
        cb.channels[0] = 0; cb.channels[1] = 1; cb.channels[2] = 2;
        for (int q = 0; q < 3; q++) cb.bias[q] = k[q];
        bs.push_back(cb); files.push_back("");
      };
      if (binary) {
        const XmlInput* a = findInput(node, "in1");
        const XmlInput* c = findInput(node, "in2");
        if (category == "multiply") {                                               // (a * c) * mulK + addK
          if (!chain(a, paramIndex, count, mulK, zero, GTLC_COMBINE_SET) || !chain(c, paramIndex, count, one, zero, GTLC_COMBINE(GTLC_COMBINE_MULTIPLY, 0))) return false;
          if (addK[0] != 0.0f || addK[1] != 0.0f || addK[2] != 0.0f) constant(addK, GTLC_COMBINE(GTLC_COMBINE_ADD, 0));
          return true;
        }
        float neg[3] = { -mulK[0], -mulK[1], -mulK[2] };                             // add: a mulK + addK + c mulK; subtract: ... - c mulK
        return chain(a, paramIndex, count, mulK, addK, GTLC_COMBINE_SET) && chain(c, paramIndex, count, category == "add" ? mulK : neg, zero, GTLC_COMBINE(GTLC_COMBINE_ADD, 0));
      }
      if (category != "mix") return false;
      const XmlInput* bg = findInput(node, "bg");
      const XmlInput* fg = findInput(node, "fg");
      const XmlInput* mx = findInput(node, "mix");
      float bgK[3], fgK[3], mK[3];
      parseConstant3(bg, 0.0f, bgK); parseConstant3(fg, 0.0f, fgK); parseConstant3(mx, 0.0f, mK);
      const bool bgConn = isConnected(bg), fgConn = isConnected(fg);
      if (!bgConn && !fgConn && !isConnected(mx)) return false;
      if (!isConnected(mx)) {                                                       // constant mask m: bg (1 - m) + fg m is affine in the two lookups
        float wb[3], wf[3], add[3];
        for (int q = 0; q < 3; q++) {
          wb[q] = (1.0f - mK[q]) * mulK[q]; wf[q] = mK[q] * mulK[q];
          add[q] = addK[q] + (bgConn ? 0.0f : bgK[q] * wb[q]) + (fgConn ? 0.0f : fgK[q] * wf[q]);
        }
        if (bgConn && fgConn) return chain(bg, paramIndex, count, wb, add, GTLC_COMBINE_SET) && chain(fg, paramIndex, count, wf, zero, GTLC_COMBINE(GTLC_COMBINE_ADD, 0));
        return bgConn ? chain(bg, paramIndex, count, wb, add, GTLC_COMBINE_SET) : chain(fg, paramIndex, count, wf, add, GTLC_COMBINE_SET);
      }
      // an image mask: it waits in a scratch float while bg is written, then fg is mixed in (math::lerp)
      for (int q = 0; q < 3; q++) if (mulK[q] != 1.0f || addK[q] != 0.0f) return false;
      if (scratchUsed >= GTLC_BIND_SCRATCH_COUNT) return false;
      const int slot = GTLC_BIND_SCRATCH + scratchUsed;
      if (!chain(mx, slot, 1, one, zero, GTLC_COMBINE_SET)) return false;
      if (bgConn) { if (!chain(bg, paramIndex, count, one, zero, GTLC_COMBINE_SET)) return false; } else constant(bgK, GTLC_COMBINE_SET);
      if (fgConn) { if (!chain(fg, paramIndex, count, one, zero, GTLC_COMBINE(GTLC_COMBINE_MIX, slot))) return false; } else constant(fgK, GTLC_COMBINE(GTLC_COMBINE_MIX, slot));
      scratchUsed++;
      return true;
    }

    struct LoweredPrimvar { int paramIndex, componentCount; std::string name; float fallback[4]; };
    // bindings[i] reads files[i] ("" = a constant operand); groups[i] = the shader input it helps to compute (bindings of one input stand or fall together)
    struct LoweredTextures { std::vector<GtlcTextureBinding> bindings; std::vector<std::string> files; std::vector<int> groups; std::vector<LoweredPrimvar> primvars; int scratchUsed = 0, groupCount = 0; };

    // A surface-shader input fed by UsdPrimvarReader / geompropvalue / geomcolor: the MDL backend turns these into
    // scene_data_lookup_* calls (mdl_interface.glsl:266-479) on the named primvar.
    bool lowerPrimvarInput(const std::string& xml, const XmlInput& in, int paramIndex, int count, LoweredPrimvar& out)
    {
      std::string category; std::vector<XmlInput> node;
      if (in.nodename.empty() || !findNodeByName(xml, in.nodename, category, node)) return false;
      const bool usd = category == "UsdPrimvarReader";
      if (!usd && category != "geompropvalue") return false;
      const XmlInput* name = findInput(node, usd ? "varname" : "geomprop");
      if (!name || name->value.empty()) return false;
      out.paramIndex = paramIndex; out.componentCount = count; out.name = name->value;
      out.fallback[0] = out.fallback[1] = out.fallback[2] = out.fallback[3] = 0.0f;
      if (const XmlInput* fb = findInput(node, usd ? "fallback" : "default")) { float v[4]; const int n = parseFloats(fb->value, v, 4); for (int k = 0; k < n; k++) out.fallback[k] = v[k]; if (n == 1) out.fallback[1] = out.fallback[2] = v[0]; }
      return true;
    }

    struct Slot { const char* name; int index; int count; };
    const Slot kUpsSlots[] = {
      { "diffuseColor", GTLC_UPS_DIFFUSE_R, 3 }, { "emissiveColor", GTLC_UPS_EMISSIVE_R, 3 }, { "useSpecularWorkflow", GTLC_UPS_USE_SPECULAR_WORKFLOW, 1 },
      { "specularColor", GTLC_UPS_SPECULAR_R, 3 }, { "metallic", GTLC_UPS_METALLIC, 1 }, { "roughness", GTLC_UPS_ROUGHNESS, 1 },
      { "clearcoat", GTLC_UPS_CLEARCOAT, 1 }, { "clearcoatRoughness", GTLC_UPS_CLEARCOAT_ROUGHNESS, 1 }, { "opacity", GTLC_UPS_OPACITY, 1 }, { "ior", GTLC_UPS_IOR, 1 } };
    const Slot kOpbrSlots[] = {
      { "base_weight", GTLC_OPBR_BASE_WEIGHT, 1 }, { "base_color", GTLC_OPBR_BASE_COLOR_R, 3 }, { "base_diffuse_roughness", GTLC_OPBR_BASE_DIFFUSE_ROUGHNESS, 1 },
      { "base_metalness", GTLC_OPBR_BASE_METALNESS, 1 }, { "specular_weight", GTLC_OPBR_SPECULAR_WEIGHT, 1 }, { "specular_color", GTLC_OPBR_SPECULAR_COLOR_R, 3 },
      { "specular_roughness", GTLC_OPBR_SPECULAR_ROUGHNESS, 1 }, { "specular_ior", GTLC_OPBR_SPECULAR_IOR, 1 },
      { "specular_roughness_anisotropy", GTLC_OPBR_SPECULAR_ROUGHNESS_ANISOTROPY, 1 }, { "transmission_weight", GTLC_OPBR_TRANSMISSION_WEIGHT, 1 },
      { "transmission_color", GTLC_OPBR_TRANSMISSION_COLOR_R, 3 }, { "transmission_depth", GTLC_OPBR_TRANSMISSION_DEPTH, 1 },
      { "coat_weight", GTLC_OPBR_COAT_WEIGHT, 1 }, { "coat_color", GTLC_OPBR_COAT_COLOR_R, 3 }, { "coat_roughness", GTLC_OPBR_COAT_ROUGHNESS, 1 },
      { "coat_ior", GTLC_OPBR_COAT_IOR, 1 }, { "coat_darkening", GTLC_OPBR_COAT_DARKENING, 1 }, { "fuzz_weight", GTLC_OPBR_FUZZ_WEIGHT, 1 },
      { "fuzz_color", GTLC_OPBR_FUZZ_COLOR_R, 3 }, { "fuzz_roughness", GTLC_OPBR_FUZZ_ROUGHNESS, 1 }, { "emission_luminance", GTLC_OPBR_EMISSION_LUMINANCE, 1 },
      { "emission_color", GTLC_OPBR_EMISSION_COLOR_R, 3 }, { "thin_film_weight", GTLC_OPBR_THIN_FILM_WEIGHT, 1 }, { "thin_film_thickness", GTLC_OPBR_THIN_FILM_THICKNESS, 1 },
      { "thin_film_ior", GTLC_OPBR_THIN_FILM_IOR, 1 }, { "geometry_opacity", GTLC_OPBR_GEOMETRY_OPACITY, 1 }, { "geometry_thin_walled", GTLC_OPBR_GEOMETRY_THIN_WALLED, 1 },
      { "subsurface_weight", GTLC_OPBR_SUBSURFACE_WEIGHT, 1 }, { "subsurface_color", GTLC_OPBR_SUBSURFACE_COLOR_R, 3 },
      { "transmission_scatter", GTLC_OPBR_TRANSMISSION_SCATTER_R, 3 }, { "transmission_scatter_anisotropy", GTLC_OPBR_TRANSMISSION_SCATTER_ANISOTROPY, 1 } };

    template <size_t N> void applyInputs(const std::string& xml, const std::vector<XmlInput>& inputs, const Slot (&slots)[N], const char* normalInput, bool ups,
                                         GtlcMaterialDesc& d, const char* matName, LoweredTextures* tex)
    {
      for (const XmlInput& in : inputs) {
        const Slot* slot = nullptr;
        for (const Slot& s : slots) if (in.name == s.name) { slot = &s; break; }
        const bool isNormal = in.name == normalInput;
        if (!slot && !isNormal) continue;   // inputs the closures do not model (displacement, occlusion, ...)
        if (in.connected) {
          std::vector<GtlcTextureBinding> bs; std::vector<std::string> files;
          const bool srgbInput = ups && (in.name == "diffuseColor" || in.name == "emissiveColor" || in.name == "specularColor");
          LoweredPrimvar pv;
          if (tex && lowerTextureInput(xml, in, isNormal ? GTLC_BIND_NORMAL : slot->index, isNormal ? 3 : slot->count, srgbInput, bs, files, tex->scratchUsed)) {
            for (size_t k = 0; k < bs.size(); k++) { tex->bindings.push_back(bs[k]); tex->files.push_back(files[k]); tex->groups.push_back(tex->groupCount); }
            tex->groupCount++;
          }
          else if (tex && !isNormal && tex->primvars.size() < GTLC_MAX_PRIMVAR_BINDINGS && lowerPrimvarInput(xml, in, slot->index, slot->count, pv)) tex->primvars.push_back(pv);
          else fprintf(stderr, "[gi] warning: material '%s': input '%s' is connected to a graph this build cannot lower; using the default value\n", matName ? matName : "", in.name.c_str());
          continue;
        }
        if (isNormal) continue;
        float v[4];
        if (parseFloats(in.value, v, slot->count) == slot->count) memcpy(&d.params[slot->index], v, sizeof(float) * slot->count);
      }
    }

    bool lowerMaterial(const char* mtlxSrc, const char* name, GtlcMaterialDesc* out, LoweredTextures* tex)
    {
      std::string xml(mtlxSrc ? mtlxSrc : "");
      std::vector<XmlInput> inputs;
      if (findShaderNode(xml, "UsdPreviewSurface", inputs)) {
        gtlc_material_desc_defaults(GTLC_CLOSURE_USD_PREVIEW_SURFACE, out);
        applyInputs(xml, inputs, kUpsSlots, "normal", true, *out, name, tex);
      } else if (findShaderNode(xml, "open_pbr_surface", inputs)) {
        gtlc_material_desc_defaults(GTLC_CLOSURE_OPENPBR, out);
        applyInputs(xml, inputs, kOpbrSlots, "geometry_normal", false, *out, name, tex);
      } else {
        return false;
      }
      out->name = name;
      return true;
    }

    // ---- image files: imgio is out of scope; the facade decodes binary PPM (P6, 8 bit) and PFM itself ----
    bool readFileBytes(const char* path, std::string& bytes)
    {
      bytes.clear();
      if (s_assetReader) {
        GiAsset* a = s_assetReader->open(path);
        if (a) { bytes.assign((const char*) s_assetReader->data(a), s_assetReader->size(a)); s_assetReader->close(a); }
      }
      if (bytes.empty()) { std::ifstream f(path, std::ios::binary); bytes.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>()); }
      return !bytes.empty();
    }

    // rows come out bottom-first (v = 0), as the reference's decoders deliver them (PngDecoder.cpp:78)
    GtlcHandle loadTextureFile(const char* path)
    {
      std::string bytes;
      if (!readFileBytes(path, bytes) || bytes.size() < 8) return 0;
      int w = 0, h = 0, maxval = 0, consumed = 0; float scale = 0.0f; char magic[3] = { 0, 0, 0 };
      GtlcTextureDesc td; memset(&td, 0, sizeof(td));
      if (bytes[0] == 'P' && bytes[1] == '6') {
        if (sscanf(bytes.c_str(), "%2s %d %d %d%n", magic, &w, &h, &maxval, &consumed) != 4 || w <= 0 || h <= 0 || maxval != 255) return 0;
        const size_t off = (size_t) consumed + 1;
        if (bytes.size() < off + (size_t) w * h * 3) return 0;
        std::vector<uint8_t> rgba((size_t) w * h * 4);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
          const uint8_t* s = (const uint8_t*) bytes.data() + off + ((size_t) (h - 1 - y) * w + x) * 3;     // PPM rows run top to bottom
          uint8_t* d = &rgba[((size_t) y * w + x) * 4];
          d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 255;
        }
        td.width = (uint32_t) w; td.height = (uint32_t) h; td.format = GTLC_TEX_RGBA8_UNORM; td.texels = rgba.data();
        return gtlc_texture_create(&td);
      }
      if (bytes[0] == 'P' && (bytes[1] == 'F' || bytes[1] == 'f')) {
        if (sscanf(bytes.c_str(), "%2s %d %d %f%n", magic, &w, &h, &scale, &consumed) != 4 || w <= 0 || h <= 0 || scale >= 0.0f) return 0;   // little endian only
        const int ch = bytes[1] == 'F' ? 3 : 1;
        const size_t off = (size_t) consumed + 1;
        if (bytes.size() < off + (size_t) w * h * ch * 4) return 0;
        std::vector<float> payload((size_t) w * h * ch);                            // PFM rows already run bottom to top
        memcpy(payload.data(), bytes.data() + off, payload.size() * 4);             // (the header leaves the floats unaligned)
        const float* src = payload.data();
        std::vector<float> texels((size_t) w * h * (ch == 3 ? 4 : 1));
        for (size_t i = 0; i < (size_t) w * h; i++) {
          if (ch == 1) texels[i] = src[i];
          else { texels[i * 4 + 0] = src[i * 3 + 0]; texels[i * 4 + 1] = src[i * 3 + 1]; texels[i * 4 + 2] = src[i * 3 + 2]; texels[i * 4 + 3] = 1.0f; }
        }
        td.width = (uint32_t) w; td.height = (uint32_t) h; td.format = ch == 3 ? GTLC_TEX_RGBA32_FLOAT : GTLC_TEX_R32_FLOAT; td.texels = texels.data();
        return gtlc_texture_create(&td);
      }
      return 0;
    }
  }

  // Exposed for tests: lowers MaterialX XML to a closure + parameter block (constant inputs only / with the texture
  // bindings and the image files they reference).  Returns false if no supported shader node is present.
  bool giLowerMtlxToClosure(const char* mtlxSrc, const char* name, GtlcMaterialDesc* out) { return lowerMaterial(mtlxSrc, name, out, nullptr); }
  bool giLowerMtlxToClosureWithTextures(const char* mtlxSrc, const char* name, GtlcMaterialDesc* out, std::vector<GtlcTextureBinding>* bindings, std::vector<std::string>* files)
  {
    LoweredTextures t;
    if (!lowerMaterial(mtlxSrc, name, out, &t)) return false;
    *bindings = t.bindings; *files = t.files;
    out->primvarCount = (uint32_t) t.primvars.size();      // test hook: the names are reported through `files` after the image files
    for (const LoweredPrimvar& p : t.primvars) files->push_back("primvar:" + p.name + ":" + std::to_string(p.paramIndex) + ":" + std::to_string(p.componentCount));
    return true;
  }

  const char* giGetLastError() { std::lock_guard<std::mutex> g(s_errMutex); static thread_local std::string c; c = s_lastError; return c.c_str(); }
  void giRegisterMtlxDocSerializer(GiMtlxDocToXmlFn fn) { s_docSerializer = fn; }

  GiStatus giInitialize(const GiInitParams& /* shader / MDL paths are meaningless here: closures are compiled ahead of time */)
  {
    GtlcInitParams p; p.deviceIndex = -1; p.reserved = 0;
    if (const char* ov = getenv("GTL_DEVICE_INDEX_OVERRIDE")) p.deviceIndex = atoi(ov);   // CgpuVk.cpp:975
    if (gtlc_initialize(&p) != GTLC_OK) { logError(gtlc_last_error()); return GiStatus::Error; }
    return GiStatus::Ok;
  }
  void giTerminate() { gtlc_terminate(); }
  void giRegisterAssetReader(GiAssetReader* reader) { s_assetReader = reader; }

  GiScene* giCreateScene() { return wrap<GiScene>(gtlc_scene_create()); }
  void giDestroyScene(GiScene* s) { if (!s) return; gtlc_scene_destroy(s->h); delete s; }

  GiMaterial* giCreateMaterialFromMtlxStr(GiScene* scene, const char* name, const char* mtlxSrc)
  {
    if (!scene) return nullptr;
    GtlcMaterialDesc d; LoweredTextures t;
    if (!lowerMaterial(mtlxSrc, name, &d, &t)) { logError(std::string("material '") + (name ? name : "") + "': no UsdPreviewSurface / open_pbr_surface node found (other MaterialX shading models need the mc/MDL path, out of scope)"); return nullptr; }
    // images: a file that cannot be decoded drops its binding (the input keeps its constant), like the reference's
    // 1x1 black fallback image keeps rendering going (TextureManager.cpp:245-262)
    // (an input computed from several lookups loses all of its bindings when one of its images is missing)
    std::vector<GtlcTextureBinding> bound; std::vector<GtlcHandle> textures;
    std::vector<char> groupFailed((size_t) t.groupCount, 0);
    for (size_t i = 0; i < t.bindings.size(); i++) {
      if (t.files[i].empty()) continue;                                             // a constant operand
      GtlcHandle h = loadTextureFile(t.files[i].c_str());
      if (!h) { fprintf(stderr, "[gi] warning: material '%s': failed to read image %s (this build decodes binary .ppm / .pfm; imgio is out of scope)\n", name ? name : "", t.files[i].c_str()); groupFailed[(size_t) t.groups[i]] = 1; continue; }
      t.bindings[i].texture = h; textures.push_back(h);
    }
    for (size_t i = 0; i < t.bindings.size(); i++) if (!groupFailed[(size_t) t.groups[i]]) bound.push_back(t.bindings[i]);
    d.textures = bound.empty() ? nullptr : bound.data(); d.textureCount = (uint32_t) bound.size();
    std::vector<GtlcPrimvarBinding> pvs(t.primvars.size());
    for (size_t i = 0; i < pvs.size(); i++) { pvs[i].paramIndex = t.primvars[i].paramIndex; pvs[i].componentCount = t.primvars[i].componentCount; pvs[i].name = t.primvars[i].name.c_str(); memcpy(pvs[i].fallback, t.primvars[i].fallback, 16); }
    d.primvars = pvs.empty() ? nullptr : pvs.data(); d.primvarCount = (uint32_t) pvs.size();
    GiMaterial* m = wrap<GiMaterial>(gtlc_material_create(scene->h, &d));
    for (GtlcHandle h : textures) gtlc_texture_destroy(h);     // the material holds its own references
    return m;
  }
  GiMaterial* giCreateMaterialFromMtlxDoc(GiScene* scene, const char* name, const std::shared_ptr<void> doc)
  {
    if (!s_docSerializer) { logError("giCreateMaterialFromMtlxDoc: no MaterialX serializer registered (giRegisterMtlxDocSerializer); see INTEGRATION.md"); return nullptr; }
    std::string xml = s_docSerializer(doc);
    return giCreateMaterialFromMtlxStr(scene, name, xml.c_str());
  }
  GiMaterial* giCreateMaterialFromMdlFile(GiScene*, const char* name, const char* filePath, const char*, const GiMaterialParameters&)
  {
    logError(std::string("giCreateMaterialFromMdlFile('") + (name ? name : "") + "', '" + (filePath ? filePath : "") + "'): arbitrary MDL needs the MDL SDK (not available; SURVEY.md section 2 #8)");
    return nullptr;
  }
  void giDestroyMaterial(GiMaterial* m) { if (!m) return; gtlc_material_destroy(m->h); delete m; }

  static std::vector<GtlcPrimvarDesc> toPrimvarDescs(const std::vector<GiPrimvarData>& primvars)
  {
    static_assert((int) GiPrimvarType::Int4 == GTLC_PRIMVAR_INT4 && (int) GiPrimvarInterpolation::Vertex == GTLC_INTERP_VERTEX, "enum values must match");
    std::vector<GtlcPrimvarDesc> out(primvars.size());
    for (size_t i = 0; i < primvars.size(); i++) {
      out[i].name = primvars[i].name.c_str(); out[i].type = (int32_t) primvars[i].type; out[i].interpolation = (int32_t) primvars[i].interpolation;
      out[i].data = primvars[i].data.data(); out[i].dataSize = primvars[i].data.size();
    }
    return out;
  }

  GiMesh* giCreateMesh(GiScene* scene, const GiMeshDesc& desc)
  {
    if (!scene) return nullptr;
    static_assert(sizeof(GiVertex) == sizeof(GtlcVertex) && sizeof(GiFace) == sizeof(GtlcFace), "layouts must match");
    GtlcMeshDesc d;
    d.faceCount = desc.faceCount; d.faces = reinterpret_cast<const GtlcFace*>(desc.faces.data());
    d.faceIds = desc.faceIds.size() == desc.faceCount ? reinterpret_cast<const int32_t*>(desc.faceIds.data()) : nullptr;
    d.id = desc.id; d.isDoubleSided = desc.isDoubleSided; d.isLeftHanded = desc.isLeftHanded; d.name = desc.name; d.maxFaceId = desc.maxFaceId;
    d.vertexCount = desc.vertexCount; d.vertices = reinterpret_cast<const GtlcVertex*>(desc.vertices.data());
    std::vector<GtlcPrimvarDesc> pv = toPrimvarDescs(desc.primvars);
    d.primvars = pv.empty() ? nullptr : pv.data(); d.primvarCount = (uint32_t) pv.size(); d.reserved = 0;
    return wrap<GiMesh>(gtlc_mesh_create(scene->h, &d));
  }
  void giSetMeshTransform(GiMesh* m, const float* mat4x4) { if (m) gtlc_mesh_set_transform(m->h, mat4x4); }
  void giSetMeshInstanceTransforms(GiMesh* m, uint32_t count, const float (*t)[4][4]) { if (m) gtlc_mesh_set_instance_transforms(m->h, count, reinterpret_cast<const float*>(t)); }
  void giSetMeshInstancerPrimvars(GiMesh* m, const std::vector<GiPrimvarData>& primvars)
  {
    if (!m) return;
    std::vector<GtlcPrimvarDesc> pv = toPrimvarDescs(primvars);
    gtlc_mesh_set_instancer_primvars(m->h, (uint32_t) pv.size(), pv.empty() ? nullptr : pv.data());
  }
  void giSetMeshInstanceIds(GiMesh* m, uint32_t count, int* ids) { if (m) gtlc_mesh_set_instance_ids(m->h, count, ids); }
  void giSetMeshMaterial(GiMesh* m, GiMaterial* mat) { if (m) gtlc_mesh_set_material(m->h, mat ? mat->h : 0); }
  void giSetMeshVisibility(GiMesh* m, bool visible) { if (m) gtlc_mesh_set_visibility(m->h, visible); }
  void giDestroyMesh(GiMesh* m) { if (!m) return; gtlc_mesh_destroy(m->h); delete m; }

  // _EncodeRenderBufferAsHeatmap (Gi.cpp:327-343): x / max(x) over the image -> a 256-step Turbo colour, alpha 255.  The
  // reference indexes Google's 256-entry Turbo table; here the published polynomial fit of the same map (Mikhailov 2019) is
  // evaluated at the 256 table positions -- within ~1/255 of the table, and this is a debug visualisation, not a parity target.
  static void turboColor(float x, float* rgb)
  {
    const float r4[4] = { 0.13572138f, 4.61539260f, -42.66032258f, 132.13108234f }, g4[4] = { 0.09140261f, 2.19418839f, 4.84296658f, -14.18503333f },
                b4[4] = { 0.10667330f, 12.64194608f, -60.58204836f, 110.36276771f };
    const float r2[2] = { -152.94239396f, 59.28637943f }, g2[2] = { 4.27729857f, 2.82956604f }, b2[2] = { -89.90310912f, 27.34824973f };
    x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
    const float v4[4] = { 1.0f, x, x * x, x * x * x };
    const float v2[2] = { v4[2] * v4[2], v4[3] * v4[2] };
    auto dot4 = [&](const float* c) { return v4[0] * c[0] + v4[1] * c[1] + v4[2] * c[2] + v4[3] * c[3]; };
    rgb[0] = dot4(r4) + v2[0] * r2[0] + v2[1] * r2[1];
    rgb[1] = dot4(g4) + v2[0] * g2[0] + v2[1] * g2[1];
    rgb[2] = dot4(b4) + v2[0] * b2[0] + v2[1] * b2[1];
    for (int k = 0; k < 3; k++) rgb[k] = rgb[k] < 0.0f ? 0.0f : (rgb[k] > 1.0f ? 1.0f : rgb[k]);
  }
  static void encodeRenderBufferAsHeatmap(GiRenderBuffer* rb)
  {
    float* rgba = (float*) gtlc_render_buffer_mem(rb->h);
    const uint32_t w = rb->width, h = rb->height;
    if (!rgba || !w || !h) return;
    const size_t channelCount = (size_t) w * h * 4;
    float maxValue = 0.0f;
    for (size_t i = 0; i < channelCount; i += 4) maxValue = std::max(maxValue, rgba[i]);          // only the first channel counts
    for (size_t i = 0; i < channelCount && maxValue > 0.0f; i += 4) {
      const int valIndex = std::min(int((rgba[i] / maxValue) * 255.0), 255);
      turboColor((float) valIndex / 255.0f, &rgba[i]);
      rgba[i + 3] = 255.0f;
    }
  }

  GiStatus giRender(const GiRenderParams& params)
  {
    std::vector<GtlcAovBinding> bindings(params.aovBindings.size());
    for (size_t i = 0; i < bindings.size(); i++) {
      bindings[i].aovId = (int32_t) params.aovBindings[i].aovId;
      memcpy(bindings[i].clearValue, params.aovBindings[i].clearValue, GI_MAX_AOV_COMP_SIZE);
      bindings[i].renderBuffer = params.aovBindings[i].renderBuffer ? params.aovBindings[i].renderBuffer->h : 0;
    }
    GtlcRenderParams p;
    p.aovBindingCount = (uint32_t) bindings.size(); p.aovBindings = bindings.data();
    static_assert(sizeof(GiCameraDesc) == sizeof(GtlcCameraDesc), "camera layouts must match");
    memcpy(&p.camera, &params.camera, sizeof(GtlcCameraDesc));
    p.domeLight = params.domeLight ? params.domeLight->h : 0;
    const GiRenderSettings& s = params.renderSettings;
    GtlcRenderSettings& r = p.renderSettings;
    r.clippingPlanes = s.clippingPlanes; r.depthOfField = s.depthOfField; r.domeLightCameraVisible = s.domeLightCameraVisible;
    r.filterImportanceSampling = s.filterImportanceSampling; r.frame = s.frame; r.jitteredSampling = s.jitteredSampling;
    r.lightIntensityMultiplier = s.lightIntensityMultiplier; r.maxBounces = s.maxBounces; r.maxSampleValue = s.maxSampleValue;
    r.maxVolumeWalkLength = s.maxVolumeWalkLength; r.mediumStackSize = s.mediumStackSize; r.metersPerSceneUnit = s.metersPerSceneUnit;
    r.nextEventEstimation = s.nextEventEstimation; r.progressiveAccumulation = s.progressiveAccumulation; r.rrBounceOffset = s.rrBounceOffset;
    r.rrInvMinTermProb = s.rrInvMinTermProb; r.spp = s.spp; r.time = s.time;
    p.scene = params.scene ? params.scene->h : 0;
    if (gtlc_render(&p) != GTLC_OK) { logError(gtlc_last_error()); return GiStatus::Error; }
    // Gi.cpp:2504-2512: the ClockCycles AOV leaves giRender as a heat map of the per-pixel cycle counts
    for (size_t i = 0; i < bindings.size(); i++)
      if (bindings[i].aovId == GTLC_AOV_CLOCK_CYCLES && params.aovBindings[i].renderBuffer) encodeRenderBufferAsHeatmap(params.aovBindings[i].renderBuffer);
    return GiStatus::Ok;
  }

#define GI_LIGHT_BOILERPLATE(Type, prefix)                                                                           \
  Gi##Type##Light* giCreate##Type##Light(GiScene* scene) { return scene ? wrap<Gi##Type##Light>(gtlc_##prefix##_light_create(scene->h)) : nullptr; } \
  void giDestroy##Type##Light(GiScene* scene, Gi##Type##Light* l) { if (!scene || !l) return; gtlc_##prefix##_light_destroy(scene->h, l->h); delete l; } \
  void giSet##Type##LightBaseEmission(Gi##Type##Light* l, float* rgb) { if (l) gtlc_##prefix##_light_set_base_emission(l->h, rgb); } \
  void giSet##Type##LightDiffuseSpecular(Gi##Type##Light* l, float d, float s) { if (l) gtlc_##prefix##_light_set_diffuse_specular(l->h, d, s); }
  GI_LIGHT_BOILERPLATE(Sphere, sphere)
  GI_LIGHT_BOILERPLATE(Distant, distant)
  GI_LIGHT_BOILERPLATE(Rect, rect)
  GI_LIGHT_BOILERPLATE(Disk, disk)
  void giSetSphereLightPosition(GiSphereLight* l, float* p) { if (l) gtlc_sphere_light_set_position(l->h, p); }
  void giSetSphereLightRadius(GiSphereLight* l, float x, float y, float z) { if (l) gtlc_sphere_light_set_radius(l->h, x, y, z); }
  void giSetDistantLightDirection(GiDistantLight* l, float* d) { if (l) gtlc_distant_light_set_direction(l->h, d); }
  void giSetDistantLightAngle(GiDistantLight* l, float a) { if (l) gtlc_distant_light_set_angle(l->h, a); }
  void giSetRectLightOrigin(GiRectLight* l, float* o) { if (l) gtlc_rect_light_set_origin(l->h, o); }
  void giSetRectLightTangents(GiRectLight* l, float* t0, float* t1) { if (l) gtlc_rect_light_set_tangents(l->h, t0, t1); }
  void giSetRectLightDimensions(GiRectLight* l, float w, float h) { if (l) gtlc_rect_light_set_dimensions(l->h, w, h); }
  void giSetDiskLightOrigin(GiDiskLight* l, float* o) { if (l) gtlc_disk_light_set_origin(l->h, o); }
  void giSetDiskLightTangents(GiDiskLight* l, float* t0, float* t1) { if (l) gtlc_disk_light_set_tangents(l->h, t0, t1); }
  void giSetDiskLightRadius(GiDiskLight* l, float x, float y) { if (l) gtlc_disk_light_set_radius(l->h, x, y); }

  // Dome light: the reference decodes any image format through imgio (out of scope).  The facade reads
  // Portable Float Maps (.pfm, "PF" header, bottom-to-top rows) itself -- through the registered
  // GiAssetReader when there is one -- and reports other formats as an error.
  GiDomeLight* giCreateDomeLight(GiScene* scene, const char* filePath)
  {
    if (!scene || !filePath) return nullptr;
    std::string bytes;
    if (s_assetReader) {
      GiAsset* a = s_assetReader->open(filePath);
      if (a) { bytes.assign((const char*) s_assetReader->data(a), s_assetReader->size(a)); s_assetReader->close(a); }
    }
    if (bytes.empty()) { std::ifstream f(filePath, std::ios::binary); bytes.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>()); }
    int w = 0, h = 0, consumed = 0; float scale = 0.0f; char magic[3] = { 0, 0, 0 };
    if (bytes.size() < 8 || sscanf(bytes.c_str(), "%2s %d %d %f%n", magic, &w, &h, &scale, &consumed) != 4 || strcmp(magic, "PF") != 0 || w <= 0 || h <= 0) {
      logError(std::string("unable to load dome light texture at ") + filePath + " (this build decodes .pfm only; imgio is out of scope)");
      return nullptr;
    }
    size_t off = (size_t) consumed + 1;
    if (bytes.size() < off + (size_t) w * h * 12) { logError("truncated .pfm dome light texture"); return nullptr; }
    if (!(scale < 0.0f)) { logError(std::string("big-endian .pfm dome light texture is not supported: ") + filePath); return nullptr; }   // scale < 0 marks little-endian samples
    std::vector<float> rgb((size_t) w * h * 3);                       // the payload starts at an arbitrary byte offset: copy, do not alias
    memcpy(rgb.data(), bytes.data() + off, rgb.size() * sizeof(float));
    std::vector<float> rgba((size_t) w * h * 4);
    const float* src = rgb.data();
    for (int y = 0; y < h; y++)          // PFM rows run bottom to top; the sampler's v = 0 row is the top of the image
      for (int x = 0; x < w; x++) {
        const float* s = src + ((size_t) (h - 1 - y) * w + x) * 3;
        float* d = &rgba[((size_t) y * w + x) * 4];
        d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 1.0f;
      }
    return wrap<GiDomeLight>(gtlc_dome_light_create(scene->h, (uint32_t) w, (uint32_t) h, rgba.data()));
  }
  void giDestroyDomeLight(GiDomeLight* l) { if (!l) return; gtlc_dome_light_destroy(l->h); delete l; }
  void giSetDomeLightRotation(GiDomeLight* l, float* q) { if (l) gtlc_dome_light_set_rotation(l->h, q); }
  void giSetDomeLightBaseEmission(GiDomeLight* l, float* rgb) { if (l) gtlc_dome_light_set_base_emission(l->h, rgb); }
  void giSetDomeLightDiffuseSpecular(GiDomeLight* l, float d, float s) { if (l) gtlc_dome_light_set_diffuse_specular(l->h, d, s); }

  GiRenderBuffer* giCreateRenderBuffer(uint32_t w, uint32_t h, GiRenderBufferFormat f)
  {
    GiRenderBuffer* rb = wrap<GiRenderBuffer>(gtlc_render_buffer_create(w, h, (int32_t) f));
    if (rb) { rb->width = w; rb->height = h; }
    return rb;
  }
  void giDestroyRenderBuffer(GiRenderBuffer* rb) { if (!rb) return; gtlc_render_buffer_destroy(rb->h); delete rb; }
  void* giGetRenderBufferMem(GiRenderBuffer* rb) { return rb ? gtlc_render_buffer_mem(rb->h) : nullptr; }
}
