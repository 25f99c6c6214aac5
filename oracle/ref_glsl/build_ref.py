#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Builds oracle/_ref/libglslref.so from the reference's OWN shader sources, where they lie under
/root/reference/src/gi/shaders, so that tests/test_reference_glsl_cpu.py can hold the oracle's restatements against them.

GLSL is not C++, but the pure functions of these files (no ray-tracing built-ins, no descriptors) become C++ after three
mechanical rewrites, done here on the fly into a temporary directory (nothing derived from the reference's text is written into
the repository; only the shared object is kept, git-ignored):
  * parameter qualifiers:  `in T x` -> `T x`,  `out T x` / `inout T x` -> `T& x`
  * swizzles used as values:  `.xy` `.yx` `.zw` `.xyz` `.rgb` -> `.xy()` ...
  * vector constructors:  `vec4(a, b, c, d)` -> `vec4{a, b, c, d}` (left-to-right evaluation of the arguments, as GLSL specifies)
  * a line range is cut out of the files that also hold Vulkan-only code (rp_main.rgen / .chit / .miss, mdl_interface.glsl)
The types and built-ins come from oracle/ref_glsl/glsl_shim.h; g++ runs with -fsingle-precision-constant (GLSL literals are
floats) and -ffp-contract=off.  A second translation unit (ref_host.cpp, built without that flag) holds line ranges of the reference's
C++ host code, src/gi/impl/Gi.cpp -- the light encoders -- which need no rewriting at all."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/gi/shaders"
OUT = os.path.join(os.path.dirname(HERE), "_ref")

# (source, first line, last line (inclusive, None = whole file), generated name)
FRAGMENTS = [
    ("aovs.glsl", None, None, "aovs.glsl"),
    ("common.glsl", None, None, "common.glsl"),
    ("colormap.glsl", None, None, "colormap.glsl"),
    ("mdl_types.glsl", None, None, "mdl_types.glsl"),
    ("rp_main_payload.glsl", 1, 98, "rp_main_payload.glsl"),
    ("rp_main.rgen", 48, 130, "rgen_functions.inc"),           # sampleDistance .. fisGauss
    ("rp_main.rgen", 191, 288, "rgen_camera.inc"),             # main(): pixel -> camera ray, RNG draws, clip range (leaves the sample loop open)
    ("mdl_shading_state.glsl", None, None, "mdl_shading_state.glsl"),
    ("rp_main.rgen", 132, 521, "rgen_main.inc"),               # clearAovs + main(): the whole ray-generation shader
    ("rp_main.chit", 132, 493, "chit_main.inc"),               # main(): the whole closest-hit shader
    ("rp_main.miss", 15, 87, "miss_all.inc"),                  # stepVolume, quatRotateDir, sampleDomeLight, main()
    ("rp_main_shadow.miss", 14, 17, "shadow_miss_main.inc"),   # main()
    ("rp_main.chit", 29, 131, "chit_sample_light.inc"),        # sampleLight
    ("rp_main.miss", 36, 53, "miss_functions.inc"),            # quatRotateDir, sampleDomeLight
    ("mdl_interface.glsl", 8, 38, "mdl_wrap.inc"),             # apply_wrap_and_crop
    ("mdl_interface.glsl", 238, 256, "mdl_adapt_normal.inc"),  # mdl_adapt_normal
    ("mdl_interface.glsl", 40, 205, "mdl_textures.inc"),       # tex_lookup_* / tex_texel_* for 2-D and 3-D textures
    ("mdl_interface.glsl", 266, 452, "mdl_scene_data.inc"),    # scene_data_isvalid, get_scene_data_indices, scene_data_lookup_float{,2,3,4} / _color / _int
    ("mdl_renderer_state.glsl", None, None, "mdl_renderer_state.glsl"),
]


def brace_constructors(text: str) -> str:
    """`vec4(a, b, c, d)` -> `vec4{a, b, c, d}`: GLSL evaluates constructor arguments left to right (the RNG draws of rng1d_next4f
    depend on it), C++ only promises that for braced initialisers."""
    out, i = [], 0
    pat = re.compile(r"\b(?:vec[234]|ivec[234]|uvec[234])\(")
    while True:
        m = pat.search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        j, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        out.append(text[i:m.end() - 1] + "{" + brace_constructors(text[m.end():j - 1]) + "}")
        i = j


def translate(text: str) -> str:
    text = re.sub(r"(?<=[(,])\s*in\s+(?=\w+\s+\w+\s*[,)])", " ", text)
    text = re.sub(r"(?<=[(,])\s*(?:inout|out)\s+(\w+)\s+(\w+)(?=\s*[,)])", r" \1& \2", text)
    text = re.sub(r"\.(xyz|rgb|xy|yx|zw)\b(?!\s*\()", r".\1()", text)
    return brace_constructors(text)


# C++ of the reference's host side (no rewriting): (path below /root/reference/src, first line, last line, generated name)
HOST_FRAGMENTS = [
    ("gi/impl/Gi.cpp", 287, 300, "gi_encode_direction.inc"),   # _EncodeOctahedral, _EncodeDirection
    ("gi/impl/Gi.cpp", 2573, 2941, "gi_lights.inc"),           # gi{Create,Destroy,Set}{Sphere,Distant,Rect,Disk}Light*
    ("gi/impl/Gi.cpp", 2375, 2425, "gi_uniforms.inc"),         # giRender: camera vectors, lens radius, the UniformData initialiser
    ("gi/impl/Gi.cpp", 845, 859, "gi_pack_vertices.inc"),      # _giBuildGeometryStructures: GiVertex -> rp::FVertex (the loop over a mesh's vertices)
    ("hdGatling/main.cpp", 136, 143, "harness_srgb.inc"),      # _AccurateLinearToSrgb
    ("hdGatling/mesh.cpp", 83, 86, "mesh_bitangent_sign.inc"), # _CalculateBitangentSign
    ("hdGatling/mesh.cpp", 231, 261, "mesh_fallback_tangents.inc"),  # _DuffOrthonormalBasis, _CalculateFallbackTangents
    ("hdGatling/main.cpp", 465, 487, "harness_bytes.inc"),     # the graphical tests' float -> 8-bit RGBA loop
]


def build() -> str:
    if not os.path.isdir(REF):
        raise SystemExit("build_ref: %s does not exist (the reference sources are only present in the build container)" % REF)
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="glslref_")
    try:
        for src, a, b, name in FRAGMENTS:
            with open(os.path.join(REF, src)) as f:
                lines = f.read().split("\n")
            if a is not None:
                lines = lines[a - 1:b]
            with open(os.path.join(tmp, name), "w") as f:
                f.write(translate("\n".join(lines)) + "\n")
        for src, a, b, name in HOST_FRAGMENTS:
            with open(os.path.join(os.path.dirname(os.path.dirname(REF)), src)) as f:
                lines = f.read().split("\n")
            with open(os.path.join(tmp, name), "w") as f:
                f.write("\n".join(lines[a - 1:b]) + "\n")
        so = os.path.join(OUT, "libglslref.so")
        src_root = os.path.dirname(os.path.dirname(REF))
        common = ["g++", "-std=c++20", "-fopenmp", "-I", os.path.join(src_root, "gi"), "-I", os.path.join(src_root, "gb"), "-O1", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w", "-I", tmp, "-I", HERE, "-I", REF, "-c"]
        objs = []
        for unit, extra in (("ref_exports.cpp", ["-fsingle-precision-constant"]),      # GLSL: every literal is a float
                            ("ref_host.cpp", [])):                                      # C++ host code: float / double as written
            obj = os.path.join(tmp, unit + ".o")
            r = subprocess.run(common + extra + [os.path.join(HERE, unit), "-o", obj], capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stderr[-6000:])
                raise SystemExit("build_ref: g++ failed on " + unit)
            objs.append(obj)
        r = subprocess.run(["g++", "-shared", "-fopenmp", "-o", so] + objs + ["-lm"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-3000:])
            raise SystemExit("build_ref: link failed")
        return so
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build())
