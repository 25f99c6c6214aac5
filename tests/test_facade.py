"""The C++ facade (include/gtl/gi/Gi.h + gatling_b200/gi_cpp/Gi.cpp): builds with plain g++ against the C-ABI,
lowers MaterialX on the CPU, and -- on the GPU -- renders bit-identically to the same scene driven through
the Python binding."""
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from gatling_b200 import build as B
from gatling_b200 import scenes

EXE = os.path.join(os.path.dirname(__file__), "cpp", "gi_facade_main")


def _ensure_built():
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc") or os.path.exists(B.LIB):
        B.build_facade()
    assert os.path.exists(EXE)


def test_facade_builds_and_lowers_materialx():
    _ensure_built()
    out = subprocess.run([EXE, "lower"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "lower ok" in out.stdout


def test_facade_object_lifecycle_without_gpu():
    """Meshes with primvars, instancer primvars, instance ids, materials with primvar readers, lights: everything Hydra's
    prim sync does before the first giRender works on the host and reaches the C-ABI with fully initialised structs."""
    _ensure_built()
    out = subprocess.run([EXE, "objects"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "objects ok" in out.stdout


def test_facade_header_is_self_contained(tmp_path):
    """hdGatling only sees <gtl/gi/Gi.h>: it must compile on its own and keep the reference's aggregate layouts."""
    src = tmp_path / "t.cpp"
    src.write_text('#include <gtl/gi/Gi.h>\nusing namespace gtl;\nstatic_assert(sizeof(GiVertex) == 48, "");\nstatic_assert(sizeof(GiFace) == 12, "");\n'
                   'static_assert(int(GiAovId::COUNT) == 17, "");\nstatic_assert(sizeof(GiCameraDesc) == 64, "");\nint main() { GiRenderSettings s{}; (void) s; return 0; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++20", "-fsyntax-only", "-I" + inc, str(src)])


def _python_twin(textured=False):
    def vert(p, n):
        v = np.zeros(1, scenes.VERTEX_DTYPE)
        v["pos"], v["norm"] = p, n
        t, _ = scenes.duff_basis(np.array([n], np.float32))
        v["tangent"] = t[0]
        v["bitangentSign"] = 1.0
        return v
    fv = np.concatenate([vert(p, (0, 0, 1)) for p in [(-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0)]])
    lv = np.concatenate([vert(p, (0, 0, -1)) for p in [(-0.3, -0.3, 1.5), (-0.3, 0.3, 1.5), (0.3, 0.3, 1.5), (0.3, -0.3, 1.5)]])
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    inst = np.stack([np.eye(4, dtype=np.float32), np.array([[0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 0.5, 0], [1.2, 0.4, 0.6, 1]], np.float32)])
    mats = [scenes.material_usd_preview_surface("Red", diffuse=(0.9, 0.1, 0.05), roughness=0.35, metallic=0.25),
            scenes.material_usd_preview_surface("Light", diffuse=(0, 0, 0), emissive=(6, 5, 4))]
    meshes = [scenes.Mesh(fv, faces, 0, object_id=1, name="floor"), scenes.Mesh(lv, faces, 1, instance_transforms=inst, object_id=2, name="light")]
    cam = scenes.Camera(position=(0, -4, 1.5), forward=(0, 0.9701425, -0.2425356), up=(0, 0.2425356, 0.9701425), vfov=0.7, focalLength=0.05)
    st = scenes.RenderSettings(maxBounces=5, nextEventEstimation=True, spp=4)
    d = scenes.SceneDesc("twin", meshes, mats, cam, st, 48, 32, clear_color=(0.25, 0.5, 1.0, 1.0))
    d.sphere_lights.append(scenes.SphereLight(position=(-0.8, 0.2, 0.7), base_emission=(3, 3, 5), radius=(0.1, 0.1, 0.1)))
    if textured:      # what tests/cpp/gi_facade_main.cpp::texturedUps lowers to; data.pfm / n.ppm do not exist -> those bindings are dropped
        fv["u"] = (fv["pos"][:, 0] + 1) * 0.5
        fv["v"] = (fv["pos"][:, 1] + 1) * 0.5
        d.textures = [scenes.procedural_texture("checker")]
        mats[0] = scenes.material_usd_preview_surface("Red", metallic=0.25)
        mats[0].textures = [scenes.TextureBinding(scenes.UPS["diffuse"], 0, (0, 1, 2), scenes.WRAP_MIRRORED_REPEAT, scenes.WRAP_CLIP, True,
                                                  fallback=(0.05, 0.6, 0.1, 1.0), uv_rotation=30.0, uv_scale=(3.0, 2.0), uv_translation=(0.13, -0.4))]
    return d


def write_ppm(path, texture):
    t = texture.texels[::-1, :, :3]          # PPM rows run top to bottom; Texture rows bottom to top
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (t.shape[1], t.shape[0]))
        f.write(np.ascontiguousarray(t).tobytes())


@pytest.mark.gpu
def test_facade_renders_like_the_binding(gpu, tmp_path):
    _ensure_built()
    out = str(tmp_path / "img.bin")
    res = subprocess.run([EXE, "render", out], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    cpp = np.fromfile(out, np.float32).reshape(32, 48, 4)
    d = _python_twin()
    b = gpu.BuiltScene(d)
    try:
        py = b.render().copy()
    finally:
        b.close()
    assert np.array_equal(cpp, py), "C++ facade and Python binding disagree (max diff %g)" % np.abs(cpp - py).max()
    from oracle import orc
    ref, _, _ = orc.OracleScene(d).render()
    assert np.sqrt(np.mean((py[..., :3].astype(np.float64) - ref[..., :3]) ** 2)) < 1e-3


@pytest.mark.gpu
def test_facade_textured_material_renders_like_the_binding(gpu, tmp_path):
    """MaterialX UsdUVTexture + UsdTransform2d network -> image file read by the facade -> texture binding: identical to
    the same binding made by hand through the Python harness."""
    _ensure_built()
    tex = str(tmp_path / "checker.ppm")
    write_ppm(tex, scenes.procedural_texture("checker"))
    out = str(tmp_path / "img.bin")
    res = subprocess.run([EXE, "render", out, tex], capture_output=True, text=True, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout + res.stderr
    assert "failed to read image data.pfm" in res.stderr          # missing files drop their binding, loudly
    cpp = np.fromfile(out, np.float32).reshape(32, 48, 4)
    d = _python_twin(textured=True)
    b = gpu.BuiltScene(d)
    try:
        py = b.render().copy()
    finally:
        b.close()
    assert np.array_equal(cpp, py), "C++ facade and Python binding disagree (max diff %g)" % np.abs(cpp - py).max()
    plain = gpu.BuiltScene(_python_twin())
    try:
        assert np.abs(plain.render() - py).mean() > 1e-3
    finally:
        plain.close()
