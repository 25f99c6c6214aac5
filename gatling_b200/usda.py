"""Tiny `.usda` subset reader -- enough for /root/reference/cornell.usda (BASELINE.json config C1).

The real caller of the gi API is hdGatling on top of OpenUSD (out of scope);
this reader stands in for it in the USD-free harness: Xform-with-matrix > Mesh (points,
faceVertexCounts, faceVertexIndices, faceVarying normals, doubleSided, material:binding),
Material > UsdPreviewSurface shader constants, and one Camera.  Triangulation / de-indexing /
tangents follow hdGatling (see scenes.polygon_mesh).
"""
from __future__ import annotations

import math
import re

import numpy as np

from . import scenes as S

_NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)"


def _numbers(text):
    return [float(x) for x in re.findall(_NUM, text)]


def _block(text, start):
    """Returns the text between the '{' at/after `start` and its matching '}'."""
    i = text.index("{", start)
    depth, j = 0, i
    while True:
        c = text[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[i + 1:j], j + 1
        j += 1


def _attr(body, pattern):
    m = re.search(pattern + r"\s*=\s*(\[[^\]]*\]|\([^=]*?\)\s*\)|\([^)]*\)|[^\n]*)", body)
    return m.group(1) if m else None


def load(path, width=512, height=512, spp=64, max_bounces=4) -> S.SceneDesc:
    with open(path) as f:
        text = f.read()
    materials, mat_names = [], []
    mm = re.search(r'def\s+"Materials"', text)
    if mm:
        mbody, _ = _block(text, mm.end())
        for m in re.finditer(r'def\s+Material\s+"(\w+)"', mbody):
            body, _ = _block(mbody, m.end())
            kw = {}
            for key, name in (("diffuseColor", "diffuse"), ("emissiveColor", "emissive"), ("specularColor", "specular")):
                v = _attr(body, r"inputs:" + key)
                if v:
                    kw[name] = tuple(_numbers(v)[:3])
            for key, name in (("metallic", "metallic"), ("roughness", "roughness"), ("clearcoat", "clearcoat"),
                              ("clearcoatRoughness", "clearcoatRoughness"), ("opacity", "opacity"), ("ior", "ior"),
                              ("useSpecularWorkflow", "useSpecularWorkflow")):
                v = _attr(body, r"inputs:" + key + r"\b")
                if v:
                    kw[name] = _numbers(v)[0]
            materials.append(S.material_usd_preview_surface(m.group(1), **kw))
            mat_names.append(m.group(1))
    meshes, camera = [], None
    for m in re.finditer(r'def\s+Xform\s+"(\w+)"', text):
        if m.group(1) == "Root":
            continue
        body, _ = _block(text, m.end())
        xf = np.eye(4)
        t = _attr(body, r"matrix4d\s+xformOp:transform")
        if t:
            xf = np.array(_numbers(t)[:16], np.float64).reshape(4, 4)
        cm = re.search(r'def\s+Camera\s+"(\w+)"', body)
        if cm:
            cbody, _ = _block(body, cm.end())
            pos = (np.array([0, 0, 0, 1.0]) @ xf)[:3]
            fwd = (np.array([0, 0, -1.0, 0]) @ xf)[:3]
            up = (np.array([0, 1.0, 0, 0]) @ xf)[:3]
            fwd /= np.linalg.norm(fwd)
            up /= np.linalg.norm(up)
            focal = np.float32(_numbers(_attr(cbody, r"float\s+focalLength"))[0]) * np.float32(0.1)
            aperture = np.float32(_numbers(_attr(cbody, r"float\s+verticalAperture"))[0]) * np.float32(0.1)
            clip = _numbers(_attr(cbody, r"float2\s+clippingRange"))
            vfov = np.float32(2.0) * np.float32(math.atan(float(aperture / (np.float32(2.0) * focal))))
            camera = S.Camera(position=tuple(float(np.float32(x)) for x in pos), forward=tuple(float(np.float32(x)) for x in fwd),
                              up=tuple(float(np.float32(x)) for x in up), vfov=float(vfov), focalLength=float(focal),
                              clipStart=clip[0], clipEnd=clip[1])
            continue
        gm = re.search(r'def\s+Mesh\s+"(\w+)"', body)
        if not gm:
            continue
        gbody, _ = _block(body, gm.end())
        counts = [int(x) for x in _numbers(_attr(gbody, r"int\[\]\s+faceVertexCounts"))]
        indices = [int(x) for x in _numbers(_attr(gbody, r"int\[\]\s+faceVertexIndices"))]
        points = np.array(_numbers(_attr(gbody, r"point3f\[\]\s+points")), np.float32).reshape(-1, 3)
        normals = np.array(_numbers(_attr(gbody, r"normal3f\[\]\s+normals")), np.float32).reshape(-1, 3)
        double_sided = bool(int(_numbers(_attr(gbody, r"uniform\s+bool\s+doubleSided") or "0")[0]))
        bind = re.search(r"rel\s+material:binding\s*=\s*<[^>]*/(\w+)>", gbody)
        v, f = S.polygon_mesh(points, counts, indices, normals)
        meshes.append(S.Mesh(v, f, mat_names.index(bind.group(1)), xf.astype(np.float32), None, double_sided, False, len(meshes) + 1, gm.group(1)))
    st = S.RenderSettings(spp=spp, maxBounces=max_bounces)
    return S.SceneDesc("usda:" + path, meshes, materials, camera, st, width, height)
