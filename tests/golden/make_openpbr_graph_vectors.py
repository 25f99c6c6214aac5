#!/usr/bin/env python
"""Golden vectors from the reference's own OpenPBR definition, /root/reference/src/gi/mtlx/open_pbr_surface.mtlx (the node graph the
reference hands to MaterialX's MDL generator): a small interpreter evaluates every scalar / colour node of the graph in float32 for seeded
random parameter sets and stores what the graph feeds into its BSDF / EDF / VDF nodes.  tests/test_openpbr_graph_cpu.py compares the
oracle's lowering (orc_openpbr.h opbr_lobes, orc_bsdf.h volume coefficients) with these values.  Run in the build container (the reference
tree is not present anywhere else); the output tests/golden/openpbr_graph.npz is committed.

Node semantics: MaterialX stdlib (multiply / add / subtract / divide component-wise with scalar broadcast; mix = bg (1 - m) + fg m;
ifgreater = value1 > value2 ? in1 : in2; convert float -> vector broadcast; ln, power, sqrt, sign, min, max, clamp, extract)."""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gatling_b200 import scenes  # noqa: E402

MTLX = "/root/reference/src/gi/mtlx/open_pbr_surface.mtlx"
MATH_TYPES = ("float", "color3", "vector3", "boolean", "integer")
F = np.float32
# node outputs worth freezing: everything that enters a BSDF / EDF / VDF node or scales one
KEEP = ["effective_specular_roughness", "modulated_eta_s", "metal_reflectivity", "metal_edgecolor", "if_transmission_tint", "Kcoat", "Ebase", "base_darkening",
        "modulated_base_darkening", "coat_attenuation", "emission_weight", "if_volume_absorption", "if_volume_scattering", "base_color_nonnegative",
        "subsurface_color_nonnegative", "coat_ior_to_F0", "eta_s", "thin_film_thickness_nm", "one_minus_coat_F0"]


def parse_value(text, typ):
    if typ == "boolean":
        return F(1.0 if text.strip() == "true" else 0.0)
    v = np.array([float(x) for x in text.replace(" ", "").split(",")], F)
    return v[0] if typ in ("float", "integer") else (np.repeat(v, 3) if len(v) == 1 else v)


def vec(x):
    x = np.asarray(x, F)
    return np.repeat(x, 3) if x.ndim == 0 else x


def evaluate(graph, iface):
    """iface: interface name -> float32 scalar or 3-vector.  Returns {node name: value} for every math node."""
    out = {}
    nodes = {n.get("name"): n for n in graph}

    def arg(node, name, default=None):
        for i in node:
            if i.get("name") == name:
                if i.get("nodename"):
                    if i.get("nodename") not in out:
                        compute(nodes[i.get("nodename")])          # the file does not list its nodes in dependency order
                    return out[i.get("nodename")]
                if i.get("interfacename"):
                    return iface[i.get("interfacename")]
                return parse_value(i.get("value"), i.get("type"))
        return default

    def compute(n):
        t, typ = n.tag, n.get("type")
        wide = typ in ("color3", "vector3")
        if t in ("multiply", "add", "subtract", "divide", "min", "max", "power"):
            a, b = arg(n, "in1"), arg(n, "in2")
            if wide:
                a, b = vec(a), vec(b)
            r = {"multiply": lambda: a * b, "add": lambda: a + b, "subtract": lambda: a - b, "divide": lambda: a / b, "min": lambda: np.minimum(a, b),
                 "max": lambda: np.maximum(a, b), "power": lambda: np.power(a, b)}[t]()
        elif t == "mix":
            fg, bg, m = arg(n, "fg"), arg(n, "bg"), arg(n, "mix")
            if wide:
                fg, bg = vec(fg), vec(bg)
            r = bg * (F(1.0) - m) + fg * m
        elif t == "convert":
            r = arg(n, "in")
            r = vec(r) if wide else F(r)
        elif t == "ln":
            r = np.log(arg(n, "in"))
        elif t == "sqrt":
            r = np.sqrt(arg(n, "in"))
        elif t == "sign":
            r = np.sign(arg(n, "in"))
        elif t == "clamp":
            r = np.minimum(np.maximum(arg(n, "in"), arg(n, "low")), arg(n, "high"))
        elif t == "extract":
            r = arg(n, "in")[int(arg(n, "index"))]
        elif t == "ifgreater":
            a, b = arg(n, "in1"), arg(n, "in2")
            if wide:
                a, b = vec(a), vec(b)
            r = a if arg(n, "value1") > arg(n, "value2") else b
        else:
            raise SystemExit("node type not interpreted: " + t)
        out[n.get("name")] = np.asarray(r, F)

    with np.errstate(divide="ignore", invalid="ignore"):           # depth 0 divides by zero on purpose: the graph selects around it (ifgreater)
        for n in graph:
            if n.get("type") in MATH_TYPES and n.get("name") not in out:
                compute(n)
    return out


# interface name -> (index into the OpenPBR parameter block, components); names as in the nodedef of the same file
O = scenes.OPBR
IFACE = {"base_weight": ("base_weight", 1), "base_color": ("base_color", 3), "base_diffuse_roughness": ("base_diffuse_roughness", 1), "base_metalness": ("base_metalness", 1),
         "specular_weight": ("specular_weight", 1), "specular_color": ("specular_color", 3), "specular_roughness": ("specular_roughness", 1), "specular_ior": ("specular_ior", 1),
         "specular_roughness_anisotropy": ("specular_roughness_anisotropy", 1), "transmission_weight": ("transmission_weight", 1), "transmission_color": ("transmission_color", 3),
         "transmission_depth": ("transmission_depth", 1), "transmission_scatter": ("transmission_scatter", 3), "transmission_scatter_anisotropy": ("transmission_scatter_anisotropy", 1),
         "coat_weight": ("coat_weight", 1), "coat_color": ("coat_color", 3), "coat_roughness": ("coat_roughness", 1), "coat_ior": ("coat_ior", 1), "coat_darkening": ("coat_darkening", 1),
         "fuzz_weight": ("fuzz_weight", 1), "fuzz_color": ("fuzz_color", 3), "fuzz_roughness": ("fuzz_roughness", 1), "emission_luminance": ("emission_luminance", 1),
         "emission_color": ("emission_color", 3), "thin_film_weight": ("thin_film_weight", 1), "thin_film_thickness": ("thin_film_thickness", 1), "thin_film_ior": ("thin_film_ior", 1),
         "subsurface_weight": ("subsurface_weight", 1), "subsurface_color": ("subsurface_color", 3), "geometry_thin_walled": ("geometry_thin_walled", 1)}
DEFAULTS = {"subsurface_radius": vec(F(1.0)), "subsurface_radius_scale": np.array([1.0, 0.5, 0.25], F), "subsurface_scatter_anisotropy": F(0.0), "coat_roughness_anisotropy": F(0.0)}


def random_params(rng, k):
    """One OpenPBR parameter block (scenes.material_openpbr layout) inside the ranges the spec allows; every 4th has depth 0 (no volume)."""
    m = scenes.material_openpbr("g")
    p = np.array(m.params, F)
    u = lambda lo=0.0, hi=1.0: F(rng.uniform(lo, hi))
    for name in ("base_weight", "base_metalness", "specular_weight", "specular_roughness", "transmission_weight", "coat_weight", "coat_roughness", "coat_darkening", "subsurface_weight"):
        p[O[name]] = u()
    for name in ("base_color", "specular_color", "coat_color", "subsurface_color", "emission_color"):
        p[O[name]:O[name] + 3] = rng.uniform(0.0, 1.0, 3)
    p[O["transmission_color"]:O["transmission_color"] + 3] = rng.uniform(0.05, 1.0, 3)
    p[O["transmission_scatter"]:O["transmission_scatter"] + 3] = rng.uniform(0.0, 0.8, 3) * (k % 3 != 0)
    p[O["specular_ior"]] = u(1.05, 2.5); p[O["coat_ior"]] = u(1.05, 2.2)
    p[O["transmission_depth"]] = F(0.0) if k % 4 == 0 else u(0.05, 3.0)
    p[O["emission_luminance"]] = u(0.0, 10.0); p[O["thin_film_thickness"]] = u(0.0, 1.0)
    if k % 5 == 0:
        p[O["coat_weight"]] = F(0.0)
    if k % 7 == 0:
        p[O["coat_ior"]] = p[O["specular_ior"]] * F(1.3)               # specular_ior / coat_ior < 1: the TIR fix branch
    return p


def main():
    graph = [e for e in ET.parse(MTLX).getroot().iter("nodegraph")][0]
    rng = np.random.default_rng(2024)
    K = 96
    params = np.stack([random_params(rng, k) for k in range(K)])
    rows = {name: [] for name in KEEP}
    for p in params:
        iface = dict(DEFAULTS)
        for name, (key, n) in IFACE.items():
            iface[name] = F(p[O[key]]) if n == 1 else np.array(p[O[key]:O[key] + 3], F)
        vals = evaluate(graph, iface)
        for name in KEEP:
            rows[name].append(np.atleast_1d(vals[name]))
    np.savez_compressed(os.path.join(HERE, "openpbr_graph.npz"), params=params, **{k: np.stack(v).astype(F) for k, v in rows.items()})
    print("wrote openpbr_graph.npz:", K, "parameter sets,", len(KEEP), "node outputs from", MTLX)


if __name__ == "__main__":
    main()
