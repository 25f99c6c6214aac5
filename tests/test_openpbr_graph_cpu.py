"""CPU: the oracle's OpenPBR lowering against the reference's own definition of OpenPBR.

tests/golden/openpbr_graph.npz holds, for 96 seeded parameter sets, the values the reference's node graph
(/root/reference/src/gi/mtlx/open_pbr_surface.mtlx, evaluated node by node in float32 by tests/golden/make_openpbr_graph_vectors.py in the
build container) feeds into its BSDF / EDF / VDF nodes.  The oracle's hit-independent closure constants (orc_openpbr.h opbr_lobes; the CUDA
path's lower_materials_kernel computes the same record) must agree: the graph evaluates powers with pow() where the oracle multiplies, and
associates the coat's two attenuation factors the other way round, so the bar is a few float32 ulp, not bits."""
import ctypes as C
import os

import numpy as np

from gatling_b200 import scenes
from oracle import orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "openpbr_graph.npz")
O = scenes.OPBR


def lowered(params):
    L = orc.lib()
    L.orc_openpbr_lowered.argtypes = [C.c_void_p, C.c_void_p]
    out = np.zeros(32, np.float32)
    p = np.ascontiguousarray(params, np.float32)
    L.orc_openpbr_lowered(p.ctypes.data, out.ctypes.data)
    return out


def close(a, b, ulps=8):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.all(np.abs(a.astype(np.float64) - b) <= ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)) + 1e-12)


def test_lowering_follows_the_reference_graph():
    g = np.load(GOLDEN)
    n = len(g["params"])
    assert n == 96
    branches = {"tir_fix": 0, "no_volume": 0, "no_coat": 0, "shifted_absorption": 0}
    for k in range(n):
        p = g["params"][k]
        o = lowered(p)
        a, a_c, eta = o[0], o[1], o[2]
        m_f0, m_f82, tint, coat_att, Kd, sa, ss, em = o[8:11], o[11:14], o[14:17], o[17:20], o[20:23], o[23:26], o[26:29], o[29:32]
        r = np.float32(g["effective_specular_roughness"][k][0])
        assert close(a, max(r * r, np.float32(1e-4)), 16), (k, a, r * r)                     # mtlx:102-137; open_pbr_anisotropy with anisotropy 0: alpha = r^2
        rc = np.float32(p[O["coat_roughness"]])
        assert close(a_c, max(rc * rc, np.float32(1e-4))), k
        assert close(eta, g["modulated_eta_s"][k][0], 16), (k, eta, g["modulated_eta_s"][k])  # mtlx:306-373
        assert close(m_f0, g["metal_reflectivity"][k]) and close(m_f82, g["metal_edgecolor"][k]), k     # mtlx:434-470
        assert np.array_equal(tint, g["if_transmission_tint"][k]), k                           # mtlx:374-396
        assert close(coat_att, g["modulated_base_darkening"][k] * g["coat_attenuation"][k], 16), (k, coat_att)   # mtlx:472-550 (two BSDF multiplies)
        ssw = np.float32(min(max(p[O["subsurface_weight"]], 0.0), 1.0))                        # mtlx:195-218: mix(diffuse(base_weight base_color), subsurface, subsurface_weight)
        want_kd = g["base_color_nonnegative"][k] * np.float32(p[O["base_weight"]]) * (np.float32(1) - ssw) + g["subsurface_color_nonnegative"][k] * ssw
        assert close(Kd, want_kd), k
        assert close(ss, g["if_volume_scattering"][k]), (k, ss)                                # mtlx:238-245, 286-292
        assert close(sa, g["if_volume_absorption"][k], 64), (k, sa, g["if_volume_absorption"][k])   # mtlx:221-285 (difference of two quotients: more ulps)
        # mtlx:582-623: uniform_edf(emission_weight) mixed by coat_weight with its coat_color-tinted copy; MaterialX's uniform_edf is
        # df::diffuse_edf() with intensity color * PI (the generalized_schlick_edf factor of the coated copy is not modelled, orc_openpbr.h)
        cw = np.float32(min(max(p[O["coat_weight"]], 0.0), 1.0))
        tinted = np.float32(1.0) * (np.float32(1.0) - cw) + np.asarray(p[O["coat_color"]:O["coat_color"] + 3], np.float32) * cw
        assert close(em, g["emission_weight"][k] * np.float32(np.pi) * tinted, 16), (k, em)
        branches["tir_fix"] += p[O["specular_ior"]] / p[O["coat_ior"]] <= 1.0
        branches["no_volume"] += p[O["transmission_depth"]] == 0.0
        branches["no_coat"] += p[O["coat_weight"]] == 0.0
        branches["shifted_absorption"] += bool(np.any(g["if_volume_absorption"][k] == 0.0) and p[O["transmission_depth"]] > 0.0)
    assert all(v > 3 for v in branches.values()), branches                                     # every branch of the graph's ifgreater nodes was exercised
