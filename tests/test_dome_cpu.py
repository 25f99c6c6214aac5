"""Dome light (R11, rp_main.miss:39-87): lat-long mapping, quaternion rotation and the camera-invisible fallback, checked
against an independent numpy restatement on a dome that is NOT constant -- a two-axis gradient with one hot texel -- so that
a swapped axis, a flipped v, a wrong wrap or an ignored rotation changes pixels."""
import dataclasses

import numpy as np

from gatling_b200 import scenes
from oracle import orc


def gradient_dome(w=16, h=8, hot=None):
    t = np.zeros((h, w, 4), np.float32)
    x = (np.arange(w, dtype=np.float32) + 0.5) / w
    y = (np.arange(h, dtype=np.float32) + 0.5) / h
    t[..., 0] = x[None, :]                    # red grows with u (longitude)
    t[..., 1] = y[:, None]                    # green grows with v (row index)
    t[..., 2] = 0.25
    hy, hx = hot if hot is not None else (h // 4, (3 * w) // 4)
    t[hy, hx, :3] = (6.0, 3.0, 9.0)           # one hot texel off every symmetry axis
    t[..., 3] = 1.0
    return t


def quat_rotate(q, d):
    """rp_main.miss:39-44 (GLM's cross-product form)."""
    qv = np.asarray(q[:3], np.float64)
    a = np.cross(qv, d)
    b = np.cross(qv, a)
    return d + (a * q[3] + b) * 2.0


def latlong_lookup(texels, d):
    """rp_main.miss:46-53 + a linear/linear REPEAT sampler (Gi.cpp:388-392), float64."""
    h, w = texels.shape[:2]
    u = (np.arctan2(d[:, 2], d[:, 0]) + 0.5 * np.pi) / (2.0 * np.pi)
    v = 1.0 - np.arccos(np.clip(d[:, 1], -1, 1)) / np.pi
    x = u * w - 0.5
    y = v * h - 0.5
    x0 = np.floor(x); y0 = np.floor(y)
    ax = (x - x0)[:, None]; ay = (y - y0)[:, None]
    xa = np.mod(x0.astype(np.int64), w); xb = np.mod(x0.astype(np.int64) + 1, w)
    ya = np.mod(y0.astype(np.int64), h); yb = np.mod(y0.astype(np.int64) + 1, h)
    t = texels.astype(np.float64)[..., :3]
    top = t[ya, xa] * (1 - ax) + t[ya, xb] * ax
    bot = t[yb, xa] * (1 - ax) + t[yb, xb] * ax
    return top * (1 - ay) + bot * ay


def empty_scene(w, h, dome, **settings):
    d = scenes.cornell(w, h, spp=1)
    d.meshes = []
    d.camera = dataclasses.replace(d.camera, position=(0.0, 0.0, 0.0), forward=(0.3, 0.2, -1.0), up=(0.0, 1.0, 0.0), vfov=1.6)
    d.settings = dataclasses.replace(d.settings, jitteredSampling=False, maxSampleValue=100.0, **settings)
    d.dome_light = dome
    return d


def test_dome_mapping_against_numpy():
    q = np.array([0.18, -0.37, 0.11, 0.0]); q[3] = np.sqrt(1.0 - (q[:3] ** 2).sum())          # unit quaternion, all axes
    # the hot texel sits where the rotated view direction lands, so it is in the picture
    f = np.array([[0.3, 0.2, -1.0]]); f /= np.linalg.norm(f)
    rf = quat_rotate(q, f); rf /= np.linalg.norm(rf)
    hu = (np.arctan2(rf[0, 2], rf[0, 0]) + 0.5 * np.pi) / (2.0 * np.pi); hv = 1.0 - np.arccos(rf[0, 1]) / np.pi
    dome = scenes.DomeLight(gradient_dome(hot=(int(hv * 8) % 8, int(np.floor(hu * 16)) % 16)), rotation=tuple(float(x) for x in q), base_emission=(0.5, 2.0, 1.5))
    d = empty_scene(48, 32, dome)
    img, _, rays = orc.OracleScene(d).render()
    dirs = orc.camera_rays(d, 0)[:, 4:7].astype(np.float64)
    rot = quat_rotate(q, dirs)
    rot /= np.linalg.norm(rot, axis=1, keepdims=True)
    want = latlong_lookup(dome.texels, rot) * np.array(dome.base_emission)
    got = img[..., :3].reshape(-1, 3)
    assert np.abs(got - want).max() < 2e-4, np.abs(got - want).max()
    assert got.max() > 2.0                            # the hot texel is in view: the comparison is not over a flat image
    # the rotation matters: the unrotated lookup differs grossly
    plain = latlong_lookup(dome.texels, dirs / np.linalg.norm(dirs, axis=1, keepdims=True)) * np.array(dome.base_emission)
    assert np.abs(plain - want).max() > 0.5


def test_dome_identity_rotation_axes():
    """+x maps to u = 0.25, +z to u = 0.5, towards +y v grows to 1 (the LAST texel rows): pins the orientation convention."""
    t = gradient_dome(64, 32)
    d = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, np.cos(np.radians(10.0)), np.sin(np.radians(10.0))]], np.float64)
    c = latlong_lookup(t, d / np.linalg.norm(d, axis=1, keepdims=True))
    assert abs(c[0, 0] - 0.25) < 1e-6 and abs(c[1, 0] - 0.5) < 1e-6
    assert abs(c[2, 1] - (1.0 - 10.0 / 180.0)) < 1e-6


def test_dome_camera_invisible_uses_the_fallback_for_primary_rays_only():
    """rp_main.miss:71-82: without DOME_LIGHT_CAMERA_VISIBLE a primary ray sees texture 0 (the clear colour, quantised to 8
    bits, Gi.cpp:2183-2199) while bounce rays still see the dome."""
    dome = scenes.DomeLight(gradient_dome(), rotation=(0.0, 0.3826834, 0.0, 0.9238795), base_emission=(2.0, 2.0, 2.0))
    d = empty_scene(16, 8, dome, domeLightCameraVisible=False)
    d.clear_color = (0.2, 0.4, 0.6, 1.0)
    img, _, _ = orc.OracleScene(d).render()
    assert np.allclose(img[..., :3], 2.0 * np.array([51, 102, 153]) / 255.0, atol=1e-6)     # miss:84: the multiplier applies to texture 0 too
    # with a floor under the camera the floor pixels are lit by the dome (bounce rays), the sky pixels stay the fallback
    c = scenes.cornell(16, 16, spp=8, max_bounces=3)
    floor = [m for m in c.meshes if True][0:1]
    c.meshes = floor
    c.dome_light = dome
    c.clear_color = (0.0, 0.0, 0.0, 1.0)
    c.settings = dataclasses.replace(c.settings, domeLightCameraVisible=False)
    lit, _, _ = orc.OracleScene(c).render()
    c.dome_light = None
    dark, _, _ = orc.OracleScene(c).render()
    assert lit[..., :3].sum() > dark[..., :3].sum() + 1.0
