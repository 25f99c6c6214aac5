"""CPU: the oracle's restatement of the medium stack / volume random walk (rp_main.rgen:48-97,317-346,462-477,
rp_main.miss:16-34, rp_main.chit:160-186,447-480) behaves like a participating medium."""
import dataclasses

import numpy as np

from gatling_b200 import scenes
from oracle import orc


def _render(desc):
    img, _, rays = orc.OracleScene(desc).render()
    assert np.isfinite(img).all()
    return img, rays


def test_medium_stack_changes_the_image_and_is_deterministic():
    imgs = {}
    for n in (0, 1, 2):
        d = scenes.volume_cornell(40, 40, spp=4, medium_stack_size=n)
        imgs[n], rays = _render(d)
        again, rays2 = _render(d)
        assert np.array_equal(imgs[n], again) and rays == rays2
    assert np.abs(imgs[0] - imgs[1]).mean() > 1e-3      # Beer-only toggle vs scattering walk
    assert np.abs(imgs[1] - imgs[2]).mean() > 1e-4      # the nested core needs a second stack level


def test_scattering_without_absorption_conserves_energy():
    """A closed, non-absorbing, index-matched scattering sphere in a uniform white environment must stay white:
    every volume step multiplies the throughput by sigma_s T / (pdf . sigma_t T) = albedo = 1 in expectation (miss:16-34)."""
    d = scenes.volume_cornell(32, 32, spp=16, medium_stack_size=1, nee=False, max_bounces=200)
    d.meshes = [m for m in d.meshes if m.name == "sphere2"]
    m = scenes.material_openpbr("Fog", specular_weight=0.0, specular_ior=1.0, transmission_weight=1.0, transmission_color=(1.0, 1.0, 1.0),
                                transmission_depth=0.2, transmission_scatter=(0.6, 0.6, 0.6))
    # extinction = -ln(1)/depth = 0 -> absorption = -scatter, shifted to 0 (open_pbr_surface.mtlx:263-279): sigma_t = sigma_s
    d.materials = list(d.materials)
    d.materials[d.meshes[0].material] = m
    d.settings = dataclasses.replace(d.settings, rrBounceOffset=1000, maxSampleValue=1e9, maxVolumeWalkLength=4095)
    img, _ = _render(d)
    assert abs(img[..., :3].mean() - 1.0) < 0.02, img[..., :3].mean()
    # with real absorption the sphere darkens
    d.materials[d.meshes[0].material] = scenes.material_openpbr("Ink", specular_weight=0.0, specular_ior=1.0, transmission_weight=1.0,
                                                               transmission_color=(0.2, 0.2, 0.2), transmission_depth=0.2, transmission_scatter=(0.6, 0.6, 0.6))
    dark, _ = _render(d)
    assert dark[..., :3].mean() < img[..., :3].mean() - 0.01
