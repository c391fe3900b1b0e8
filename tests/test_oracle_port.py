"""Pins oracle/pbr_oracle.cc (the CPU restatement of the reference path) against the golden vectors generated from
the compiled, unmodified reference (tests/golden/make_golden.py): an oracle that has not passed these is not an
oracle.  CPU only."""
import os

import numpy as np
import pytest

import checks
import common
from conftest import GOLDEN, golden


@pytest.fixture(scope="module")
def cornell_oracle(cornell_host):
    import oraclebind
    if not oraclebind.available():
        pytest.skip("oracle/libpbr_oracle.so not built (make -C oracle port)")
    return oraclebind.Oracle(cornell_host.flat())


@pytest.fixture(scope="module")
def hair_oracle(hair_host):
    import oraclebind
    if not oraclebind.available():
        pytest.skip("oracle/libpbr_oracle.so not built (make -C oracle port)")
    return oraclebind.Oracle(hair_host.flat())


def test_rays_vs_embree_golden(cornell_oracle):
    """hit / primID agreement with Embree >= 99.99 %, t within 1e-5 relative, u/v/Ng, occlusion"""
    agree = checks.check_rays(cornell_oracle, golden("cornell_rays.npz"))
    assert agree >= 0.9999


def test_paths_vs_reference_golden(cornell_oracle):
    """GetRadiance() per path with the reference's PCG32 streams: Lambert, GGX, clearcoat fall-through, NEE + MIS,
    roulette and the random walk (incl. g++'s right-to-left draw order) all have to match for a path to agree"""
    frac = checks.check_radiance(cornell_oracle, golden("cornell_paths.npz"), min_agree=0.99)
    assert frac >= 0.99


def test_hair_scene_vs_reference_golden(hair_oracle):
    g = golden("hair_scene.npz")
    rays = common.rays_from_f8(g["rays"])
    hits = hair_oracle.trace(rays)
    ids = g["hit_ids"]; f = g["hit_f"]
    same = (hits["instance_id"] == ids[:, 0]) & (hits["prim_id"] == ids[:, 2])
    assert same.mean() >= 0.999
    curve = same & (ids[:, 0] == 9)
    assert curve.sum() > 3000
    assert np.all(np.abs(hits["t"][curve] - f[curve, 0]) <= 2e-5 * np.abs(f[curve, 0]))
    assert np.abs(hits["v"][curve] - f[curve, 2]).max() < 2e-3          # the hair BSDF's h
    assert np.abs(hits["normal_g"][curve] - f[curve, 3:6]).max() < 1e-3  # tangent
    assert (hair_oracle.occluded(rays) == g["occluded"]).mean() >= 0.999
    frac = common.path_agreement(hair_oracle.radiance(rays, g["seeds"]), g["radiance"], rel=1e-3)
    assert frac >= 0.97, frac


def test_image_mean_vs_reference_render(cornell_oracle):
    """Render(): mean linear RGB of a short render against the reference's 4096-spp image (loose: 16 spp noise)"""
    if not os.path.exists(os.path.join(GOLDEN, "cornell_image_256.npz")):
        pytest.skip("image fixture not generated")
    g = golden("cornell_image_256.npz")
    ref = g["mean_8192"].reshape(-1, 3).mean(axis=0)
    rgba, count, sec, rays = cornell_oracle.render(256, 256, 16, seed=7)
    assert np.all(count == 16) and np.all(rgba[..., 3] == 16.0) and not np.isnan(rgba).any()
    mean = (rgba[..., :3] / 16.0).mean(axis=(0, 1))
    assert np.all(np.abs(mean - ref) <= 0.03 * ref), (mean, ref)
    assert 3.0 < rays.sum() / (256 * 256 * 16) < 9.0   # ~6 rays per sample on this scene
