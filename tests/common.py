"""Shared helpers and parameter sets for the tests and the golden generator."""
import numpy as np

import pbrlab_b200 as pb

# (alpha_x, alpha_y, distrib): isotropic GGX, near-delta (Monkey: roughness 0.01 -> alpha 1e-4), anisotropic, clearcoat
GGX_CASES = [(0.04, 0.04, 2), (1e-4, 1e-4, 2), (0.25, 0.25, 2), (0.3, 0.05, 2), (0.0009, 0.0009, 1), (0.09, 0.09, 1),
             (1.0, 1.0, 1)]


def principled(**kw):
    """23 floats in CyclesPrincipledBsdfParameter order with the reference defaults (src/material-param.h:24-49)."""
    d = dict(base_color=(0.8, 0.8, 0.8), subsurface=0.0, subsurface_radius=(1.0, 1.0, 1.0),
             subsurface_color=(0.7, 0.1, 0.1), metallic=0.0, specular=0.5, specular_tint=0.0, roughness=0.5,
             anisotropic=0.0, anisotropic_rotation=0.0, sheen=0.0, sheen_tint=0.5, clearcoat=0.0,
             clearcoat_roughness=0.03, ior=1.45, transmission=0.0, transmission_roughness=0.0)
    d.update(kw)
    out = []
    for k in ["base_color", "subsurface", "subsurface_radius", "subsurface_color", "metallic", "specular",
              "specular_tint", "roughness", "anisotropic", "anisotropic_rotation", "sheen", "sheen_tint", "clearcoat",
              "clearcoat_roughness", "ior", "transmission", "transmission_roughness"]:
        v = d[k]
        out.extend(v if isinstance(v, (tuple, list)) else [v])
    return np.array(out, np.float32)


PRINCIPLED_CASES = [
    principled(),                                                                      # defaults
    principled(base_color=(0.8, 0.5, 0.2), specular=1.0, roughness=0.01),              # Monkey
    principled(base_color=(1.0, 0.8, 0.8), subsurface=1.0, subsurface_radius=(1.0, 0.2, 0.1),
               subsurface_color=(1.0, 0.8, 0.8), specular=1.0, roughness=0.2),         # Lucy
    principled(base_color=(0.0, 0.0, 0.0), specular=0.0),                              # Light: no closure at all
    principled(base_color=(0.2, 0.2, 0.8), specular=0.0),                              # diffuse wall
    principled(metallic=1.0, roughness=0.3, anisotropic=0.8, base_color=(0.9, 0.6, 0.2)),
    principled(clearcoat=1.0, clearcoat_roughness=0.1, specular_tint=0.5, base_color=(0.6, 0.1, 0.1)),
    principled(subsurface=0.5, subsurface_radius=(0.5, 0.0, 0.3), transmission=0.4, metallic=0.2),
]


def hair_params(**kw):
    """20 floats of pbrgpu_material type 1 with the reference defaults (src/material-param.h:51-72)."""
    d = dict(melanin_mode=1.0, base_color=(0.18, 0.06, 0.02), melanin=0.5, redness=0.8, randomize=0.0, roughness=0.2,
             azimuthal=0.3, ior=1.55, shift=2.0, tint0=(1, 1, 1), tint1=(1, 1, 1), tint2=(1, 1, 1))
    d.update(kw)
    out = []
    for k in ["melanin_mode", "base_color", "melanin", "redness", "randomize", "roughness", "azimuthal", "ior", "shift",
              "tint0", "tint1", "tint2"]:
        v = d[k]
        out.extend(v if isinstance(v, (tuple, list)) else [v])
    return np.array(out, np.float32)


HAIR_CASES = [hair_params(), hair_params(melanin_mode=0.0, base_color=(0.6, 0.4, 0.1), roughness=0.4, azimuthal=0.6),
              hair_params(melanin=0.9, redness=0.2, roughness=0.05, shift=5.0, tint1=(0.9, 0.8, 0.7))]


def sphere_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def hemisphere_dirs(rng, n):
    v = sphere_dirs(rng, n)
    v[:, 2] = np.abs(v[:, 2])
    # a share of grazing and near-normal directions
    v[: n // 16, 2] = np.float32(1e-3)
    v[: n // 16] /= np.linalg.norm(v[: n // 16], axis=1, keepdims=True)
    return v.astype(np.float32)


def camera_rays(cam, n, rng, window=(0.0, 1.0, 0.0, 1.0), w=512, h=512):
    """n random pinhole rays through the image window (fractions of the frame), cam as ref_camera returns it."""
    cam = np.asarray(cam, np.float32)
    px = rng.random((n, 2)).astype(np.float32)
    fx = (window[0] + (window[1] - window[0]) * px[:, 0]) * w
    fy = (window[2] + (window[3] - window[2]) * px[:, 1]) * h
    tgt = np.stack([cam[3] + cam[6] * fx, cam[4] - cam[7] * fy, np.full(n, cam[5], np.float32)], 1).astype(np.float32)
    d = tgt - cam[:3]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return pb.make_rays(np.tile(cam[:3], (n, 1)), d.astype(np.float32))


def rays_from_f8(f8):
    r = np.zeros(len(f8), pb.RAY_DTYPE)
    r["org"] = f8[:, 0:3]; r["tmin"] = f8[:, 3]; r["dir"] = f8[:, 4:7]; r["tmax"] = f8[:, 7]
    return r


def path_agreement(a, b, rel=1e-4, floor=1e-2):
    """fraction of per-path RGB radiances equal within rel * max(floor, |b|)"""
    err = np.abs(a - b).max(axis=1)
    return float((err <= rel * np.maximum(floor, np.abs(b).max(axis=1))).mean())


def dump_report(name, obj):
    """evidence for profiles/: written under gpurun_out/ when the tests run on the GPU box (scratch, merged back)"""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f, indent=1, sort_keys=True)
    except OSError:
        pass
