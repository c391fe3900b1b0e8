"""Generates the golden fixtures under tests/golden/ from the COMPILED REFERENCE (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the build container:  python tests/golden/make_golden.py [--image]
The fixtures are what the tests use when the reference library is not around, and they pin the restated oracle."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refbind  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402
import pbrlab_b200 as pb  # noqa: E402
import common  # noqa: E402

R = refbind.RefLib()
rng = np.random.default_rng(20261017)


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items()})


# ---------------------------------------------------------------- closure / math known-answer vectors
def kat():
    out = {}
    n = 4096
    out["rng_seeds"] = np.array([[0, 0], [42, 54], [1234567890, 7], [2**40 + 3, 2**33 + 1]], np.uint64)
    out["rng_draws"] = np.stack([R.rng_draws(int(a), int(b), 64) for a, b in out["rng_seeds"]])
    x = np.concatenate([rng.uniform(-8, 8, n), rng.uniform(-300, 300, n // 4), [0.0, -0.0, 1e-8, 3.14159265, 6.2831853]]).astype(np.float32)
    xpos = np.concatenate([rng.uniform(1e-6, 10, n), 10 ** rng.uniform(-30, 30, n // 4), [1.0, 0.5, 2.0]]).astype(np.float32)
    xu = np.concatenate([rng.uniform(-1, 1, n), [-1.0, 1.0, 0.0, 0.99999, -0.99999, 1.5, -2.0]]).astype(np.float32)
    xe = np.concatenate([rng.uniform(-90, 90, n), [-126.5, 126.5, 0.0, -200.0, 200.0]]).astype(np.float32)
    y = rng.uniform(-4, 4, len(x)).astype(np.float32)
    out["fm_x"] = x; out["fm_xpos"] = xpos; out["fm_xu"] = xu; out["fm_xe"] = xe; out["fm_y"] = y
    out["fm_sin"] = R.fastmath(0, x); out["fm_cos"] = R.fastmath(1, x)
    out["fm_sincos_s"] = R.fastmath(8, x); out["fm_sincos_c"] = R.fastmath(9, x)
    out["fm_exp2"] = R.fastmath(2, xe); out["fm_exp"] = R.fastmath(3, xe)
    out["fm_log2"] = R.fastmath(4, xpos); out["fm_log"] = R.fastmath(5, xpos)
    out["fm_atan2"] = R.fastmath(6, x, y); out["fm_asin"] = R.fastmath(7, xu)
    u = rng.random((n, 2)).astype(np.float32)
    out["u2"] = u
    out["cosine_hemisphere"] = R.cosine_hemisphere(u)
    out["uniform_sphere"] = R.uniform_sphere(u)
    a = (10 ** rng.uniform(-6, 6, n)).astype(np.float32); b = (10 ** rng.uniform(-6, 6, n)).astype(np.float32)
    a[:8] = b[:8]; a[8] = 0; b[9] = 0
    out["mis_a"] = a; out["mis_b"] = b; out["mis_w"] = R.power_heuristic(a, b)
    c = rng.uniform(-1, 1, n).astype(np.float32); eta = rng.uniform(1.0, 2.5, n).astype(np.float32); eta[:4] = 0
    out["fr_cos"] = c; out["fr_eta"] = eta; out["fr"] = R.fresnel_dielectric_cos(c, eta)
    wo = common.hemisphere_dirs(rng, n); wi = common.hemisphere_dirs(rng, n)
    wi[:64, 2] *= -1                       # below the surface -> (0, 0)
    out["wo"] = wo; out["wi"] = wi
    ggx = []
    for (ax, ay, d) in common.GGX_CASES:
        ggx.append(np.concatenate([R.ggx_eval(wi, wo, ax, ay, d), R.ggx_sample(wo, ax, ay, u, d)], 1))
    out["ggx"] = np.stack(ggx)
    pr = []
    for p in common.PRINCIPLED_CASES:
        ev, bsdf = R.principled_eval(p, wi, wo)
        pr.append(dict(ev=ev, bsdf=bsdf, w=R.principled_weights(p, wo)))
    out["principled_eval"] = np.stack([q["ev"] for q in pr])
    out["principled_bsdf"] = np.stack([q["bsdf"] for q in pr])
    out["principled_w"] = np.stack([q["w"] for q in pr])
    # hair: local frame x = tangent, full sphere of directions
    hwo = common.sphere_dirs(rng, n); hwi = common.sphere_dirs(rng, n)
    h = rng.uniform(-1, 1, n).astype(np.float32); h[:4] = [-1, 1, 0, 0.999999]
    us = rng.random((n, 4)).astype(np.float32)
    out["hair_wo"] = hwo; out["hair_wi"] = hwi; out["hair_h"] = h; out["hair_us"] = us
    out["hair_eval"] = np.stack([R.hair_eval(p, h, hwi, hwo) for p in common.HAIR_CASES])
    out["hair_sample"] = np.stack([R.hair_sample(p, h, hwo, us) for p in common.HAIR_CASES])
    out["hair_setup"] = np.stack([R.hair_setup(p) for p in common.HAIR_CASES])
    k = 512
    sss_in = np.concatenate([rng.uniform(0.01, 1.0, (k, 3)), 10 ** rng.uniform(-4, 0, (k, 3)), rng.uniform(0, 1, (k, 3))], 1).astype(np.float32)
    sss_in[0, :3] = 1.0; sss_in[1, :3] = 0.0
    out["sss_in"] = sss_in
    out["sss_coeff"] = np.stack([R.sss_coefficients(r[:3], r[3:6], r[6:9]) for r in sss_in])
    sd_in = np.concatenate([rng.uniform(0, 2, (k, 3)), 10 ** rng.uniform(-1, 3, (k, 3)), 10 ** rng.uniform(0, 3, (k, 3)), rng.random((k, 2))], 1).astype(np.float32)
    sd_in[:, 3:6] = np.minimum(sd_in[:, 3:6], sd_in[:, 6:9])
    sd_in[0, :3] = 0
    out["sss_dist_in"] = sd_in
    out["sss_dist"] = np.stack([R.sss_sample_distance(r[:3], r[3:6], r[6:9], r[9:11]) for r in sd_in])
    save("kat_closures.npz", **out)


# ---------------------------------------------------------------- cornell: loader, rays, vertices, paths
def cornell():
    obj = scenes.cornell()
    L = R.obj_load(obj)
    verts = L.vertices()
    meta = {"num_shapes": L.num_shapes(), "num_materials": L.num_materials(), "num_vertices": int(len(verts)),
            "vertices_sum": float(verts.astype(np.float64).sum()), "shapes": [], "materials": []}
    sn_prim = {}
    for i in range(L.num_shapes()):
        name, vid, mid = L.shape(i)
        prim = rng.integers(0, len(vid), 64).astype(np.uint32)
        uv = rng.random((64, 2)).astype(np.float32) * 0.5
        sn = L.shading_normal(i, prim, uv)
        meta["shapes"].append({"name": name, "faces": int(len(vid)), "vid_sum": int(vid.astype(np.int64).sum()),
                               "mid": sorted(set(int(m) for m in mid))})
        sn_prim["sn_prim_%d" % i] = prim; sn_prim["sn_uv_%d" % i] = uv; sn_prim["sn_%d" % i] = sn
    for i in range(L.num_materials()):
        kind, p, tex, name = L.material(i)
        meta["materials"].append({"name": name, "kind": kind, "p": [float(v) for v in p], "tex": [int(t) for t in tex]})
    S = R.scene([obj])
    bmin, bmax = S.aabb()
    meta["bmin"] = [float(v) for v in bmin]; meta["bmax"] = [float(v) for v in bmax]
    meta["camera_512"] = [float(v) for v in S.camera(512, 512)]
    meta["camera_1920x1080"] = [float(v) for v in S.camera(1920, 1080)]
    with open(os.path.join(HERE, "cornell_loader.json"), "w") as f:
        json.dump(meta, f, indent=1)
    save("cornell_normals.npz", **sn_prim)

    n = 20000
    rays = common.camera_rays(S.camera(512, 512), n, rng)
    f, ids = S.trace(pb.rays_to_f8(rays))
    hit = ids[:, 0] != 0xFFFFFFFF
    P = rays["org"][hit] + f[hit, 0:1] * rays["dir"][hit]
    d2 = common.sphere_dirs(rng, len(P))
    rays2 = pb.make_rays(P, d2, tmin=1e-3)
    # short rays as the random walk and the shadow test issue them
    rays3 = pb.make_rays(P, common.sphere_dirs(rng, len(P)), tmin=0.0, tmax=(10 ** rng.uniform(-3, 0.5, len(P))).astype(np.float32))
    allrays = np.concatenate([rays, rays2, rays3])
    f, ids = S.trace(pb.rays_to_f8(allrays))
    occ = S.occluded(pb.rays_to_f8(allrays))
    surf = S.surface(pb.rays_to_f8(allrays))
    save("cornell_rays.npz", rays=pb.rays_to_f8(allrays), hit_f=f, hit_ids=ids, occluded=occ, surface=surf)

    seeds = np.stack([rng.integers(0, 2**62, n, dtype=np.uint64), rng.integers(0, 2**62, n, dtype=np.uint64)], 1)
    shade = S.shade(pb.rays_to_f8(rays), seeds)
    rad = S.radiance(pb.rays_to_f8(rays), seeds)
    ls, _ = S.sample_light(seeds[:256])
    save("cornell_paths.npz", rays=pb.rays_to_f8(rays), seeds=seeds, shade=shade, radiance=rad, light_samples=ls)
    return S


def image(S):
    w = h = 256
    a, ca, _ = S.render(w, h, 4096)
    b, cb, _ = S.render(w, h, 8192)
    save("cornell_image_256.npz", mean_4096=(a[..., :3] / ca[..., None]).astype(np.float32),
         mean_8192=(b[..., :3] / cb[..., None]).astype(np.float32))


def image_more():
    """image gates for the hair and the many-triangle configurations at fixture size: the C4 geometry at a tenth of
    the strand count, and the C5 generator at 200 k triangles (GGX + SSS blobs)"""
    w = h = 96
    for name, files in (("hair_image_96.npz", [scenes.cornell(), scenes.cyhair(5000, 21, center=(-2.5, 6.0, 0.0), radius=1.2,
                                                                               length=2.5, thickness=0.008)]),
                        ("displaced_image_96.npz", [scenes.displaced(200_000)])):
        S = R.scene(files)
        a, ca, _ = S.render(w, h, 2048)
        b, cb, _ = S.render(w, h, 4096)
        save(name, mean_2048=(a[..., :3] / ca[..., None]).astype(np.float32),
             mean_4096=(b[..., :3] / cb[..., None]).astype(np.float32))


# ---------------------------------------------------------------- hair: CyHair ingest, curve hits, hair vertices
def hair():
    hp = os.path.join(scenes.CACHE, "golden_hair.hair")
    scenes.write_cyhair(hp, n_strands=400, n_points=9, center=(-2.5, 6.0, 0.0), radius=1.0, length=2.0,
                        thickness=0.02, seed=99)
    ok, v, idx = R.hair_load(hp)
    assert ok
    obj = scenes.cornell()
    S = R.scene([obj, hp])
    bmin, bmax = S.aabb()
    n = 20000
    cam = S.camera(512, 512)
    # aim most rays at the hair ball
    rays = common.camera_rays(cam, n, rng, window=(0.05, 0.45, 0.15, 0.6))
    f, ids = S.trace(pb.rays_to_f8(rays))
    occ = S.occluded(pb.rays_to_f8(rays))
    surf = S.surface(pb.rays_to_f8(rays))
    seeds = np.stack([rng.integers(0, 2**62, n, dtype=np.uint64), rng.integers(0, 2**62, n, dtype=np.uint64)], 1)
    shade = S.shade(pb.rays_to_f8(rays), seeds)
    rad = S.radiance(pb.rays_to_f8(rays), seeds)
    print("hair hits:", int((ids[:, 0] == 9).sum()), "of", n)
    save("hair_scene.npz", bezier=v, bezier_idx=idx, bmin=bmin, bmax=bmax, rays=pb.rays_to_f8(rays), hit_f=f,
         hit_ids=ids, occluded=occ, surface=surf, seeds=seeds, shade=shade, radiance=rad)
    # hair-only bounds (camera depends on them)
    S2 = R.scene([hp])
    b0, b1 = S2.aabb()
    save("hair_bounds.npz", bmin=b0, bmax=b1)


# ---------------------------------------------------------------- textured scene: loader textures, fetches, vertices, paths
def textured():
    obj = scenes.textured()
    L = R.obj_load(obj)
    out = {"num_textures": np.array([L.num_textures()], np.uint32)}
    uv = np.concatenate([rng.uniform(-0.3, 1.3, (2048, 2)), [[0, 0], [1, 1], [0.5, 1.0], [1.0, 0.25], [-1, 2]]]).astype(np.float32)
    out["fetch_uv"] = uv
    for i in range(L.num_textures()):
        out["tex_%d" % i] = L.texture(i)
        out["fetch_%d" % i] = L.texture_fetch3(i, uv)
    mats = []
    for i in range(L.num_materials()):
        kind, p, tex, name = L.material(i)
        mats.append(np.concatenate([p, tex.astype(np.float64)]))
    out["materials"] = np.array(mats, np.float64)
    S = R.scene([obj])
    n = 20000
    rays = common.camera_rays(S.camera(512, 512), n, rng)
    f, ids = S.trace(pb.rays_to_f8(rays))
    surf = S.surface(pb.rays_to_f8(rays))
    seeds = np.stack([rng.integers(0, 2**62, n, dtype=np.uint64), rng.integers(0, 2**62, n, dtype=np.uint64)], 1)
    shade = S.shade(pb.rays_to_f8(rays), seeds)
    rad = S.radiance(pb.rays_to_f8(rays), seeds)
    save("textured_scene.npz", rays=pb.rays_to_f8(rays), hit_f=f, hit_ids=ids, occluded=S.occluded(pb.rays_to_f8(rays)),
         surface=surf, seeds=seeds, shade=shade, radiance=rad, **out)


def output_stage():
    """the CLI's output statements (pc/pbrlab-cli.cc:47-57 -> LinerToSrgb -> WritePNG) on a synthetic RenderLayer that
    covers the sRGB knee, values above 1, zeros, negatives, a NaN and a 0-count pixel; the PNG the reference wrote is
    decoded (PIL) and stored next to the inputs"""
    import tempfile
    from PIL import Image
    r = np.random.default_rng(77)
    h, w = 48, 64
    count = np.full((h, w), 16, np.uint32)
    mean = (10 ** r.uniform(-5, 0.3, (h, w, 4))).astype(np.float32)
    mean[0, :, 0] = np.linspace(0.0030, 0.0033, w, dtype=np.float32)      # around the linear/power knee
    mean[1, :, 1] = np.linspace(0.99, 1.01, w, dtype=np.float32)          # around the clamp
    mean[2, :8, 2] = [0.0, -0.0, -0.5, np.nan, np.inf, 1e-30, 255.0 / 256.0, 1.0]
    mean[..., 3] = 1.0
    rgba = mean * np.float32(16.0)
    count[3, 0] = 0                                                        # 0/0 -> NaN in every channel
    rgba[3, 0] = 0.0
    d = tempfile.mkdtemp()
    with np.errstate(all="ignore"):
        assert R.output_stage(rgba, count, d)
    png = np.array(Image.open(os.path.join(d, "rgba.png")))
    assert png.shape == (h, w, 4) and png.dtype == np.uint8
    save("output_stage.npz", rgba=rgba, count=count, png=png)


if __name__ == "__main__":
    if "--output-stage-only" in sys.argv:
        output_stage()
        sys.exit(0)
    if "--image-more-only" in sys.argv:
        image_more()
        sys.exit(0)
    if "--textured-only" in sys.argv:
        textured()
        sys.exit(0)
    kat()
    S = cornell()
    hair()
    textured()
    output_stage()
    if "--image" in sys.argv:
        image(S)
