"""The parity tests proper: the CUDA path, called through the C ABI (libpbrgpu.so), against the golden fixtures
generated from the compiled reference, against the reference itself when oracle/_ref travelled with the snapshot,
and — at full size — through size-independent properties."""
import numpy as np
import pytest

import checks
import common
import pbrlab_b200 as pb
from conftest import golden
from pbrlab_b200 import scenes

pytestmark = pytest.mark.gpu


def test_native_library_is_the_one_running(cornell_gpu):
    """the .so that answers is the in-tree CUDA library, and it really launched kernels"""
    _, ctx = cornell_gpu
    rays = common.rays_from_f8(golden("cornell_rays.npz")["rays"][:1024])
    ctx.trace(rays)
    st = ctx.stats()
    assert st["kernel_launches"] >= 1 and st["closest_rays"] == 1024
    maps = open("/proc/self/maps").read()
    assert "pbrlab_b200/lib/libpbrgpu.so" in maps


def test_closure_known_answers(cornell_gpu):
    _, ctx = cornell_gpu
    try:
        checks.check_kat(ctx, golden("kat_closures.npz"))
    finally:
        common.dump_report("kat_outliers_gpu.json", checks.KAT_REPORT)


def test_rays_vs_embree_golden(cornell_gpu):
    _, ctx = cornell_gpu
    agree = checks.check_rays(ctx, golden("cornell_rays.npz"))
    assert agree >= 0.9999


def test_shading_vertices(cornell_gpu):
    _, ctx = cornell_gpu
    checks.check_shade(ctx, golden("cornell_paths.npz"), min_agree=0.99)


def test_paths_wavefront_vs_reference_and_megakernel(cornell_gpu):
    _, ctx = cornell_gpu
    g = golden("cornell_paths.npz")
    checks.check_radiance(ctx, g, min_agree=0.99)
    rays = common.rays_from_f8(g["rays"])
    wave = ctx.radiance(rays, g["seeds"])
    mega = ctx.radiance(rays, g["seeds"], mega=True)
    # same per-vertex device functions, different scheduling: identical up to the order of the two NEE additions
    assert common.path_agreement(wave, mega, rel=1e-5) >= 0.9999


def test_hair_scene(hair_gpu):
    _, ctx = hair_gpu
    g = golden("hair_scene.npz")
    rays = common.rays_from_f8(g["rays"])
    hits = ctx.trace(rays)
    ids = g["hit_ids"]; f = g["hit_f"]
    same = (hits["instance_id"] == ids[:, 0]) & (hits["geom_id"] == ids[:, 1]) & (hits["prim_id"] == ids[:, 2])
    # the thresholds of the dense-hair gate below (north_star: >= 99.99 %, t within 1e-5 relative); this fixture is
    # small (tests/golden/hair_scene.npz), so the 99.99 % is spelt as a count of rays
    n_diff = int((~same).sum())
    assert n_diff <= max(1, len(rays) // 10000), (n_diff, len(rays))
    curve = same & (ids[:, 0] == 9)
    assert curve.sum() > 3000
    assert np.all(np.abs(hits["t"][curve] - f[curve, 0]) <= 1e-5 * np.abs(f[curve, 0]))
    assert np.abs(hits["u"][curve] - f[curve, 1]).max() < 1e-3
    # v = the hair BSDF's h: 2 * (distance from the ribbon's axis) / width - 1, a quotient of two differences of
    # ray-space coordinates ~1e3 times larger than the ribbon is wide, so 1e-5 relative on t is ~1e-3 absolute on v
    dv = np.abs(hits["v"][curve] - f[curve, 2])
    assert dv.max() < 2e-3 and np.mean(dv > 2e-4) < 1e-3, (float(dv.max()), float(np.mean(dv > 2e-4)))
    n_occ = int((ctx.occluded(rays) != g["occluded"]).sum())
    assert n_occ <= max(1, len(rays) // 10000), n_occ
    rad = ctx.radiance(rays, g["seeds"])
    frac = common.path_agreement(rad, g["radiance"], rel=1e-4)
    common.dump_report("hair_scene_gpu.json", {"rays": len(rays), "hit_mismatch": n_diff, "occlusion_mismatch": n_occ,
                                               "max_dv": float(dv.max()), "paths_within_1e-4": frac})
    # whole paths through hair: every bounce re-derives h from v (|dv| up to 4e-5 here), and the peaked hair lobes
    # turn that into a slightly different sampled direction that the following bounces amplify; measured: 96.4 % of
    # the paths agree to 1e-4 (97 % to 1e-3), the rest diverge chaotically with equal means
    assert frac >= 0.955, frac
    ma, mb = rad.mean(axis=0), g["radiance"].mean(axis=0)
    assert np.all(np.abs(ma - mb) <= 0.02 * np.maximum(mb, 1e-3)), (ma, mb)


def test_dense_hair_rays_vs_live_embree(built, ref):
    """ray gate on hair: a dense CyHair ball (5 000 strands x 20 segments, the C3/C4 geometry at a tenth of the strand
    count) inside the Cornell box; 2 Mi camera + secondary + short rays against the compiled reference: hit / primID
    agreement >= 99.99 %, t within 1e-5 relative, u and v (the hair BSDF's h) close, occlusion flags"""
    if ref is None:
        pytest.skip("oracle/_ref did not travel: tests/golden/hair_scene.npz covers the hair ray gate")
    files = [scenes.cornell(), scenes.cyhair(5000, 21, center=(-2.5, 6.0, 0.0), radius=1.2, length=2.5, thickness=0.008)]
    sc = pb.Scene(files)
    ctx = sc.context()
    S = ref.scene(files)
    rng = np.random.default_rng(23)
    n = 1 << 20
    rays = common.camera_rays(S.camera(512, 512), n, rng)
    f, ids = S.trace(pb.rays_to_f8(rays))
    hit = ids[:, 0] != 0xFFFFFFFF
    P = rays["org"][hit] + f[hit, 0:1] * rays["dir"][hit]
    # secondary rays leave from what the camera sees; a further batch starts inside the hair volume
    inside = (np.array([-2.5, 5.0, 0.0]) + rng.uniform(-1.5, 1.5, (n // 2, 3))).astype(np.float32)
    org = np.concatenate([P, inside]).astype(np.float32)
    k = len(org)
    tmax = np.where(rng.random(k) < 0.3, 10 ** rng.uniform(-3, 0.5, k), 1.844e18).astype(np.float32)
    tmin = np.where(rng.random(k) < 0.5, 1e-3, 0.0).astype(np.float32)
    rays2 = pb.make_rays(org, common.sphere_dirs(rng, k), tmin=tmin, tmax=tmax)
    f2, ids2 = S.trace(pb.rays_to_f8(rays2))
    R = np.concatenate([rays, rays2]); F = np.concatenate([f, f2]); I = np.concatenate([ids, ids2])
    H = ctx.trace(R)
    same = (H["instance_id"] == I[:, 0]) & (H["geom_id"] == I[:, 1]) & (H["prim_id"] == I[:, 2])
    assert len(H) >= 2_000_000
    assert same.mean() >= 0.9999, same.mean()
    both = same & (I[:, 0] != 0xFFFFFFFF)
    assert np.all(np.abs(H["t"][both] - F[both, 0]) <= 1e-5 * np.abs(F[both, 0]))
    curve = both & (I[:, 0] == 9)
    assert curve.sum() > 100_000, curve.sum()
    assert np.abs(H["u"][curve] - F[curve, 1]).max() < 1e-3
    assert np.abs(H["v"][curve] - F[curve, 2]).max() < 2e-3
    assert (ctx.occluded(rays2) == S.occluded(pb.rays_to_f8(rays2))).mean() >= 0.9999
    sc.close()


def test_fixed_16m_ray_batch_vs_embree(cornell_gpu, ref):
    """north_star ray gate: 8 Mi camera rays (4096 x 2048 through the C1 camera) + 8 Mi secondary / shadow / walk rays,
    hit flag + primID agreement >= 99.99 %, t within 1e-5 relative."""
    if ref is None:
        pytest.skip("oracle/_ref did not travel: the golden batch (tests/golden/cornell_rays.npz) covers this gate")
    _, ctx = cornell_gpu
    S = ref.scene([scenes.cornell()])
    cam = S.camera(512, 512)
    n = 1 << 23
    yy, xx = np.meshgrid(np.arange(2048, dtype=np.float32) + 0.5, np.arange(4096, dtype=np.float32) + 0.5, indexing="ij")
    tgt = np.stack([cam[3] + cam[6] * xx.ravel() * (512.0 / 4096.0), cam[4] - cam[7] * yy.ravel() * (512.0 / 2048.0),
                    np.full(n, cam[5], np.float32)], 1).astype(np.float32)
    d = tgt - cam[:3]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = pb.make_rays(np.tile(cam[:3], (n, 1)), d.astype(np.float32))
    hits = ctx.trace(rays)
    f, ids = S.trace(pb.rays_to_f8(rays))
    rng = np.random.default_rng(16)
    hit = ids[:, 0] != 0xFFFFFFFF
    P = (rays["org"][hit] + f[hit, 0:1] * rays["dir"][hit])[: n]
    k = len(P)
    d2 = common.sphere_dirs(rng, k)
    tmax = np.where(rng.random(k) < 0.3, 10 ** rng.uniform(-3, 0.5, k), 1.844e18).astype(np.float32)
    tmin = np.where(rng.random(k) < 0.5, 1e-3, 0.0).astype(np.float32)
    rays2 = pb.make_rays(P, d2, tmin=tmin, tmax=tmax)
    hits2 = ctx.trace(rays2)
    f2, ids2 = S.trace(pb.rays_to_f8(rays2))
    H = np.concatenate([hits, hits2]); F = np.concatenate([f, f2]); I = np.concatenate([ids, ids2])
    same = (H["instance_id"] == I[:, 0]) & (H["geom_id"] == I[:, 1]) & (H["prim_id"] == I[:, 2])
    assert len(H) >= 15_000_000
    assert same.mean() >= 0.9999, same.mean()
    both = same & (I[:, 0] != 0xFFFFFFFF)
    assert np.all(np.abs(H["t"][both] - F[both, 0]) <= 1e-5 * np.abs(F[both, 0]))
    occ = ctx.occluded(rays2)
    assert (occ == S.occluded(pb.rays_to_f8(rays2))).mean() >= 0.9999


def test_image_statistics_vs_reference_render(cornell_gpu):
    """north_star image gate: mean luminance within 0.5 %, per-pixel RMSE no larger than the reference's own
    4096-vs-8192-spp noise floor (fixture: tests/golden/cornell_image_256.npz, rendered by oracle/_ref)."""
    import os
    from conftest import GOLDEN
    if not os.path.exists(os.path.join(GOLDEN, "cornell_image_256.npz")):
        pytest.skip("image fixture not generated")
    _, ctx = cornell_gpu
    g = golden("cornell_image_256.npz")
    rgba, count = ctx.render(256, 256, 4096, seed=20261017)
    assert np.all(count == 4096) and np.all(rgba[..., 3] == 4096.0)
    img = rgba[..., :3] / count[..., None]
    assert not np.isnan(img).any()
    lum = lambda x: (0.212671 * x[..., 0] + 0.715160 * x[..., 1] + 0.072169 * x[..., 2])
    ref8, ref4 = g["mean_8192"], g["mean_4096"]
    assert abs(lum(img).mean() - lum(ref8).mean()) <= 0.005 * lum(ref8).mean()
    floor = np.sqrt(np.mean((ref4 - ref8) ** 2))
    rmse = np.sqrt(np.mean((img - ref8) ** 2))
    assert rmse <= floor * 1.05, (rmse, floor)


def test_render_properties_at_full_size(cornell_gpu):
    """C1 at full size: counts, alpha, sample-split additivity, host-API and C-ABI agreement, seed dependence"""
    scene, ctx = cornell_gpu
    w = h = 512
    rgba, count = ctx.render(w, h, 64, seed=1)
    assert np.all(count == 64) and np.all(rgba[..., 3] == 64.0) and not np.isnan(rgba).any()
    a, ca = ctx.render(w, h, 64, seed=1, sample_offset=0, sample_stride=2)
    b, cb = ctx.render(w, h, 64, seed=1, sample_offset=1, sample_stride=2)
    assert np.all(ca == 32) and np.all(cb == 32)
    assert np.allclose(a + b, rgba, rtol=1e-4, atol=1e-4)        # interleaved sample split is exact up to fp order
    r2, c2, _ = scene.render(w, h, 64, seed=1)                   # pbrlab::Render() through the C++ entry point
    assert np.array_equal(c2, count) and np.allclose(r2, rgba, rtol=1e-4, atol=1e-4)
    other, _ = ctx.render(w, h, 64, seed=2)
    assert not np.allclose(other, rgba, rtol=1e-3, atol=1e-3)
    m1 = (rgba[..., :3] / 64).mean(); m2 = (other[..., :3] / 64).mean()
    assert abs(m1 - m2) < 0.01 * m1                              # different streams, same estimator
    st = ctx.stats()
    assert st["paths"] == w * h * 64
    # ray queries per path as the reference counts them (every walk segment is an rtcIntersect1 there; here the clearance
    # field answers most of them without a traversal): reference 5.995
    queries = st["closest_rays"] + st["shadow_rays"] + st["sss_rays"] + st["sss_skipped"]
    assert 5.5 < queries / st["paths"] < 6.5, queries / st["paths"]
    assert st["sss_rays"] > 0 and st["sss_skipped"] > 0.5 * st["sss_rays"]


def test_live_material_edit(cornell_gpu):
    """materials can be changed between renders without a re-commit (GUI contract)"""
    scene, ctx = cornell_gpu
    base, _ = ctx.render(128, 128, 16, seed=3)
    words = scene.flat().materials.copy()
    edited = words.copy()
    p = edited[:, 4:27].view(np.float32)
    p[:, 0:3] *= 0.25                                             # darken every base colour
    edited[:, 4:27] = p.view(np.uint32)
    ctx.set_materials(edited)
    dark, _ = ctx.render(128, 128, 16, seed=3)
    ctx.set_materials(words)
    again, _ = ctx.render(128, 128, 16, seed=3)
    assert dark[..., :3].sum() < 0.8 * base[..., :3].sum()
    assert np.allclose(again, base, rtol=1e-4, atol=1e-4)


def test_frame_schedule_switches_do_not_change_the_image(built, monkeypatch):
    """what only changes WHEN a sample or a vertex runs leaves the frame as it is, up to the order of the float sums:
    raster order vs longest paths first (PBRGPU_ORDER), the walk kernels beside closest hit + shading (PBRGPU_OVERLAP),
    thin spreading of short launches (PBRGPU_THIN), and the last paths run to their end by one thread each
    (PBRGPU_FINISH_PATHS: FinishPathsKernel takes paths and walks up in the middle) vs one iteration per vertex"""
    frames = []
    for env in ({}, {"PBRGPU_ORDER": "0"}, {"PBRGPU_OVERLAP": "0"}, {"PBRGPU_THIN": "0"}, {"PBRGPU_FINISH_PATHS": "0"},
                {"PBRGPU_FINISH_PATHS": "1000000"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        sc = pb.Scene([scenes.cornell()])
        rgba, count = sc.context().render(256, 192, 24, seed=77)
        st = sc.context().stats()
        assert np.all(count == 24) and st["paths"] == 256 * 192 * 24
        frames.append(rgba)
        sc.close()
        for k in env:
            monkeypatch.delenv(k)
    for f in frames[1:]:
        # (one path in a few hundred thousand sums its two NEE terms in the other order: compare per-pixel sums loosely,
        # the whole image tightly)
        assert np.allclose(frames[0], f, rtol=2e-3, atol=1e-3), float(np.abs(frames[0] - f).max())
        assert abs(float(f[..., :3].sum()) - float(frames[0][..., :3].sum())) <= 1e-5 * float(frames[0][..., :3].sum())


def test_render_into_a_layer_kept_across_frames(cornell_gpu):
    """pbrlab::Render() into ONE RenderLayer held across frames (how the reference's GUI / CLI call it, and what
    bench.py's e2e times): same sums as a fresh layer, the buffers are reused, a smaller and a larger frame in between
    resize it, and a live edit that shrinks the material table below an id in use is refused without a scan"""
    scene, ctx = cornell_gpu
    fresh, fcount, _ = scene.render(96, 64, 8, seed=9)
    a, ac, _ = scene.render_layer(96, 64, 8, seed=9)
    assert np.allclose(a, fresh, rtol=1e-4, atol=1e-5) and np.array_equal(ac, fcount)
    addr = a.ctypes.data
    small, sc_, _ = scene.render_layer(32, 32, 4, seed=9)
    assert np.all(sc_ == 4) and np.all(small[..., 3] == 4.0)
    b, bc, _ = scene.render_layer(96, 64, 8, seed=9)
    assert b.ctypes.data == addr                                  # same host buffers
    assert np.allclose(b, fresh, rtol=1e-4, atol=1e-5) and np.array_equal(bc, fcount)
    words = scene.flat().materials.copy()
    with pytest.raises(RuntimeError):
        ctx.set_materials(words[:1])                              # ids up to len(words) - 1 are in use
    ctx.set_materials(words)


def test_edge_cases(built):
    """empty / degenerate inputs: errors, not crashes"""
    ctx = pb.Context()
    with pytest.raises(RuntimeError):
        ctx.render(16, 16, 1)                                     # not committed
    assert ctx.lib.pbrgpu_commit(ctx.h, None, None) != 0           # empty scene
    # a single triangle, no material, no light: every path is absorbed, image stays black, counts still add up
    v = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [0, 1, 0, 1]], np.float32)
    import ctypes as C
    idx = np.array([0, 1, 2], np.uint32); z = np.zeros(1, np.uint32); none = np.full(1, 0xFFFFFFFF, np.uint32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    assert ctx.lib.pbrgpu_set_triangles(ctx.h, P(v), 3, P(idx), None, 0, None, None, 0, None, P(none), P(z), P(z), P(z), C.c_uint64(1)) == 0
    assert ctx.lib.pbrgpu_commit(ctx.h, None, None) == 0
    rgba, count = ctx.render(8, 8, 2)
    assert np.all(count == 2) and np.all(rgba[..., :3] == 0) and np.all(rgba[..., 3] == 2)
    hits = ctx.trace(pb.make_rays([[0.2, 0.2, 1.0]], [[0, 0, -1.0]]))
    assert hits["prim_id"][0] == 0 and abs(hits["t"][0] - 1.0) < 1e-6
    assert ctx.trace(pb.make_rays([[0.2, 0.2, 1.0]], [[0, 0, 1.0]]))["instance_id"][0] == 0xFFFFFFFF
    # zero-sample render clears the layer
    rgba, count = ctx.render(8, 8, 0)
    assert np.all(count == 0) and np.all(rgba == 0)
    ctx.close()


def test_multi_device_context_matches_one_device(built):
    """SURVEY §8(e) inside the library, one process: a context over two devices renders interleaved sample shares with
    the scene replicated and sums the accumulators with one grouped ncclReduce (ncclCommInitAll); per-path streams are
    keyed by (global sample, pixel), so the frame equals the single-device frame up to the order of the float
    additions.  PBRGPU_NO_NCCL=1 takes the peer-copy fallback."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one = pb.Scene([scenes.cornell()], device_ids=[0])
    two = pb.Scene([scenes.cornell()], device_ids=[0, 1])
    a, ca, _ = one.render(192, 160, 16, seed=99)
    b, cb, _ = two.render(192, 160, 16, seed=99)
    assert np.array_equal(ca, cb) and np.all(cb == 16)
    assert np.all(b[..., 3] == 16.0)
    assert np.allclose(a, b, rtol=1e-4, atol=1e-4), float(np.abs(a - b).max())
    b3, cb3, _ = two.render(192, 160, 3, seed=99)                         # uneven shares: 2 + 1 samples
    a3, ca3, _ = one.render(192, 160, 3, seed=99)
    assert np.array_equal(ca3, cb3) and np.allclose(a3, b3, rtol=1e-4, atol=1e-4)
    one.close(); two.close()


@pytest.mark.parametrize("fixture,files", [
    ("hair_image_96.npz", lambda: [scenes.cornell(), scenes.cyhair(5000, 21, center=(-2.5, 6.0, 0.0), radius=1.2,
                                                                   length=2.5, thickness=0.008)]),
    ("displaced_image_96.npz", lambda: [scenes.displaced(200_000)]),
])
def test_image_statistics_hair_and_displaced(built, fixture, files):
    """north_star image gate on the other two scene families (fixtures rendered by oracle/_ref at 2048 and 4096 spp,
    tests/golden/make_golden.py --image-more-only): the C4 geometry at a tenth of the strand count (Principled Hair +
    Lucy SSS) and the C5 generator at 200 k triangles (GGX + SSS blobs): mean luminance within 0.5 %, per-pixel RMSE
    no larger than the reference's own 2048-vs-4096-spp noise floor"""
    g = golden(fixture)
    sc = pb.Scene(files())
    rgba, count, _ = sc.render(96, 96, 4096, seed=20261018)
    assert np.all(count == 4096)
    img = rgba[..., :3] / count[..., None]
    assert not np.isnan(img).any()
    lum = lambda x: (0.212671 * x[..., 0] + 0.715160 * x[..., 1] + 0.072169 * x[..., 2])
    ref4, ref2 = g["mean_4096"], g["mean_2048"]
    assert abs(lum(img).mean() - lum(ref4).mean()) <= 0.005 * lum(ref4).mean(), (lum(img).mean(), lum(ref4).mean())
    floor = np.sqrt(np.mean((ref2 - ref4) ** 2))
    rmse = np.sqrt(np.mean((img - ref4) ** 2))
    assert rmse <= floor * 1.05, (rmse, floor)
    sc.close()


def test_curve_part_splits_report_the_same_hits(built, hair_file, monkeypatch):
    """the curve BVH over whole segments, halves and quarters (PBRGPU_CURVE_SPLIT = 1, 2, 4) and the engine with and
    without the line-distance rejection report bit-identical hits and occlusion flags on the GPU"""
    g = golden("hair_scene.npz")
    rays = common.rays_from_f8(g["rays"])
    rng = np.random.default_rng(31)
    inside = (np.array([-2.5, 5.5, 0.0]) + rng.uniform(-1.2, 1.2, (60000, 3))).astype(np.float32)
    more = pb.make_rays(inside, common.sphere_dirs(rng, len(inside)), tmin=1e-3)
    rays = np.concatenate([rays, more])
    results = []
    for split, nocull in (("4", False), ("2", False), ("1", False), ("4", True)):
        monkeypatch.setenv("PBRGPU_CURVE_SPLIT", split)
        if nocull:
            monkeypatch.setenv("PBRGPU_NO_CURVE_CULL", "1")
        sc = pb.Scene([scenes.cornell(), hair_file])
        ctx = sc.context()
        results.append((ctx.trace(rays), ctx.occluded(rays)))
        sc.close()
    monkeypatch.delenv("PBRGPU_CURVE_SPLIT")
    monkeypatch.delenv("PBRGPU_NO_CURVE_CULL")
    h0, o0 = results[0]
    assert (h0["instance_id"] == 9).sum() > 5000
    for h, o in results[1:]:
        for k in ("t", "u", "v", "instance_id", "geom_id", "prim_id"):
            assert np.array_equal(h0[k], h[k]), k
        assert np.array_equal(o0, o)


def test_cancel_mid_frame(cornell_gpu):
    """Render() with the flag raised from a second thread mid-frame (reference src/render.cc:217-231,
    pc/glfw-window.cc:621-625): returns true, every pixel holds whole samples only (rgba.a == count), finish_pass
    rose monotonically and stayed <= num_sample, and the next Render() on the same scene is clean."""
    scene, ctx = cornell_gpu
    w = h = 512
    spp = 4096
    rgba, count, info = scene.render_cancelled(w, h, spp, cancel_at_pass=8, seed=5)
    st = ctx.stats()
    assert info["returned"] is True
    assert not info["progress_violation"]
    assert info["raised_at"] >= 8 and info["finish_pass"] >= info["raised_at"] and info["finish_pass"] <= spp
    assert np.array_equal(rgba[..., 3], count.astype(np.float32))          # alpha and count incremented together
    assert int(count.max()) <= spp and not np.isnan(rgba).any()
    total = int(count.sum(dtype=np.uint64))
    assert 0 < total < w * h * spp // 2, total                              # it really stopped early
    assert total == st["paths"]                                            # everything accumulated is reported
    assert info["finish_pass"] <= total // (w * h) + 1
    # cancelled before the first batch: a cleared layer, finish_pass 0, still `true`
    r0, c0, i0 = scene.render_cancelled(w, h, 64, cancel_at_pass=0, seed=5)
    assert i0["returned"] is True and i0["finish_pass"] == 0 and not c0.any() and not r0.any()
    # and the scene renders normally afterwards
    r2, c2, _ = scene.render(w, h, 16, seed=5)
    assert np.all(c2 == 16) and np.all(r2[..., 3] == 16.0) and not np.isnan(r2).any()
    r1, c1 = ctx.render(w, h, 16, seed=5)
    assert np.allclose(r1, r2, rtol=1e-4, atol=1e-4)                       # nothing of the cancelled frame leaked in
    # the C ABI's own flag, raised before the call
    import ctypes as C
    flag = C.c_int(1)
    r3, c3 = ctx.render(64, 64, 8, cancel=flag)
    assert not c3.any() and not r3.any()


def test_clearance_field_never_changes_a_walk(built, monkeypatch):
    """PBRGPU_SSS_SKIP=0 traces every random-walk segment, =1 answers the segments the clearance field can vouch for
    without a traversal.  On the GPU, one Shader() call per path through the single-vertex hook (the walk runs to its
    end): exit position, direction, throughput, pdf are bit-identical; the NEE column only up to the order of its two
    atomic additions."""
    g = golden("cornell_paths.npz")
    rays = common.rays_from_f8(g["rays"])
    out = []
    skipped = []
    for skip in ("1", "0"):
        monkeypatch.setenv("PBRGPU_SSS_SKIP", skip)
        sc = pb.Scene([scenes.cornell()])
        ctx = sc.context()
        out.append(ctx.shade(rays, g["seeds"]))
        st = ctx.stats()
        skipped.append(st["sss_skipped"])
        assert st["sss_rays"] + st["sss_skipped"] > 10_000
        sc.close()
    monkeypatch.delenv("PBRGPU_SSS_SKIP")
    a, b = out
    assert skipped[0] > 3_000 and skipped[1] == 0, skipped
    exact = [0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13, 14, 15]
    assert np.array_equal(a[:, exact], b[:, exact])
    assert np.allclose(a[:, 7:10], b[:, 7:10], rtol=1e-6, atol=1e-9)


def test_hair_material_on_triangles_is_shaded(built):
    """the C ABI allows a type-1 (hair) material on triangles and the reference's Shader() dispatches on the material
    type, not the geometry (src/shader/shader.cc:8-34): in a scene WITHOUT curves those hits must still be shaded and
    retired (they were parked in the hair queue for ever: count < spp)"""
    import ctypes as C
    ctx = pb.Context()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    v = np.array([[-1, 0, -1, 1], [1, 0, -1, 1], [1, 0, 1, 1], [-1, 0, 1, 1],
                  [-1, 0, -1, 1], [1, 0, -1, 1], [1, 2, -1, 1], [-1, 2, -1, 1]], np.float32)
    idx = np.array([0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7], np.uint32)
    mats = np.zeros((2, 28), np.uint32)
    mats[0, 0] = 1; mats[0, 1:3] = 0xFFFFFFFF
    mats[0, 4:24] = common.hair_params().view(np.uint32)
    mats[1, 0] = 0; mats[1, 1:3] = 0xFFFFFFFF
    mats[1, 4:27] = common.principled().view(np.uint32)
    mid = np.array([0, 0, 1, 1], np.uint32); z = np.zeros(4, np.uint32); prim = np.arange(4, dtype=np.uint32)
    assert ctx.lib.pbrgpu_set_materials(ctx.h, P(mats), 2) == 0
    assert ctx.lib.pbrgpu_set_triangles(ctx.h, P(v), 8, P(idx), None, 0, None, None, 0, None, P(mid), P(z), P(z), P(prim), C.c_uint64(4)) == 0
    assert ctx.lib.pbrgpu_commit(ctx.h, None, None) == 0
    rgba, count = ctx.render(64, 64, 32)
    assert np.all(count == 32) and np.all(rgba[..., 3] == 32.0)
    st = ctx.stats()
    assert st["paths"] == 64 * 64 * 32 and st["closest_rays"] > st["paths"]     # the hair vertices bounced on
    ctx.close()


def _big_gate(ctx, S, rng, n_cam, box_lo, box_hi, curve_instance=None):
    """camera rays + secondary / short rays from what they hit + rays born inside the geometry's box, against the live
    reference; returns (rays, agreement, t_ok, occlusion agreement, curve hits)"""
    rays = common.camera_rays(S.camera(512, 512), n_cam, rng)
    f, ids = S.trace(pb.rays_to_f8(rays))
    hit = ids[:, 0] != 0xFFFFFFFF
    P = rays["org"][hit] + f[hit, 0:1] * rays["dir"][hit]
    inside = rng.uniform(box_lo, box_hi, (n_cam, 3)).astype(np.float32)
    org = np.concatenate([P, inside]).astype(np.float32)
    k = len(org)
    tmax = np.where(rng.random(k) < 0.3, 10 ** rng.uniform(-3, 0.5, k), 1.844e18).astype(np.float32)
    tmin = np.where(rng.random(k) < 0.5, 1e-3, 0.0).astype(np.float32)
    rays2 = pb.make_rays(org, common.sphere_dirs(rng, k), tmin=tmin, tmax=tmax)
    f2, ids2 = S.trace(pb.rays_to_f8(rays2))
    R = np.concatenate([rays, rays2]); F = np.concatenate([f, f2]); I = np.concatenate([ids, ids2])
    H = ctx.trace(R)
    same = (H["instance_id"] == I[:, 0]) & (H["geom_id"] == I[:, 1]) & (H["prim_id"] == I[:, 2])
    both = same & (I[:, 0] != 0xFFFFFFFF)
    rel = np.zeros(len(R))
    rel[both] = np.abs(H["t"][both] - F[both, 0]) / np.abs(F[both, 0])
    # north_star: "the hit/primID flag agrees ... at >= 99.99 % with t within 1e-5 relative": a ray counts as agreeing
    # when the same primitive (or a miss) is reported AND t is within 1e-5
    agree = same & (rel <= 1e-5)
    occ = float((ctx.occluded(rays2) == S.occluded(pb.rays_to_f8(rays2))).mean())
    curves = int((both & (I[:, 0] == curve_instance)).sum()) if curve_instance is not None else 0
    return {"rays": len(R), "same_primitive": float(same.mean()), "agreement": float(agree.mean()),
            "hits_compared": int(both.sum()), "t_outside_1e-5": int((rel > 1e-5).sum()), "max_rel_t": float(rel.max()),
            "occlusion_agreement": occ, "curve_hits": curves}


@pytest.mark.parametrize("which", ["displaced_2m", "hair_1m_segments", "displaced_20m"])
def test_ray_gate_at_scale(built, ref, which):
    """north_star ray gate (>= 99.99 % hit / primID agreement, t within 1e-5 relative, occlusion flags) at the scale of
    the large configurations, against the live compiled reference (Embree, src/scene.h:89-91):
      displaced_2m      the C5 generator at 2 M triangles — the builder's chunked parallel passes and its stable
                        partitions run on every range >= 2^18 primitives (bvh_builder.cc), 4 Mi rays
      hair_1m_segments  the C3 CyHair file itself (50 000 strands x 20 segments) over the light stage, 4 Mi rays
      displaced_20m     the C5 OBJ itself (19.9 M triangles); minutes of OBJ parsing on both sides, so it runs only with
                        PBR_RUN_SLOW=1 (the log of the builder's run is under profiles/)"""
    import os
    if ref is None:
        pytest.skip("oracle/_ref did not travel")
    if which == "displaced_20m" and os.environ.get("PBR_RUN_SLOW", "0") != "1":
        pytest.skip("set PBR_RUN_SLOW=1 (parses a 1.5 GB OBJ twice)")
    rng = np.random.default_rng(2026)
    if which == "hair_1m_segments":
        files = [scenes.light_stage(), scenes.cyhair(50000, 21, center=(-2.5, 3.5, 0.0), radius=1.2, length=2.5, thickness=0.008)]
        lo, hi, curve_inst = (-5.0, 0.5, -2.5), (0.0, 6.0, 2.5), 3
    else:
        files = [scenes.displaced(2_000_000 if which == "displaced_2m" else 20_000_000)]
        lo, hi, curve_inst = (-9.0, 0.5, -9.0), (9.0, 17.0, 7.0), None
    sc = pb.Scene(files)
    ctx = sc.context()
    S = ref.scene(files)
    r = _big_gate(ctx, S, rng, 1 << 21, lo, hi, curve_inst)
    r["builder"] = os.environ.get("PBRGPU_BVH", "auto")
    common.dump_report("ray_gate_%s.json" % which, r)
    assert r["rays"] >= 4_000_000
    assert r["agreement"] >= 0.9999, r
    assert r["hits_compared"] > 1_000_000 and r["max_rel_t"] < 1e-3, r
    assert r["occlusion_agreement"] >= 0.9999, r
    if curve_inst is not None:
        assert r["curve_hits"] > 200_000, r
    sc.close()


JOB_WORKER = r"""
import os, sys, time
sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
import numpy as np
import pbrlab_b200 as pb
from pbrlab_b200 import scenes
rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
if rank == 0:
    ident = pb.nccl_unique_id()
    with open(idfile + ".tmp", "wb") as f: f.write(ident)
    os.replace(idfile + ".tmp", idfile)
else:
    while not os.path.exists(idfile): time.sleep(0.05)
    ident = open(idfile, "rb").read()
sc = pb.Scene([scenes.cornell()], device_ids=[rank])
ctx = sc.context()
ctx.nccl_init(ident, rank, world)
rgba, count = ctx.render(192, 160, 18, seed=99)          # collective: the library splits and reduces
np.savez(out, rgba=rgba, count=count, paths=ctx.stats()["paths"])
rgba2, count2 = ctx.render(192, 160, 5, seed=3)           # fewer samples than 2 * ranks leave uneven shares
np.savez(out.replace(".npz", "_b.npz"), rgba=rgba2, count=count2)
sc.close()
"""


def test_multi_process_job_matches_one_device(built, tmp_path):
    """north_star / SURVEY §8(e): a frame partitioned over the GPUs of one box, one process per GPU, the accumulators
    summed by ONE ncclReduce inside the library (pbrgpu_nccl_init + pbrgpu_render).  Rank 0's frame equals the
    single-device frame up to the order of the float additions; the other ranks keep their partial sums."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n < 4 else 4
    here = os.path.dirname(os.path.abspath(__file__))
    script = tmp_path / "job_worker.py"
    script.write_text(JOB_WORKER.format(root=os.path.dirname(here), here=here))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), idfile, str(tmp_path / ("r%d.npz" % r))],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    one = pb.Scene([scenes.cornell()], device_ids=[0])
    a, ca, _ = one.render(192, 160, 18, seed=99)
    r0 = np.load(tmp_path / "r0.npz")
    assert np.array_equal(r0["count"], ca) and np.all(ca == 18)
    assert np.allclose(r0["rgba"], a, rtol=1e-4, atol=1e-4), float(np.abs(r0["rgba"] - a).max())
    shares = [int(np.load(tmp_path / ("r%d.npz" % r))["paths"]) for r in range(world)]
    assert sum(shares) == 192 * 160 * 18 and max(shares) - min(shares) <= 192 * 160
    r1 = np.load(tmp_path / "r1.npz")
    assert np.all(r1["count"] == len(range(1, 18, world)))                 # a non-root rank keeps its own share
    b, cb, _ = one.render(192, 160, 5, seed=3)
    rb = np.load(tmp_path / "r0_b.npz")
    assert np.array_equal(rb["count"], cb) and np.allclose(rb["rgba"], b, rtol=1e-4, atol=1e-4)
    one.close()


def test_device_built_bvh_is_the_default_and_reports_the_host_builders_hits(built, monkeypatch):
    """pbrgpu_commit builds the triangle BVH on the GPU (csrc/bvh_device.cuh) unless told otherwise;
    the host binned-SAH builder stays as the parity reference: both trees report the same hits on the golden batch
    (different trees may break an exact tie between two triangles sharing an edge differently: counted)"""
    g = golden("cornell_rays.npz")
    rays = common.rays_from_f8(g["rays"])
    res = {}
    for mode in ("auto", "sah"):
        monkeypatch.setenv("PBRGPU_BVH", mode)
        sc = pb.Scene([scenes.cornell()])
        ctx = sc.context()
        ci = ctx.commit_info()
        assert ci["tri_builder"] == ("ploc (device)" if mode == "auto" else "sah (host)"), ci
        assert 0 < ci["tri_nodes"] < 362620 and 0 < ci["tri_depth"] <= 31 and ci["commit_s"] > 0
        res[mode] = (ctx.trace(rays), ctx.occluded(rays), ci)
        sc.close()
    monkeypatch.delenv("PBRGPU_BVH")
    (ha, oa, ca), (hs, osah, cs) = res["auto"], res["sah"]
    same = (ha["instance_id"] == hs["instance_id"]) & (ha["geom_id"] == hs["geom_id"]) & (ha["prim_id"] == hs["prim_id"])
    assert (~same).sum() <= max(1, len(rays) // 10000), int((~same).sum())
    assert np.array_equal(ha["t"][same], hs["t"][same]) and np.array_equal(ha["u"][same], hs["u"][same])
    assert np.array_equal(oa, osah)
    common.dump_report("bvh_builders_cornell.json", {"device_ploc": ca, "host_sah": cs, "rays": len(rays),
                                                     "different_primitive": int((~same).sum())})
