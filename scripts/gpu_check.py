"""First-contact GPU check (run under gpurun): ray / shading / path parity against the compiled reference when
oracle/_ref travelled with the snapshot, then a timed render.  Prints a JSON summary to gpurun_out/gpu_check.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbrlab_b200 as pb  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402
import refbind  # noqa: E402

out = {}
obj = scenes.cornell()
t = time.time()
S = pb.Scene([obj])
out["scene_commit_s"] = time.time() - t
ctx = S.context()
R = refbind.RefLib() if refbind.available() else None
RS = R.scene([obj]) if R else None


def camera_rays(n, seed, w=512, h=512):
    cam = np.zeros(8, np.float32)
    bmin, bmax = ctx.bounds()
    hs = bmax[0] - bmin[0]
    vs = bmax[1] - bmin[1]
    if hs > vs:
        vs = hs * h / w
    else:
        hs = vs * w / h
    eye = np.array([(bmax[0] + bmin[0]) * 0.5, (bmax[1] + bmin[1]) * 0.5, bmax[2] + hs * 0.5 * np.sqrt(3.0)], np.float32)
    rng = np.random.default_rng(seed)
    px = rng.random((n, 2)).astype(np.float32)
    tgt = np.stack([eye[0] - hs * 0.5 + hs * px[:, 0], eye[1] + vs * 0.5 - vs * px[:, 1], np.full(n, bmax[2], np.float32)], 1)
    d = tgt - eye
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return pb.make_rays(np.tile(eye, (n, 1)), d.astype(np.float32))


n = 1 << 20
rays = camera_rays(n, 1)
t = time.time(); hits = ctx.trace(rays); out["trace_1M_wall_s"] = time.time() - t
st = ctx.stats()
out["trace_1M_kernel_ms"] = st["trace_closest_ms"]
out["trace_nodes_per_ray"] = st["nodes_visited"] / n
out["trace_prims_per_ray"] = st["prims_tested"] / n
if RS:
    rf, rid = RS.trace(pb.rays_to_f8(rays))
    same = (hits["instance_id"] == rid[:, 0]) & (hits["geom_id"] == rid[:, 1]) & (hits["prim_id"] == rid[:, 2])
    out["trace_agree"] = float(same.mean())
    hit = same & (rid[:, 0] != 0xFFFFFFFF)
    out["trace_t_relerr_max"] = float(np.max(np.abs(hits["t"][hit] - rf[hit, 0]) / rf[hit, 0]))

# secondary rays: bounce the camera hits diffusely
m = hits["instance_id"] != 0xFFFFFFFF
P = rays["org"][m] + hits["t"][m, None] * rays["dir"][m]
rng = np.random.default_rng(5)
d2 = rng.normal(size=P.shape).astype(np.float32)
d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
rays2 = pb.make_rays(P, d2, tmin=1e-3)
h2 = ctx.trace(rays2)
occ = ctx.occluded(rays2)
out["secondary_rays"] = int(len(rays2))
if RS:
    rf2, rid2 = RS.trace(pb.rays_to_f8(rays2))
    same2 = (h2["instance_id"] == rid2[:, 0]) & (h2["prim_id"] == rid2[:, 2])
    out["secondary_agree"] = float(same2.mean())
    rocc = RS.occluded(pb.rays_to_f8(rays2))
    out["occluded_agree"] = float((occ == rocc).mean())

# per-path radiance: wavefront vs megakernel vs reference
k = 1 << 18
rr = camera_rays(k, 9)
seeds = np.stack([np.arange(k, dtype=np.uint64) + 7, np.arange(k, dtype=np.uint64) * 3 + 1], 1)
t = time.time(); a = ctx.radiance(rr, seeds); out["radiance_wavefront_s"] = time.time() - t
st = ctx.stats()
out["radiance_rays_per_path"] = (st["closest_rays"] + st["shadow_rays"] + st["sss_rays"]) / k
t = time.time(); b = ctx.radiance(rr, seeds, mega=True); out["radiance_mega_s"] = time.time() - t
err = np.abs(a - b).max(axis=1)
out["wavefront_vs_mega_agree_1e-5"] = float((err <= 1e-5 * np.maximum(1e-2, np.abs(b).max(axis=1))).mean())
out["mean_radiance_gpu"] = a.mean(axis=0).tolist()
if RS:
    c = RS.radiance(pb.rays_to_f8(rr), seeds)
    err = np.abs(a - c).max(axis=1)
    out["wavefront_vs_ref_agree_1e-4"] = float((err <= 1e-4 * np.maximum(1e-2, np.abs(c).max(axis=1))).mean())
    out["mean_radiance_ref"] = c.mean(axis=0).tolist()

# timed renders
for (w, h, spp) in [(512, 512, 64), (1920, 1080, 16)]:
    ctx.render(w, h, 4)
    t = time.time(); rgba, count = ctx.render(w, h, spp); dt = time.time() - t
    st = ctx.stats()
    rays_total = st["closest_rays"] + st["shadow_rays"] + st["sss_rays"]
    key = "%dx%dx%d" % (w, h, spp)
    out[key] = {"seconds": dt, "Msamples_s": w * h * spp / dt * 1e-6, "Mrays_s": rays_total / dt * 1e-6,
                "rays_per_sample": rays_total / (w * h * spp), "launches": st["kernel_launches"],
                "mean_rgb": (rgba[..., :3] / count[..., None]).mean(axis=(0, 1)).tolist(),
                "nan_pixels": int(np.isnan(rgba).any(axis=2).sum())}
if RS:
    t = time.time(); rr_, rc_, sec = RS.render(512, 512, 16)
    out["ref_512x512x16"] = {"seconds": sec, "Msamples_s": 512 * 512 * 16 / sec * 1e-6, "threads": R.num_threads(),
                             "mean_rgb": (rr_[..., :3] / rc_[..., None]).mean(axis=(0, 1)).tolist()}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
