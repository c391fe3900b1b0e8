"""Traversal statistics (nodes visited / primitives tested per ray) for camera rays of a workload:
python scripts/trace_stats.py c2|c3"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrlab_b200 as pb
from pbrlab_b200 import scenes
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
files = [scenes.cornell()] if wl == "c2" else [scenes.light_stage(), scenes.cyhair(50000, 21, center=(-2.5, 3.5, 0.0), radius=1.2, length=2.5, thickness=0.008)]
S = pb.Scene(files); ctx = S.context()
bmin, bmax = ctx.bounds()
w, h = 1920, 1080
hs = bmax[0] - bmin[0]; vs = bmax[1] - bmin[1]
if hs > vs: vs = hs * h / w
else: hs = vs * w / h
eye = np.array([(bmax[0] + bmin[0]) * 0.5, (bmax[1] + bmin[1]) * 0.5, bmax[2] + hs * 0.5 * np.sqrt(3.0)], np.float32)
rng = np.random.default_rng(1)
n = 1 << 20
px = rng.random((n, 2)).astype(np.float32)
tgt = np.stack([eye[0] - hs * 0.5 + hs * px[:, 0], eye[1] + vs * 0.5 - vs * px[:, 1], np.full(n, bmax[2], np.float32)], 1)
d = tgt - eye; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = pb.make_rays(np.tile(eye, (n, 1)), d.astype(np.float32))
hits = ctx.trace(rays); st = ctx.stats()
hit = hits["instance_id"] != 0xFFFFFFFF
print("camera rays: hit %.3f  nodes/ray %.2f  prims/ray %.2f  kernel %.3f ms" % (hit.mean(), st["nodes_visited"] / n, st["prims_tested"] / n, st["trace_closest_ms"]))
ids, cnt = np.unique(hits["instance_id"][hit], return_counts=True); print("instances hit", dict(zip(ids.tolist(), cnt.tolist())))
P = rays["org"][hit] + hits["t"][hit, None] * rays["dir"][hit]
d2 = rng.normal(size=P.shape).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
r2 = pb.make_rays(P, d2, tmin=1e-3)
h2 = ctx.trace(r2); st = ctx.stats(); m = len(r2)
print("secondary rays: hit %.3f nodes/ray %.2f  prims/ray %.2f  kernel %.3f ms (%d rays)" % ((h2["instance_id"] != 0xFFFFFFFF).mean(), st["nodes_visited"] / m, st["prims_tested"] / m, st["trace_closest_ms"], m))
# rays that start on hair only
hh = hits["instance_id"][hit] == hits["instance_id"].max() if wl == "c3" else None
if hh is not None and hh.sum() > 0:
    r3 = r2[hh]; h3 = ctx.trace(r3); st = ctx.stats(); m = len(r3)
    print("secondary rays from hair: nodes/ray %.2f prims/ray %.2f kernel %.3f ms (%d rays)" % (st["nodes_visited"] / m, st["prims_tested"] / m, st["trace_closest_ms"], m))
