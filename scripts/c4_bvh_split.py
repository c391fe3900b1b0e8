"""C4 (Cornell + hair ball): how many of a ray's node visits and primitive tests belong to the triangle BVH and how many
to the curve BVH — what one merged top level could save at most.  Same camera / secondary ray batches as bench.py's
traversal_counters, traced through (a) the whole scene, (b) the triangles alone, (c) the hair alone."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrlab_b200 as pb  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402

hair = scenes.cyhair(50000, 21, center=(-2.5, 6.0, 0.0), radius=1.2, length=2.5, thickness=0.008)
full = pb.Scene([scenes.cornell(), hair])
ctx = full.context()
w, h = 3840, 2160
rng = np.random.default_rng(7)
bmin, bmax = ctx.bounds()
hs = bmax[0] - bmin[0]; vs = bmax[1] - bmin[1]
if hs > vs: vs = hs * h / w
else: hs = vs * w / h
eye = np.array([(bmax[0] + bmin[0]) * 0.5, (bmax[1] + bmin[1]) * 0.5, bmax[2] + hs * 0.5 * np.sqrt(3.0)], np.float32)
nr = 1 << 20
px = rng.random((nr, 2)).astype(np.float32)
tgt = np.stack([eye[0] - hs * 0.5 + hs * px[:, 0], eye[1] + vs * 0.5 - vs * px[:, 1], np.full(nr, bmax[2], np.float32)], 1)
d = tgt - eye; d /= np.linalg.norm(d, axis=1, keepdims=True)
cam = pb.make_rays(np.tile(eye, (nr, 1)), d.astype(np.float32))
hits = ctx.trace(cam)
hit = hits["instance_id"] != 0xFFFFFFFF
P = cam["org"][hit] + hits["t"][hit, None] * cam["dir"][hit]
d2 = rng.normal(size=P.shape).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
sec = pb.make_rays(P, d2, tmin=1e-3)
on_hair = (hits["geom_id"][hit] >= 0) & (hits["t"][hit] > 0)   # placeholder mask (all hits)
# secondary rays that START on the hair ball: origin within its bounding sphere (centre, radius + length)
c = np.array([-2.5, 6.0, 0.0], np.float32)
inside = np.linalg.norm(P - c, axis=1) < 1.2 + 2.5
batches = {"camera": cam, "secondary": sec, "secondary from inside the hair ball": sec[inside]}
scn = {"whole scene": full, "triangles alone": pb.Scene([scenes.cornell()]), "hair alone": pb.Scene([hair])}
for bname, rays in batches.items():
    print("%s rays (%d):" % (bname, len(rays)))
    for sname, S in scn.items():
        cx = S.context()
        cx.trace(rays)
        st = cx.stats()
        print("   %-16s nodes/ray %6.2f  prims/ray %6.2f" % (sname, st["nodes_visited"] / len(rays), st["prims_tested"] / len(rays)))
