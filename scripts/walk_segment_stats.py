"""MEASUREMENT (CPU, no GPU needed): the random-walk segments that reach the ray engine on the Cornell
scene — how long they are, how many BVH nodes a closest-hit query tests for them, and how many of those nodes lie on
the chain from the root on which exactly one inner child is hit (what a traversal that starts below the root saves).
    PBRGPU_BVH_TRIS=ploc python scripts/walk_segment_stats.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pbrlab_b200 as pb  # noqa: E402
import emulbind  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402

host = pb.Scene([scenes.cornell()], commit_to_device=False)
E = emulbind.Emul(host.flat())
# camera rays aimed at Lucy: paths that enter the medium
lo, hi = E.bounds()
rng = np.random.default_rng(11)
n = 60000
eye = np.array([(lo[0] + hi[0]) * 0.5, (lo[1] + hi[1]) * 0.5, hi[2] + (hi[1] - lo[1]) * 0.5 * np.sqrt(3.0)], np.float32)
tgt = np.stack([rng.uniform(lo[0], hi[0], n), rng.uniform(lo[1], hi[1], n), np.full(n, hi[2])], 1).astype(np.float32)
d = tgt - eye; d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = pb.make_rays(np.tile(eye, (n, 1)), d.astype(np.float32))
seeds = np.stack([rng.integers(0, 2**62, n, dtype=np.uint64)] * 2, 1)
cap = 400000
E.lib.emul_record_segments(C.c_uint64(cap))
E.radiance(rays, seeds)
E.lib.emul_fetch_segments.restype = C.c_uint64
seg = np.zeros((cap, 8), np.float32)
got = int(E.lib.emul_fetch_segments(seg.ctypes.data_as(C.c_void_p), C.c_uint64(cap)))
E.lib.emul_record_segments(C.c_uint64(0))
seg = seg[:min(got, cap)]
print("traced walk segments recorded: %d" % len(seg))
r = pb.make_rays(seg[:, 0:3].copy(), seg[:, 4:7].copy(), tmin=0.0)
r["tmin"][:] = seg[:, 3]; r["tmax"][:] = seg[:, 7]
hits, st = E.trace(r, stats=True)
m = len(r)
print("TraverseBvh: nodes/segment %.2f  triangle tests/segment %.2f  hit fraction %.3f" % (st[0] / m, st[1] / m, (hits["instance_id"] != 0xFFFFFFFF).mean()))
out = np.zeros(4, np.uint64)
E.lib.emul_tri_chain_stats(E.h, r.ctypes.data_as(C.c_void_p), C.c_uint64(m), out.ctypes.data_as(C.c_void_p))
print("depth-first count: nodes/segment %.2f, of which on the one-child chain from the root %.2f (%.0f %%); chains that end in a node with leaf hits only: %.0f %%" % (
    out[1] / m, out[2] / m, 100.0 * out[2] / out[1], 100.0 * out[3] / m))
L = seg[:, 7]
print("segment length quantiles (scene units): 50 %% %.4f  90 %% %.4f  99 %% %.4f; scene height %.2f" % (np.quantile(L, 0.5), np.quantile(L, 0.9), np.quantile(L, 0.99), hi[1] - lo[1]))

# ---- how many of the traced segments a finer clearance representation could answer: distance from the segment's origin
# to the mesh (nearest of ~16 points sampled on every triangle, so an over-estimate by up to the sample spacing) against
# the segment's length
try:
    from scipy.spatial import cKDTree
    flat = host.flat()
    V = np.asarray(flat.verts)[:, :3].astype(np.float64)
    T = np.asarray(flat.vidx).reshape(-1, 3)
    A, B, Cc = V[T[:, 0]], V[T[:, 1]], V[T[:, 2]]
    pts = []
    k = 5
    for i in range(k + 1):
        for j in range(k + 1 - i):
            u, v = i / k, j / k
            pts.append(A * (1 - u - v) + B * u + Cc * v)
    P = np.concatenate(pts)
    edge = np.median(np.linalg.norm(B - A, axis=1))
    tree = cKDTree(P)
    dist, _ = tree.query(seg[:, 0:3].astype(np.float64), workers=-1)
    spacing = edge / k
    miss = hits["instance_id"] == 0xFFFFFFFF
    for margin in (0.0, spacing):
        ok = (L * 1.02 < dist - margin)
        print("median triangle edge %.4f, sample spacing %.4f: segments shorter than the distance to the mesh (margin %.4f): %.1f %% of the traced ones (%.1f %% of the traced ones miss)" % (
            edge, spacing, margin, 100.0 * ok.mean(), 100.0 * miss.mean()))
except ImportError:
    print("scipy not available: skipped the distance statistics")

# ---- the same question for a concrete representation: bricks of F x F x F fine cells under the coarse cells, each fine
# cell holding  (distance from its centre to the mesh) - (half its diagonal),  one sphere step from the segment's origin
try:
    cell = (hi[1] - lo[1]) / 501.0          # the clearance field's cell on this scene (505 cells incl. 4 spare layers)
    O = seg[:, 0:3].astype(np.float64)
    for F in (1, 2, 4, 8):
        f = cell / F
        centre = (np.floor(O / f) + 0.5) * f
        dc, _ = tree.query(centre, workers=-1)
        bound = dc - 0.866 * f - spacing
        ok = L * 1.02 <= bound
        print("fine cell = cell / %d (%.4f): one sphere step answers %.1f %% of the segments that are traced today" % (F, f, 100.0 * ok.mean()))
except NameError:
    pass
