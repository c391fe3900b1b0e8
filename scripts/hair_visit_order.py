"""MEASUREMENT (CPU, no GPU needed): node visits and candidate tests per hair ray of the curve BVH under
different child-visit orders — the static octant order the traversal kernels use, best-first over the whole tree (the
lower bound), depth-first with a node's hit children sorted by entry distance, and that without a stored entry distance
per child.  Rays: secondary rays leaving the hair ball of the C3 / C4 workloads in random directions.
    python scripts/hair_visit_order.py"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, pbrlab_b200 as pb, emulbind
from pbrlab_b200 import scenes
hair = scenes.cyhair(50000, 21, center=(-2.5, 6.0, 0.0), radius=1.2, length=2.5, thickness=0.008)
host = pb.Scene([hair], commit_to_device=False)
E = emulbind.Emul(host.flat())
rng = np.random.default_rng(7)
# secondary rays from hair surfaces: camera rays first
lo, hi = E.bounds()
n = 200000
c = np.array([-2.5, 6.0, 0.0], np.float32)
eye = np.array([c[0], c[1], 14.0], np.float32)
tgt = c + rng.normal(size=(n,3)).astype(np.float32) * 1.5
d = tgt - eye; d /= np.linalg.norm(d, axis=1, keepdims=True)
cam = pb.make_rays(np.tile(eye,(n,1)), d.astype(np.float32))
hits, st = E.trace(cam, stats=True)
hit = hits["instance_id"] != 0xFFFFFFFF
print("camera rays: hit frac %.3f nodes/ray %.2f prims/ray %.2f" % (hit.mean(), st[0]/n, st[1]/n))
P = cam["org"][hit] + hits["t"][hit,None]*cam["dir"][hit]
d2 = rng.normal(size=P.shape).astype(np.float32); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
sec = pb.make_rays(P, d2, tmin=1e-3)
h2, st2 = E.trace(sec, stats=True)
m = len(sec)
print("secondary rays from hair (%d): the traversal as built (children around the origin first) nodes/ray %.2f prims/ray %.2f" % (m, st2[0]/m, st2[1]/m))
st3 = np.zeros(2, np.uint64); t3 = np.zeros(m, np.float32)
for mode, name in ((0, "best-first (global heap)"), (1, "depth-first, children sorted per node"), (2, "same, a node's leaves before its inner children"), (3, "same, but a child is only culled by its own boxes (no stored entry distance)"),
                   (4, "the static octant order (the engine's), this tool's count"), (5, "octant order, children whose box holds the ray origin first")):
    E.lib.emul_trace_curves_best_first(E.h, sec.ctypes.data_as(C.c_void_p), C.c_uint64(m), t3.ctypes.data_as(C.c_void_p), st3.ctypes.data_as(C.c_void_p), C.c_int(mode))
    print("   %-40s nodes/ray %.2f prims/ray %.2f" % (name, st3[0]/m, st3[1]/m))
