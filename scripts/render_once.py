"""One small render for profiling: python scripts/render_once.py W H SPP [warmup] [c3|c4|c5|files...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrlab_b200 as pb  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402

w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
warm = int(sys.argv[4]) if len(sys.argv) > 4 else 0
files = sys.argv[5:] if len(sys.argv) > 5 else [scenes.cornell()]
hair = dict(n_strands=50000, n_points=21, radius=1.2, length=2.5, thickness=0.008)
if files in (["c1"], ["c2"]):
    files = [scenes.cornell()]
elif files == ["c3"]:
    files = [scenes.light_stage(), scenes.cyhair(center=(-2.5, 3.5, 0.0), **hair)]
elif files == ["c4"]:
    files = [scenes.cornell(), scenes.cyhair(center=(-2.5, 6.0, 0.0), **hair)]
elif files == ["c5"]:
    files = [scenes.displaced(20_000_000)]
S = pb.Scene(files)
ctx = S.context()
for _ in range(warm):
    ctx.render(w, h, spp)
t = time.time()
rgba, count = ctx.render(w, h, spp)
dt = time.time() - t
st = ctx.stats()
rays = st["closest_rays"] + st["shadow_rays"] + st["sss_rays"]
print("render %dx%dx%d: %.4f s  %.2f Msamples/s  %.2f Mrays/s  launches %d" % (w, h, spp, dt, w * h * spp / dt * 1e-6, rays / dt * 1e-6, st["kernel_launches"]))
