"""One small render for profiling: python scripts/render_once.py W H SPP [warmup]"""
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
