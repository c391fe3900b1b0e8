"""Timing sweep helper: python scripts/sweep.py W H SPP pool_spp[,pool_spp...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrlab_b200 as pb
from pbrlab_b200 import scenes
w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pools = [int(x) for x in sys.argv[4].split(",")]
S = pb.Scene([scenes.cornell()])
ctx = S.context()
ctx.render(w, h, 4)
for pool in pools:
    ctx.set_wave_spp(pool)
    best = 1e9
    for rep in range(3):
        t = time.time(); ctx.render(w, h, spp); dt = time.time() - t
        best = min(best, dt)
    st = ctx.stats()
    rays = st["closest_rays"] + st["shadow_rays"] + st["sss_rays"]
    print("pool_spp %3d (%.1fM slots): %.4f s  %.1f Msamples/s  %.1f Mrays/s  launches %d" % (pool, pool * w * h / 1e6, best, w * h * spp / best * 1e-6, rays / best * 1e-6, st["kernel_launches"]), flush=True)
