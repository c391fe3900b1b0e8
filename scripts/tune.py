"""Tuning sweep over the PBRGPU_* launch knobs: python scripts/tune.py W H SPP KEY=v1,v2,... [KEY=...]
Each combination runs in this process with a fresh context (the knobs are read at pbrgpu_create)."""
import itertools, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrlab_b200 as pb
from pbrlab_b200 import scenes
w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
keys = []; vals = []
files = [scenes.cornell()]
for a in sys.argv[4:]:
    k, v = a.split("=")
    if k == "FILES":
        files = v.split(","); continue
    if k == "SCENE" and v == "c3":
        files = [scenes.light_stage(), scenes.cyhair(50000, 21, center=(-2.5, 3.5, 0.0), radius=1.2, length=2.5, thickness=0.008)]; continue
    if k == "SCENE" and v == "c4":
        files = [scenes.cornell(), scenes.cyhair(50000, 21, center=(-2.5, 6.0, 0.0), radius=1.2, length=2.5, thickness=0.008)]; continue
    if k == "SCENE" and v.startswith("c5"):
        files = [scenes.displaced(int(v.split(":")[1]) if ":" in v else 20_000_000)]; continue
    keys.append(k); vals.append(v.split(","))
for combo in itertools.product(*vals):
    for k, v in zip(keys, combo):
        os.environ["PBRGPU_" + k] = v
    S = pb.Scene(files)
    ctx = S.context()
    ctx.render(w, h, 2)
    best = 1e9
    for rep in range(3):
        t = time.time(); ctx.render(w, h, spp); best = min(best, time.time() - t)
    ctx.set_profiling(True); ctx.render(w, h, spp); st = ctx.stats(); ctx.set_profiling(False)
    rays = st["closest_rays"] + st["shadow_rays"] + st["sss_rays"]
    print(" ".join("%s=%s" % kv for kv in zip(keys, combo)), "| %.4f s %.1f Msamples/s %.1f Mrays/s | closest %.1f any %.1f shade %.1f sss %.1f regen %.1f ms | launches %d | rays c %.1fM s %.1fM w %.1fM skipped %.1fM" % (
        best, w * h * spp / best * 1e-6, rays / best * 1e-6, st["trace_closest_ms"], st["trace_any_ms"], st["shade_ms"], st["sss_ms"], st["regen_ms"], st["kernel_launches"], st["closest_rays"] * 1e-6, st["shadow_rays"] * 1e-6, st["sss_rays"] * 1e-6, st["sss_skipped"] * 1e-6), flush=True)
    S.close()
