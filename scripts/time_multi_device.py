"""One process, one context over N devices (pbrgpu_create with N ids: the library splits the frame's samples over
its devices, replicates the scene and sums with one grouped ncclReduce): python scripts/time_multi_device.py N [spp]
Prints Msamples/s of pbrlab::Render() (host RenderLayer) on the C2 frame, best of 3, next to the 1-device figure."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pbrlab_b200 as pb  # noqa: E402
from pbrlab_b200 import scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
w, h = 1920, 1080
res = {}
for devs in ([0], list(range(n))):
    S = pb.Scene([scenes.cornell()], device_ids=devs)
    S.render_layer(w, h, 8)
    best = 1e9
    for _ in range(3):
        t = time.time()
        rgba, count, sec = S.render_layer(w, h, spp)
        best = min(best, time.time() - t)
    assert int(count.min()) == spp and int(count.max()) == spp
    res[len(devs)] = w * h * spp / best * 1e-6
    print("devices %d: %.4f s  %.1f Msamples/s (pbrlab::Render(), host RenderLayer, %d spp)" % (len(devs), best, res[len(devs)], spp), flush=True)
    S.close()
print("multi-device context: %.2fx of one device on %d devices (efficiency %.3f)" % (res[n] / res[1], n, res[n] / res[1] / n))
