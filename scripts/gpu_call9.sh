#!/bin/bash
TAG=${1:-r1p}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 1920 1080 32 SCENE=c3 RIBBON_LANES=4,8,12 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 3840 2160 8 SCENE=c4 RIBBON_LANES=8 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
cat gpurun_out/${TAG}_tune.log
python - <<'PY' 2>&1 | grep -v "add shape\|num face\|^$" | tee gpurun_out/${TAG}_gather.log
import pbrlab_b200 as pb
from pbrlab_b200 import scenes
S = pb.Scene([scenes.cornell()]); ctx = S.context()
for ws in (8 << 20, 32 << 20, 64 << 20, 96 << 20, 256 << 20, 1 << 30, 4 << 30):
    print("working set %6d MB:" % (ws >> 20), " ".join("chains=%d %.0f GB/s" % (c, ctx.measure_gather(ws, 2048, c)) for c in (1, 2, 4, 8)))
PY
