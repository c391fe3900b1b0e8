#!/bin/bash
# device BVH builder: parity tests (triangle BVHs now come from the GPU), the 20 M-triangle ray gate, C5 commit times
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
grep -v "^$" gpurun_out/${TAG}_tests.log | tail -14
: > gpurun_out/${TAG}_tune.log
run() { echo "## $*" >> gpurun_out/${TAG}_tune.log; timeout 600 python scripts/tune.py "$@" 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log; }
run 1920 1080 128 BVH_TRIS=sah,ploc,auto
run 1920 1080 32 SCENE=c4 BVH_TRIS=sah,auto BVH_CURVES=sah,ploc
run 1920 1080 32 SCENE=c3 BVH_CURVES=sah,ploc
cat gpurun_out/${TAG}_tune.log
# C5: scene load and commit with the device builder vs the host builder, then the 20 M-triangle gate
PBRGPU_VERBOSE_COMMIT=1 timeout 900 python scripts/render_once.py 3840 2160 16 1 c5 > gpurun_out/${TAG}_c5_device.log 2>&1
grep -i "commit\|device BVH\|render " gpurun_out/${TAG}_c5_device.log
PBRGPU_BVH=sah PBRGPU_VERBOSE_COMMIT=1 timeout 900 python scripts/render_once.py 3840 2160 16 1 c5 > gpurun_out/${TAG}_c5_host.log 2>&1
grep -i "commit\|BuildBvh8\|render " gpurun_out/${TAG}_c5_host.log
PBR_RUN_SLOW=1 timeout 1500 python -m pytest tests -m gpu -q -k "displaced_20m" > gpurun_out/${TAG}_gate20m.log 2>&1
tail -3 gpurun_out/${TAG}_gate20m.log
cat gpurun_out/ray_gate_displaced_20m.json
