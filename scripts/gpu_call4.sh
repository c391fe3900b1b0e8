#!/bin/bash
TAG=${1:-r1k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 1920 1080 32 SCENE=c3 RIBBON_LANES=1,4,8,12 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 3840 2160 8 SCENE=c4 RIBBON_LANES=8 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
cat gpurun_out/${TAG}_tune.log
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 12 --launch-count 2 \
    -k regex:'TraceClosest|TraceAny' -f -o gpurun_out/${TAG}_c3_full \
    python scripts/render_once.py 1920 1080 16 0 c3 > gpurun_out/${TAG}_ncu_c3.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_c3.log
