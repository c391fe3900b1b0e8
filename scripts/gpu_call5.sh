#!/bin/bash
TAG=${1:-r1l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 1920 1080 32 SCENE=c3 CURVE_SPLIT=1,2,4 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 3840 2160 8 SCENE=c4 CURVE_SPLIT=1,2,4 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
cat gpurun_out/${TAG}_tune.log
