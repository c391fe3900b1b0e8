#!/bin/bash
TAG=${1:-r1n}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 1920 1080 32 CLEAR_MARCH=1,3,6,12 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
PBRGPU_CLEAR_CELLS=64e6 timeout 600 python scripts/tune.py 1920 1080 32 CLEAR_MARCH=1,6 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 1920 1080 32 REFILL_SSS=12,16,20,24,28 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
timeout 600 python scripts/tune.py 1920 1080 32 REFILL=8,12,16,20 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
cat gpurun_out/${TAG}_tune.log
# C5: one full ncu capture of the HBM-resident traversal
timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 12 --launch-count 2 \
    -k regex:'TraceClosest|TraceAny' -f -o gpurun_out/${TAG}_c5_full \
    python scripts/render_once.py 3840 2160 4 0 c5 > gpurun_out/${TAG}_ncu_c5.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_c5.log
