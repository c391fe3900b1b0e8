#!/bin/bash
# one mid-frame launch of the random-walk kernel under ncu for each PBRGPU_WALK_SLOTS in $SLOTS (default "1 2")
TAG=${1:-r2v}
mkdir -p gpurun_out
for k in ${SLOTS:-1 2}; do
  PBRGPU_WALK_SLOTS=$k PBRGPU_OVERLAP=0 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 12 --launch-count 1 \
    -k regex:'SssWalk' -f -o gpurun_out/${TAG}_walk_slots$k python scripts/render_once.py 1920 1080 128 0 > gpurun_out/${TAG}_ncu_slots$k.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_slots$k.log
done
ls -la gpurun_out | tail -5
