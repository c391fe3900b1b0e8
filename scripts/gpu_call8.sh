#!/bin/bash
TAG=${1:-r1o}
mkdir -p gpurun_out
for wl in c3 c4; do
  if [ $wl = c3 ]; then W=1920; H=1080; S=16; else W=3840; H=2160; S=4; fi
  timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 12 --launch-count 2 \
    -k regex:'TraceClosest|TraceAny' -f -o gpurun_out/${TAG}_${wl}_full \
    python scripts/render_once.py $W $H $S 0 $wl > gpurun_out/${TAG}_ncu_${wl}.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu_${wl}.log
done
