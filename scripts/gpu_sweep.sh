#!/bin/bash
# knob sweeps on a workload: scripts/gpu_sweep.sh TAG "W H SPP" "SCENE=c3 KEY=a,b" ["SCENE=c4 KEY=..."] ...
TAG=$1; DIM=$2; shift; shift
mkdir -p gpurun_out
: > gpurun_out/${TAG}_tune.log
for sw in "$@"; do
  timeout 600 python scripts/tune.py $DIM $sw 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
done
cat gpurun_out/${TAG}_tune.log
