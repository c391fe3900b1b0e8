#!/bin/bash
# compute-sanitizer over small renders (1 Mi-slot pool): memcheck on the Cornell + hair and the Cornell scene (8 spp: the
# sample order, the walk stream, thin spreading and FinishPathsKernel all take part), racecheck (shared-memory hazards
# of the warp-cooperative candidate rejection and of the block reservations) on both
TAG=${1:-san}
mkdir -p gpurun_out
export PBRGPU_POOL_MI=1
timeout 1000 compute-sanitizer --tool memcheck --print-limit 5 python scripts/render_once.py 160 120 8 0 c4 > gpurun_out/${TAG}_mem_c4.log 2>&1; tail -4 gpurun_out/${TAG}_mem_c4.log
timeout 1000 compute-sanitizer --tool racecheck --print-limit 5 python scripts/render_once.py 96 64 6 0 c4 > gpurun_out/${TAG}_race_c4.log 2>&1; tail -4 gpurun_out/${TAG}_race_c4.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/render_once.py 160 120 8 0 > gpurun_out/${TAG}_mem_c2.log 2>&1; tail -3 gpurun_out/${TAG}_mem_c2.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python scripts/render_once.py 96 64 6 0 > gpurun_out/${TAG}_race_c2.log 2>&1; tail -3 gpurun_out/${TAG}_race_c2.log
