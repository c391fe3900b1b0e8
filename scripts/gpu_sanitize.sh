#!/bin/bash
# compute-sanitizer over small renders (1 Mi-slot pool): memcheck on the Cornell + hair and the Cornell scene,
# racecheck (shared-memory hazards of the warp-cooperative candidate rejection) on Cornell + hair
mkdir -p gpurun_out
export PBRGPU_POOL_MI=1
timeout 1000 compute-sanitizer --tool memcheck --print-limit 5 python scripts/render_once.py 160 120 2 0 c4 > gpurun_out/san_mem_c4.log 2>&1; tail -4 gpurun_out/san_mem_c4.log
timeout 1000 compute-sanitizer --tool racecheck --print-limit 5 python scripts/render_once.py 96 64 1 0 c4 > gpurun_out/san_race_c4.log 2>&1; tail -4 gpurun_out/san_race_c4.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/render_once.py 160 120 2 0 > gpurun_out/san_mem_c2.log 2>&1; tail -3 gpurun_out/san_mem_c2.log
