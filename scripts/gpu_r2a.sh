#!/bin/bash
# Round 2, first GPU session: parity tests (new: cancel, clearance identity, scale gates, output stage, KAT counts),
# knob sweeps for the drain tail / pool size / pipelined diffuse shading, the default bench line with other_configs,
# launch list + one full ncu capture.  Usage: scripts/gpu_r2a.sh TAG [steps...]
TAG=${1:-r2a}; shift
STEPS=${@:-tests tune bench launches ncu}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
nproc >> gpurun_out/${TAG}_smi.log
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  tail -40 gpurun_out/${TAG}_tests.log
fi
if has tune; then
  : > gpurun_out/${TAG}_tune.log
  run() { echo "## $*" >> gpurun_out/${TAG}_tune.log; timeout 600 python scripts/tune.py "$@" 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log; }
  # the per-GPU load of the 8-GPU strong split (128 spp) and of 2 GPUs (512 spp): drain tail + pool size
  run 1920 1080 128 DIFFUSE_PIPE=0,1 WALK_TARGET_MI=0,16
  run 1920 1080 128 WALK_TARGET_MI=4,8,32,64
  run 1920 1080 128 POOL_DIV=2,4,16 
  run 1920 1080 128 POOL_DIV=1 POOL_MI=8,16
  run 1920 1080 128 WALK_BOUNCES_MAX=256,8192
  run 1920 1080 512 DIFFUSE_PIPE=0,1
  run 1920 1080 128 DIFFUSE_PIPE=1 DIFFUSE_BLOCKS=4,5,6
  run 1920 1080 128 DIFFUSE_PIPE=1 DIFFUSE_THREADS=256 DIFFUSE_BLOCKS=2,3
  # C1: half a pool of samples
  run 512 512 64 WALK_TARGET_MI=0,16 POOL_MIN_MI=1,2,4
  cat gpurun_out/${TAG}_tune.log
fi
if has bench; then
  timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; cut -c1-3000 gpurun_out/${TAG}_bench.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench_ref.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --no-other-configs \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
  tail -2 gpurun_out/${TAG}_launches_bench.log | cut -c1-300
fi
if has ncu; then
  # one launch of every kernel family out of the middle of a frame (skip the first iterations: the pool is filling)
  timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 44 --launch-count 8 \
    -k regex:'TraceClosest|ShadeSurface|ShadeDiffuse|SssWalk|SssExit|TraceAny' -f -o gpurun_out/${TAG}_full \
    python scripts/render_once.py 1920 1080 128 0 > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log
fi
ls -la gpurun_out | tail -12
