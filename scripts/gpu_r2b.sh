#!/bin/bash
# Round 2 GPU session: parity tests, knob sweeps, default bench (other_configs), launch list, ncu capture.
# Usage: scripts/gpu_r2b.sh TAG [steps...]   steps: tests tune bench benchfast launches ncu
TAG=${1:-r2b}; shift
STEPS=${@:-tests tune bench launches ncu}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  grep -v "^$" gpurun_out/${TAG}_tests.log | tail -${TEST_TAIL:-60}
fi
if has tune; then
  : > gpurun_out/${TAG}_tune.log
  run() { echo "## $*" >> gpurun_out/${TAG}_tune.log; timeout 600 python scripts/tune.py "$@" 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log; }
  IFS=';' read -ra SWEEPS <<< "${TUNE_ARGS:-1920 1080 128 POOL_MI=32}"
  for sw in "${SWEEPS[@]}"; do run $sw; done
  cat gpurun_out/${TAG}_tune.log
fi
if has benchfast; then
  timeout 900 python bench.py --steps 3 --warmup 3 --no-other-configs > gpurun_out/${TAG}_benchfast.json 2> gpurun_out/${TAG}_benchfast.err
  echo "benchfast exit $?"; cut -c1-1500 gpurun_out/${TAG}_benchfast.json
fi
if has bench; then
  timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; cut -c1-1500 gpurun_out/${TAG}_bench.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench_ref.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --no-other-configs \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
  tail -2 gpurun_out/${TAG}_launches_bench.log | cut -c1-300
fi
if has ncu; then
  # one launch of every kernel family out of the middle of a frame (skip the first iterations: the pool is filling)
  timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip ${NCU_SKIP:-56} --launch-count 9 \
    -k regex:'TraceClosest|ShadeSurface|SssWalk|SssExit|TraceAny|Retire' -f -o gpurun_out/${TAG}_full \
    python scripts/render_once.py 1920 1080 128 0 ${NCU_SCENE:-} > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log
fi
ls -la gpurun_out | tail -8
