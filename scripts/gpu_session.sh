#!/bin/bash
# One gpurun call: GPU parity tests, knob comparison, the default bench line, the ncu launch list and one full ncu
# capture per kernel family.  Everything lands in gpurun_out/<tag>_*.  Usage: scripts/gpu_session.sh TAG [steps...]
# steps: tests tune bench launches ncu (default: all)
TAG=${1:-run}; shift
STEPS=${@:-tests tune bench launches ncu}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
if has tune; then
  # TUNE_ARGS: one or more knob sweeps separated by ';' (each is a product of KEY=v1,v2 lists)
  : > gpurun_out/${TAG}_tune.log
  IFS=';' read -ra SWEEPS <<< "${TUNE_ARGS:-SSS_SKIP=0,1}"
  for sw in "${SWEEPS[@]}"; do
    timeout 600 python scripts/tune.py 1920 1080 32 $sw 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
  done
  cat gpurun_out/${TAG}_tune.log
fi
if has bench; then
  timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
  cat gpurun_out/${TAG}_bench_ref.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
  tail -2 gpurun_out/${TAG}_launches_bench.log | cut -c1-300
fi
if has ncu; then
  # one launch of every kernel family out of the middle of a frame (skip the first iterations: the pool is filling)
  timeout 1200 ncu --set full --clock-control none --import-source on --launch-skip 44 --launch-count 8 \
    -k regex:'TraceClosest|ShadeSurface|SssWalk|SssExit|TraceAny|Regenerate' -f -o gpurun_out/${TAG}_full \
    python scripts/render_once.py 1920 1080 128 0 > gpurun_out/${TAG}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu.log
fi
ls -la gpurun_out | tail -20
