#!/bin/bash
# Evidence of a HEAD in one gpurun call: GPU tests, the ncu launch list of a short bench run, one full ncu capture per
# kernel family for C2 / C5 / C4 (summarised ON the box: the reports are too large to bring back together), the
# default bench line and the reference arm.  Usage: scripts/gpu_final.sh TAG [steps...]   steps: tests launches ncu bench
TAG=${1:-final}; shift
STEPS=${@:-tests launches ncu bench}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
if has tests; then
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline --no-other-configs \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
  tail -1 gpurun_out/${TAG}_launches_bench.log | cut -c1-200
fi
if has ncu; then
  KRE='TraceClosest|ShadeSurface|ShadeHair|SssWalk|SssExit|TraceAny'
  cap() {  # workload w h spp skip count
    PBRGPU_OVERLAP=0 timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip $5 --launch-count $6 \
      -k regex:"$KRE" -f -o gpurun_out/${TAG}_full_$1 python scripts/render_once.py $2 $3 $4 0 $1 > gpurun_out/${TAG}_ncu_$1.log 2>&1
    tail -1 gpurun_out/${TAG}_ncu_$1.log
    python scripts/ncu_families.py gpurun_out/${TAG}_full_$1.ncu-rep $1 "$7" > /dev/null 2>&1
    python scripts/ncu_summary.py gpurun_out/${TAG}_full_$1.ncu-rep "ncu summary ${TAG} $1: $7" > gpurun_out/${TAG}_ncu_$1.md 2>/dev/null
  }
  HEADREV=$(cat gpurun_out/../.head_rev 2>/dev/null || echo "this HEAD")
  cap c2 1920 1080 128 44 8 "$HEADREV, 1920x1080x128 spp, one launch per kernel mid-frame"
  cap c5 3840 2160 16 44 8 "$HEADREV, 3840x2160x16 spp, one launch per kernel mid-frame"
  cap c4 3840 2160 16 49 9 "$HEADREV, 3840x2160x16 spp, one launch per kernel mid-frame"
  cp profiles/ncu_families.json gpurun_out/${TAG}_ncu_families.json
  rm -f gpurun_out/${TAG}_full_c5.ncu-rep gpurun_out/${TAG}_full_c4.ncu-rep   # (64 MiB comes back at most; the C2 report is kept)
fi
if has bench; then
  # with the freshly written profiles/ncu_families.json in place (roofline.traffic of this HEAD)
  timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  cut -c1-400 gpurun_out/${TAG}_bench.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
  cut -c1-300 gpurun_out/${TAG}_bench_ref.json
fi
ls -la gpurun_out | tail -15
