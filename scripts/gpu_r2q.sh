#!/bin/bash
# parity tests + knob sweeps: TUNE_ARGS = sweeps separated by ';', RES = "W H SPP" (default 1920 1080 128), RES2 optional second size
TAG=${1:-r2q}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
if [ -z "$NO_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
: > gpurun_out/${TAG}_tune.log
IFS=';' read -ra SWEEPS <<< "${TUNE_ARGS:-OVERLAP=1}"
for res in "${RES:-1920 1080 ${SPP:-128}}" "$RES2" "$RES3"; do
  [ -z "$res" ] && continue
  for sw in "${SWEEPS[@]}"; do
    echo "## $res $sw" >> gpurun_out/${TAG}_tune.log
    timeout 900 python scripts/tune.py $res $sw 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
  done
done
cat gpurun_out/${TAG}_tune.log
