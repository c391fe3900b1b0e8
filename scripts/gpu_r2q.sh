#!/bin/bash
# triangle grid for random-walk segments: parity tests + on/off sweep
TAG=${1:-r2q}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_tune.log
IFS=';' read -ra SWEEPS <<< "${TUNE_ARGS:-WALK_GRID=0,1}"
for sw in "${SWEEPS[@]}"; do
  echo "## ${SPP:-128} $sw" >> gpurun_out/${TAG}_tune.log
  timeout 900 python scripts/tune.py 1920 1080 ${SPP:-128} $sw 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
done
cat gpurun_out/${TAG}_tune.log
