#!/bin/bash
# longest-paths-first sample order: parity tests + on/off sweep at the per-GPU frame sizes of the 8-, 2- and 1-GPU split
TAG=${1:-r2p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
: > gpurun_out/${TAG}_tune.log
for spp in 128 512 1024; do
  echo "## 1920 1080 $spp" >> gpurun_out/${TAG}_tune.log
  timeout 900 python scripts/tune.py 1920 1080 $spp ${TUNE_ARGS:-ORDER=0,1} 2>&1 | grep "|" >> gpurun_out/${TAG}_tune.log
done
cat gpurun_out/${TAG}_tune.log
PBRGPU_TRACE_ITERATIONS=1 timeout 300 python scripts/render_once.py 1920 1080 128 0 2>&1 | grep "^iter" | tail -80 > gpurun_out/${TAG}_iters.log
tail -60 gpurun_out/${TAG}_iters.log | cut -c1-150
