#!/bin/bash
# multi-GPU bench line, as the driver launches it: scripts/gpu_scale.sh TAG N [extra bench args]
TAG=${1:-scale}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 2 --warmup 3 "$@" > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
echo "exit $?"; grep '^{' gpurun_out/${TAG}_n$N.json | cut -c1-1500; tail -5 gpurun_out/${TAG}_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_n${N}_ref.json 2>> gpurun_out/${TAG}_n$N.err
grep '^{' gpurun_out/${TAG}_n${N}_ref.json | cut -c1-600
