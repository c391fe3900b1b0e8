#!/bin/bash
# strong scaling of the fixed C2 frame on 1, 2, 4, 8 GPUs of one box as the driver launches it, + the multi-GPU tests
TAG=${1:-r2x}; NS=${2:-"1 2 4 8"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "multi_process or multi_device" > gpurun_out/${TAG}_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$n.json 2> gpurun_out/${TAG}_bench_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_bench_n$n.json") if l.startswith("{")][-1])
    print("n=%d value %.1f e2e %.1f ms/step %.1f scaling %s spp/gpu %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], d["config"]["spp_per_gpu"]))
except Exception as e:
    print("n=$n: no line", e)
PY
done
# one process, one context over several devices (ncclCommInitAll inside the library)
if [ -n "$MULTI_DEVICE" ]; then
  for n in $MULTI_DEVICE; do
    timeout 600 python scripts/time_multi_device.py $n 1024 2>&1 | grep "devices\|multi-device" | tee -a gpurun_out/${TAG}_multi_device.log
  done
fi
