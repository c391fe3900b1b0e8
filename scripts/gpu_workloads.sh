#!/bin/bash
# One gpurun call: GPU parity tests, then bench lines (reduced spp; throughput is spp-independent) for the C3, C4 and
# C5 workloads with the CPU reference timed beside each.  Usage: scripts/gpu_workloads.sh TAG [steps...]
TAG=${1:-wl}; shift
STEPS=${@:-tests c1 c3 c4 c5}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1
  echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
  tail -3 gpurun_out/${TAG}_tests.log
fi
run() {  # workload spp ref_spp
  timeout 1500 python bench.py --workload $1 --spp $2 --ref-spp $3 --steps 2 --warmup 3 \
    > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
  echo "bench $1 exit $?"; grep '^{' gpurun_out/${TAG}_bench_$1.json | cut -c1-1800
}
if has c1; then run c1 64 16; fi
if has c3; then run c3 64 8; fi
if has c4; then run c4 16 4; fi
if has c5; then run c5 16 4; fi
ls -la gpurun_out | tail
