#!/bin/bash
# A/B of library builds: scripts/gpu_variants.sh TAG "name1 name2 ..." (scratch/<name>.so replaces lib/libpbrgpu.so for one tune.py line each)
TAG=${1:-var}; NAMES=${2:-base}
mkdir -p gpurun_out; : > gpurun_out/${TAG}_variants.log
cp pbrlab_b200/lib/libpbrgpu.so /tmp/libpbrgpu.orig.so
for n in $NAMES; do
  cp scratch/$n.so pbrlab_b200/lib/libpbrgpu.so
  echo "## $n" >> gpurun_out/${TAG}_variants.log
  timeout 600 python scripts/tune.py ${RES:-1920 1080 128} ${TUNE_ARGS:-OVERLAP=1} 2>&1 | grep "|" >> gpurun_out/${TAG}_variants.log
done
cp /tmp/libpbrgpu.orig.so pbrlab_b200/lib/libpbrgpu.so
cat gpurun_out/${TAG}_variants.log
