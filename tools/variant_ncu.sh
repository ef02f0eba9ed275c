#!/bin/bash
# usage: variant_ncu.sh <ne> name...   ncu --set full capture (with source) of one launch of each k_tile op per kernel variant
ne=$1; shift
for v in "$@"; do
  TSE_CUDA_LIB=$( [ $v = main ] && echo $PWD/transport_se_b200/libtse_cuda.so || echo $PWD/build/variants/libtse_$v.so ) ncu --set full --clock-control none --import-source on -k regex:k_pipe --launch-skip 12 --launch-count 6 -o gpurun_out/ncu_${v}_ne$ne -f python bench.py --ne $ne --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_$v.log 2>&1; tail -2 gpurun_out/ncu_$v.log
done
