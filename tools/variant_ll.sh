#!/bin/bash
# usage: variant_ll.sh <ne> name...   ncu launch list (durations + dram bytes) of one remap cycle per kernel variant
ne=$1; shift
for v in "$@"; do
  TSE_CUDA_LIB=$( [ $v = main ] && echo $PWD/transport_se_b200/libtse_cuda.so || echo $PWD/build/variants/libtse_$v.so ) ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ll_${v}_ne$ne.csv python bench.py --ne $ne --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2> gpurun_out/ll_$v.err || tail -3 gpurun_out/ll_$v.err
  echo "== $v ne$ne"; python tools/launch_list.py gpurun_out/ll_${v}_ne$ne.csv | sed -n 1,9p
done
