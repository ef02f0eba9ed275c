#!/bin/bash
# usage: variant_ll_q.sh <ne> <qsize> name...   launch list per kernel variant at a given tracer count (items per CTA = qsize/2):
# two tracer counts separate a kernel's per-item rate from its per-CTA overhead
ne=$1; q=$2; shift 2
for v in "$@"; do
  TSE_CUDA_LIB=$( [ $v = main ] && echo $PWD/transport_se_b200/libtse_cuda.so || echo $PWD/build/variants/libtse_$v.so ) ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ll_${v}_ne${ne}_q$q.csv python bench.py --ne $ne --qsize $q --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2> gpurun_out/ll_$v.err || tail -3 gpurun_out/ll_$v.err
  echo "== $v ne$ne q$q"; python tools/launch_list.py gpurun_out/ll_${v}_ne${ne}_q$q.csv | sed -n 1,6p
done
