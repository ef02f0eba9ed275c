#!/bin/bash
# usage: gpu_ncu_pipe.sh <ne> <regex> <skip> <count> <tag>
ne=$1; rx=$2; skip=$3; cnt=$4; tag=$5
python bench.py --ne $ne --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip --launch-count $cnt -o gpurun_out/ncu_$tag -f python bench.py --ne $ne --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_$tag.log 2>&1; tail -2 gpurun_out/ncu_$tag.log
