#!/bin/bash
# usage: gpu_sanitize.sh <tool> variants...   compute-sanitizer on the smoke case (ne=4, qsize=4: one tracer step + remap) per library variant
tool=$1; shift
for v in "$@"; do
  lib=$( [ $v = main ] && echo $PWD/transport_se_b200/libtse_cuda.so || echo $PWD/build/variants/libtse_$v.so )
  TSE_CUDA_LIB=$lib timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_${tool}_$v.log 2>&1
  echo "== $tool $v: exit $?"; grep -E "smoke:|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error:" gpurun_out/san_${tool}_$v.log | sort | uniq -c | head -12
done
