#!/bin/bash
# usage: variant_bench.sh <ne> <steps> name...   quick bench of kernel variants built by build_variant.sh
ne=$1; steps=$2; shift 2
for v in "$@"; do
  TSE_CUDA_LIB=$( [ $v = main ] && echo $PWD/transport_se_b200/libtse_cuda.so || echo $PWD/build/variants/libtse_$v.so ) python bench.py --ne $ne --steps $steps --warmup 2 --no-cpu --no-e2e > gpurun_out/vb_$v.json 2> gpurun_out/vb_$v.err || tail -3 gpurun_out/vb_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/vb_$v.json')); print('$v ne$ne ms/tracer-step', round(d['ms_per_tracer_step'],2), 'step frac', round(d['step_hbm']['frac'],3), 'stage avg ms', round(d['roofline']['avg_launch_ms'],3), 'euler', round(d['timers_ms']['euler_step'],1), 'remap', round(d['timers_ms']['vertical_remap'],1), 'mass drift', d['mass_drift_rel'], d['mass_drift_rel_checkerboard'])"
done
