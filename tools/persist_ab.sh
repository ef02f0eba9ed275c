#!/bin/bash
# A/B of the persistent k_pipe launch (TSE_PIPE_PERSIST=1) against one CTA per chunk (0): parity tests, bench, launch list
mkdir -p gpurun_out
for p in 1 0; do
  echo "== TSE_PIPE_PERSIST=$p"
  TSE_PIPE_PERSIST=$p timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_limiter.py "tests/test_gpu_driver.py::test_device_driver_matches_oracle" "tests/test_gpu_driver.py::test_large_mesh_matches_oracle" -m gpu -x -q 2>&1 | tail -2
  TSE_PIPE_PERSIST=$p timeout 300 python bench.py --ne 120 --steps 6 --warmup 2 --no-cpu --no-e2e > gpurun_out/pab_$p.json 2> gpurun_out/pab_$p.err || tail -3 gpurun_out/pab_$p.err
  python -c "
import json; d=json.load(open('gpurun_out/pab_$p.json')); print('persist=$p ms/tracer-step', round(d['ms_per_tracer_step'],2), 'stage avg ms', round(d['roofline']['avg_launch_ms'],3), 'hash', d.get('field_hash_all'), 'mass drift', d['mass_drift_rel'])"
done
TSE_PIPE_PERSIST=1 timeout 400 bash tools/variant_ll.sh 120 main | head -14
