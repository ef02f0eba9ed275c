#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_limiter.py tests/test_gpu_parity.py "tests/test_gpu_driver.py::test_device_driver_matches_oracle" "tests/test_gpu_driver.py::test_large_mesh_matches_oracle" "tests/test_gpu_driver.py::test_negative_thickness_is_reported" -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/s2_tests.log
tail -8 gpurun_out/s2_tests.log
bash tools/variant_bench.sh 120 6 main
python bench.py --ne 30 --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/plain_ne30.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_vertical_remap|k_nbr_minmax" --launch-skip 6 --launch-count 3 -o gpurun_out/ncu_r2a_ne30 -f python bench.py --ne 30 --steps 1 --warmup 2 --no-e2e --no-cpu > gpurun_out/ncu_r2a.log 2>&1; tail -2 gpurun_out/ncu_r2a.log
