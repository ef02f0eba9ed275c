#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py "tests/test_gpu_driver.py::test_device_driver_matches_oracle" "tests/test_gpu_driver.py::test_mass_is_order_independent" -m gpu -x -q 2>&1 | tail -5
python bench.py --ne 30 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/plain_s4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_nbr_minmax" --launch-skip 2 --launch-count 2 -o gpurun_out/ncu_r2c_ne30 -f python bench.py --ne 30 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_r2c.log 2>&1; tail -2 gpurun_out/ncu_r2c.log
bash tools/variant_ll.sh 120 main
