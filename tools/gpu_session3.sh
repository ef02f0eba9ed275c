#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_limiter.py tests/test_gpu_parity.py "tests/test_gpu_driver.py::test_device_driver_matches_oracle" "tests/test_gpu_driver.py::test_negative_thickness_is_reported" "tests/test_gpu_driver.py::test_mass_is_order_independent" -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_tests.log
tail -6 gpurun_out/s3_tests.log
bash tools/variant_bench.sh 120 6 main $@
bash tools/variant_ll.sh 120 main
