#!/bin/bash
# One GPU session: quick parity subset, variant benches at ne120, ncu launch list of the main library.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_limiter.py tests/test_gpu_parity.py "tests/test_gpu_driver.py::test_device_driver_matches_oracle" "tests/test_gpu_driver.py::test_large_mesh_matches_oracle" "tests/test_gpu_driver.py::test_negative_thickness_is_reported" -m gpu -x -q -s 2>&1 | tail -25 > gpurun_out/s1_tests.log
tail -12 gpurun_out/s1_tests.log
bash tools/variant_bench.sh 120 6 main $@
bash tools/variant_ll.sh 120 main
