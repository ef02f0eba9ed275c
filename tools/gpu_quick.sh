#!/bin/bash
# quick check: stage parity + bench + launch list at ne120 for the main library (and variants given as args)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_limiter.py tests/test_gpu_parity.py "tests/test_gpu_driver.py::test_device_driver_matches_oracle" -m gpu -x -q 2>&1 | tail -4
bash tools/variant_bench.sh 120 6 main $@
bash tools/variant_ll.sh 120 main | head -12
