#!/bin/bash
# usage: build_variant.sh <name> <nvcc -D flags...>   -> build/variants/libtse_<name>.so (select with TSE_CUDA_LIB)
name=$1; shift
mkdir -p build/variants
TSE_NVCC_FLAGS="$*" python -c "
import sys; sys.path.insert(0, 'transport_se_b200'); import _build
_build.build_cuda(force=True, verbose=True, out='build/variants/libtse_$name.so')" 2>&1 | grep -E "error|Used 2[0-9][0-9]|spill stores, [1-9]" | sort | uniq -c | head -20
