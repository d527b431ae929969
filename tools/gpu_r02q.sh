#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02q}
timeout 1500 python -m pytest tests -m gpu -q -rfEs --no-header -p no:cacheprovider -s > gpurun_out/${tag}_gpu_tests.log 2>&1
grep -E "passed|failed|FAILED|ERROR|\[bottleneck\]|\[transposed\]" gpurun_out/${tag}_gpu_tests.log | tail -30
