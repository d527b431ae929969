#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02e}
timeout 300 python tools/probe_accum_error.py > gpurun_out/${tag}_accum_error.log 2>&1
cat gpurun_out/${tag}_accum_error.log | tail -20
timeout 600 python tools/probe_parity_depth.py 32 64 128 > gpurun_out/${tag}_parity_depth.log 2>&1
grep "====" gpurun_out/${tag}_parity_depth.log
timeout 1200 python -m pytest tests -m gpu -q -rfEs --no-header -p no:cacheprovider -s > gpurun_out/${tag}_gpu_tests.log 2>&1
grep -E "passed|failed|FAILED|ERROR|\[full-size\]|\[bf16 growth\]|\[golden\]|\[fwd\]|\[bwd\]" gpurun_out/${tag}_gpu_tests.log | tail -60
