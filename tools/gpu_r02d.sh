#!/bin/bash
mkdir -p gpurun_out
tag=${1:-r02d}
timeout 300 python tools/probe_accum_error.py > gpurun_out/${tag}_accum_error.log 2>&1
cat gpurun_out/${tag}_accum_error.log | tail -20
timeout 600 python -m pytest tests/test_widen_gpu.py -m gpu -q -rfEs --no-header -p no:cacheprovider -x -k "graphed or resume or capturable or sliding" > gpurun_out/${tag}_graph_tests.log 2>&1
tail -5 gpurun_out/${tag}_graph_tests.log
