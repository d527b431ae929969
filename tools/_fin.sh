#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEsxX --no-header -p no:cacheprovider > gpurun_out/r02i_gpu_tests.log 2>&1
tail -3 gpurun_out/r02i_gpu_tests.log
timeout 200 python __graft_entry__.py smoke >> gpurun_out/r02i_gpu_tests.log 2>&1
grep smoke gpurun_out/r02i_gpu_tests.log | cut -c1-200
timeout 900 python tools/bench_medformer.py --batch 2 --side 128 --schedule graph > gpurun_out/r02i_medformer_bench_b2.json 2> /dev/null
cat gpurun_out/r02i_medformer_bench_b2.json
