#!/bin/bash
# One GPU visit: tests, bench, ncu launch list, ncu full capture of the dominant kernel.
mkdir -p gpurun_out
tag=${1:-r01}
timeout 900 python -m pytest tests -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/${tag}_gpu_tests.log 2>&1
echo "pytest exit=$?" >> gpurun_out/${tag}_gpu_tests.log
timeout 200 python __graft_entry__.py smoke >> gpurun_out/${tag}_gpu_tests.log 2>&1
timeout 900 python bench.py --trace gpurun_out/${tag}_trace.txt > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit=$?" >> gpurun_out/${tag}_bench.err
if [ "$2" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
fi
if [ "$2" != "noncu" ] && [ "$2" != "lite" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_fprop -s 136 -c 3 -o gpurun_out/${tag}_conv3_fprop_full \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_wgrad_kernel -s 68 -c 3 -o gpurun_out/${tag}_conv3_wgrad_full \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/${tag}_ncu_full.log 2>&1
fi
grep -E "passed|failed|FAILED|\[fwd\]|\[bwd\]|\[golden\]|smoke" gpurun_out/${tag}_gpu_tests.log | tail -40
cat gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
