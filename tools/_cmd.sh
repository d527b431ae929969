mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
python bench.py --no-cpu-baseline --trace gpurun_out/trace_y.txt
python bench.py --no-cpu-baseline
} > gpurun_out/run_y.log 2>&1
cat gpurun_out/run_y.log | cut -c1-260
