mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --trace gpurun_out/trace_i.txt
} > gpurun_out/run_i.log 2>&1
cat gpurun_out/run_i.log | cut -c1-400
